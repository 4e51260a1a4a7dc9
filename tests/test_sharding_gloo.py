"""The N > 1 path on CPU: two gloo ranks shard trials exactly like bench.py / the multi-GPU driver does
(contiguous trial ranges, no data-path collective, one gather of results; max-over-ranks timing reduction)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_trial, n_frame, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from seqikpy_b200 import synthetic as S
    from seqikpy_b200.batch import shard_range
    import hostsim_build as H
    import model_trf2 as M
    lo, hi = shard_range(n_trial, rank, world)
    size, bounds, init = S.chain_constants()
    # this rank's shard, solved by the host build of the solver core (stands in for the GPU kernel on CPU)
    res = np.zeros((hi - lo, 6, n_frame, 7), dtype=np.float32)
    for i, tr in enumerate(range(lo, hi)):
        pose = S.make_trial(tr, n_frame)
        for li, leg in enumerate(S.LEGS):
            seg = [size[f"{leg}_{s}"] for s in ("Coxa", "Femur", "Tibia", "Tarsus")]
            from oracle.seqik_oracle import DOF_ORDER
            lb = [bounds[f"{leg}_{d}"][0] for d in DOF_ORDER]
            ub = [bounds[f"{leg}_{d}"][1] for d in DOF_ORDER]
            res[i, li] = H.solve_chain(pose[:, li], seg, lb, ub, M.null_sq_from_seeds(init[leg]), M.seeds7(init[leg]), gn_mask=63)[0]
    # the only communication: scalar statistics (timing max, leg-frame count), then a gather of results on rank 0
    ms = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    count = torch.tensor([res.shape[0] * 6 * n_frame], dtype=torch.int64)
    dist.all_reduce(count)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((lo, hi, res), gathered, dst=0)
    if rank == 0:
        full = np.concatenate([g[2] for g in sorted(gathered, key=lambda g: g[0])])
        np.savez(Path(out_dir) / "gathered.npz", angles=full, ms=ms.numpy(), count=count.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharding_equals_single_process(tmp_path):
    n_trial, n_frame, world = 3, 40, 2
    mp.spawn(_worker, args=(world, _free_port(), n_trial, n_frame, str(tmp_path)), nprocs=world, join=True)
    out = np.load(tmp_path / "gathered.npz")
    assert out["angles"].shape == (n_trial, 6, n_frame, 7)
    assert float(out["ms"][0]) == 11.0 and int(out["count"][0]) == n_trial * 6 * n_frame
    # single-process result for the same trials: shards must concatenate to it bit for bit
    sys.path.insert(0, str(ROOT / "tests"))
    from seqikpy_b200 import synthetic as S
    from oracle.seqik_oracle import DOF_ORDER
    import hostsim_build as H
    import model_trf2 as M
    size, bounds, init = S.chain_constants()
    for tr in range(n_trial):
        pose = S.make_trial(tr, n_frame)
        for li, leg in enumerate(S.LEGS):
            seg = [size[f"{leg}_{s}"] for s in ("Coxa", "Femur", "Tibia", "Tarsus")]
            lb = [bounds[f"{leg}_{d}"][0] for d in DOF_ORDER]
            ub = [bounds[f"{leg}_{d}"][1] for d in DOF_ORDER]
            ref = H.solve_chain(pose[:, li], seg, lb, ub, M.null_sq_from_seeds(init[leg]), M.seeds7(init[leg]), gn_mask=63)[0]
            assert np.array_equal(out["angles"][tr, li], ref)
