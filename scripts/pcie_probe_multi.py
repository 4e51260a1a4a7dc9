"""Aggregate host <-> device copy bandwidth with k of the node's GPUs copying at once (GPU box, torchrun, one rank per
GPU): contiguous pinned copies, device -> host, host -> device and both together, for k = 1, 2, 4, 8 participants.
Evidence for DESIGN.md 7: what bounds the END-TO-END number when several GPUs of one host return results at once.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/pcie_probe_multi.py
"""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import torch.distributed as dist

from seqikpy_b200.batch import bind_to_gpu_numa

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
bound = None if "--no-numa-bind" in sys.argv else bind_to_gpu_numa(local)
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_out, n_in = 6000 * 1000 * 34, 6000 * 1000 * 15               # one step of the benchmark: 816 MB out, 360 MB in
h_out = torch.empty(n_out, dtype=torch.float32, pin_memory=True)
d_out = torch.empty(n_out, dtype=torch.float32, device="cuda")
h_in = torch.empty(n_in, dtype=torch.float32, pin_memory=True)
d_in = torch.empty(n_in, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
cur = torch.cuda.current_stream()


def run(fn, active, reps=4):
    if active:
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if active:
        for _ in range(reps):
            fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps if active else 0.0], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


def d2h():
    h_out.copy_(d_out, non_blocking=True)


def h2d():
    d_in.copy_(h_in, non_blocking=True)


def both():
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        h_out.copy_(d_out, non_blocking=True)
    with torch.cuda.stream(s2):
        d_in.copy_(h_in, non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)


k = 1
while k <= world:
    active = rank < k
    t_out, t_in, t_both = run(d2h, active), run(h2d, active), run(both, active)
    if rank == 0:
        gb_out, gb_in = k * n_out * 4 / 1e9, k * n_in * 4 / 1e9
        print(json.dumps({"gpus_copying": k, "d2h_GBs_aggregate": round(gb_out / t_out * 1e3, 1), "h2d_GBs_aggregate": round(gb_in / t_in * 1e3, 1),
                          "both_ms": round(t_both, 2), "both_d2h_GBs_aggregate": round(gb_out / t_both * 1e3, 1),
                          "both_h2d_GBs_aggregate": round(gb_in / t_both * 1e3, 1), "numa_bound_cpus_rank0": None if bound is None else len(bound)}), flush=True)
    k *= 2
dist.destroy_process_group()
