"""Kernel time of schedule 3 (frame-parallel blocks) against schedule 2 over batch sizes and resident warps. Dev tool (GPU box)."""
import sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import os
import numpy as np, torch
from seqikpy_b200 import _native as _N
if os.environ.get("SEQIK_LIB"):
    _N.LIB_PATH = Path(os.environ["SEQIK_LIB"]).resolve()
from seqikpy_b200 import synthetic as S, engine
from seqikpy_b200.batch import chain_param_table
from seqikpy_b200.kinematic_chain import KinematicChainSeq

size, bounds, init = S.chain_constants()
chain = KinematicChainSeq(bounds, list(S.LEGS), size)
base = S.to_chains(torch.from_numpy(S.make_trials(range(32), 1000)).cuda())          # 192 chains x 1000 frames


def run(n_trial, n_frame, sched, r=0, reps=5, variant=0):
    n_chain = n_trial * 6
    reps_c = (n_chain + 191) // 192
    pose = base[:, :min(n_frame, 1000)].repeat(reps_c, max(1, n_frame // 1000), 1, 1)[:n_chain].contiguous()
    n_frame = pose.shape[1]
    params = torch.from_numpy(chain_param_table(chain, init, S.LEGS, reps_c * 32)[:n_chain]).cuda()
    ang = torch.empty((n_chain, n_frame, 7), device="cuda"); fk = torch.empty((n_chain, n_frame, 9, 3), device="cuda")
    for _ in range(2):
        engine.leg_solve(pose, params, angles=ang, fk=fk, schedule=sched, chains_per_warp=r, want_stats=False, block_variant=variant)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        engine.leg_solve(pose, params, angles=ang, fk=fk, schedule=sched, chains_per_warp=r, want_stats=False, block_variant=variant)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"trials": n_trial, "chains": n_chain, "frames": n_frame, "sched": sched, "variant": variant, "r": r, "ms": round(ms, 4),
                      "Glf_s": round(n_chain * n_frame / ms / 1e6, 2), "GBs": round(n_chain * n_frame * 196 / ms / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
    if mode == "quick":
        for n_trial in (100, 1000, 1250, 10000):
            run(n_trial, 1000, 2)
            run(n_trial, 1000, 3)
    elif mode == "s3":
        for n_trial in (100, 1000, 1250, 10000):
            run(n_trial, 1000, 3)
    elif mode == "variants":
        for n_trial in (100, 400, 1000, 10000):
            for variant in (1, 2):
                run(n_trial, 1000, 3, variant=variant)
    elif mode == "r":
        for n_trial in (1000, 1250, 10000):
            for r in (10, 12, 13, 14, 15, 16, 17, 18):
                run(n_trial, 1000, 3, r)
    elif mode == "one":
        run(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]) if len(sys.argv) > 5 else 0, reps=2)
    elif mode == "c5":
        run(100, 100000, 2, reps=2)
        run(100, 100000, 3, reps=2)
