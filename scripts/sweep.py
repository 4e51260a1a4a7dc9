"""Kernel-time sweep over (trials, frames, schedule, chains per warp) on one GPU. Dev tool (GPU box)."""
import sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from seqikpy_b200 import synthetic as S, engine
from seqikpy_b200.batch import chain_param_table
from seqikpy_b200.kinematic_chain import KinematicChainSeq

size, bounds, init = S.chain_constants()
chain = KinematicChainSeq(bounds, list(S.LEGS), size)
base = S.to_chains(torch.from_numpy(S.make_trials(range(16), 1000)).cuda())          # 96 chains x 1000 frames


def run(n_trial, n_frame, sched, cpw, reps=3, gate=0, flags=None, trip=0):
    kw = {} if flags is None else {"flags": flags}
    n_chain = n_trial * 6
    reps_c = (n_chain + 95) // 96
    pose = base[:, :min(n_frame, 1000)].repeat(reps_c, max(1, n_frame // 1000), 1, 1)[:n_chain].contiguous()
    n_frame = pose.shape[1]
    params = torch.from_numpy(chain_param_table(chain, init, S.LEGS, reps_c * 16)[:n_chain]).cuda()
    ang = torch.empty((n_chain, n_frame, 7), device="cuda"); fk = torch.empty((n_chain, n_frame, 9, 3), device="cuda")
    for _ in range(2):
        engine.leg_solve(pose, params, angles=ang, fk=fk, schedule=sched, chains_per_warp=cpw, want_stats=False, gate=gate, trip_period=trip, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        engine.leg_solve(pose, params, angles=ang, fk=fk, schedule=sched, chains_per_warp=cpw, want_stats=False, gate=gate, trip_period=trip, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"trials": n_trial, "chains": n_chain, "frames": n_frame, "sched": sched, "cpw": cpw, "gate": gate, "trip": trip, "flags": flags, "ms": round(ms, 3),
                      "Mlf_s": round(n_chain * n_frame / ms / 1e3, 1)}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "gate":
        for n_trial in (100, 1000, 1250, 10000):
            for gate in (1, 2, 3, 4, 5, 6, 8):
                run(n_trial, 500, 2, 0, gate=gate)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "newton":
        for n_trial in (100, 1000, 1250, 10000):
            for flags in (0x7F, 0xFF):
                for gate in (1, 2, 3, 4):
                    run(n_trial, 500, 2, 0, gate=gate, flags=flags)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "trip":
        for n_trial in (100, 1000, 1250, 10000):
            for gate, trip in ((2, 1), (1, 1), (1, 2), (1, 3), (2, 2), (2, 3), (3, 2), (3, 3), (2, 4), (1, 4)):
                run(n_trial, 500, 2, 0, gate=gate, trip=trip)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        for n_trial in (3000, 5000, 10000):
            for cpw in (4, 5, 6, 8):
                run(n_trial, 500, 2, cpw)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "cpw":
        for n_trial in (100, 200, 400, 1000, 1250, 2500):
            for cpw in (1, 2, 3, 4, 5, 6, 7, 8):
                run(n_trial, 500, 2, cpw)
        sys.exit(0)
    for n_trial in (100, 400, 1000, 1250, 2500, 5000, 10000):
        for sched, cpws in ((2, (1, 2, 4, 8)), (1, (0,))):
            for cpw in cpws:
                if sched == 2 and n_trial * 6 / cpw > 40000:
                    continue
                run(n_trial, 500, sched, cpw)
