"""Generic 7-DOF leg IK (seqik_leg_solve_generic_f32/_f64) on the synthetic workload: kernel-side timing with CUDA
events for several chain counts, one JSON line each.  (GPU box)

    python scripts/generic_bench.py [n_frame=100] [chain counts ...=6,600,6000,60000]
    GEN_CPW=1,2,4,8,0  the same with these chains-per-warp settings (0 = the launcher's own rule), GEN_DTYPES=float32,float64
"""
import os
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

from seqikpy_b200 import engine, synthetic as S
from seqikpy_b200.kinematic_chain import KinematicChainGeneric

n_frame = int(sys.argv[1]) if len(sys.argv) > 1 else 100
counts = [int(a) for a in sys.argv[2:]] or [6, 600, 6000, 60000]
size, bounds, init = S.chain_constants()
chain = KinematicChainGeneric(bounds, list(S.LEGS), size)
# the stage-4 seed vector is in yaw, pitch, roll order; the reference hands it to the generic chain positionally, which for the
# locomotion constants puts some slots outside their bounds (scipy raises, and so do we) -- permute it into chain order here
perm = [0, 3, 1, 2, 4, 5, 6, 7, 8]
rows6 = np.stack([chain.pack_chain_params(leg, np.asarray(init[leg]["stage_4"], dtype=float)[perm]) for leg in S.LEGS])
n_unique = 8
pose = np.stack([S.make_trial(tr, 1000)[:n_frame] for tr in range(n_unique)]).transpose(0, 2, 1, 3, 4)   # (trial, leg, frame, 5, 3)
pose = pose.reshape(n_unique * 6, n_frame, 5, 3)
cpws = [int(c) for c in os.environ.get("GEN_CPW", "0").split(",")]
sched = int(os.environ.get("GEN_SCHED", "0"))
dtypes = [getattr(torch, d) for d in os.environ.get("GEN_DTYPES", "float32,float64").split(",")]
for dtype, n_chain, cpw in ((d, n, c) for d in dtypes for n in counts for c in cpws):
    rep = (n_chain + pose.shape[0] - 1) // pose.shape[0]
    d_pose = torch.from_numpy(np.ascontiguousarray(pose)).to("cuda", dtype).repeat(rep, 1, 1, 1)[:n_chain].contiguous()
    d_rows = torch.from_numpy(np.tile(rows6, (n_unique * rep, 1))[:n_chain]).to("cuda", dtype).contiguous()
    for _ in range(2):
        out = engine.leg_solve_generic(d_pose, d_rows, chains_per_warp=cpw, schedule=sched)
    torch.cuda.synchronize()
    reps = 3
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e[0].record()
    for _ in range(reps):
        out = engine.leg_solve_generic(d_pose, d_rows, chains_per_warp=cpw, schedule=sched)
    e[1].record()
    torch.cuda.synchronize()
    ms = e[0].elapsed_time(e[1]) / reps
    ang, fk, status, nfev = out
    resid = (fk[:, :, 8] - d_pose[:, :, 4]).norm(dim=-1)
    lf = n_chain * n_frame
    print(json.dumps({"kernel": "leg_solve_generic", "dtype": str(dtype).split(".")[1], "chains": n_chain, "cpw": cpw, "schedule": sched, "frames": n_frame,
                      "ms": ms, "leg_frames_per_s": lf / ms * 1e3, "evals_per_leg_frame": float(nfev.double().sum()) / lf,
                      "us_per_eval_per_chain": ms * 1e3 / (float(nfev.double().sum()) / n_chain),
                      "claw_residual_max_mm": float(resid.max()), "claw_residual_mean_mm": float(resid.mean()),
                      "status_ok": bool((status == 1).all())}), flush=True)
