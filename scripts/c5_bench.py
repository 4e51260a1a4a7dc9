"""BASELINE config 5: long-sequence stress, 100 trials x 100k frames x 6 legs, alignment + leg IK/FK + head IK fused
on the device (FusedPipeline).  Kernel-side timing with CUDA events; prints one JSON line.  (GPU box)"""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from seqikpy_b200 import synthetic as S, data as D
from seqikpy_b200.batch import FusedPipeline
from seqikpy_b200.kinematic_chain import KinematicChainSeq
from seqikpy_b200.utils import calculate_body_size

n_trial = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n_frame = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
n_unique = min(n_trial, 4)
size, bounds, init = S.chain_constants()
tmpl = dict(D.TEMPLATE_NMF_LOCOMOTION)
for k in ("R_Antenna_base", "L_Antenna_base", "R_Antenna_edge", "L_Antenna_edge", "Neck", "Thorax_mid", "R_wing", "L_wing"):
    tmpl[k] = D.NMF_TEMPLATE[k]
size = calculate_body_size(tmpl, list(S.LEGS))
chain = KinematicChainSeq(bounds, list(S.LEGS), size)
t0 = time.time()
legs = np.stack([S.to_raw(S.make_trial(tr, n_frame).astype(np.float32)).transpose(1, 0, 2, 3) for tr in range(n_unique)])
heads = [S.make_head_trial(tr, n_frame, dtype=np.float32) for tr in range(n_unique)]
rep = (n_trial + n_unique - 1) // n_unique
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda().repeat(rep, *([1] * (a.ndim - 1)))[:n_trial].contiguous()
d_legs = dev(legs)
d_r, d_l, d_th = (dev(S.to_raw(np.stack([h[i] for h in heads]))) for i in range(3))
gen_s = time.time() - t0
pipe = FusedPipeline(chain, init, S.LEGS, tmpl, size, n_trial, n_frame)
for _ in range(2):
    out = pipe.run(d_legs, d_r, d_l, d_th)
torch.cuda.synchronize()
reps = 3
e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
e[0].record()
for _ in range(reps):
    out = pipe.run(d_legs, d_r, d_l, d_th)
e[1].record(); torch.cuda.synchronize()
ms = e[0].elapsed_time(e[1]) / reps
# split: solver alone
sess = pipe.session
e[0].record()
for _ in range(reps):
    sess.solve_device(d_legs, affine=out["leg_affine"].view(-1, 8), want_stats=False)
e[1].record(); torch.cuda.synchronize()
ms_solver = e[0].elapsed_time(e[1]) / reps
lf = n_trial * 6 * n_frame
fk_err = sess.mean_fk_error(torch.empty(0)) if False else None
print(json.dumps({"workload": f"{n_trial} trials x {n_frame} frames x 6 legs, alignment + leg IK/FK + head IK fused", "ms": ms, "ms_solver_only": ms_solver,
                  "leg_frames_per_s": lf / ms * 1e3, "chains": n_trial * 6, "us_per_frame_per_chain": ms * 1e3 / n_frame,
                  "head_frames_per_s": n_trial * n_frame / ms * 1e3, "nfev_per_leg_frame": (sess.nfev.double().sum(0) / lf).tolist() if sess.nfev is not None else None,
                  "data_gen_s": gen_s, "scale_mean": float(out["leg_affine"][..., 3].mean()), "head_roll_std": float(out["head_angles"][:, 0].std())}))
