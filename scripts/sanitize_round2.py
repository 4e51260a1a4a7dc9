"""Small invocations of the kernels added in round 2's second half, for compute-sanitizer (memcheck / racecheck): the fused
alignment statistics, the warp and the long-series quantile kernels (fast path, few-valued shortcut, general fallback), the
float32 pchip resampler and the eight-lanes-per-chain generic solver.  Dev tool (GPU box):
    compute-sanitizer --tool racecheck python scripts/sanitize_round2.py"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
from seqikpy_b200 import engine, synthetic as S
from seqikpy_b200.kinematic_chain import KinematicChainGeneric

rng = np.random.default_rng(3)
t = torch
pose = t.from_numpy((rng.normal(size=(5, 1, 5, 3)) + 0.02 * rng.normal(size=(5, 300, 5, 3))).astype(np.float32)).cuda()
consts = t.from_numpy((1.0 + rng.random((5, 4))).astype(np.float32)).cuda()
a = engine.leg_affine(pose, consts)
b = engine.leg_affine_unfused(pose, consts)
assert t.equal(a, b)
long_pose = t.from_numpy((rng.normal(size=(2, 1, 5, 3)) + 0.02 * rng.normal(size=(2, 5000, 5, 3))).astype(np.float32)).cuda()
engine.leg_affine(long_pose, consts[:2])
x = rng.normal(size=(6, 9000)).astype(np.float32)
x[1] = 2.5                                   # constant: few-valued shortcut
x[2] = np.round(x[2], 1)                     # heavy ties: general kernel
x[3, 4000:] = np.inf                         # +inf padding with counts
m = t.tensor([9000, 9000, 9000, 4000, 9000, 1], dtype=t.int32).cuda()
got = engine.mid_quantile(t.from_numpy(x).cuda(), m).cpu().numpy()
ref = [0.5 * (np.quantile(np.sort(x[r].astype(np.float64))[:int(m[r])], 0.45) + np.quantile(np.sort(x[r].astype(np.float64))[:int(m[r])], 0.55)) for r in range(6)]
assert np.allclose(got, ref, rtol=1e-6, atol=1e-6), (got, ref)
sig = t.from_numpy(rng.normal(size=(3, 200, 7)).astype(np.float32)).cuda()
engine.pchip_resample(sig, 0.01, 0.001)
size, bounds, init = S.chain_constants()
chain = KinematicChainGeneric(bounds, list(S.LEGS), size)
perm = [0, 3, 1, 2, 4, 5, 6, 7, 8]
rows = np.stack([chain.pack_chain_params(leg, np.asarray(init[leg]["stage_4"], dtype=float)[perm]) for leg in S.LEGS])
gp = S.make_trial(0, 1000)[:12].transpose(1, 0, 2, 3)          # (6 legs, 12 frames, 5, 3)
for dt in (t.float32, t.float64):
    out = engine.leg_solve_generic(t.from_numpy(np.ascontiguousarray(gp)).to("cuda", dt), t.from_numpy(rows).to("cuda", dt), schedule=2)
    assert bool((out[2] == 1).all())
t.cuda.synchronize()
print("sanitize_round2: ok")
