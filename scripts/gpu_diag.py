"""Diagnostic dump (GPU box): per-leg parity of the grooming trial for each kernel schedule. Dev tool."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
from seqikpy_b200 import data as D, engine, _native as N
from seqikpy_b200.kinematic_chain import KinematicChainSeq
from helpers import residual_of_angles, fk_residual
from oracle import seqik_oracle as O

g = dict(np.load(ROOT / "tests/golden/grooming_leg.npz"))
chain = KinematicChainSeq(D.BOUNDS, ["RF", "LF"])
params = torch.from_numpy(np.stack([chain.pack_chain_params(l, D.INITIAL_ANGLES[l]) for l in ("RF", "LF")]).astype(np.float32)).cuda()
pose = torch.from_numpy(g["pose"].astype(np.float32)).cuda()
np.set_printoptions(precision=2, linewidth=200)
for sched in (1, 2):
    for flags in (N.FLAG_DEFAULT,):
        ang, fk, st, nf = engine.leg_solve(pose, params, schedule=sched, flags=flags)
        torch.cuda.synchronize()
        a = ang.cpu().numpy().astype(np.float64); f = fk.cpu().numpy().astype(np.float64)
        print(f"== schedule {sched} flags {flags}: nfev/frame {nf.cpu().numpy() / 6000}")
        for li, leg in enumerate(("RF", "LF")):
            err = np.abs(a[li] - g["ref_angles"][li]); erro = np.abs(a[li] - g["oracle_angles"][li])
            bad = np.where(err.max(1) > 1e-3)[0]; bado = np.where(erro.max(1) > 1e-3)[0]
            seg = [chain.body_size[f"{leg}_{s}"] for s in O.SEGMENTS]
            r_o = fk_residual(f[li], g["pose"][li]); r_r = residual_of_angles(g["ref_angles"][li], seg, g["pose"][li])
            worse = np.where(((r_o - r_r) > 1e-4 + 2e-6).any(1))[0]
            print(leg, "bad vs ref", len(bad), bad[:40], "bad vs oracle", len(bado), bado[:40])
            print("   max err per dof", err.max(0), "median", np.median(err), "p99.9", np.quantile(err, 0.999))
            print("   fk worse frames", len(worse), worse[:40], "mean res ours/ref", r_o.mean(), r_r.mean())
            for t in bad[:6]:
                print("   frame", t, "ours", a[li][t], "ref", g["ref_angles"][li][t], "res ours", r_o[t], "res ref", r_r[t])
