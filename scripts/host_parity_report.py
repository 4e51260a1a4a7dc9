"""Parity report of the pipeline kernel's per-lane arithmetic WITHOUT a GPU: the carried solves (tests/hostsim:
hostsim_carried_f32 = StageSolve::restart / trip exactly as leg_solve_pipe_kernel runs them, serially) over the bundled
grooming trial (6000 frames x 2 legs, against the reference's shipped angles) and synthetic trials 2-4 (against the
oracle fixture), for each solver flag set given in hex.  Prints bad frames in / outside the reference's singular
windows, max / median angle differences, FK-residual deltas, FK-vs-angles consistency and evaluations per solve.

    python scripts/host_parity_report.py 3f 7f ff

Development aid (CPU only, uses tests/ and oracle/ like the tests do; never part of the product path)."""
import sys

import numpy as np
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import hostsim_build as H, model_trf2 as M
from helpers import *
from oracle import seqik_oracle as O
from seqikpy_b200 import data as D, synthetic as S
from seqikpy_b200.kinematic_chain import KinematicChainSeq
from test_host_core import leg_consts
g=np.load(str(ROOT) + '/tests/golden/grooming_leg.npz', allow_pickle=True)
size = O.calculate_body_size(D.NMF_TEMPLATE, ["RF", "LF"])
chain_g = KinematicChainSeq(D.BOUNDS, ["RF","LF"], size)
sz, bounds, init = S.chain_constants()
chain_s = KinematicChainSeq(bounds, list(S.LEGS), sz)
w=np.load(str(ROOT) + '/tests/golden/synthetic_wide.npz', allow_pickle=True)
for flags in [int(a,16) for a in sys.argv[1:]]:
    print("==== flags", hex(flags))
    for li,leg in enumerate(["RF","LF"]):
        seg, lb, ub, nsq, seed = leg_consts(size, D.BOUNDS, D.INITIAL_ANGLES, leg)
        prm = chain_g.pack_chain_params(leg, D.INITIAL_ANGLES[leg])
        ang, fk, nfev = H.run_carried_f32(g["pose"][li], prm, flags)
        bad = bad_frames(ang, g["ref_angles"][li]); allowed = singular_windows(g["ref_angles"][li])
        d = np.abs(ang-g["ref_angles"][li]); ok=np.array([t not in allowed for t in range(len(ang))])
        r_ours = fk_residual(fk, g["pose"][li]); r_ref = residual_of_angles(g["ref_angles"][li], seg, g["pose"][li])
        r_chk = residual_of_angles(ang.astype(np.float64), seg, g["pose"][li])
        worse = np.where(((r_ours - r_ref) > FK_TOL + 2e-6).any(axis=1))[0]
        worse2 = np.where(((r_chk - r_ref) > FK_TOL + 2e-6).any(axis=1))[0]
        print(leg, "bad", len(bad), "outside windows", sorted(set(bad)-allowed)[:10], "max|d| outside", d[ok].max(), "median", np.median(d[ok]),
              "worse-FK frames", len(worse), "outside", sorted(set(worse)-allowed)[:10], "by angles:", len(worse2), sorted(set(worse2)-allowed)[:10], "mean res ours/ref", r_ours.mean(), r_ref.mean(), "fk-vs-angles consistency", np.abs(r_ours-r_chk).max(), "nfev", nfev.mean(0), "max nfev", nfev.max(0))
    mx=0; worst_fk=-1; nf=[]; cons=0
    for ti,tr in enumerate(w["trials"][:3]):
        pose=S.make_trial(int(tr), int(w["n_frame"]))
        for li,leg in enumerate(S.LEGS):
            seg, lb, ub, nsq, seed = leg_consts(sz, bounds, init, leg)
            ang, fk, nfev = H.run_carried_f32(pose[:, li], chain_s.pack_chain_params(leg, init[leg]), flags)
            mx=max(mx, np.abs(ang - w["oracle_angles"][ti, li]).max())
            worst_fk=max(worst_fk,(fk_residual(fk, pose[:, li]) - w["oracle_fk_residual"][ti, li]).max())
            cons=max(cons, np.abs(fk_residual(fk, pose[:, li]) - residual_of_angles(ang.astype(np.float64), seg, pose[:,li])).max())
            nf.append(nfev)
    nf=np.stack(nf)
    print("synthetic wide: max|d|", mx, "worst fk delta", worst_fk, "consistency", cons, "nfev mean", nf.mean((0,1)), "trips hist s1", np.round(np.bincount(nf[:,:,0].ravel()-1)/nf[:,:,0].size,3))
