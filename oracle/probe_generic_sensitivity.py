"""Evidence for DESIGN.md section 5.4: the reference's GENERIC 7-DOF leg IK (LegInvKinGeneric + KinematicChainGeneric,
seqikpy/leg_inverse_kinematics.py:406-613, seqikpy/kinematic_chain.py:424-532) has no reproducible answer to be a
drop-in for.  TEST INFRASTRUCTURE / analysis script; run in the build container:

    python oracle/probe_generic_sensitivity.py

One claw target (3 residuals) against 7 free joint angles is under-determined; scipy's TRF then moves along the
4-dimensional self-motion manifold driven by the rounding noise of its SVD null space.  Perturbing the aligned pose
by 1e-12 mm changes the returned joint angles by up to ~0.5 rad from frame ~40 on (110 of 150 frames differ by more
than the 1e-3 rad parity tolerance) while the claw residual stays the same.  The sequential solver (the hot path of
this repository) is not affected: each of its stages is a well-posed <= 2-variable problem."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import seqik_oracle as O            # noqa: E402
from seqikpy_b200 import data as D              # noqa: E402


def generic_chain(leg, size, bounds):
    return O.build_generic_chain(leg, size, bounds)


def run(links, pose, seed):
    x, out = np.array(seed, dtype=float), []
    for t in range(pose.shape[0]):
        x = O.inverse_kinematics(links, pose[t, 4] - pose[t, 0], x)
        out.append(x.copy())
    return np.array(out)


if __name__ == "__main__":
    g = dict(np.load(ROOT / "tests" / "golden" / "grooming_leg.npz"))
    size = O.calculate_body_size(D.NMF_TEMPLATE, ["RF", "LF"])
    links = generic_chain("RF", size, D.BOUNDS)
    pose = g["pose"][0][:150]
    t0 = time.time()
    a0 = run(links, pose, D.INITIAL_ANGLES["RF"]["stage_4"])
    a1 = run(links, pose + 1e-12 * np.random.default_rng(0).normal(size=pose.shape), D.INITIAL_ANGLES["RF"]["stage_4"])
    d = np.abs(a1 - a0)
    print(f"{time.time() - t0:.1f} s; max |angle difference| per chain slot under a 1e-12 mm input perturbation:")
    print(np.array2string(d.max(0), precision=3))
    print("frames (of 150) differing by more than 1e-3 rad:", int((d.max(1) > 1e-3).sum()))
