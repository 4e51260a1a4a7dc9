"""Shared helpers of the parity tests."""
import numpy as np

from oracle import seqik_oracle as O

ANGLE_TOL = 1e-3      # rad, BASELINE.json north_star: joint angles within 1e-3 rad of the reference
FK_TOL = 1e-4         # mm,  north_star: FK residual never worse than the reference's by more than 1e-4 mm per joint
F32_FK_NOISE = 2e-6   # mm,  float32 rounding of ~2.5 mm coordinates (the device computes in FP32)


def angles_dict_to_array(angles, leg):
    return np.stack([np.asarray(angles[f"Angle_{leg}_{d}"]) for d in O.DOF_ORDER], 1)


def fk_residual(fk9, pose5):
    """(N, 4) distance between FK rows 5..8 and the target key points 1..4."""
    return np.linalg.norm(np.asarray(fk9)[:, [5, 6, 7, 8]] - np.asarray(pose5)[:, 1:5], axis=2)


def residual_of_angles(ang7, seg, pose5):
    """FK residual of any angle set, evaluated with the oracle's float64 closed-form FK."""
    fk = O.fk_closed_form(ang7, seg, np.asarray(pose5)[:, 0])
    return fk_residual(fk, pose5)


def bad_frames(a, b, tol=ANGLE_TOL):
    return np.where(np.abs(np.asarray(a) - np.asarray(b)).max(axis=1) > tol)[0]


def singular_windows(ref_angles7, margin=20, tol=1e-6):
    """Frames within `margin` of a frame where the REFERENCE's own solution sits on the CTr_pitch = 0 kinematic
    singularity (there d(end point)/d(ThC_roll) vanishes; the reference leaves such a corner only through the
    rounding noise of its finite-difference Jacobian, bound-to-bound, so its path is not reproducible -- SURVEY.md
    finding 4).  On the bundled grooming trial: none for RF, two episodes for LF (frames 280-287 and 3347-3350)."""
    sing = np.where(np.abs(np.asarray(ref_angles7)[:, 3]) < tol)[0]
    frames = set()
    for t in sing:
        frames.update(range(max(0, t - margin), t + margin + 1))
    return frames
