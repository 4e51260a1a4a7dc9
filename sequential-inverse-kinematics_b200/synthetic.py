"""Synthetic pose data of the benchmark configurations (BASELINE.json configs 3-5).

Geometry, joint limits and seeds are the six-leg locomotion set of the reference
(examples/example_leg_inv_kinematics_parallel.py:21-140).  Every trial is generated from its own
``numpy.random.default_rng(20240611 + trial)`` so that any subset of trials (a GPU shard, the few
trials the CPU oracle checks) is reproducible independently of the rest.

Per trial and leg (order RF, RM, RH, LF, LM, LH): frequencies f ~ U(1, 4) and phases phi ~ U(0, 2 pi)
for the 7 DOFs, ground-truth angles theta_d(t) = clip(theta0_d + 0.3 sin(2 pi f_d t / 1000 + phi_d),
lb_d + 0.05, ub_d - 0.05), key points = closed-form forward kinematics + template coxa position,
plus N(0, 0.02 mm) noise on key points 1-4.
"""
import numpy as np

from .data import BOUNDS_LOCOMOTION, INITIAL_ANGLES_LOCOMOTION, TEMPLATE_NMF_LOCOMOTION
from .kinematic_chain import DOF_ORDER, SEGMENTS
from .utils import calculate_body_size

LEGS = ("RF", "RM", "RH", "LF", "LM", "LH")
SEED_BASE = 20240611
NOISE_MM = 0.02


def _rot_apply(axis, ang, v):
    """Rotate vectors v (..., 3) about a coordinate axis by ang (...)."""
    c, s = np.cos(ang), np.sin(ang)
    x, y, z = v[..., 0], v[..., 1], v[..., 2]
    if axis == 0:
        return np.stack([x, c * y - s * z, s * y + c * z], -1)
    if axis == 1:
        return np.stack([c * x + s * z, y, -s * x + c * z], -1)
    return np.stack([c * x - s * y, s * x + c * y, z], -1)


def leg_key_points(theta, seg):
    """Closed-form FK of the leg: theta (F, 7), seg (4,) -> (F, 4, 3) positions of the
    Coxa-Femur, Femur-Tibia, Tibia-Tarsus joints and the claw relative to the Thorax-Coxa joint."""
    yaw, pitch, roll, cp, cr, fp, tp = (theta[:, i] for i in range(7))
    f = theta.shape[0]

    def down(length):
        v = np.zeros((f, 3))
        v[:, 2] = -length
        return v
    # innermost rotation first: p = Rx(yaw) Ry(pitch) Rz(roll) [ (0,0,-Cx) + Ry(cp) [ (0,0,-Fe) + Rz(cr) Ry(fp) [ ... ] ] ]
    p_claw = _rot_apply(1, tp, down(seg[3]))
    p_tars = down(seg[2])
    l3 = _rot_apply(2, cr, _rot_apply(1, fp, np.stack([p_tars, p_tars + p_claw], 1).reshape(-1, 3).reshape(f, 2, 3).transpose(1, 0, 2)))
    # l3: (2, F, 3) = tarsus joint and claw in the femur frame (relative to the femur-tibia joint)
    p_tib = down(seg[1])
    pts = np.stack([p_tib, p_tib + l3[0], p_tib + l3[1]], 0)            # relative to coxa-femur joint, femur frame
    pts = _rot_apply(1, cp, pts)
    p_fem = down(seg[0])
    pts = np.concatenate([p_fem[None], p_fem[None] + pts], 0)           # (4, F, 3) in the coxa frame
    pts = _rot_apply(0, yaw, _rot_apply(1, pitch, _rot_apply(2, roll, pts)))
    return pts.transpose(1, 0, 2)


def chain_constants(legs=LEGS):
    """(body_size, bounds, initial_angles) of the synthetic workload."""
    return calculate_body_size(TEMPLATE_NMF_LOCOMOTION, list(legs)), BOUNDS_LOCOMOTION, INITIAL_ANGLES_LOCOMOTION


def make_trial(trial: int, n_frame: int = 1000, legs=LEGS, dtype=np.float64, return_truth: bool = False):
    """Pose of one trial: (n_frame, n_leg, 5, 3)."""
    rng = np.random.default_rng(SEED_BASE + int(trial))
    size, bounds, init = chain_constants(legs)
    t = np.arange(n_frame, dtype=float)
    pose = np.empty((n_frame, len(legs), 5, 3))
    truth = np.empty((n_frame, len(legs), 7))
    for li, leg in enumerate(legs):
        freq = rng.uniform(1.0, 4.0, 7)
        phase = rng.uniform(0.0, 2 * np.pi, 7)
        theta0 = np.array(init[leg]["stage_4"][1:8], dtype=float)
        theta0[6] = -0.6
        lb = np.array([bounds[f"{leg}_{d}"][0] for d in DOF_ORDER]) + 0.05
        ub = np.array([bounds[f"{leg}_{d}"][1] for d in DOF_ORDER]) - 0.05
        theta = np.clip(theta0 + 0.3 * np.sin(2 * np.pi * freq * t[:, None] / 1000.0 + phase), lb, ub)
        seg = np.array([size[f"{leg}_{s}"] for s in SEGMENTS])
        pts = leg_key_points(theta, seg) + rng.normal(0.0, NOISE_MM, (n_frame, 4, 3))
        coxa = np.asarray(TEMPLATE_NMF_LOCOMOTION[f"{leg}_Coxa"], dtype=float)
        pose[:, li, 0] = coxa
        pose[:, li, 1:] = pts + coxa
        truth[:, li] = theta
    pose = pose.astype(dtype, copy=False)
    return (pose, truth) if return_truth else pose


def make_trials(trials, n_frame: int = 1000, legs=LEGS, dtype=np.float32, out=None):
    """Stack of trials -> (n_trial, n_frame, n_leg, 5, 3).  ``trials`` is an iterable of trial indices."""
    trials = list(trials)
    if out is None:
        out = np.empty((len(trials), n_frame, len(legs), 5, 3), dtype=dtype)
    for i, tr in enumerate(trials):
        out[i] = make_trial(tr, n_frame, legs)
    return out


def to_chains(pose):
    """(n_trial, n_frame, n_leg, 5, 3) -> chain-major (n_trial * n_leg, n_frame, 5, 3), the solver's layout."""
    n_trial, n_frame, n_leg = pose.shape[:3]
    if hasattr(pose, "permute"):
        return pose.permute(0, 2, 1, 3, 4).reshape(n_trial * n_leg, n_frame, 5, 3).contiguous()
    return np.ascontiguousarray(pose.transpose(0, 2, 1, 3, 4).reshape(n_trial * n_leg, n_frame, 5, 3))


# ---------------------------------------------------------------------------------------------
# Head / antenna key points and raw (un-aligned) poses for the fused-pipeline configuration (BASELINE config 5)
# ---------------------------------------------------------------------------------------------
HEAD_SEED_BASE = 30240611
HEAD_NOISE_MM = 0.005
RAW_SCALE = 1.0 / 1.05
RAW_OFFSET = np.array([0.1, -0.2, 0.3])


def _rot_xyz(roll, pitch, yaw, v):
    """Rz(yaw) Ry(pitch) Rx(roll) v for per-frame angles (F,) and vectors (F, 3)."""
    return _rot_apply(2, yaw, _rot_apply(1, pitch, _rot_apply(0, roll, v)))


def make_head_trial(trial: int, n_frame: int = 1000, template=None, dtype=np.float64):
    """Antenna bases/tips and thorax key points of one trial, already in template scale:
    r_head, l_head (n_frame, 2, 3), thorax (n_frame, 3, 3), neck (3,).  The template antenna points are rotated about
    the neck by head roll/pitch/yaw sinusoids (0.2 rad), the tips additionally pitched about their base (0.3 rad)."""
    from .data import NMF_TEMPLATE
    tmpl = NMF_TEMPLATE if template is None else template
    rng = np.random.default_rng(HEAD_SEED_BASE + int(trial))
    t = np.arange(n_frame, dtype=float)
    freq = rng.uniform(0.5, 3.0, 5)
    phase = rng.uniform(0.0, 2 * np.pi, 5)
    ang = [amp * np.sin(2 * np.pi * f * t / 1000.0 + p) for amp, f, p in zip((0.2, 0.2, 0.2, 0.3, 0.3), freq, phase)]
    neck = np.asarray(tmpl["Neck"], dtype=float)
    out = {}
    for side, tip_pitch in (("R", ang[3]), ("L", ang[4])):
        base0 = np.asarray(tmpl[f"{side}_Antenna_base"], dtype=float) - neck
        stalk0 = np.asarray(tmpl[f"{side}_Antenna_edge"], dtype=float) - np.asarray(tmpl[f"{side}_Antenna_base"], dtype=float)
        base = _rot_xyz(ang[0], ang[1], ang[2], np.tile(base0, (n_frame, 1)))
        stalk = _rot_xyz(ang[0], ang[1], ang[2], _rot_apply(1, tip_pitch, np.tile(stalk0, (n_frame, 1))))
        pts = np.stack([base + neck, base + stalk + neck], 1)
        out[side] = pts + rng.normal(0.0, HEAD_NOISE_MM, pts.shape)
    thorax0 = np.stack([np.asarray(tmpl[k], dtype=float) for k in ("R_wing", "Thorax_mid", "L_wing")])
    thorax = np.tile(thorax0, (n_frame, 1, 1)) + rng.normal(0.0, HEAD_NOISE_MM, (n_frame, 3, 3))
    return out["R"].astype(dtype), out["L"].astype(dtype), thorax.astype(dtype), neck.astype(dtype)


def to_raw(points):
    """Template-scale key points -> "raw" recording coordinates (what the alignment has to undo): p / 1.05 + offset."""
    return points * RAW_SCALE + RAW_OFFSET.astype(points.dtype)
