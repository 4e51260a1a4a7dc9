"""CPU oracle for the sequential leg-IK hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  The product
package (``sequential-inverse-kinematics_b200``) never does; it fails loudly
when the CUDA library is missing.

What is restated here
---------------------
The reference's leg path is ``LegInvKinSeq.run_ik_and_fk``
(/root/reference/seqikpy/leg_inverse_kinematics.py:324-403) ->
``calculate_ik_stage`` (:200-322) -> ``ikpy.chain.Chain.inverse_kinematics``
(:62-69) -> ``scipy.optimize.least_squares``.  The arithmetic lives in two
third-party packages:

* ``ikpy==3.3.4`` (pinned in /root/reference/setup.py:14) -- NOT installed in
  the build image and not installable (no network).  Its published algorithm
  is restated below: ``URDFLink.get_link_frame_matrix`` =
  Trans(origin_translation) . RPY(origin_orientation) . Rot(axis, theta) with
  RPY(r,p,y) = Rz(y) Ry(p) Rx(r) and Rot = Rodrigues with the axis used as
  given; ``Chain.forward_kinematics`` = cumulative product of the link
  frames; ``Chain.inverse_kinematics(target_position, initial_position)`` =
  ``scipy.optimize.least_squares(lambda x: fk(x)[:3, 3] - target, x0,
  bounds=(lb, ub))`` over ALL links (ikpy's default ``active_links_mask`` is
  all-True and its default optimizer is "least_squares").
* ``scipy.optimize.least_squares`` (un-pinned; Trust-Region-Reflective,
  2-point finite-difference Jacobian, ftol=xtol=gtol=1e-8) -- installed
  (1.18.1) and CALLED here, not restated.

The chain topology follows /root/reference/seqikpy/kinematic_chain.py:152-421
(stage builders) and the frame loop / warm start / column extraction follow
/root/reference/seqikpy/leg_inverse_kinematics.py:239-322.

The generic single-target path (``LegInvKinGeneric`` / ``KinematicChainGeneric``,
leg_inverse_kinematics.py:406-613, kinematic_chain.py:424-532) is restated by
``build_generic_chain`` / ``run_generic_leg`` / ``run_generic_ik_and_fk`` on the same ikpy
glue.  The reference ships no output for it; its fixture (tests/golden/generic_leg.npz) is
this oracle's own output together with a measurement of how reproducible that output is.

Parity pinning
--------------
The reference's own tests hold no numeric check for this path
(tests/test_kin_chain.py checks link names only).  The oracle is pinned
against the reference's shipped outputs
``data/anipose_220525_aJO_Fly001_001/pose-3d/{leg_joint_angles,forward_kinematics}.pkl``
(committed as fixtures in tests/golden/, see oracle/make_golden.py and
tests/test_oracle_golden.py).
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import least_squares

LEGS = ("RF", "LF", "RM", "LM", "RH", "LH")
DOF_ORDER = ("ThC_yaw", "ThC_pitch", "ThC_roll", "CTr_pitch", "CTr_roll", "FTi_pitch", "TiTa_pitch")
SEGMENTS = ("Coxa", "Femur", "Tibia", "Tarsus")

X_AXIS = (1.0, 0.0, 0.0)
Y_AXIS = (0.0, 1.0, 0.0)
Z_AXIS = (0.0, 0.0, 1.0)


# ----------------------------------------------------------------------------
# ikpy 3.3.4 geometry (restated from its published behaviour)
# ----------------------------------------------------------------------------
def axis_rotation(axis, theta):
    """ikpy.utils.geometry.axis_rotation_matrix: Rodrigues, axis NOT normalised."""
    x, y, z = axis
    c, s = np.cos(theta), np.sin(theta)
    return np.array([
        [x * x + (1 - x * x) * c, x * y * (1 - c) - z * s, x * z * (1 - c) + y * s],
        [x * y * (1 - c) + z * s, y * y + (1 - y * y) * c, y * z * (1 - c) - x * s],
        [x * z * (1 - c) - y * s, y * z * (1 - c) + x * s, z * z + (1 - z * z) * c],
    ])


def rpy_matrix(roll, pitch, yaw):
    """ikpy.utils.geometry.rpy_matrix = Rz(yaw) . Ry(pitch) . Rx(roll)."""
    return axis_rotation(Z_AXIS, yaw) @ axis_rotation(Y_AXIS, pitch) @ axis_rotation(X_AXIS, roll)


class Link:
    """One ikpy ``URDFLink`` (or ``OriginLink`` when ``origin`` is True)."""

    def __init__(self, name, translation=(0, 0, 0), orientation=(0, 0, 0), rotation=None,
                 bounds=(-np.inf, np.inf), origin=False):
        self.name = name
        self.bounds = (float(bounds[0]), float(bounds[1]))
        self.rotation = None if rotation is None else tuple(float(a) for a in rotation)
        self.is_origin = origin
        # constant part of the frame: Trans . RPY
        base = np.eye(4)
        base[:3, 3] = translation
        base[:3, :3] = rpy_matrix(*orientation)
        self._base = base

    def frame(self, theta):
        if self.is_origin or self.rotation is None:
            return self._base
        m = np.eye(4)
        m[:3, :3] = axis_rotation(self.rotation, theta)
        return self._base @ m


def forward_kinematics(links, q, full=False):
    """ikpy ``Chain.forward_kinematics``: cumulative product of link frames."""
    m = np.eye(4)
    out = []
    for link, theta in zip(links, q):
        m = m @ link.frame(theta)
        if full:
            out.append(m)
    return out if full else m


def inverse_kinematics(links, target, x0, return_result=False):
    """ikpy ``Chain.inverse_kinematics`` (position only, all links 'active')."""
    lb = np.array([l.bounds[0] for l in links])
    ub = np.array([l.bounds[1] for l in links])

    def residual(x):
        return forward_kinematics(links, x)[:3, 3] - target

    res = least_squares(residual, np.asarray(x0, dtype=float), bounds=(lb, ub))
    return res if return_result else res.x


# ----------------------------------------------------------------------------
# KinematicChainSeq stage builders  (kinematic_chain.py:152-421)
# ----------------------------------------------------------------------------
def build_chain(stage, leg, body_size, bounds, angles=None, t=0):
    """Links of ``create_leg_chain(leg, stage=stage, angles=angles, t=t)``.

    ``angles`` is the reference's ``joint_angles_dict`` ("Angle_{leg}_{dof}" ->
    (N,) array).  Earlier-stage DOFs become ``fixed`` links carrying their
    angle in the rpy slot of ``origin_orientation``.
    """
    if leg not in LEGS:
        raise ValueError(f"Unknown leg name ({leg}) is provided!")
    if not 1 <= stage <= 4:
        raise ValueError(f"Unknown stage number ({stage}) number is provided!")

    def b(dof):
        return bounds[f"{leg}_{dof}"]

    def a(dof):
        return float(angles[f"Angle_{leg}_{dof}"][t])

    def L(seg):
        return float(body_size[f"{leg}_{seg}"])

    def rev(dof, axis, trans=(0, 0, 0)):
        return Link(f"{leg}_{dof}", trans, (0, 0, 0), axis, b(dof))

    def fix(dof, slot, trans=(0, 0, 0)):
        orient = [0.0, 0.0, 0.0]
        orient[slot] = a(dof)
        return Link(f"{leg}_{dof}", trans, orient, None, b(dof))

    links = [Link("Base link", origin=True)]
    if stage == 1:
        links += [
            rev("ThC_yaw", X_AXIS), rev("ThC_pitch", Y_AXIS),
            rev("CTr_pitch", Y_AXIS, (0, 0, -L("Coxa"))),
        ]
    elif stage == 2:
        links += [
            fix("ThC_yaw", 0), fix("ThC_pitch", 1), rev("ThC_roll", Z_AXIS),
            rev("CTr_pitch", Y_AXIS, (0, 0, -L("Coxa"))),
            rev("FTi_pitch", Y_AXIS, (0, 0, -L("Femur"))),
        ]
    elif stage == 3:
        links += [
            fix("ThC_yaw", 0), fix("ThC_pitch", 1), fix("ThC_roll", 2),
            fix("CTr_pitch", 1, (0, 0, -L("Coxa"))), rev("CTr_roll", Z_AXIS),
            rev("FTi_pitch", Y_AXIS, (0, 0, -L("Femur"))),
            rev("TiTa_pitch", Y_AXIS, (0, 0, -L("Tibia"))),
        ]
    else:
        links += [
            fix("ThC_yaw", 0), fix("ThC_pitch", 1), fix("ThC_roll", 2),
            fix("CTr_pitch", 1, (0, 0, -L("Coxa"))), fix("CTr_roll", 2),
            fix("FTi_pitch", 1, (0, 0, -L("Femur"))),
            rev("TiTa_pitch", Y_AXIS, (0, 0, -L("Tibia"))),
            Link(f"{leg}_Claw", (0, 0, -L("Tarsus")), (0, 0, 0), (0.0, 0.0, 0.0), (-np.pi, np.pi)),
        ]
    return links


STAGE_DOFS = {1: ("ThC_yaw", "ThC_pitch"), 2: ("ThC_roll", "CTr_pitch"),
              3: ("CTr_roll", "FTi_pitch"), 4: ("TiTa_pitch",)}


def calculate_body_size(template, legs):
    """utils.calculate_body_size (utils.py:89-123)."""
    size = {}
    names = SEGMENTS + ("Claw",)
    for i, seg in enumerate(names):
        for leg in legs:
            if seg == "Claw":
                size[leg] = sum(size[f"{leg}_{s}"] for s in SEGMENTS)
            else:
                size[f"{leg}_{seg}"] = np.linalg.norm(
                    np.asarray(template[f"{leg}_{seg}"]) - np.asarray(template[f"{leg}_{names[i + 1]}"]))
    if "R_Antenna_base" in template:
        size["Antenna"] = np.linalg.norm(template["R_Antenna_base"] - template["R_Antenna_edge"])
        size["Antenna_mid_thorax"] = np.linalg.norm(template["R_Antenna_base"] - template["Thorax_mid"])
    return size


# ----------------------------------------------------------------------------
# LegInvKinSeq  (leg_inverse_kinematics.py:200-403)
# ----------------------------------------------------------------------------
def run_stage(leg, stage, end_effector, origin, init, body_size, bounds, angles, stats=None):
    """``calculate_ik_stage``: serial, warm-started frame loop of one stage."""
    n = end_effector.shape[0]
    target = end_effector - origin
    init = np.asarray(init, dtype=float)
    ja = np.empty((n, len(init)))
    fk = np.empty((n, len(init), 3))
    links = build_chain(1, leg, body_size, bounds) if stage == 1 else None
    for t in range(n):
        if stage != 1:
            links = build_chain(stage, leg, body_size, bounds, angles, t)
        x0 = init if t == 0 else ja[t - 1]
        res = inverse_kinematics(links, target[t], x0, return_result=True)
        ja[t] = res.x
        if stats is not None:
            stats.append((leg, stage, t, res.status, res.nfev, res.cost))
        if stage == 4:
            mats = forward_kinematics(links, ja[t], full=True)
            fk[t] = np.array([m[:3, 3] for m in mats]) + origin[t]
    names = [l.name for l in links]
    for dof in STAGE_DOFS[stage]:
        angles[f"Angle_{leg}_{dof}"] = ja[:, names.index(f"{leg}_{dof}")]
    return fk


def run_ik_and_fk(aligned_pos, body_size, bounds, initial_angles, stages=(1, 2, 3, 4), stats=None):
    """``LegInvKinSeq.run_ik_and_fk``: returns (joint_angles_dict, fk_dict)."""
    stages = list(stages)
    if max(stages) > 4 or not all(np.diff(stages) == 1):
        raise ValueError("Maximum stage number is 4 and the list should be strictly incremental.")
    angles, fk = {}, {}
    for name, arr in aligned_pos.items():
        if "leg" not in name.lower():
            continue
        leg = name.split("_")[0]
        if leg not in body_size:
            continue
        origin = arr[:, 0, :]
        for stage in stages:
            fk[name] = run_stage(leg, stage, arr[:, stage, :], origin,
                                 initial_angles[leg][f"stage_{stage}"], body_size, bounds, angles, stats)
    return angles, fk


# ----------------------------------------------------------------------------
# KinematicChainGeneric + LegInvKinGeneric  (kinematic_chain.py:424-532, leg_inverse_kinematics.py:406-613)
# ----------------------------------------------------------------------------
GENERIC_DOF_ORDER = ("ThC_roll", "ThC_yaw", "ThC_pitch", "CTr_pitch", "CTr_roll", "FTi_pitch", "TiTa_pitch")


def build_generic_chain(leg, body_size, bounds):
    """Links of ``KinematicChainGeneric.create_leg_chain(leg)`` (kinematic_chain.py:464-530): all joints revolute,
    ThC in roll-yaw-pitch order, the claw as a degenerate last link."""
    if leg not in LEGS:
        raise ValueError(f"Unknown leg name ({leg}) is provided!")

    def L(seg):
        return float(body_size[f"{leg}_{seg}"])

    def rev(dof, axis, trans=(0, 0, 0)):
        return Link(f"{leg}_{dof}", trans, (0, 0, 0), axis, bounds[f"{leg}_{dof}"])

    return [Link("Base link", origin=True), rev("ThC_roll", Z_AXIS), rev("ThC_yaw", X_AXIS), rev("ThC_pitch", Y_AXIS),
            rev("CTr_pitch", Y_AXIS, (0, 0, -L("Coxa"))), rev("CTr_roll", Z_AXIS),
            rev("FTi_pitch", Y_AXIS, (0, 0, -L("Femur"))), rev("TiTa_pitch", Y_AXIS, (0, 0, -L("Tibia"))),
            Link(f"{leg}_Claw", (0, 0, -L("Tarsus")), (0, 0, 0), (0.0, 0.0, 0.0), (-np.pi, np.pi))]


def run_generic_leg(leg, end_effector, origin, init, body_size, bounds, teacher=None, stats=None):
    """``LegInvKinGeneric.calculate_ik_stage`` (leg_inverse_kinematics.py:473-547): one 9-slot solve per frame against
    the claw, warm-started from the previous frame.  Returns (joint_angles (N, 9), forward_kinematics (N, 9, 3)).
    ``teacher`` (N, 9), if given, replaces the warm start of every frame (single solves from prescribed seeds)."""
    links = build_generic_chain(leg, body_size, bounds)
    n = end_effector.shape[0]
    target = end_effector - origin
    ja = np.empty((n, len(links)))
    fk = np.empty((n, len(links), 3))
    x0 = np.asarray(init, dtype=float)
    for t in range(n):
        if teacher is not None:
            x0 = teacher[t]
        res = inverse_kinematics(links, target[t], x0, return_result=True)
        ja[t] = res.x
        if stats is not None:
            stats.append((leg, t, res.status, res.nfev, res.cost))
        fk[t] = np.array([m[:3, 3] for m in forward_kinematics(links, ja[t], full=True)]) + origin[t]
        x0 = ja[t]
    return ja, fk


def run_generic_ik_and_fk(aligned_pos, body_size, bounds, initial_angles):
    """``LegInvKinGeneric.run_ik_and_fk`` (leg_inverse_kinematics.py:549-613): returns (joint_angles_dict, fk_dict)."""
    angles, fk = {}, {}
    for name, arr in aligned_pos.items():
        if "leg" not in name.lower():
            continue
        leg = name.split("_")[0]
        if leg not in body_size:
            continue
        ja, fk[name] = run_generic_leg(leg, arr[:, -1, :], arr[:, 0, :], initial_angles[leg]["stage_4"], body_size, bounds)
        for i, dof in enumerate(GENERIC_DOF_ORDER):
            angles[f"Angle_{leg}_{dof}"] = ja[:, 1 + i]
    return angles, fk


def fk_generic(angles7, seg_len, origin):
    """9-row FK of the generic chain for (N,7) angles in GENERIC_DOF_ORDER (rows: Base, 7 joint origins, claw)."""
    angles7 = np.asarray(angles7, dtype=float)
    n = angles7.shape[0]
    out = np.zeros((n, 9, 3))
    for t in range(n):
        r, y, p, cp, cr, fp, tp = angles7[t]
        R = axis_rotation(Z_AXIS, r) @ axis_rotation(X_AXIS, y) @ axis_rotation(Y_AXIS, p)
        pos = R @ np.array([0, 0, -seg_len[0]])
        out[t, 4] = out[t, 5] = pos
        R = R @ axis_rotation(Y_AXIS, cp) @ axis_rotation(Z_AXIS, cr)
        pos = pos + R @ np.array([0, 0, -seg_len[1]])
        out[t, 6] = pos
        R = R @ axis_rotation(Y_AXIS, fp)
        pos = pos + R @ np.array([0, 0, -seg_len[2]])
        out[t, 7] = pos
        R = R @ axis_rotation(Y_AXIS, tp)
        out[t, 8] = pos + R @ np.array([0, 0, -seg_len[3]])
    return out + np.asarray(origin).reshape(-1, 1, 3)


def fk_closed_form(angles7, seg_len, origin):
    """Closed-form 9-row FK of the stage-4 chain for (N,7) angles.

    Used by tests to evaluate FK residuals of any angle set (ours or the
    reference's) without the solver.
    """
    angles7 = np.asarray(angles7, dtype=float)
    n = angles7.shape[0]
    out = np.zeros((n, 9, 3))
    for t in range(n):
        y, p, r, cp, cr, fp, tp = angles7[t]
        R = axis_rotation(X_AXIS, y) @ axis_rotation(Y_AXIS, p) @ axis_rotation(Z_AXIS, r)
        pos = np.zeros(3)
        pos = pos + R @ np.array([0, 0, -seg_len[0]])
        out[t, 4] = out[t, 5] = pos
        R = R @ axis_rotation(Y_AXIS, cp) @ axis_rotation(Z_AXIS, cr)
        pos = pos + R @ np.array([0, 0, -seg_len[1]])
        out[t, 6] = pos
        R = R @ axis_rotation(Y_AXIS, fp)
        pos = pos + R @ np.array([0, 0, -seg_len[2]])
        out[t, 7] = pos
        R = R @ axis_rotation(Y_AXIS, tp)
        pos = pos + R @ np.array([0, 0, -seg_len[3]])
        out[t, 8] = pos
    return out + np.asarray(origin).reshape(-1, 1, 3)
