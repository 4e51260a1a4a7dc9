"""Helpers on the leg-IK path: template sizes and pickle I/O.

Mirrors the functions of the reference's ``seqikpy/utils.py`` that the hot path and its input converters use:
``calculate_body_size`` (:89-123), ``save_file`` (:235-238), ``load_file`` (:241-245), ``dict_to_nparray_pose`` /
``dict_to_nparray_angle`` (:293-329), ``interpolate_signal`` / ``interpolate_joint_angles`` (:332-359, on the device).
The video / stimulus helpers of that module are outside the path.
"""
import pickle
from typing import Dict, List

import numpy as np

_LEG_NAMES = ("RF", "LF", "RM", "LM", "RH", "LH")
_JOINTS = ("Coxa", "Femur", "Tibia", "Tarsus", "Claw")


def calculate_body_size(
    body_template: Dict[str, np.ndarray],
    legs_list: List[str] = ["RF", "LF", "RM", "LM", "RH", "LH"],
) -> Dict[str, np.ndarray]:
    """Segment lengths (distance between consecutive template joints), the whole-leg length
    ``body_size[leg]`` and, when the template has antennae, ``Antenna`` and
    ``Antenna_mid_thorax``.  Key order follows the reference (segment-major)."""
    unknown = set(legs_list) - set(_LEG_NAMES)
    if unknown:
        raise NameError(
            f"legs_list could only contain {list(_LEG_NAMES)}, currently, it contains {legs_list}")
    size = {}
    for proximal, distal in zip(_JOINTS[:-1], _JOINTS[1:]):
        for leg in legs_list:
            size[f"{leg}_{proximal}"] = np.linalg.norm(
                body_template[f"{leg}_{proximal}"] - body_template[f"{leg}_{distal}"])
    for leg in legs_list:
        size[leg] = (size[f"{leg}_Coxa"] + size[f"{leg}_Femur"] + size[f"{leg}_Tibia"] + size[f"{leg}_Tarsus"])
    if "R_Antenna_base" in body_template:
        size["Antenna"] = np.linalg.norm(body_template["R_Antenna_base"] - body_template["R_Antenna_edge"])
        size["Antenna_mid_thorax"] = np.linalg.norm(body_template["R_Antenna_base"] - body_template["Thorax_mid"])
    return size


def dict_to_nparray_pose(pose_dict, claw_is_end_effector):
    """df3dPP leg dictionary ({"Coxa": {"raw_pos_aligned": (N, 3)}, ...}) -> (N, 5 or 4, 3) array
    (reference utils.py:293-310)."""
    key_points = _JOINTS if claw_is_end_effector else _JOINTS[:-1]
    return np.stack([np.asarray(pose_dict[kp]["raw_pos_aligned"], dtype=float) for kp in key_points], axis=1)


def dict_to_nparray_angle(angle_dict, leg, claw_is_end_effector):
    """df3dPP angle dictionary -> (N, 7 or 6) array in the order roll, yaw, pitch, ... (reference utils.py:313-329)."""
    dofs = ["ThC_roll", "ThC_yaw", "ThC_pitch", "CTr_pitch", "CTr_roll", "FTi_pitch", "TiTa_pitch"]
    if not claw_is_end_effector:
        dofs = dofs[:-1]
    return np.stack([np.asarray(angle_dict[f"{leg}_leg"][d], dtype=float) for d in dofs], axis=1)


def save_file(out_fname, data):
    """Pickle ``data`` to ``out_fname`` (same on-disk layout as the reference)."""
    with open(out_fname, "wb") as f:
        pickle.dump(data, f)


def load_file(output_fname):
    """Load a pickle written by :func:`save_file` (or by the reference)."""
    with open(output_fname, "rb") as f:
        return pickle.load(f)


def interpolate_signal(signal, original_ts, new_ts):
    """Resamples a signal sampled every ``original_ts`` onto a ``new_ts`` grid with a shape-preserving cubic (pchip)
    -- the hand-off of joint angles to a simulation time step (reference utils.py:332-349, which calls
    ``scipy.interpolate.pchip_interpolate``).  ``signal``: (n,) or (n, k) along axis 0.  Runs the FP64 device kernel
    (``seqik_pchip_resample_f64``, equal to scipy's result to ~1e-12 relative); like every compute path of this package
    it needs the CUDA library and raises without it.  +-inf samples are zeroed together with the last sample, as the
    reference's retry does; NaN raises ``ValueError`` like scipy."""
    from . import _native as N
    from . import engine
    torch = N.require_cuda()
    arr = np.array(signal, dtype=np.float64)
    if arr.ndim not in (1, 2):
        raise ValueError("signal must be (n,) or (n, k)")
    x = torch.from_numpy(np.ascontiguousarray(arr.reshape(1, arr.shape[0], -1))).cuda()
    out = engine.pchip_resample(x, original_ts, new_ts).cpu().numpy()[0]
    return out[:, 0] if arr.ndim == 1 else out


def interpolate_joint_angles(joint_angles_dict, **kwargs):
    """``interpolate_signal`` over every entry of a joint-angle dictionary (reference utils.py:352-359); entries of equal
    length are resampled by one kernel launch."""
    from . import _native as N
    from . import engine
    torch = N.require_cuda()
    original_ts, new_ts = kwargs["original_ts"], kwargs["new_ts"]
    arrays = {dof: np.array(values, dtype=np.float64) for dof, values in joint_angles_dict.items()}
    out = {}
    by_len = {}
    for dof, arr in arrays.items():
        if arr.ndim != 1:
            out[dof] = interpolate_signal(arr, original_ts, new_ts)
        else:
            by_len.setdefault(arr.shape[0], []).append(dof)
    for n, dofs in by_len.items():
        x = torch.from_numpy(np.stack([arrays[d] for d in dofs])).cuda()
        res = engine.pchip_resample(x, original_ts, new_ts).cpu().numpy()
        for i, d in enumerate(dofs):
            out[d] = res[i]
    return {dof: out[dof] for dof in joint_angles_dict}
