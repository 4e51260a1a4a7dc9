"""Helpers on the leg-IK path: template sizes and pickle I/O.

Mirrors the three functions of the reference's ``seqikpy/utils.py`` that the hot path uses:
``calculate_body_size`` (:89-123), ``save_file`` (:235-238), ``load_file`` (:241-245).  The
format converters and video helpers of that module are outside the path.
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


def save_file(out_fname, data):
    """Pickle ``data`` to ``out_fname`` (same on-disk layout as the reference)."""
    with open(out_fname, "wb") as f:
        pickle.dump(data, f)


def load_file(output_fname):
    """Load a pickle written by :func:`save_file` (or by the reference)."""
    with open(output_fname, "rb") as f:
        return pickle.load(f)
