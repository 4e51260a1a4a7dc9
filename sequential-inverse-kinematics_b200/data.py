"""Constants of the leg-IK path: seeds, joint limits, NeuroMechFly template and sizes.

Same names and values as the reference's ``seqikpy/data.py`` (:4-167) -- they are part of
the public API (``from seqikpy.data import BOUNDS, INITIAL_ANGLES, NMF_TEMPLATE``) -- plus
the six-leg locomotion set that the reference only ships inside
``examples/example_leg_inv_kinematics_parallel.py`` (:21-140).  Tables that are mirror
images left/right are generated from the right side.
"""
import numpy as np

_LEGS = ("RF", "LF", "RM", "LM", "RH", "LH")


def _stage_seeds(yaw, pitch, roll, ctr_pitch=-2.14, ctr_roll=-1.25, fti_pitch=1.48, tita_pitch=0.0, s2_last=1.4):
    """The four stage seed vectors in chain order (Base link first, see kinematic_chain.py)."""
    return {
        "stage_1": np.array([0.0, yaw, pitch, ctr_pitch]),
        "stage_2": np.array([0.0, yaw, pitch, roll, ctr_pitch, s2_last]),
        "stage_3": np.array([0.0, yaw, pitch, roll, ctr_pitch, ctr_roll, fti_pitch, tita_pitch]),
        "stage_4": np.array([0.0, yaw, pitch, roll, ctr_pitch, ctr_roll, fti_pitch, tita_pitch, 0.0]),
    }


# Seeds of the optimisation (reference data.py:4-22).
INITIAL_ANGLES = {
    "RF": _stage_seeds(0.45, -0.07, -0.32),
    "LF": _stage_seeds(-0.45, -0.07, 0.32, ctr_roll=1.25),
}

_deg = np.deg2rad
# Joint limits (reference data.py:26-41); left side mirrors roll-type DOFs.
BOUNDS = {
    "RF_ThC_roll": (_deg(-130), _deg(50)),
    "RF_ThC_yaw": (_deg(-50), _deg(50)),
    "RF_ThC_pitch": (_deg(-40), _deg(60)),
    "RF_CTr_pitch": (_deg(-180), _deg(0)),
    "RF_CTr_roll": (_deg(-150), _deg(0)),
    "RF_FTi_pitch": (_deg(0), _deg(170)),
    "RF_TiTa_pitch": (_deg(-150), _deg(0)),
    "LF_ThC_roll": (_deg(-50), _deg(130)),
    "LF_ThC_yaw": (_deg(-50), _deg(50)),
    "LF_ThC_pitch": (_deg(-40), _deg(60)),
    "LF_CTr_pitch": (_deg(-180), _deg(0)),
    "LF_CTr_roll": (_deg(0), _deg(150)),
    "LF_FTi_pitch": (_deg(0), _deg(170)),
    "LF_TiTa_pitch": (_deg(-150), _deg(0)),
}

# Segment sizes of the template (reference data.py:44-77).
_SEG_SIZE = {
    "F": {"Coxa": 0.40, "Femur": 0.69, "Tibia": 0.54, "Tarsus": 0.63, "leg": 2.26},
    "M": {"Coxa": 0.182, "Femur": 0.7829999999999999, "Tibia": 0.668, "Tarsus": 0.6949999999999998, "leg": 2.328},
    "H": {"Coxa": 0.199, "Femur": 0.8360000000000001, "Tibia": 0.6849999999999998, "Tarsus": 0.7950000000000002, "leg": 2.515},
}
NMF_SIZE = {}
for _seg in ("Coxa", "Femur", "Tibia", "Tarsus"):
    for _pos in ("F", "M", "H"):
        NMF_SIZE[f"R{_pos}_{_seg}"] = _SEG_SIZE[_pos][_seg]
    for _pos in ("F", "M", "H"):
        NMF_SIZE[f"L{_pos}_{_seg}"] = _SEG_SIZE[_pos][_seg]
for _side in ("R", "L"):
    for _pos in ("F", "M", "H"):
        NMF_SIZE[f"{_side}{_pos}"] = _SEG_SIZE[_pos]["leg"]
NMF_SIZE["Antenna"] = 0.2745906043549196
NMF_SIZE["Antenna_mid_thorax"] = 0.9355746896961248

# Key points to align (reference data.py:80-98).
PTS2ALIGN = {
    "R_head": ["base_anten_R", "tip_anten_R"],
    "RF_leg": ["thorax_coxa_R", "coxa_femur_R", "femur_tibia_R", "tibia_tarsus_R", "claw_R"],
    "Thorax": ["thorax_wing_R", "thorax_midpoint_tether", "thorax_wing_L"],
    "L_head": ["base_anten_L", "tip_anten_L"],
    "LF_leg": ["thorax_coxa_L", "coxa_femur_L", "femur_tibia_L", "tibia_tarsus_L", "claw_L"],
}


def get_pts2align(path: str):
    """PTS2ALIGN without the legs named in the path (reference data.py:101-112)."""
    pts = PTS2ALIGN.copy()
    if "_RF" in path:
        del pts["RF_leg"]
    elif "_LF" in path:
        del pts["LF_leg"]
    elif "_RLF" in path or "_LRF" in path:
        del pts["LF_leg"]
        del pts["RF_leg"]
    return pts


# Key points used in alignment (reference data.py:116-134).
SKELETON = (
    PTS2ALIGN["R_head"] + PTS2ALIGN["RF_leg"] + PTS2ALIGN["Thorax"] + PTS2ALIGN["L_head"] + PTS2ALIGN["LF_leg"]
)


def _mirror(d):
    """Adds the left-side twin (y -> -y) of every R* key."""
    out = {}
    for k, v in d.items():
        out[k] = np.array(v, dtype=float)
        out["L" + k[1:]] = np.array([v[0], -v[1], v[2]], dtype=float)
    return out


# NeuroMechFly v0.0.6 landmark positions (reference data.py:139-167), same key order.
_R = _mirror({
    "RF_Coxa": (0.33, -0.17, 1.07), "RF_Femur": (0.33, -0.17, 0.67), "RF_Tibia": (0.33, -0.17, -0.02),
    "RF_Tarsus": (0.33, -0.17, -0.56), "RF_Claw": (0.33, -0.17, -1.19),
    "R_Antenna_base": (1.01, -0.10, 1.41), "R_Antenna_edge": (1.06, -0.10, 1.14),
    "R_post_vertical": (0.7, -0.2, 1.59), "R_wing": (0.08, -0.4, 1.43), "R_dorsal_hum": (0.41, -0.37, 1.32),
})
NMF_TEMPLATE = {k: _R[k] for k in (
    "RF_Coxa", "RF_Femur", "RF_Tibia", "RF_Tarsus", "RF_Claw",
    "LF_Coxa", "LF_Femur", "LF_Tibia", "LF_Tarsus", "LF_Claw",
    "R_Antenna_base", "L_Antenna_base", "R_Antenna_edge", "L_Antenna_edge",
    "R_post_vertical", "L_post_vertical", "R_wing", "L_wing")}
NMF_TEMPLATE["Neck"] = np.array([0.53, 0.0, 1.3])
NMF_TEMPLATE["Thorax_mid"] = np.array([0.08, 0.0, 1.43])
NMF_TEMPLATE["L_dorsal_hum"] = _R["L_dorsal_hum"]
NMF_TEMPLATE["R_dorsal_hum"] = _R["R_dorsal_hum"]

# ---------------------------------------------------------------------------------------------
# Six-leg locomotion set (reference examples/example_leg_inv_kinematics_parallel.py:21-140)
# ---------------------------------------------------------------------------------------------
_LOCO_R = {
    "RF": ((0.35, -0.27), (0.400, -0.025, -0.731, -1.249, -1.912)),
    "RM": ((0.0, -0.125), (0.0, -0.182, -0.965, -1.633, -2.328)),
    "RH": ((-0.215, -0.087), (-0.073, -0.272, -1.108, -1.793, -2.588)),
}
TEMPLATE_NMF_LOCOMOTION = {}
for _leg in ("RF", "LF", "RM", "LM", "RH", "LH"):
    (_x, _y), _zs = _LOCO_R["R" + _leg[1]]
    for _seg, _z in zip(("Coxa", "Femur", "Tibia", "Tarsus", "Claw"), _zs):
        TEMPLATE_NMF_LOCOMOTION[f"{_leg}_{_seg}"] = np.array([_x, _y if _leg[0] == "R" else -_y, _z])

_LOCO_PITCH = {"F": -0.07, "M": 0.37, "H": 0.07}
INITIAL_ANGLES_LOCOMOTION = {}
for _leg in ("RF", "LF", "RM", "LM", "RH", "LH"):
    _sgn = 1.0 if _leg[0] == "R" else -1.0
    INITIAL_ANGLES_LOCOMOTION[_leg] = _stage_seeds(
        0.45 * _sgn, _LOCO_PITCH[_leg[1]], -0.32 * _sgn, ctr_roll=-1.25 * _sgn)

_PI = 3.141592653589793
BOUNDS_LOCOMOTION = {}
for _side in ("R", "L"):
    _roll = (-_PI, 0) if _side == "R" else (0, _PI)
    BOUNDS_LOCOMOTION.update({
        f"{_side}F_ThC_yaw": (-_PI, _PI), f"{_side}F_ThC_pitch": (_deg(-90), _deg(90)), f"{_side}F_ThC_roll": (-_PI, _PI),
        f"{_side}F_CTr_pitch": (-_PI, _PI), f"{_side}F_FTi_pitch": (-_PI, _PI), f"{_side}F_CTr_roll": (-_PI, _PI),
        f"{_side}F_TiTa_pitch": (-_PI, _deg(0)),
        f"{_side}M_ThC_yaw": (_deg(-50), _deg(50)), f"{_side}M_ThC_pitch": (-_PI, _PI), f"{_side}M_ThC_roll": _roll,
        f"{_side}M_CTr_pitch": (-_PI, _PI), f"{_side}M_FTi_pitch": (-_PI, _PI), f"{_side}M_CTr_roll": (-_PI, _PI),
        f"{_side}M_TiTa_pitch": (-_PI, _deg(0)),
        f"{_side}H_ThC_yaw": (_deg(-50), _deg(50)), f"{_side}H_ThC_pitch": (_deg(-50), _deg(50)), f"{_side}H_ThC_roll": _roll,
        f"{_side}H_CTr_pitch": (_deg(-180), _deg(0)), f"{_side}H_FTi_pitch": (-_PI, _PI), f"{_side}H_CTr_roll": (-_PI, _PI),
        f"{_side}H_TiTa_pitch": (-_PI, _deg(0)),
    })
