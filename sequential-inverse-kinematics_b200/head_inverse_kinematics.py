"""Head and antenna joint angles on the GPU (drop-in for ``seqikpy.head_inverse_kinematics``).

``HeadInverseKinematics(aligned_pos, body_template).compute_head_angles()`` keeps the reference's
signature, key order and pickle name (reference seqikpy/head_inverse_kinematics.py:53-339).  The
seven per-frame angles -- signed angles between projected vectors, the antenna vectors first
de-rotated by the head roll (:163-307) -- are one elementwise CUDA kernel
(``seqik_head_angles_f32``); only the two rest angles of the template (:309-329, two scalars)
are evaluated on the host.  No CPU path for the per-frame work.
"""
from collections import namedtuple
from pathlib import Path
from typing import Dict, Literal, Optional, Union
import logging

import numpy as np

from . import _native as N
from . import engine
from .utils import save_file

AxesTuple = namedtuple("AxesTuple", "X_AXIS Y_AXIS Z_AXIS")
Axes = AxesTuple(X_AXIS=np.array([1, 0, 0]), Y_AXIS=np.array([0, 1, 0]), Z_AXIS=np.array([0, 0, 1]))

logging.basicConfig(format=" %(asctime)s - %(levelname)s- %(message)s", handlers=[logging.StreamHandler()])

_KEYS = ("Angle_head_roll", "Angle_head_pitch", "Angle_head_yaw", "Angle_antenna_yaw_L", "Angle_antenna_pitch_L",
         "Angle_antenna_yaw_R", "Angle_antenna_pitch_R")


def _signed_angle_xz(v1, v2):
    """angle_between_segments(v1, v2, Y) for two vectors of the x-z plane (template rest angles only)."""
    v1 = np.asarray(v1, dtype=float)
    v2 = np.asarray(v2, dtype=float)
    cosang = np.dot(v1, v2) / (np.linalg.norm(v1) * np.linalg.norm(v2))
    det = v1[2] * v2[0] - v1[0] * v2[2]          # Y . (v1 x v2)
    return np.arccos(cosang) * (1.0 if det > 0 else -1.0)


class HeadInverseKinematics:
    """Head roll/pitch/yaw and antenna pitch/yaw from the aligned antenna bases/tips and the neck."""

    def __init__(
        self,
        aligned_pos: Dict[str, np.ndarray],
        body_template: Dict[str, np.ndarray],
        log_level: Literal["DEBUG", "INFO", "WARNING", "ERROR"] = "INFO",
        device: str = "cuda",
    ) -> None:
        self.aligned_pos = aligned_pos
        self.body_template = body_template
        self.device = device
        if not all(key in self.aligned_pos for key in ["R_head", "L_head", "Neck"]):
            raise ValueError(
                """self.aligned_pos must have R_head, L_head, Neck as keys,
                at least one of them is missing in the current data""")
        r, l = np.asarray(self.aligned_pos["R_head"]), np.asarray(self.aligned_pos["L_head"])
        assert r.ndim == 3 and l.ndim == 3 and r.shape[2] == 3 and l.shape[2] == 3, f"""
                One of head vectors
                (R_head: {r.shape}, L_head: {l.shape})
                does not have the right shape (N,k,3).
                """
        self.rest_head_pitch = self.get_rest_head_pitch()
        self.rest_antenna_pitch = self.get_rest_antenna_pitch()
        self.logger = logging.getLogger(self.__class__.__name__)
        self.logger.setLevel(getattr(logging, log_level.upper(), None))
        self._cache = None

    # ------------------------------------------------------------------ template rest pose (host, two scalars)
    def get_rest_antenna_pitch(self) -> np.ndarray:
        head_vector = np.array(self.body_template["Neck"] - self.body_template["R_Antenna_base"], dtype=float)
        head_vector[1] = 0
        antenna_vector = np.array(self.body_template["R_Antenna_edge"] - self.body_template["R_Antenna_base"], dtype=float)
        antenna_vector[1] = 0
        return np.array([_signed_angle_xz(head_vector, antenna_vector)])

    def get_rest_head_pitch(self) -> np.ndarray:
        head_vector = np.array((self.body_template["R_Antenna_base"] + self.body_template["L_Antenna_base"]) * 0.5
                               - self.body_template["Neck"], dtype=float)
        head_vector[1] = 0
        return np.array([_signed_angle_xz(head_vector, Axes.X_AXIS)])

    # ------------------------------------------------------------------ device round trip
    def _all_angles(self) -> np.ndarray:
        """(7, N) float64, computed once per instance by the CUDA kernel."""
        if self._cache is not None:
            return self._cache
        torch = N.require_cuda()
        N.load_library()
        r = np.asarray(self.aligned_pos["R_head"], dtype=float)
        l = np.asarray(self.aligned_pos["L_head"], dtype=float)
        neck = np.asarray(self.aligned_pos["Neck"], dtype=float)
        n = r.shape[0]
        if l.shape[0] != n:
            raise ValueError("R_head and L_head must have the same number of frames")

        def two_points(a):
            # a head segment with a single key point has no antenna tip: repeat the base (antenna angles undefined)
            return a[:, :2] if a.shape[1] >= 2 else np.concatenate([a[:, :1], a[:, :1]], axis=1)
        dev = torch.device(self.device)
        d_r = torch.from_numpy(np.ascontiguousarray(two_points(r), dtype=np.float32)[None]).to(dev)
        d_l = torch.from_numpy(np.ascontiguousarray(two_points(l), dtype=np.float32)[None]).to(dev)
        # the reference takes aligned_pos["Neck"][:, 0, :] (head_inverse_kinematics.py:156-161): first key point per frame
        neck = neck[:, 0, :] if neck.ndim == 3 else neck.reshape(-1, 3)
        if neck.shape[0] == 1:
            d_neck = torch.from_numpy(neck.astype(np.float32)).to(dev)                 # (1, 3): one point per trial
        elif neck.shape[0] == n:
            d_neck = torch.from_numpy(np.ascontiguousarray(neck, dtype=np.float32)[None]).to(dev)
        else:
            raise ValueError(f"Neck must have 1 or {n} frames, got {neck.shape[0]}")
        rest = torch.tensor([[float(self.rest_head_pitch[0]), float(self.rest_antenna_pitch[0])]], dtype=torch.float32, device=dev)
        out = engine.head_angles(d_r, d_l, d_neck, rest)
        self._cache = out[0].cpu().numpy().astype(np.float64)
        return self._cache

    # ------------------------------------------------------------------ reference API
    def compute_head_angles(
        self,
        export_path: Union[str, Path] = None,
        compute_ant_angles: Optional[bool] = True,
    ) -> Dict[str, np.ndarray]:
        """Head (and antenna) joint angles; pickled as ``head_joint_angles.pkl`` when ``export_path`` is given."""
        ang = self._all_angles()
        head_angles = {}
        n_keys = 7 if compute_ant_angles else 3
        if compute_ant_angles and (np.asarray(self.aligned_pos["R_head"]).shape[1] < 2
                                   or np.asarray(self.aligned_pos["L_head"]).shape[1] < 2):
            raise IndexError("antenna angles need two key points (base, tip) in R_head and L_head")
        for i in range(n_keys):
            head_angles[_KEYS[i]] = ang[i].copy()
        if export_path is not None:
            save_file(Path(export_path) / "head_joint_angles.pkl", head_angles)
            self.logger.info("Head joint angles are saved at %s!", export_path)
        return head_angles

    def compute_head_roll(self) -> np.ndarray:
        return self._all_angles()[0].copy()

    def compute_head_pitch(self) -> np.ndarray:
        return self._all_angles()[1].copy()

    def compute_head_yaw(self) -> np.ndarray:
        return self._all_angles()[2].copy()

    def _check_head_roll(self, head_roll) -> None:
        """The kernel de-rotates the antenna vectors by the head roll IT computes (what ``compute_head_angles`` hands
        to these methods in the reference, head_inverse_kinematics.py:127-134).  A caller-supplied roll that differs
        from it cannot be honoured by the fused kernel: refuse loudly instead of returning angles for another roll."""
        if head_roll is None:
            return
        roll = np.asarray(head_roll, dtype=float)
        own = self._all_angles()[0]
        if roll.shape != own.shape or not np.allclose(roll, own, rtol=0.0, atol=1e-5):
            raise ValueError("head_roll differs from the head roll computed from aligned_pos; the device kernel "
                             "de-rotates by the computed roll (pass compute_head_roll() or None)")

    def compute_antenna_yaw(self, side: Literal["R", "L"], head_roll: np.ndarray = None) -> np.ndarray:
        """``head_roll``: None, or the roll of ``compute_head_roll()`` (anything else raises, see _check_head_roll)."""
        side = side.upper()
        if side not in {"R", "L"}:
            raise ValueError("Side should be either R or L")
        self._check_head_roll(head_roll)
        return self._all_angles()[3 if side == "L" else 5].copy()

    def compute_antenna_pitch(self, side: Literal["R", "L"], head_roll: np.ndarray = None) -> np.ndarray:
        side = side.upper()
        if side not in {"R", "L"}:
            raise ValueError("Side should be either R or L")
        self._check_head_roll(head_roll)
        return self._all_angles()[4 if side == "L" else 6].copy()
