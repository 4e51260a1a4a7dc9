"""Pose alignment on the GPU (drop-in for the ``AlignPose`` class of ``seqikpy.alignment``).

``AlignPose(pose_data_dict, legs_list, ...).align_pose()`` keeps the reference's signature, key
order, ``Neck`` entry and pickle name (reference seqikpy/alignment.py:229-555).  The statistics --
mid-quantiles of per-frame series (:83-87, 392-415, 425-434) -- run as a segmented radix select
on the device, the scale/translate map (:471-485, 547-553) as an elementwise kernel; for the
batched solver the same map is fused into the solver's pose load (``engine.leg_solve(affine=...)``).
The raw-format converters (:103-226) are host-side reshapes kept under the reference's names;
``from_file_path`` takes a ``convert_func`` exactly like the reference.  No CPU path for the per-frame work.
"""
from pathlib import Path
from typing import Callable, Dict, List, Literal, Optional, Union
import logging
import pickle

import numpy as np

from . import _native as N
from . import engine
from .data import NMF_TEMPLATE, PTS2ALIGN
from .utils import calculate_body_size, dict_to_nparray_pose, save_file

logging.basicConfig(format=" %(asctime)s - %(levelname)s- %(message)s", handlers=[logging.StreamHandler()])


def _leg_length_model(body_size: Dict[str, float], leg_name: str, claw_is_ee: bool) -> float:
    """Model leg length, without the tarsus unless the claw is aligned (reference :90-95)."""
    if claw_is_ee:
        return body_size[leg_name]
    return body_size[leg_name] - body_size[f"{leg_name}_Tarsus"]


# ---------------------------------------------------------------------------------------------
# Raw-format converters (host-side gather/reshape; reference alignment.py:103-226).  They feed AlignPose through
# ``from_file_path(convert_func=...)`` exactly like the reference's.
# ---------------------------------------------------------------------------------------------
def convert_from_anipose_to_dict(pose_3d: Dict[str, np.ndarray], pts2align: Dict[str, List[str]]) -> Dict[str, np.ndarray]:
    """anipose table ({"<keypoint>_x|_y|_z": (N,)}) -> {"<segment>": (N, n_key_points, 3)} for the segments of ``pts2align``."""
    out = {}
    for segment, key_points in pts2align.items():
        out[segment] = np.stack(
            [np.stack([np.asarray(pose_3d[f"{kp}_{ax}"], dtype=float) for ax in "xyz"], axis=-1) for kp in key_points], axis=1)
    return out


def convert_from_df3d_to_dict(pose_3d: np.ndarray, pts2align: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """DeepFly3D array (N, n_key_points, 3) -> {"<segment>": (N, k, 3)} with ``pts2align`` mapping segments to indices."""
    return {segment: np.array(pose_3d[:, idx, :], dtype=float) for segment, idx in pts2align.items()}


def convert_from_df3dpp_to_dict(pose_3d: Dict[str, Dict[str, np.ndarray]],
                                pts2align: Optional[List[str]] = None) -> Dict[str, np.ndarray]:
    """DeepFly3DPostProcessing dictionary -> {"<leg>_leg": (N, 5, 3)} for the legs in ``pts2align`` (default: all)."""
    segments = list(pose_3d.keys()) if pts2align is None else pts2align
    return {segment: dict_to_nparray_pose(pose_3d[segment], claw_is_end_effector=True) for segment in segments}


class AlignPose:
    """Scales and translates raw 3D key points onto the biomechanical template."""

    def __init__(
        self,
        pose_data_dict: Dict[str, np.ndarray],
        legs_list: List[str],
        include_claw: Optional[bool] = False,
        body_template: Optional[Dict[str, np.ndarray]] = None,
        body_size: Optional[Dict[str, float]] = None,
        log_level: Literal["DEBUG", "INFO", "WARNING", "ERROR"] = "INFO",
        device: str = "cuda",
    ) -> None:
        self.pose_data_dict = pose_data_dict
        self.include_claw = include_claw
        self.body_template = NMF_TEMPLATE if body_template is None else body_template
        self.body_size = calculate_body_size(self.body_template, legs_list) if body_size is None else body_size
        self.device = device
        self.logger = logging.getLogger(self.__class__.__name__)
        self.logger.setLevel(getattr(logging, log_level.upper(), None))

    @classmethod
    def from_file_path(
        cls, main_dir: Union[str, Path],
        file_name: Optional[str] = "pose3d.*",
        convert_func: Optional[Callable] = None,
        pts2align: Optional[Dict[str, List[str]]] = None,
        **kwargs
    ):
        """Loads the newest pickle matching ``file_name`` under ``main_dir`` (FileNotFoundError if none),
        optionally converts it with ``convert_func(pose_3d, pts2align)`` (reference :291-343)."""
        paths = list(Path(main_dir).rglob(file_name))
        if len(paths) > 0:
            with open(paths[-1].as_posix(), "rb") as f:
                pose_3d = pickle.load(f)
        else:
            raise FileNotFoundError(f"{file_name} does not exits in {main_dir}")
        if convert_func is not None:
            pts2align = PTS2ALIGN if pts2align is None else pts2align
            return cls(convert_func(pose_3d, pts2align), **kwargs)
        return cls(pose_3d, **kwargs)

    def align_pose(self, export_path: Optional[Union[str, Path]] = None) -> Dict[str, np.ndarray]:
        """Aligns every ``*leg*`` and ``*head*`` entry; other entries are dropped; adds the template ``Neck``."""
        aligned_pose = {}
        for segment, segment_array in self.pose_data_dict.items():
            if "leg" in segment:
                aligned_pose[segment] = self.align_leg(leg_array=segment_array, leg_name=segment[:2])
            elif "head" in segment:
                aligned_pose[segment] = self.align_head(head_array=segment_array, side=segment[0])
            else:
                self.logger.debug("%s is not aligned", segment)
                continue
        if "Neck" in self.body_template:
            aligned_pose["Neck"] = self.body_template["Neck"].reshape((-1, 1, 3))
        if export_path is not None:
            export_full_path = Path(export_path) / "pose3d_aligned.pkl"
            save_file(out_fname=export_full_path, data=aligned_pose)
            self.logger.info("Aligned pose is saved at %s", export_path)
        return aligned_pose

    @property
    def thorax_mid_pts(self) -> np.ndarray:
        """Middle point of the right and left wing hinges."""
        assert "Thorax" in self.pose_data_dict, "To align the head, you need to have a `Thorax` key point"
        thorax_pts = self.pose_data_dict["Thorax"]
        return 0.5 * (thorax_pts[:, 0, :] + thorax_pts[:, -1, :])

    # ------------------------------------------------------------------ statistics (device)
    def _to_device(self, arr):
        torch = N.require_cuda()
        N.load_library()
        return torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).to(torch.device(self.device))

    def _leg_affine(self, leg_array, leg_name):
        torch = N.require_cuda()
        leg_array = np.asarray(leg_array)
        if leg_array.ndim != 3 or leg_array.shape[1] < 5 or leg_array.shape[2] != 3:
            raise ValueError(f"leg array must be (N, 5, 3), got {leg_array.shape}")
        d_pose = self._to_device(leg_array[None, :, :5])
        tmpl = np.asarray(self.body_template[f"{leg_name}_Coxa"], dtype=float)
        consts = np.concatenate([tmpl, [_leg_length_model(self.body_size, leg_name, self.include_claw)]])
        d_consts = torch.from_numpy(consts.astype(np.float32)[None]).to(d_pose.device)
        d_aff = engine.leg_affine(d_pose, d_consts, include_claw=self.include_claw)
        return d_pose, d_aff

    def get_mean_length(self, segment_array: np.ndarray, segment_is_leg: bool) -> Dict[str, float]:
        """Mid-quantile length of each segment (reference :402-415)."""
        torch = N.require_cuda()
        segment_array = np.asarray(segment_array, dtype=float)
        lengths = np.linalg.norm(np.diff(segment_array, axis=1), axis=2)       # series construction only
        names = ["coxa", "femur", "tibia", "tarsus"] if segment_is_leg else ["antenna"]
        d_series = self._to_device(lengths.T[:len(names)])
        vals = engine.mid_quantile(d_series).cpu().numpy().astype(float)
        return {s: float(vals[i]) for i, s in enumerate(names)}

    @staticmethod
    def get_fixed_pos(points_3d: np.ndarray, device: str = "cuda") -> np.ndarray:
        """Per-coordinate mid-quantile of a steady key point (reference :392-400)."""
        torch = N.require_cuda()
        N.load_library()
        d = torch.from_numpy(np.ascontiguousarray(np.asarray(points_3d, dtype=np.float32).T)).to(torch.device(device))
        return engine.mid_quantile(d).cpu().numpy().astype(float)

    def find_scale_leg(self, leg_name: str, mean_length: Dict) -> float:
        """Model size / fly size (reference :417-423)."""
        nmf_size = _leg_length_model(self.body_size, leg_name, self.include_claw)
        fly_leg_size = mean_length["coxa"] + mean_length["femur"] + mean_length["tibia"]
        fly_leg_size += mean_length["tarsus"] if self.include_claw else 0
        return nmf_size / fly_leg_size

    def find_stationary_indices(self, array: np.ndarray, threshold: Optional[float] = 5e-5) -> np.ndarray:
        """Indices where the second difference is below ``threshold`` (signed, reference :425-434)."""
        indices_stat = np.where((np.diff(np.diff(array)) < threshold))
        assert indices_stat, f"Threshold ({threshold}) is too low to find stationary points, please increase it."
        return indices_stat[0]

    # ------------------------------------------------------------------ maps
    def align_leg(self, leg_array: np.ndarray, leg_name: Literal["RF", "LF", "RM", "LM", "RH", "LH"]) -> np.ndarray:
        """Coxa moved to the template coxa, the other four key points scaled about the fixed coxa (reference :436-487)."""
        d_pose, d_aff = self._leg_affine(leg_array, leg_name)
        self.logger.info("Scale factor for %s leg: %s", leg_name, float(d_aff[0, 3]))
        out = engine.align_apply(d_pose, d_aff)[0].cpu().numpy().astype(np.float64)
        return out

    def align_head(self, head_array: np.ndarray, side: str) -> np.ndarray:
        """Antenna base/tip scaled about the stationary antenna base (reference :489-555)."""
        torch = N.require_cuda()
        assert "Thorax" in self.pose_data_dict, "To align the head, you need to have a `Thorax` key point"
        if self.body_size.get("Antenna_mid_thorax") and self.body_size.get("Antenna"):
            antbase2thoraxmid_tmp = self.body_size["Antenna_mid_thorax"]
            ant_tmp = self.body_size["Antenna"]
        else:
            raise KeyError(
                """Nmf template dictionary does not contain
                a key name <Antenna_mid_thorax> or <Antenna>
                Please check the dictionary you provided.""")
        head_array = np.asarray(head_array)
        dev = torch.device(self.device)
        # statistics from the float64 key points (the stationarity threshold is ill-conditioned in float32)
        d_head64 = torch.from_numpy(np.ascontiguousarray(head_array[None, :, :2], dtype=np.float64)).to(dev)
        d_thorax = torch.from_numpy(np.ascontiguousarray(np.asarray(self.pose_data_dict["Thorax"])[None], dtype=np.float64)).to(dev)
        d_head = d_head64.to(torch.float32)
        tmpl = np.asarray(self.body_template[f"{side}_Antenna_base"], dtype=float)
        consts = np.concatenate([tmpl, [antbase2thoraxmid_tmp, ant_tmp]]).astype(np.float32)
        d_consts = torch.from_numpy(consts[None]).to(d_head.device)
        d_aff, d_counts = engine.head_affine(d_head64, d_thorax, d_consts)
        aff = d_aff[0].cpu().numpy()
        assert int(d_counts[0, 0]) > 0, "Threshold (5e-05) is too low to find stationary points, please increase it."
        self.logger.info("Scale factor antenna base %s: %s, ant itself: %s", side, aff[3], aff[7])
        return engine.head_apply(d_head, d_aff)[0].cpu().numpy().astype(np.float64)
