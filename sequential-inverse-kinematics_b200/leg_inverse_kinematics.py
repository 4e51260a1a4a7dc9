"""Sequential leg inverse kinematics on the GPU (drop-in for ``seqikpy.leg_inverse_kinematics``).

``LegInvKinSeq(aligned_pos, kinematic_chain_class, initial_angles).run_ik_and_fk()`` keeps the
reference's signature, dictionary layout, key order, dtypes (float64 out) and pickle names
(reference seqikpy/leg_inverse_kinematics.py:136-403).  The per-(leg, stage, frame) Python loop
of the reference -- ``create_leg_chain`` + ikpy ``inverse_kinematics`` + scipy ``least_squares``
(:259-282) -- is replaced by ONE launch of the CUDA solver over all legs: stages 1-4 run back
to back on the device, frames stay serial inside a chain (warm start, :272), chains run in
parallel.  There is no CPU path: without ``libseqik_sm100.so`` and a CUDA device the calls raise.
"""
from abc import ABC, abstractmethod
from pathlib import Path
from typing import Dict, Literal, Optional, Tuple, Union
import logging

import numpy as np

from . import _native as N
from . import engine
from .data import INITIAL_ANGLES
from .kinematic_chain import (DOF_ORDER, GENERIC_DOF_ORDER, STAGE_ACTIVE_DOFS, KinematicChainBase, KinematicChainGeneric,
                              KinematicChainSeq)
from .utils import save_file

logging.basicConfig(format=" %(asctime)s - %(levelname)s- %(message)s", handlers=[logging.StreamHandler()])

_LEGS = ("RF", "LF", "RM", "LM", "RH", "LH")
# rows of the 9-row stage-4 FK that make up the shorter chains of stages 1-3 (kinematic_chain.py link lists)
_FK_ROWS = {1: (0, 1, 2, 4), 2: (0, 1, 2, 3, 4, 6), 3: (0, 1, 2, 3, 4, 5, 6, 7), 4: tuple(range(9))}


class LegInvKinBase(ABC):
    """Holds the aligned pose, the chain definitions and the seeds (reference :25-133)."""

    def __init__(
        self,
        aligned_pos: Dict[str, np.ndarray],
        kinematic_chain_class: KinematicChainBase,
        initial_angles: Optional[Dict[str, np.ndarray]] = None,
        log_level: Literal["DEBUG", "INFO", "WARNING", "ERROR"] = "INFO",
    ) -> None:
        self.aligned_pos = aligned_pos
        self.kinematic_chain_class = kinematic_chain_class
        self.initial_angles = INITIAL_ANGLES if initial_angles is None else initial_angles
        self.logger = logging.getLogger(self.__class__.__name__)
        self.logger.setLevel(getattr(logging, log_level.upper(), None))

    def get_scale_factor(self, vector: np.ndarray, length: float) -> float:
        """Ratio between ``length`` and the summed segment lengths of ``vector`` (reference :79-83)."""
        return length / np.sum(np.linalg.norm(np.diff(vector, axis=0), axis=1))

    @abstractmethod
    def calculate_ik_stage(self, end_effector_pos, origin, initial_angles, segment_name, **kwargs) -> np.ndarray:
        """Inverse kinematics of one stage over all frames."""

    @abstractmethod
    def run_ik_and_fk(self, export_path=None, **kwargs) -> Tuple[Dict[str, np.ndarray], Dict[str, np.ndarray]]:
        """Inverse + forward kinematics of all legs."""


class LegInvKinSeq(LegInvKinBase):
    """Sequential (coxa -> femur -> tibia -> tarsus) leg IK, see the module docstring.

    Extra, optional arguments over the reference: ``device`` (default ``"cuda"``) and the solver mode --
    ``reference_iterates=True`` (or an explicit ``flags`` word of include/seqik.h) makes the kernel walk the
    reference's own TRF iterates (``SEQIK_FLAG_REFERENCE_ITERATES``: same evaluation counts and termination
    statuses as scipy's).  The default, ``SEQIK_FLAG_DEFAULT``, adds Newton steps and the closed-form warm step:
    it ends at the box-constrained minimiser the reference's iteration converges to and stays within 1e-3 rad of
    the reference's shipped angles everywhere except inside the reference's own singular episodes (CTr_pitch = 0,
    where its path is driven by finite-difference rounding noise), max 1e-4 rad elsewhere.
    """

    def __init__(
        self,
        aligned_pos: Dict[str, np.ndarray],
        kinematic_chain_class: KinematicChainSeq,
        initial_angles: Optional[Dict[str, np.ndarray]] = None,
        log_level: Literal["DEBUG", "INFO", "WARNING", "ERROR"] = "INFO",
        device: str = "cuda",
        reference_iterates: bool = False,
        flags: Optional[int] = None,
    ) -> None:
        super().__init__(aligned_pos, kinematic_chain_class, initial_angles, log_level)
        self.joint_angles_dict = {}
        self.device = device
        #: solver flags handed to seqik_leg_solve_f32 (include/seqik.h)
        self.flags = int(flags) if flags is not None else (N.FLAG_REFERENCE_ITERATES if reference_iterates else N.FLAG_DEFAULT)
        #: solver statistics of the last call: {leg: {"nfev": (4,) evaluations per stage, "status": int}}
        self.solver_stats = {}

    # ------------------------------------------------------------------ device round trip
    def _solve(self, legs, pose, seeds, stages, angles_in=None):
        """legs: names; pose (n_leg, N, 5, 3) float; seeds: per leg {"stage_k": vector}.
        Returns float64 angles (n_leg, N, 7) and fk (n_leg, N, 9, 3)."""
        torch = N.require_cuda()
        N.load_library()
        chain = self.kinematic_chain_class
        params = np.stack([chain.pack_chain_params(leg, seeds[i], stages=stages) for i, leg in enumerate(legs)])
        dev = torch.device(self.device)
        d_pose = torch.from_numpy(np.ascontiguousarray(pose, dtype=np.float32)).to(dev)
        d_params = torch.from_numpy(params.astype(np.float32)).to(dev)
        d_angles = None
        if angles_in is not None:
            d_angles = torch.from_numpy(np.ascontiguousarray(angles_in, dtype=np.float32)).to(dev)
        d_angles, d_fk, d_status, d_nfev = engine.leg_solve(d_pose, d_params, stages=stages, angles=d_angles, want_fk=True,
                                                              flags=self.flags)
        angles = d_angles.cpu().numpy().astype(np.float64)
        fk = d_fk.cpu().numpy().astype(np.float64)
        status, nfev = d_status.cpu().numpy(), d_nfev.cpu().numpy()
        for i, leg in enumerate(legs):
            self.solver_stats[leg] = {"nfev": nfev[i].astype(np.int64), "status": int(status[i])}
            if status[i] < 0:       # what scipy.optimize.least_squares raises inside the reference's frame loop
                raise ValueError(f"Residuals are not finite in the initial point (leg {leg}: NaN/inf key points).")
            if status[i] == 0:
                self.logger.warning("Leg %s: at least one solve stopped at the evaluation limit", leg)
        return angles, fk

    def _frozen_angles(self, leg, first_stage, n_frames):
        """(N, 7) array holding the DOFs of the stages before ``first_stage`` from joint_angles_dict."""
        buf = np.zeros((n_frames, 7))
        for stage in range(1, first_stage):
            for dof in STAGE_ACTIVE_DOFS[stage]:
                key = f"Angle_{leg}_{dof}"
                if key not in self.joint_angles_dict:
                    raise KeyError(f"{key}: stage {stage} must be computed before stage {first_stage}")
                col = np.asarray(self.joint_angles_dict[key], dtype=float)
                if col.shape[0] != n_frames:
                    raise ValueError(f"{key} has {col.shape[0]} frames, the pose has {n_frames}")
                buf[:, DOF_ORDER.index(dof)] = col
        return buf

    def _store(self, leg, stages, angles):
        for stage in stages:
            for dof in STAGE_ACTIVE_DOFS[stage]:
                self.joint_angles_dict[f"Angle_{leg}_{dof}"] = angles[:, DOF_ORDER.index(dof)].copy()
            self.logger.debug("Stage %d is completed!", stage)

    # ------------------------------------------------------------------ reference API
    def calculate_ik_stage(
        self,
        end_effector_pos: np.ndarray,
        origin: np.ndarray,
        initial_angles: np.ndarray,
        segment_name: str,
        **kwargs
    ) -> np.ndarray:
        """One stage of one leg over all frames (reference :200-322).

        The DOFs solved by earlier stages are read from ``self.joint_angles_dict`` and frozen;
        the stage's angles are stored there.  Returns the joint positions of the stage's chain,
        shape ``(N, len(initial_angles), 3)`` (the reference returns uninitialised memory for
        stages 1-3; here those stages return the positions of their own, shorter chain).
        """
        stage = kwargs.get("stage", 1)
        if segment_name not in _LEGS:
            raise ValueError(f"Segment name ({segment_name}) is not valid.")
        if not 1 <= stage <= 4:
            raise ValueError(f"Stage ({stage}) should be between 1 and 4.")
        end_effector_pos = np.asarray(end_effector_pos, dtype=float).reshape(-1, 3)
        n_frames = end_effector_pos.shape[0]
        origin = np.asarray(origin, dtype=float)
        if origin.size == 3:
            origin = np.tile(origin.reshape(1, 3), (n_frames, 1))
        pose = np.zeros((1, n_frames, 5, 3))
        pose[0, :, 0] = origin
        pose[0, :, stage] = end_effector_pos
        seeds = [{f"stage_{stage}": np.asarray(initial_angles, dtype=float)}]
        angles_in = None if stage == 1 else self._frozen_angles(segment_name, stage, n_frames)[None]
        # (key points of the frozen stages are not needed: their joints follow from the frozen angles)
        angles, fk = self._solve([segment_name], pose, seeds, [stage], angles_in)
        self._store(segment_name, [stage], angles[0])
        return fk[0][:, _FK_ROWS[stage], :]

    def run_ik_and_fk(
        self,
        export_path: Union[Path, str] = None,
        **kwargs
    ) -> Tuple[Dict[str, np.ndarray], Dict[str, np.ndarray]]:
        """Joint angles and forward kinematics of every ``*_leg`` entry (reference :324-403).

        kwargs: ``stages`` (default [1, 2, 3, 4], must be consecutive), ``hide_progress_bar``
        (accepted for compatibility; a single kernel launch has no per-frame progress).
        Returns ``(joint_angles_dict, forward_kinematics_dict)``; with ``export_path`` both are
        pickled as ``leg_joint_angles.pkl`` / ``forward_kinematics.pkl``.
        """
        stages = list(kwargs.get("stages", [1, 2, 3, 4]))
        if max(stages) > 4 or not all(np.diff(stages) == 1):
            raise ValueError("Maximum stage number is 4 and the list should be strictly incremental.")
        if min(stages) < 1:
            raise ValueError("Maximum stage number is 4 and the list should be strictly incremental.")
        forward_kinematics_dict = {}
        self.logger.info("Computing joint angles and forward kinematics...")

        names, legs, arrays = [], [], []
        for segment_name, segment_array in self.aligned_pos.items():
            if "leg" not in segment_name.lower():
                self.logger.debug("Segment %s is not a leg, continuing...", segment_name)
                continue
            leg_name = segment_name.split("_")[0]
            if leg_name not in self.kinematic_chain_class.body_size:
                self.logger.warning("Leg %s is not in the kinematic chain, continuing...", leg_name)
                continue
            if leg_name not in _LEGS:
                raise ValueError(f"Segment name ({leg_name}) is not valid.")
            arr = np.asarray(segment_array)
            if arr.ndim != 3 or arr.shape[1] <= max(stages) or arr.shape[2] != 3:
                raise ValueError(f"{segment_name}: expected (N, >={max(stages) + 1}, 3), got {arr.shape}")
            names.append(segment_name)
            legs.append(leg_name)
            arrays.append(arr)

        # one launch per group of legs that share a frame count (normally: one launch)
        by_len = {}
        for i, arr in enumerate(arrays):
            by_len.setdefault(arr.shape[0], []).append(i)
        for n_frames, idx in by_len.items():
            pose = np.zeros((len(idx), n_frames, 5, 3))
            for j, i in enumerate(idx):
                k = min(5, arrays[i].shape[1])
                pose[j, :, :k] = arrays[i][:, :k]
            seeds = [self.initial_angles[legs[i]] for i in idx]
            angles_in = None
            if stages[0] > 1:
                angles_in = np.stack([self._frozen_angles(legs[i], stages[0], n_frames) for i in idx])
            if n_frames == 0:
                angles = np.zeros((len(idx), 0, 7))
                fk = np.zeros((len(idx), 0, 9, 3))
            else:
                angles, fk = self._solve([legs[i] for i in idx], pose, seeds, stages, angles_in)
            for j, i in enumerate(idx):
                self._store(legs[i], stages, angles[j])
                forward_kinematics_dict[names[i]] = fk[j][:, _FK_ROWS[stages[-1]], :]
        # restore the reference's insertion order (aligned_pos order)
        forward_kinematics_dict = {n: forward_kinematics_dict[n] for n in names}

        self.logger.debug("Joint angles and forward kinematics are computed.")
        if export_path is not None:
            save_file(Path(export_path) / "forward_kinematics.pkl", forward_kinematics_dict)
            save_file(Path(export_path) / "leg_joint_angles.pkl", self.joint_angles_dict)
            self.logger.info("Joint angles and forward kinematics are saved at %s", export_path)
        return self.joint_angles_dict, forward_kinematics_dict


class LegInvKinGeneric(LegInvKinBase):
    """Generic leg IK: all seven joints solved at once against ONE target, the claw (drop-in for the reference's
    ``LegInvKinGeneric``, seqikpy/leg_inverse_kinematics.py:406-613).

    One kernel launch solves every leg: frames serial inside a leg (warm start, :524), legs in parallel.  The device
    solver restates the reference's optimiser (scipy TRF on the chain of ``KinematicChainGeneric``) and reaches the same
    claw residual, but the problem is under-determined -- 3 equations, 7 unknowns -- and the reference's own answer
    depends on rounding noise: solve by solve from the same seed the two agree, a free-running recording
    follows its own path along the self-motion manifold.  There is no CPU path.

    Extra, optional arguments over the reference: ``device`` (default ``"cuda"``) and ``precision`` -- ``"float64"``
    (default: FP64 data and device arithmetic, which agrees with scipy solve by solve as often as scipy agrees with
    itself) or ``"float32"`` (faster, agrees a few percent less often; same claw residual bound).
    """

    def __init__(
        self,
        aligned_pos: Dict[str, np.ndarray],
        kinematic_chain_class: KinematicChainGeneric,
        initial_angles: Optional[Dict[str, np.ndarray]] = None,
        log_level: Literal["DEBUG", "INFO", "WARNING", "ERROR"] = "INFO",
        device: str = "cuda",
        precision: Literal["float64", "float32"] = "float64",
    ) -> None:
        super().__init__(aligned_pos, kinematic_chain_class, initial_angles, log_level)
        if precision not in ("float64", "float32"):
            raise ValueError(f"precision must be 'float64' or 'float32', got {precision!r}")
        self.joint_angles_dict = {}
        self.device = device
        self.precision = precision
        #: solver statistics of the last call: {leg: {"nfev": evaluations summed over frames, "status": int}}
        self.solver_stats = {}

    def _solve(self, legs, pose2, seeds):
        """legs: names; pose2 (n_leg, N, 2, 3) = ThC origin + end effector; seeds: 9-vector per leg.
        Returns float64 angles (n_leg, N, 7) in generic chain order and fk (n_leg, N, 9, 3)."""
        torch = N.require_cuda()
        N.load_library()
        chain = self.kinematic_chain_class
        params = np.stack([chain.pack_chain_params(leg, seeds[i]) for i, leg in enumerate(legs)])
        dev = torch.device(self.device)
        npdt = np.float64 if self.precision == "float64" else np.float32
        d_pose = torch.from_numpy(np.ascontiguousarray(pose2, dtype=npdt)).to(dev)
        d_params = torch.from_numpy(params.astype(npdt)).to(dev)
        d_angles, d_fk, d_status, d_nfev = engine.leg_solve_generic(d_pose, d_params, target_row=1, want_fk=True)
        angles = d_angles.cpu().numpy().astype(np.float64)
        fk = d_fk.cpu().numpy().astype(np.float64)
        status, nfev = d_status.cpu().numpy(), d_nfev.cpu().numpy()
        for i, leg in enumerate(legs):
            self.solver_stats[leg] = {"nfev": int(nfev[i]), "status": int(status[i])}
            if status[i] < 0:       # what scipy.optimize.least_squares raises inside the reference's frame loop
                raise ValueError(f"Residuals are not finite in the initial point (leg {leg}: NaN/inf key points).")
            if status[i] == 0:
                self.logger.warning("Leg %s: at least one solve stopped at the evaluation limit", leg)
        return angles, fk

    def _store(self, leg, angles):
        for i, dof in enumerate(GENERIC_DOF_ORDER):          # the reference's link order (Base and Claw skipped, :537-544)
            self.joint_angles_dict[f"Angle_{leg}_{dof}"] = angles[:, i].copy()

    def calculate_ik_stage(
        self,
        end_effector_pos: np.ndarray,
        origin: np.ndarray,
        initial_angles: np.ndarray,
        segment_name: str,
        **kwargs
    ) -> np.ndarray:
        """Generic IK of one leg over all frames (reference :473-547): ``end_effector_pos`` (N, 3) claw positions,
        ``origin`` (N, 3) or (3,) Thorax-Coxa joint, ``initial_angles`` the 9-vector seed of the chain.  Stores the
        joint angles in ``self.joint_angles_dict`` and returns the joint positions, shape (N, 9, 3)."""
        if segment_name not in _LEGS:
            raise ValueError(f"Segment name ({segment_name}) is not valid.")
        end_effector_pos = np.asarray(end_effector_pos, dtype=float).reshape(-1, 3)
        n_frames = end_effector_pos.shape[0]
        origin = np.asarray(origin, dtype=float)
        if origin.size == 3:
            origin = np.tile(origin.reshape(1, 3), (n_frames, 1))
        pose2 = np.stack([origin, end_effector_pos], axis=1)[None]
        if n_frames == 0:
            self.kinematic_chain_class.check_seed(segment_name, initial_angles)
            angles, fk = np.zeros((1, 0, 7)), np.zeros((1, 0, 9, 3))
        else:
            angles, fk = self._solve([segment_name], pose2, [np.asarray(initial_angles, dtype=float)])
        self._store(segment_name, angles[0])
        return fk[0]

    def run_ik_and_fk(
        self,
        export_path: Union[Path, str] = None,
        **kwargs
    ) -> Tuple[Dict[str, np.ndarray], Dict[str, np.ndarray]]:
        """Joint angles and forward kinematics of every ``*_leg`` entry (reference :549-613).  The end effector is the
        LAST key point of each leg array, the seed is ``initial_angles[leg]["stage_4"]`` (:582-583).  kwargs:
        ``hide_progress_bar`` (accepted for compatibility).  With ``export_path`` both dictionaries are pickled as
        ``leg_joint_angles.pkl`` / ``forward_kinematics.pkl``."""
        forward_kinematics_dict = {}
        self.logger.info("Computing joint angles and forward kinematics...")
        names, legs, arrays = [], [], []
        for segment_name, segment_array in self.aligned_pos.items():
            if "leg" not in segment_name.lower():
                self.logger.debug("Segment %s is not a leg, continuing...", segment_name)
                continue
            leg_name = segment_name.split("_")[0]
            if leg_name not in self.kinematic_chain_class.body_size:
                self.logger.warning("Leg %s is not in the kinematic chain, continuing...", leg_name)
                continue
            if leg_name not in _LEGS:
                raise ValueError(f"Segment name ({leg_name}) is not valid.")
            arr = np.asarray(segment_array)
            if arr.ndim != 3 or arr.shape[1] < 2 or arr.shape[2] != 3:
                raise ValueError(f"{segment_name}: expected (N, >=2, 3), got {arr.shape}")
            names.append(segment_name)
            legs.append(leg_name)
            arrays.append(arr)
        by_len = {}
        for i, arr in enumerate(arrays):
            by_len.setdefault(arr.shape[0], []).append(i)
        for n_frames, idx in by_len.items():
            pose2 = np.stack([arrays[i][:, (0, -1), :] for i in idx]).astype(float)
            seeds = [np.asarray(self.initial_angles[legs[i]]["stage_4"], dtype=float) for i in idx]
            if n_frames == 0:
                angles, fk = np.zeros((len(idx), 0, 7)), np.zeros((len(idx), 0, 9, 3))
            else:
                angles, fk = self._solve([legs[i] for i in idx], pose2, seeds)
            for j, i in enumerate(idx):
                self._store(legs[i], angles[j])
                forward_kinematics_dict[names[i]] = fk[j]
        forward_kinematics_dict = {n: forward_kinematics_dict[n] for n in names}
        self.logger.debug("Joint angles and forward kinematics are computed.")
        if export_path is not None:
            save_file(Path(export_path) / "forward_kinematics.pkl", forward_kinematics_dict)
            save_file(Path(export_path) / "leg_joint_angles.pkl", self.joint_angles_dict)
            self.logger.info("Joint angles and forward kinematics are saved at %s", export_path)
        return self.joint_angles_dict, forward_kinematics_dict
