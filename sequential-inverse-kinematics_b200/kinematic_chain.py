"""Leg kinematic chains of the sequential IK (drop-in for ``seqikpy.kinematic_chain``).

The reference builds an ikpy ``Chain`` of ``URDFLink`` objects per stage -- and, for
stages 2-4, per FRAME (seqikpy/kinematic_chain.py:152-421, leg_inverse_kinematics.py:259-267).
Here a chain is a plain description (link names, axes, offsets, bounds, frozen angles): the
arithmetic happens in the CUDA solver, which receives the per-leg constants packed by
:meth:`KinematicChainSeq.pack_chain_params` (layout: include/seqik.h).

Topology (reference kinematic_chain.py, proximal -> distal; translation before rotation):

  idx link            offset        axis  solved in stage   frozen in stages
  0   Base link       -             -     -                 -
  1   {leg}_ThC_yaw   0             X     1                 2,3,4
  2   {leg}_ThC_pitch 0             Y     1                 2,3,4
  3   {leg}_ThC_roll  0             Z     2                 3,4        (absent in stage 1)
  4   {leg}_CTr_pitch (0,0,-Coxa)   Y     2 (inert end of 1) 3,4
  5   {leg}_CTr_roll  0             Z     3                 4          (absent in stages 1,2)
  6   {leg}_FTi_pitch (0,0,-Femur)  Y     3 (inert end of 2) 4
  7   {leg}_TiTa_pitch(0,0,-Tibia)  Y     4 (inert end of 3) -
  8   {leg}_Claw      (0,0,-Tarsus) none  inert end of 4    -          bounds (-pi, pi)
"""
from abc import ABC, abstractmethod
from collections import namedtuple
from typing import Dict, List, Optional

import numpy as np

from .data import NMF_TEMPLATE
from .utils import calculate_body_size

AxesTuple = namedtuple("AxesTuple", "X_AXIS Y_AXIS Z_AXIS")
Axes = AxesTuple(X_AXIS=[1, 0, 0], Y_AXIS=[0, 1, 0], Z_AXIS=[0, 0, 1])

LEG_NAMES = ("RF", "LF", "RM", "LM", "RH", "LH")
DOF_ORDER = ("ThC_yaw", "ThC_pitch", "ThC_roll", "CTr_pitch", "CTr_roll", "FTi_pitch", "TiTa_pitch")
SEGMENTS = ("Coxa", "Femur", "Tibia", "Tarsus")

# (dof, rotation axis, proximal segment whose length offsets the link along -z)
_LINK_TABLE = (
    ("ThC_yaw", Axes.X_AXIS, None), ("ThC_pitch", Axes.Y_AXIS, None), ("ThC_roll", Axes.Z_AXIS, None),
    ("CTr_pitch", Axes.Y_AXIS, "Coxa"), ("CTr_roll", Axes.Z_AXIS, None), ("FTi_pitch", Axes.Y_AXIS, "Femur"),
    ("TiTa_pitch", Axes.Y_AXIS, "Tibia"),
)
# DOFs present in the chain of each stage, and how many of them (from the front) are frozen
_STAGE_LINKS = {
    1: (("ThC_yaw", "ThC_pitch", "CTr_pitch"), 0),
    2: (("ThC_yaw", "ThC_pitch", "ThC_roll", "CTr_pitch", "FTi_pitch"), 2),
    3: (("ThC_yaw", "ThC_pitch", "ThC_roll", "CTr_pitch", "CTr_roll", "FTi_pitch", "TiTa_pitch"), 4),
    4: (("ThC_yaw", "ThC_pitch", "ThC_roll", "CTr_pitch", "CTr_roll", "FTi_pitch", "TiTa_pitch"), 6),
}
# chain slots (index in the stage's seed vector) that the stage actually solves
STAGE_ACTIVE_SLOTS = {1: (1, 2), 2: (3, 4), 3: (5, 6), 4: (7,)}
STAGE_ACTIVE_DOFS = {1: ("ThC_yaw", "ThC_pitch"), 2: ("ThC_roll", "CTr_pitch"),
                     3: ("CTr_roll", "FTi_pitch"), 4: ("TiTa_pitch",)}


class Link:
    """Description of one chain link (the fields ikpy's ``URDFLink`` exposes)."""

    def __init__(self, name, origin_translation=(0, 0, 0), origin_orientation=(0, 0, 0), rotation=None,
                 joint_type="revolute", bounds=(-np.inf, np.inf)):
        self.name = name
        self.origin_translation = np.asarray(origin_translation, dtype=float)
        self.origin_orientation = np.asarray(origin_orientation, dtype=float)
        self.rotation = None if rotation is None else np.asarray(rotation, dtype=float)
        self.joint_type = joint_type
        self.bounds = tuple(bounds)

    def __repr__(self):
        return f"Link(name={self.name!r}, joint_type={self.joint_type!r}, bounds={self.bounds})"


class Chain:
    """Ordered list of links with a name (what the hot path needs of ikpy's ``Chain``)."""

    def __init__(self, name: str, links: List[Link]):
        self.name = name
        self.links = list(links)

    def __len__(self):
        return len(self.links)

    @property
    def bounds(self):
        return [link.bounds for link in self.links]


class KinematicChainBase(ABC):
    """Holds joint limits and segment sizes (reference kinematic_chain.py:27-74)."""

    def __init__(self, bounds_dof: Dict[str, np.ndarray], legs_list: List[str],
                 body_size: Optional[Dict[str, float]] = None) -> None:
        self.body_size = calculate_body_size(NMF_TEMPLATE, legs_list) if body_size is None else body_size
        self.bounds_dof = bounds_dof

    def __call__(self):
        print("Base kinematic chain is called.")

    @abstractmethod
    def create_leg_chain(self, leg_name: str, **kwargs) -> Chain:
        raise NotImplementedError


class KinematicChainSeq(KinematicChainBase):
    """Stage-by-stage chains in yaw-pitch-roll order (reference kinematic_chain.py:77-421)."""

    def __call__(self):
        print("Sequential kinematic chain is called.")

    def create_leg_chain(self, leg_name: str, **kwargs) -> Chain:
        """Chain of ``leg_name`` at ``stage`` (1-4); stages 2-4 freeze the earlier DOFs at
        ``angles["Angle_{leg}_{dof}"][t]``.  ValueError on an unknown leg or stage."""
        angles = kwargs.get("angles", None)
        stage = kwargs.get("stage", 1)
        t = kwargs.get("t", 0)
        if leg_name not in LEG_NAMES:
            raise ValueError(f"Unknown leg name ({leg_name}) is provided!")
        if not 1 <= stage <= 4:
            raise ValueError(f"Unknown stage number ({stage}) number is provided!")
        return self._build(leg_name, stage, angles, t)

    def create_leg_chain_stage_1(self, leg_name: str) -> Chain:
        """Coxa only: solves ThC yaw and pitch."""
        return self._build(leg_name, 1, None, 0)

    def create_leg_chain_stage_2(self, leg_name: str, angles: Dict[str, np.ndarray], t: int) -> Chain:
        """Coxa + femur: solves ThC roll and CTr pitch."""
        return self._build(leg_name, 2, angles, t)

    def create_leg_chain_stage_3(self, leg_name: str, angles: Dict[str, np.ndarray], t: int) -> Chain:
        """Coxa + femur + tibia: solves CTr roll and FTi pitch."""
        return self._build(leg_name, 3, angles, t)

    def create_leg_chain_stage_4(self, leg_name: str, angles: Dict[str, np.ndarray], t: int) -> Chain:
        """Whole leg: solves TiTa pitch."""
        return self._build(leg_name, 4, angles, t)

    def _build(self, leg, stage, angles, t):
        dofs, n_frozen = _STAGE_LINKS[stage]
        table = {d: (axis, seg) for d, axis, seg in _LINK_TABLE}
        links = [Link("Base link", joint_type="fixed")]
        for i, dof in enumerate(dofs):
            axis, seg = table[dof]
            offset = (0, 0, 0) if seg is None else (0, 0, -self.body_size[f"{leg}_{seg}"])
            if i < n_frozen:
                orient = [0.0, 0.0, 0.0]
                orient["XYZ".index("XYZ"[int(np.argmax(axis))])] = angles[f"Angle_{leg}_{dof}"][t]
                links.append(Link(f"{leg}_{dof}", offset, orient, None, "fixed", self.bounds_dof[f"{leg}_{dof}"]))
            else:
                links.append(Link(f"{leg}_{dof}", offset, (0, 0, 0), axis, "revolute", self.bounds_dof[f"{leg}_{dof}"]))
        if stage == 4:
            links.append(Link(f"{leg}_Claw", (0, 0, -self.body_size[f"{leg}_Tarsus"]), (0, 0, 0), [0, 0, 0],
                              "revolute", (-np.pi, np.pi)))
        return Chain(name=f"chain_stage_{stage}", links=links)

    # ------------------------------------------------------------------ device packing
    def stage_bounds(self, leg: str, stage: int):
        """(lb, ub) arrays over ALL slots of the stage chain, Base link and inert end included
        -- the arrays scipy's feasibility check sees (least_squares.py:900)."""
        dofs, _ = _STAGE_LINKS[stage]
        lb = [-np.inf] + [self.bounds_dof[f"{leg}_{d}"][0] for d in dofs]
        ub = [np.inf] + [self.bounds_dof[f"{leg}_{d}"][1] for d in dofs]
        if stage == 4:
            lb.append(-np.pi)
            ub.append(np.pi)
        return np.asarray(lb, dtype=float), np.asarray(ub, dtype=float)

    def check_seed(self, leg: str, stage: int, seed) -> None:
        """Raises the ValueError scipy raises when ANY slot of the seed is outside its bounds."""
        lb, ub = self.stage_bounds(leg, stage)
        seed = np.asarray(seed, dtype=float)
        if seed.shape != lb.shape:
            raise ValueError(
                f"Inconsistent shapes between bounds and `x0`: stage {stage} of {leg} needs {lb.size} "
                f"initial angles, got {seed.size}.")
        if not np.all((seed >= lb) & (seed <= ub)):
            raise ValueError("Initial guess is outside of provided bounds")

    def pack_chain_params(self, leg: str, initial_angles: Dict[str, np.ndarray], stages=(1, 2, 3, 4)) -> np.ndarray:
        """The 32-float per-chain constant row of include/seqik.h for one leg."""
        row = np.zeros(32, dtype=np.float64)
        row[0:4] = [self.body_size[f"{leg}_{s}"] for s in SEGMENTS]
        row[4:11] = [self.bounds_dof[f"{leg}_{d}"][0] for d in DOF_ORDER]
        row[11:18] = [self.bounds_dof[f"{leg}_{d}"][1] for d in DOF_ORDER]
        for stage in (1, 2, 3, 4):
            key = f"stage_{stage}"
            if key not in initial_angles:
                if stage in stages:      # the reference indexes initial_angles[leg][f"stage_{stage}"]: KeyError
                    raise KeyError(key)
                continue
            seed = np.asarray(initial_angles[key], dtype=float)
            if stage in stages:
                self.check_seed(leg, stage, seed)
            act = STAGE_ACTIVE_SLOTS[stage]
            for slot, dof in zip(act, STAGE_ACTIVE_DOFS[stage]):
                row[18 + DOF_ORDER.index(dof)] = seed[slot]
            inert = np.delete(seed, act)
            row[25 + stage - 1] = float(np.dot(inert, inert))
        return row


# link order of the generic chain (reference kinematic_chain.py:464-530): roll comes FIRST
GENERIC_DOF_ORDER = ("ThC_roll", "ThC_yaw", "ThC_pitch", "CTr_pitch", "CTr_roll", "FTi_pitch", "TiTa_pitch")
_GENERIC_LINKS = (
    ("ThC_roll", Axes.Z_AXIS, None), ("ThC_yaw", Axes.X_AXIS, None), ("ThC_pitch", Axes.Y_AXIS, None),
    ("CTr_pitch", Axes.Y_AXIS, "Coxa"), ("CTr_roll", Axes.Z_AXIS, None), ("FTi_pitch", Axes.Y_AXIS, "Femur"),
    ("TiTa_pitch", Axes.Y_AXIS, "Tibia"),
)


class KinematicChainGeneric(KinematicChainBase):
    """One chain for the entire leg, every joint revolute (reference kinematic_chain.py:424-532)."""

    def __call__(self):
        print("Generic kinematic chain is called.")

    def create_leg_chain(self, leg_name: str, **kwargs) -> Chain:
        """Base link, ThC roll/yaw/pitch, CTr pitch/roll, FTi pitch, TiTa pitch, Claw.  ValueError on an unknown leg."""
        if leg_name not in LEG_NAMES:
            raise ValueError(f"Unknown leg name ({leg_name}) is provided!")
        links = [Link("Base link", joint_type="fixed")]
        for dof, axis, seg in _GENERIC_LINKS:
            offset = (0, 0, 0) if seg is None else (0, 0, -self.body_size[f"{leg_name}_{seg}"])
            links.append(Link(f"{leg_name}_{dof}", offset, (0, 0, 0), axis, "revolute", self.bounds_dof[f"{leg_name}_{dof}"]))
        links.append(Link(f"{leg_name}_Claw", (0, 0, -self.body_size[f"{leg_name}_Tarsus"]), (0, 0, 0), [0, 0, 0],
                          "revolute", (-np.pi, np.pi)))
        return Chain(name="chain", links=links)

    def chain_bounds(self, leg: str):
        """(lb, ub) over the 9 chain slots -- the arrays scipy's feasibility check sees."""
        lb = [-np.inf] + [self.bounds_dof[f"{leg}_{d}"][0] for d in GENERIC_DOF_ORDER] + [-np.pi]
        ub = [np.inf] + [self.bounds_dof[f"{leg}_{d}"][1] for d in GENERIC_DOF_ORDER] + [np.pi]
        return np.asarray(lb, dtype=float), np.asarray(ub, dtype=float)

    def check_seed(self, leg: str, seed) -> None:
        """Raises the ValueError scipy raises when ANY slot of the 9-vector seed is outside its bounds."""
        lb, ub = self.chain_bounds(leg)
        seed = np.asarray(seed, dtype=float)
        if seed.shape != lb.shape:
            raise ValueError(f"Inconsistent shapes between bounds and `x0`: the generic chain of {leg} needs {lb.size} "
                             f"initial angles, got {seed.size}.")
        if not np.all((seed >= lb) & (seed <= ub)):
            raise ValueError("Initial guess is outside of provided bounds")

    def pack_chain_params(self, leg: str, seed) -> np.ndarray:
        """The 32-float constants row of ``seqik_leg_solve_generic_f32`` (include/seqik.h) for one leg.  ``seed`` is the
        9-vector the reference hands to the generic chain POSITIONALLY (``initial_angles[leg]["stage_4"]``,
        leg_inverse_kinematics.py:583): slot i seeds link i whatever DOF the stage-4 vector meant it for."""
        seed = np.asarray(seed, dtype=float)
        self.check_seed(leg, seed)
        lb, ub = self.chain_bounds(leg)
        row = np.zeros(32, dtype=np.float64)
        row[0:4] = [self.body_size[f"{leg}_{s}"] for s in SEGMENTS]
        row[4:11] = lb[1:8]
        row[11:18] = ub[1:8]
        row[18:25] = seed[1:8]
        row[25] = seed[0] ** 2 + seed[8] ** 2
        return row
