"""seqik-b200: the leg inverse-kinematics hot path of SeqIKPy on NVIDIA B200 (sm_100a).

Import as ``seqikpy_b200`` (this directory's name is not a valid identifier; the
``seqikpy_b200`` package next to it points its ``__path__`` here):

    from seqikpy_b200.alignment import AlignPose
    from seqikpy_b200.kinematic_chain import KinematicChainSeq
    from seqikpy_b200.leg_inverse_kinematics import LegInvKinSeq
    from seqikpy_b200.head_inverse_kinematics import HeadInverseKinematics
    from seqikpy_b200.data import BOUNDS, INITIAL_ANGLES, NMF_TEMPLATE

Modules mirror ``seqikpy.*`` of the reference one to one for the hot path; ``engine`` is the
batched tensor-level API and ``_native`` the ctypes binding of ``csrc/libseqik_sm100.so``.
"""
__version__ = "0.1.0"
