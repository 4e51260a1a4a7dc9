"""The C-ABI library loads and exports every symbol include/seqik.h declares (no compute without a GPU)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def header_symbols():
    hdr = (ROOT / "include" / "seqik.h").read_text()
    return re.findall(r"^(?:int|const char\*)\s+(seqik_\w+)\(", hdr, re.M)


def test_library_exports_header_symbols():
    from seqikpy_b200 import _native as N
    assert N.LIB_PATH.exists(), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(str(N.LIB_PATH))
    names = header_symbols()
    assert len(names) >= 12
    for name in names:
        assert hasattr(lib, name), name
    lib.seqik_abi_version.restype = ctypes.c_int
    hdr = (ROOT / "include" / "seqik.h").read_text()
    assert lib.seqik_abi_version() == int(re.search(r"#define SEQIK_ABI_VERSION (\d+)", hdr).group(1)) == N.ABI_VERSION


def test_binding_covers_header():
    from seqikpy_b200 import _native as N
    assert sorted(N.EXPORTED_SYMBOLS) == sorted(header_symbols())
    N.load_library()            # sets argtypes for every symbol; raises if one is missing


def test_argument_validation_without_gpu():
    """Size/NULL checks happen before any CUDA call, so they can be exercised on a CPU-only host."""
    from seqikpy_b200 import _native as N
    lib = N.load_library()
    assert lib.seqik_leg_solve_f32(0, 0, 15, 0, 0, 0, 0, 7, 0, 0, 27, 0, 0, 0, 0, -1, 10, 0xF, 0, 0) == -1
    assert b"negative" in lib.seqik_last_error()
    assert lib.seqik_leg_solve_f32(0, 0, 15, 0, 0, 0, 0, 7, 0, 0, 27, 0, 0, 0, 0, 0, 10, 0xF, 0, 0) == 0      # empty: ok
    assert lib.seqik_leg_solve_f32(0, 0, 15, 0, 0, 0, 0, 7, 0, 0, 27, 0, 0, 0, 0, 4, 10, 0xF, 0, 0) == -1     # NULL pose
    assert lib.seqik_leg_solve_f32(8, 150, 15, 0, 8, 8, 70, 7, 0, 0, 27, 0, 0, 0, 0, 4, 10, 0x5, 0, 0) == -1  # mask with a hole
    assert lib.seqik_mid_quantile_f32(8, 0, 8, 8, 3, 0, 0) == -1                                         # empty series
    assert lib.seqik_head_angles_f32(8, 8, 8, 5, 0, 0, 8, 8, 1, 1, 0) == -1                              # bad neck stride
    assert lib.seqik_memcpy2d_async(8, 4, 8, 16, 8, 2, 1, 0) == -1                                       # pitch < width
    assert lib.seqik_memcpy2d_async(8, 16, 8, 16, 8, 2, 3, 0) == -1                                      # bad direction
    gen = [0, 0, 15, 4, 0, 0, 0, 7, 0, 0, 27, 0, 0, 0, 0]
    assert lib.seqik_leg_solve_generic_f32(*gen, -1, 10, 0, 0) == -1                                       # negative size
    assert lib.seqik_leg_solve_generic_f64(*gen, 0, 10, 0, 0) == 0                                         # empty: ok
    assert lib.seqik_leg_solve_generic_f32(*gen, 4, 10, 0, 0) == -1                                        # NULL pose
    assert lib.seqik_leg_solve_generic_f32(8, 150, 15, 0, 8, 8, 70, 7, 0, 0, 27, 0, 0, 0, 0, 4, 10, 0, 0) == -1   # target_row 0
    assert lib.seqik_leg_solve_generic_f32(8, 150, 15, 5, 8, 8, 70, 7, 0, 0, 27, 0, 0, 0, 0, 4, 10, 0, 0) == -1   # row outside the stride
    assert lib.seqik_leg_solve_generic_f64(8, 150, 15, 4, 8, 8, 70, 7, 0, 0, 27, 0, 0, 0, 0, 4, 10, 1, 0) == -1   # unknown flag bit
    assert lib.seqik_pchip_resample_f32(8, 8, 1, 1, 10, 1, 0.01, 0.001, 0) == -1                          # fewer than 2 samples
    assert lib.seqik_pchip_resample_f64(8, 8, 1, 5, 10, 1, 0.01, 0.0, 0) == -1                            # non-positive step
    assert lib.seqik_pchip_resample_f64(8, 8, 0, 5, 10, 1, 0.01, 0.001, 0) == 0                           # empty: ok
    with pytest.raises(ValueError):
        N.check(-1, "x")


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import numpy as np
    from seqikpy_b200 import _native as N, data
    from seqikpy_b200.kinematic_chain import KinematicChainSeq
    from seqikpy_b200.leg_inverse_kinematics import LegInvKinSeq
    from seqikpy_b200.head_inverse_kinematics import HeadInverseKinematics
    from seqikpy_b200.alignment import AlignPose
    kc = KinematicChainSeq(data.BOUNDS, ["RF", "LF"])
    with pytest.raises(N.SeqIKNativeError):
        LegInvKinSeq({"RF_leg": np.zeros((3, 5, 3))}, kc, log_level="ERROR").run_ik_and_fk()
    with pytest.raises(N.SeqIKNativeError):
        HeadInverseKinematics({"R_head": np.ones((3, 2, 3)), "L_head": np.ones((3, 2, 3)), "Neck": np.zeros((1, 1, 3))},
                              data.NMF_TEMPLATE).compute_head_angles()
    with pytest.raises(N.SeqIKNativeError):
        AlignPose({"RF_leg": np.ones((9, 5, 3))}, ["RF"], log_level="ERROR").align_pose()


def test_product_never_imports_oracle():
    pkg = ROOT / "sequential-inverse-kinematics_b200"
    for f in list(pkg.glob("*.py")) + list((ROOT / "seqikpy_b200").glob("*.py")):
        text = f.read_text()
        assert "import oracle" not in text and "from oracle" not in text and "hostsim" not in text, f
