"""The CPU oracle (restated ikpy glue + installed scipy TRF) against the reference's shipped outputs.

This is what pins the oracle: the reference's own tests hold no numeric check for the leg path
(tests/test_kin_chain.py checks link names only), its shipped pickles do.
"""
import numpy as np
import pytest

from helpers import ANGLE_TOL, bad_frames, residual_of_angles
from oracle import seqik_oracle as O


@pytest.fixture(scope="module")
def consts():
    from seqikpy_b200 import data as D
    size = O.calculate_body_size(D.NMF_TEMPLATE, ["RF", "LF"])
    return D, size


def test_body_size_known_answers(consts):
    D, size = consts
    assert np.isclose(size["RF_Coxa"], 0.4) and np.isclose(size["RF_Femur"], 0.69)
    assert np.isclose(size["RF_Tibia"], 0.54) and np.isclose(size["RF_Tarsus"], 0.63) and np.isclose(size["RF"], 2.26)
    assert np.isclose(size["Antenna"], 0.2745906043549196) and np.isclose(size["Antenna_mid_thorax"], 0.9355746896961248)


def test_oracle_live_run_matches_shipped_angles(consts, grooming_leg):
    """Run the oracle here on the first frames of both legs: <= 2e-5 rad and <= 1e-5 mm from the shipped pickles."""
    D, size = consts
    n = 60
    pose = {"RF_leg": grooming_leg["pose"][0][:n], "LF_leg": grooming_leg["pose"][1][:n]}
    ang, fk = O.run_ik_and_fk(pose, size, D.BOUNDS, D.INITIAL_ANGLES)
    assert list(ang.keys()) == list(grooming_leg["angle_keys"])
    for li, leg in enumerate(("RF", "LF")):
        got = np.stack([ang[f"Angle_{leg}_{d}"] for d in O.DOF_ORDER], 1)
        assert np.abs(got - grooming_leg["ref_angles"][li][:n]).max() < 2e-5
        assert np.abs(got - grooming_leg["oracle_angles"][li][:n]).max() < 1e-9        # the stored oracle run is this oracle
        assert np.abs(fk[f"{leg}_leg"] - grooming_leg["ref_fk"][li][:n]).max() < 1e-5
        assert fk[f"{leg}_leg"].shape == (n, 9, 3)


def test_stored_oracle_run_vs_shipped_angles_full_trial(grooming_leg):
    """All 6000 frames: RF everywhere; LF everywhere except the reference's irreproducible frames 286-287."""
    rf = np.abs(grooming_leg["oracle_angles"][0] - grooming_leg["ref_angles"][0])
    assert rf.max() < 1e-4
    bad = bad_frames(grooming_leg["oracle_angles"][1], grooming_leg["ref_angles"][1])
    assert len(bad) <= 2 and set(bad) <= {286, 287}, bad


def test_closed_form_fk_reproduces_shipped_fk(consts, grooming_leg):
    D, size = consts
    for li, leg in enumerate(("RF", "LF")):
        seg = [size[f"{leg}_{s}"] for s in O.SEGMENTS]
        fk = O.fk_closed_form(grooming_leg["ref_angles"][li][:600], seg, grooming_leg["pose"][li][:600, 0])
        assert np.abs(fk - grooming_leg["ref_fk"][li]).max() < 1e-12
    r = residual_of_angles(grooming_leg["ref_angles"][0], [size[f"RF_{s}"] for s in O.SEGMENTS], grooming_leg["pose"][0])
    assert abs(r.mean() - 0.0662) < 2e-4                       # the reference's own mean FK error


def test_oracle_errors(consts):
    D, size = consts
    with pytest.raises(ValueError):
        O.build_chain(1, "XX", size, D.BOUNDS)
    with pytest.raises(ValueError):
        O.build_chain(5, "RF", size, D.BOUNDS)
    with pytest.raises(ValueError):
        O.run_ik_and_fk({"RF_leg": np.zeros((2, 5, 3))}, size, D.BOUNDS, D.INITIAL_ANGLES, stages=(1, 3))
    bad = {"RF": {k: np.array(v, dtype=float) for k, v in D.INITIAL_ANGLES["RF"].items()}}
    bad["RF"]["stage_1"][3] = 1.0                              # inert slot CTr_pitch above its upper bound 0
    with pytest.raises(ValueError, match="outside of provided bounds"):
        O.run_ik_and_fk({"RF_leg": np.ones((2, 5, 3))}, size, D.BOUNDS, bad)


def test_chain_link_names_match_reference_test(consts):
    """Link-name sets of reference tests/test_kin_chain.py:39-77."""
    D, size = consts
    ang = {f"Angle_RF_{d}": np.zeros(3) for d in O.DOF_ORDER}
    names = lambda st: {l.name for l in O.build_chain(st, "RF", size, D.BOUNDS, ang, 0)}
    assert names(1) == {"Base link", "RF_ThC_yaw", "RF_ThC_pitch", "RF_CTr_pitch"}
    assert names(2) == names(1) | {"RF_ThC_roll", "RF_FTi_pitch"}
    assert names(3) == names(2) | {"RF_CTr_roll", "RF_TiTa_pitch"}
    assert names(4) == names(3) | {"RF_Claw"}
