"""The solver core (csrc/seqik_core.cuh, the code the kernels run per lane) compiled for the host by g++ and
checked against the reference's shipped outputs and the oracle fixtures -- the algorithm's parity without a GPU.
The GPU tests (test_gpu_parity.py) repeat these comparisons through the CUDA library."""
import numpy as np
import pytest

import hostsim_build as H
import model_trf2 as M
from helpers import ANGLE_TOL, FK_TOL, bad_frames, fk_residual, residual_of_angles, singular_windows
from oracle import seqik_oracle as O

GN = 0b111111  # SEQIK_FLAG_DEFAULT: Gauss-Newton mode in all four stages + singularity escape + skip-confirm


def leg_consts(size, bounds, init, leg):
    seg = [size[f"{leg}_{s}"] for s in O.SEGMENTS]
    lb = [bounds[f"{leg}_{d}"][0] for d in O.DOF_ORDER]
    ub = [bounds[f"{leg}_{d}"][1] for d in O.DOF_ORDER]
    return seg, lb, ub, M.null_sq_from_seeds(init[leg]), M.seeds7(init[leg])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_grooming_rf_all_frames(grooming_leg, dtype):
    from seqikpy_b200 import data as D
    size = O.calculate_body_size(D.NMF_TEMPLATE, ["RF", "LF"])
    seg, lb, ub, nsq, seed = leg_consts(size, D.BOUNDS, D.INITIAL_ANGLES, "RF")
    ang, fk, nfev, status = H.solve_chain(grooming_leg["pose"][0], seg, lb, ub, nsq, seed, dtype=dtype, gn_mask=GN)
    assert len(bad_frames(ang, grooming_leg["ref_angles"][0])) == 0
    assert len(bad_frames(ang, grooming_leg["oracle_angles"][0])) == 0
    r_ours = fk_residual(fk, grooming_leg["pose"][0])
    r_ref = residual_of_angles(grooming_leg["ref_angles"][0], seg, grooming_leg["pose"][0])
    assert ((r_ours - r_ref) > FK_TOL + 2e-6).sum() == 0
    assert (status > 0).all() and nfev.mean() < 6


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_grooming_lf(grooming_leg, dtype):
    from seqikpy_b200 import data as D
    size = O.calculate_body_size(D.NMF_TEMPLATE, ["RF", "LF"])
    seg, lb, ub, nsq, seed = leg_consts(size, D.BOUNDS, D.INITIAL_ANGLES, "LF")
    ang, fk, _, _ = H.solve_chain(grooming_leg["pose"][1], seg, lb, ub, nsq, seed, dtype=dtype, gn_mask=GN)
    bad = bad_frames(ang, grooming_leg["ref_angles"][1])
    # mismatches only around the reference's own singular episodes (helpers.singular_windows)
    allowed = singular_windows(grooming_leg["ref_angles"][1])
    assert len(bad) <= 30 and set(bad) <= allowed, bad
    r_ours = fk_residual(fk, grooming_leg["pose"][1])
    r_ref = residual_of_angles(grooming_leg["ref_angles"][1], seg, grooming_leg["pose"][1])
    worse = np.where(((r_ours - r_ref) > FK_TOL + 2e-6).any(axis=1))[0]
    assert set(worse) <= allowed, worse


def test_locomotion_and_synthetic_vs_oracle(locomotion, synthetic_gold):
    from seqikpy_b200 import data as D, synthetic as S
    legs = list(locomotion["legs"])
    size = O.calculate_body_size(D.TEMPLATE_NMF_LOCOMOTION, legs)
    for i, leg in enumerate(legs):
        seg, lb, ub, nsq, seed = leg_consts(size, D.BOUNDS_LOCOMOTION, D.INITIAL_ANGLES_LOCOMOTION, leg)
        ang, fk, _, _ = H.solve_chain(locomotion["aligned"][i], seg, lb, ub, nsq, seed, gn_mask=GN)
        assert np.abs(ang - locomotion["oracle_angles"][i]).max() < ANGLE_TOL, leg
        r = fk_residual(fk, locomotion["aligned"][i]) - fk_residual(locomotion["oracle_fk"][i], locomotion["aligned"][i])
        assert r.max() < FK_TOL + 2e-6
    size, bounds, init = S.chain_constants()
    for tr in range(2):
        for li, leg in enumerate(S.LEGS):
            seg, lb, ub, nsq, seed = leg_consts(size, bounds, init, leg)
            ang, fk, _, _ = H.solve_chain(synthetic_gold["pose"][tr][:, li], seg, lb, ub, nsq, seed, gn_mask=GN)
            assert np.abs(ang - synthetic_gold["oracle_angles"][tr, li]).max() < ANGLE_TOL, (tr, leg)


def test_long_warm_start_chain_vs_oracle(synthetic_long):
    """2000 serially warm-started frames of two legs against the oracle."""
    from seqikpy_b200 import synthetic as S
    size, bounds, init = S.chain_constants()
    pose = S.make_trial(int(synthetic_long["trial"]), int(synthetic_long["n_frame"]))
    for i, leg in enumerate(synthetic_long["legs"]):
        seg, lb, ub, nsq, seed = leg_consts(size, bounds, init, str(leg))
        ang, _, _, status = H.solve_chain(pose[:, S.LEGS.index(str(leg))], seg, lb, ub, nsq, seed, gn_mask=GN)
        assert np.abs(ang - synthetic_long["oracle_angles"][i]).max() < ANGLE_TOL
        assert (status > 0).all()


def test_synthetic_trials_2_to_7_vs_oracle(synthetic_wide):
    """Trials 2-7 x 6 legs x all 1000 frames of the synthetic workload (36 000 leg-frames) against the oracle."""
    from seqikpy_b200 import synthetic as S
    size, bounds, init = S.chain_constants()
    n_frame = int(synthetic_wide["n_frame"])
    for ti, tr in enumerate(synthetic_wide["trials"]):
        pose = S.make_trial(int(tr), n_frame)
        for li, leg in enumerate(S.LEGS):
            seg, lb, ub, nsq, seed = leg_consts(size, bounds, init, leg)
            ang, fk, _, status = H.solve_chain(pose[:, li], seg, lb, ub, nsq, seed, gn_mask=GN)
            assert np.abs(ang - synthetic_wide["oracle_angles"][ti, li]).max() < ANGLE_TOL, (tr, leg)
            assert (fk_residual(fk, pose[:, li]) - synthetic_wide["oracle_fk_residual"][ti, li]).max() < FK_TOL + 2e-6
            assert (status > 0).all()


def test_runner_state_machine_equals_serial_composition(grooming_leg):
    """ChainRunner::step() (what a lane executes) gives bit-identical results to the serial frame solve."""
    from seqikpy_b200 import data as D
    size = O.calculate_body_size(D.NMF_TEMPLATE, ["RF", "LF"])
    seg, lb, ub, nsq, seed = leg_consts(size, D.BOUNDS, D.INITIAL_ANGLES, "LF")
    pose = grooming_leg["pose"][1][:500]
    a1, f1, nfev, _ = H.solve_chain(pose, seg, lb, ub, nsq, seed, gn_mask=GN)
    a2, f2, nf_sum, steps = H.run_runner_f32(pose, seg, lb, ub, nsq, seed, gn_mask=GN)
    assert np.array_equal(a1, a2) and np.array_equal(f1, f2)
    assert np.array_equal(nf_sum, nfev.sum(0).astype(np.uint32))
    assert steps >= (nfev - 1).sum()


def test_python_model_agrees_with_host_core(grooming_leg):
    """tests/model_trf2.py (the executable specification) and the float64 host build walk the same iterates."""
    from seqikpy_b200 import data as D
    size = O.calculate_body_size(D.NMF_TEMPLATE, ["RF", "LF"])
    seg, lb, ub, nsq, seed = leg_consts(size, D.BOUNDS, D.INITIAL_ANGLES, "RF")
    pose = grooming_leg["pose"][0][:40]
    a_host, _, nfev, _ = H.solve_chain(pose, seg, lb, ub, nsq, seed, dtype=np.float64, gn_mask=0)
    stats = []
    a_model, _ = M.solve_leg(pose, seg, lb, ub, seed, nsq, stats=stats)
    assert np.abs(np.asarray(a_model) - a_host).max() < 1e-7
