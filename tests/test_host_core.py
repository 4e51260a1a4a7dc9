"""The solver core (csrc/seqik_core.cuh, the code the kernels run per lane) compiled for the host by g++ and
checked against the reference's shipped outputs and the oracle fixtures -- the algorithm's parity without a GPU.
The GPU tests (test_gpu_parity.py) repeat these comparisons through the CUDA library."""
import numpy as np
import pytest

import hostsim_build as H
import model_trf2 as M
from helpers import ANGLE_TOL, FK_TOL, bad_frames, fk_residual, residual_of_angles, singular_windows
from oracle import seqik_oracle as O

GN_REF = 0b0111111  # SEQIK_FLAG_REFERENCE_ITERATES: Gauss-Newton mode in all four stages + singularity escape + skip-confirm
GN_NEWTON = 0b1111111  # the above + Newton steps
GN = 0b11111111     # SEQIK_FLAG_DEFAULT: the above + closed-form warm step (acts on carried solves only)


def leg_consts(size, bounds, init, leg):
    seg = [size[f"{leg}_{s}"] for s in O.SEGMENTS]
    lb = [bounds[f"{leg}_{d}"][0] for d in O.DOF_ORDER]
    ub = [bounds[f"{leg}_{d}"][1] for d in O.DOF_ORDER]
    return seg, lb, ub, M.null_sq_from_seeds(init[leg]), M.seeds7(init[leg])


@pytest.mark.parametrize("flags", [GN, GN_NEWTON, GN_REF])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_grooming_rf_all_frames(grooming_leg, dtype, flags):
    from seqikpy_b200 import data as D
    size = O.calculate_body_size(D.NMF_TEMPLATE, ["RF", "LF"])
    seg, lb, ub, nsq, seed = leg_consts(size, D.BOUNDS, D.INITIAL_ANGLES, "RF")
    ang, fk, nfev, status = H.solve_chain(grooming_leg["pose"][0], seg, lb, ub, nsq, seed, dtype=dtype, gn_mask=flags)
    assert len(bad_frames(ang, grooming_leg["ref_angles"][0])) == 0
    assert len(bad_frames(ang, grooming_leg["oracle_angles"][0])) == 0
    r_ours = fk_residual(fk, grooming_leg["pose"][0])
    r_ref = residual_of_angles(grooming_leg["ref_angles"][0], seg, grooming_leg["pose"][0])
    assert ((r_ours - r_ref) > FK_TOL + 2e-6).sum() == 0
    assert (status > 0).all() and nfev.mean() < 6


@pytest.mark.parametrize("flags", [GN, GN_NEWTON, GN_REF])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_grooming_lf(grooming_leg, dtype, flags):
    from seqikpy_b200 import data as D
    size = O.calculate_body_size(D.NMF_TEMPLATE, ["RF", "LF"])
    seg, lb, ub, nsq, seed = leg_consts(size, D.BOUNDS, D.INITIAL_ANGLES, "LF")
    ang, fk, _, _ = H.solve_chain(grooming_leg["pose"][1], seg, lb, ub, nsq, seed, dtype=dtype, gn_mask=flags)
    bad = bad_frames(ang, grooming_leg["ref_angles"][1])
    # mismatches only around the reference's own singular episodes (helpers.singular_windows)
    allowed = singular_windows(grooming_leg["ref_angles"][1])
    assert len(bad) <= 30 and set(bad) <= allowed, bad
    r_ours = fk_residual(fk, grooming_leg["pose"][1])
    r_ref = residual_of_angles(grooming_leg["ref_angles"][1], seg, grooming_leg["pose"][1])
    worse = np.where(((r_ours - r_ref) > FK_TOL + 2e-6).any(axis=1))[0]
    assert set(worse) <= allowed, worse


@pytest.mark.parametrize("flags", [GN, GN_NEWTON, GN_REF])
def test_locomotion_and_synthetic_vs_oracle(locomotion, synthetic_gold, flags):
    from seqikpy_b200 import data as D, synthetic as S
    legs = list(locomotion["legs"])
    size = O.calculate_body_size(D.TEMPLATE_NMF_LOCOMOTION, legs)
    for i, leg in enumerate(legs):
        seg, lb, ub, nsq, seed = leg_consts(size, D.BOUNDS_LOCOMOTION, D.INITIAL_ANGLES_LOCOMOTION, leg)
        ang, fk, _, _ = H.solve_chain(locomotion["aligned"][i], seg, lb, ub, nsq, seed, gn_mask=flags)
        assert np.abs(ang - locomotion["oracle_angles"][i]).max() < ANGLE_TOL, leg
        r = fk_residual(fk, locomotion["aligned"][i]) - fk_residual(locomotion["oracle_fk"][i], locomotion["aligned"][i])
        assert r.max() < FK_TOL + 2e-6
    size, bounds, init = S.chain_constants()
    for tr in range(2):
        for li, leg in enumerate(S.LEGS):
            seg, lb, ub, nsq, seed = leg_consts(size, bounds, init, leg)
            ang, fk, _, _ = H.solve_chain(synthetic_gold["pose"][tr][:, li], seg, lb, ub, nsq, seed, gn_mask=flags)
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


def test_newton_steps_save_evaluations_and_keep_the_minimiser(synthetic_gold):
    """SEQIK_FLAG_NEWTON: fewer evaluations per solve, the same minimiser (both runs within float32 noise of each other
    relative to the tolerance, and the forward-kinematics residual not worse)."""
    from seqikpy_b200 import synthetic as S
    size, bounds, init = S.chain_constants()
    for li, leg in enumerate(S.LEGS[:3]):
        seg, lb, ub, nsq, seed = leg_consts(size, bounds, init, leg)
        pose = synthetic_gold["pose"][0][:, li]
        a_gn, f_gn, n_gn, _ = H.solve_chain(pose, seg, lb, ub, nsq, seed, gn_mask=GN_REF)
        a_nt, f_nt, n_nt, st = H.solve_chain(pose, seg, lb, ub, nsq, seed, gn_mask=GN)
        assert (st > 0).all()
        assert n_nt.mean() < n_gn.mean() - 0.5, (n_nt.mean(0), n_gn.mean(0))
        assert np.abs(a_nt - a_gn).max() < 1e-4
        assert (fk_residual(f_nt, pose) - fk_residual(f_gn, pose)).max() < 1e-5


def test_carried_solves_equal_fresh_solves_to_rounding(synthetic_gold):
    """The stage-pipeline kernel carries a solve's sin/cos from frame to frame (StageSolve::restart) and re-derives them
    every SEQIK_RESYNC frames; its per-lane arithmetic, run serially on the host build, stays within float32 rounding of
    the frame-by-frame composition and of the oracle."""
    from seqikpy_b200 import synthetic as S
    from seqikpy_b200.kinematic_chain import KinematicChainSeq
    size, bounds, init = S.chain_constants()
    chain = KinematicChainSeq(bounds, list(S.LEGS), size)
    for flags in (GN_NEWTON, GN_REF):
        for li, leg in enumerate(S.LEGS[:2]):
            seg, lb, ub, nsq, seed = leg_consts(size, bounds, init, leg)
            pose = synthetic_gold["pose"][0][:, li]
            a1, f1, n1, _ = H.solve_chain(pose, seg, lb, ub, nsq, seed, gn_mask=flags)
            a2, f2, n2 = H.run_carried_f32(pose, chain.pack_chain_params(leg, init[leg]), flags)
            assert np.abs(a1 - a2).max() < 2e-5 and np.abs(f1 - f2).max() < 2e-5
            assert abs(n1.mean() - n2.mean()) < 0.1
            assert np.abs(a2 - synthetic_gold["oracle_angles"][0, li]).max() < ANGLE_TOL


@pytest.mark.parametrize("flags", [GN, GN_NEWTON, GN_REF])
def test_carried_solves_on_the_grooming_trial(grooming_leg, flags):
    """The pipeline kernel's arithmetic (carried solves; with SEQIK_FLAG_CLOSED_FORM the closed-form warm step) over all
    6000 frames of both legs against the reference's shipped angles: same acceptance as the frame-by-frame solves."""
    from seqikpy_b200 import data as D
    from seqikpy_b200.kinematic_chain import KinematicChainSeq
    size = O.calculate_body_size(D.NMF_TEMPLATE, ["RF", "LF"])
    chain = KinematicChainSeq(D.BOUNDS, ["RF", "LF"], size)
    for li, leg in enumerate(["RF", "LF"]):
        seg = [size[f"{leg}_{s}"] for s in O.SEGMENTS]
        pose = grooming_leg["pose"][li]
        ang, fk, nfev = H.run_carried_f32(pose, chain.pack_chain_params(leg, D.INITIAL_ANGLES[leg]), flags)
        allowed = singular_windows(grooming_leg["ref_angles"][li])
        bad = bad_frames(ang, grooming_leg["ref_angles"][li])
        assert set(bad) <= allowed and len(bad) <= 30, (leg, bad)
        r_ours = fk_residual(fk, pose)
        r_ref = residual_of_angles(grooming_leg["ref_angles"][li], seg, pose)
        worse = np.where(((r_ours - r_ref) > FK_TOL + 2e-6).any(axis=1))[0]
        assert set(worse) <= allowed, (leg, worse)
        # the FK rows are those of the returned angles
        assert np.abs(r_ours - residual_of_angles(ang.astype(np.float64), seg, pose)).max() < 1e-5
        if flags == GN:
            assert nfev.mean() < 1.5        # most solves end with their first evaluation


def test_closed_form_warm_step_on_synthetic_trials(synthetic_wide):
    """Default flags on the carried path, trials 2-4 x 6 legs x 1000 frames against the oracle."""
    from seqikpy_b200 import synthetic as S
    from seqikpy_b200.kinematic_chain import KinematicChainSeq
    size, bounds, init = S.chain_constants()
    chain = KinematicChainSeq(bounds, list(S.LEGS), size)
    n_frame = int(synthetic_wide["n_frame"])
    for ti, tr in enumerate(synthetic_wide["trials"][:3]):
        pose = S.make_trial(int(tr), n_frame)
        for li, leg in enumerate(S.LEGS):
            ang, fk, nfev = H.run_carried_f32(pose[:, li], chain.pack_chain_params(leg, init[leg]), GN)
            assert np.abs(ang - synthetic_wide["oracle_angles"][ti, li]).max() < ANGLE_TOL, (tr, leg)
            assert (fk_residual(fk, pose[:, li]) - synthetic_wide["oracle_fk_residual"][ti, li]).max() < FK_TOL + 2e-6
            assert nfev.mean() < 1.6


def test_refactored_pieces_equal_their_originals():
    """rotate_frame_sel == rotate_frame and restart() == restart_a + warm_step + restart_b, bit for bit (host build)."""
    import ctypes
    lib = H.load()
    rnd = np.random.default_rng(5).uniform(-1, 1, 16 * 400).astype(np.float32)
    lib.hostsim_selfcheck.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.hostsim_selfcheck.restype = ctypes.c_int
    assert lib.hostsim_selfcheck(rnd.ctypes.data, rnd.size) == 0


def test_closed_form_warm_step_is_the_box_constrained_minimiser():
    """warm_step() against scipy's TRF on the same two-variable problem min |w(a, b) - q|^2, w = -L (sb ca, sb sa, cb),
    lb <= (a, b) <= ub, started at the same point: wherever the closed form is admitted -- interior or with `a` on a
    limit -- it lands on the minimiser scipy converges to (float64 build, 1e-6 rad / 1e-10 in cost)."""
    import ctypes
    from scipy.optimize import least_squares
    lib = H.load()
    fn = lib.hostsim_warm_step_f64
    fn.argtypes = [ctypes.c_double, ctypes.c_double] + [ctypes.c_void_p] * 4
    fn.restype = None
    rng = np.random.default_rng(11)
    L = 0.54
    lbub = np.array([-0.6, 0.9, -2.4, -0.2])

    def w(x):
        a, b = x
        return -L * np.array([np.sin(b) * np.cos(a), np.sin(b) * np.sin(a), np.cos(b)])
    n_interior = n_limit = n_rejected = 0
    for _ in range(400):
        x_prev = np.array([rng.uniform(-0.55, 0.85), rng.uniform(-2.2, -0.4)])
        # the next frame's target: the previous end point moved and scaled a little (noise leaves it off the sphere);
        # every fourth case pushes `a` towards / beyond one of its limits
        x_new = x_prev + rng.normal(0, 0.12, 2)
        if rng.uniform() < 0.25:
            x_new[0] = rng.choice([lbub[0] - rng.uniform(0, 0.2), lbub[1] + rng.uniform(0, 0.2)])
            x_prev[0] = np.clip(x_new[0], lbub[0] + 0.05, lbub[1] - 0.05)
        q = w(x_new) * rng.uniform(0.8, 1.25) + rng.normal(0, 0.01, 3)
        out = np.zeros(5)
        fn(L, 1.0, x_prev.ctypes.data, q.ctypes.data, lbub.ctypes.data, out.ctypes.data)
        if out[2] == 0:
            n_rejected += 1
            continue
        ref = least_squares(lambda x: w(x) - q, x_prev, bounds=(lbub[[0, 2]], lbub[[1, 3]]), xtol=1e-14, ftol=1e-14, gtol=1e-14)
        assert abs(out[4] - ref.cost) < 1e-10, (out, ref.x, ref.cost)
        assert np.abs(out[:2] - ref.x).max() < 1e-6, (out, ref.x)
        if out[3] == 0:
            n_interior += 1
        else:
            n_limit += 1
            assert out[0] in (lbub[0], lbub[1])
    assert n_interior > 150 and n_limit > 20 and n_rejected < 200, (n_interior, n_limit, n_rejected)


def test_closed_form_warm_step_one_variable_stage():
    """The same for the one-variable stage (TiTa pitch only, has_a = 0): the minimiser in the plane of its rotation."""
    import ctypes
    from scipy.optimize import least_squares
    lib = H.load()
    fn = lib.hostsim_warm_step_f64
    fn.argtypes = [ctypes.c_double, ctypes.c_double] + [ctypes.c_void_p] * 4
    fn.restype = None
    rng = np.random.default_rng(12)
    L = 0.63
    lbub = np.array([-np.inf, np.inf, -2.6, -0.05])
    w = lambda b: -L * np.array([np.sin(b), 0.0, np.cos(b)])
    n_seeded = 0
    for _ in range(300):
        b_prev = rng.uniform(-2.4, -0.3)
        q = w(b_prev + rng.normal(0, 0.15)) * rng.uniform(0.7, 1.3) + rng.normal(0, 0.02, 3)
        ab = np.array([0.0, b_prev])
        out = np.zeros(5)
        fn(L, 0.0, ab.ctypes.data, q.ctypes.data, lbub.ctypes.data, out.ctypes.data)
        if out[2] == 0:
            continue
        n_seeded += 1
        ref = least_squares(lambda x: w(x[0]) - q, [b_prev], bounds=([lbub[2]], [lbub[3]]), xtol=1e-14, ftol=1e-14, gtol=1e-14)
        assert abs(out[1] - ref.x[0]) < 1e-6 and abs(out[4] - ref.cost) < 1e-10, (out, ref.x, ref.cost)
        assert out[0] == 0.0 and out[3] == 0
    assert n_seeded > 200


def test_newton_step_is_the_newton_step():
    """SEQIK_FLAG_NEWTON: the first step of a solve equals -(H + C)^-1 g with H the exact Hessian of 0.5 |w(a, b) - q|^2
    (finite differences of the analytic gradient) and C = diag(|g_i| / v_i) the Coleman-Li term of scipy's model; with
    the flag off it equals the Gauss-Newton step -(J^T J + C)^-1 g.  Newton converges to the minimiser in fewer trips."""
    import ctypes
    lib = H.load()
    fn = lib.hostsim_trips_f64
    fn.argtypes = [ctypes.c_double] + [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    fn.restype = None
    rng = np.random.default_rng(13)
    L = 0.69
    lbub = np.array([-0.9, 2.2, -3.0, -0.1])

    def w(x):
        a, b = x
        return -L * np.array([np.sin(b) * np.cos(a), np.sin(b) * np.sin(a), np.cos(b)])

    def jac(x):
        a, b = x
        return -L * np.array([[-np.sin(b) * np.sin(a), np.cos(b) * np.cos(a)],
                              [np.sin(b) * np.cos(a), np.cos(b) * np.sin(a)],
                              [0.0, -np.sin(b)]])

    def grad(x, q):
        return jac(x).T @ (w(x) - q)
    n_newton = 0
    for _ in range(200):
        x = np.array([rng.uniform(-0.5, 1.8), rng.uniform(-2.5, -0.6)])
        q = w(x + rng.normal(0, 0.04, 2)) * rng.uniform(0.85, 1.2) + rng.normal(0, 0.02, 3)
        g = grad(x, q)
        v = np.where(g < 0, lbub[[1, 3]] - x, x - lbub[[0, 2]])              # distance to the bound the anti-gradient heads for
        C = np.diag(np.abs(g) / v)
        h = 1e-6
        Hx = np.column_stack([(grad(x + h * e, q) - grad(x - h * e, q)) / (2 * h) for e in np.eye(2)])
        JTJ = jac(x).T @ jac(x)
        out_n, out_g = np.zeros(5), np.zeros(5)
        fn(L, x.ctypes.data, q.ctypes.data, lbub.ctypes.data, 1 | 4, 1, out_n.ctypes.data)      # Gauss-Newton mode + Newton
        fn(L, x.ctypes.data, q.ctypes.data, lbub.ctypes.data, 1, 1, out_g.ctypes.data)          # Gauss-Newton mode
        p_g = -np.linalg.solve(JTJ + C, g)
        assert np.abs(out_g[:2] - (x + p_g)).max() < 1e-9
        p_n = -np.linalg.solve(Hx + C, g)
        pd = np.all(np.linalg.eigvalsh(Hx + C) > 0)
        if pd and np.abs(out_n[:2] - (x + p_n)).max() < 1e-7:
            n_newton += 1
        else:                                                               # not admitted: the Gauss-Newton step instead
            assert np.abs(out_n[:2] - (x + p_g)).max() < 1e-9
        # both converge to the same point; Newton needs no more evaluations
        end_n, end_g = np.zeros(5), np.zeros(5)
        fn(L, x.ctypes.data, q.ctypes.data, lbub.ctypes.data, 1 | 4, 200, end_n.ctypes.data)
        fn(L, x.ctypes.data, q.ctypes.data, lbub.ctypes.data, 1, 200, end_g.ctypes.data)
        assert np.abs(end_n[:2] - end_g[:2]).max() < 2e-5 and end_n[3] <= end_g[3]
    assert n_newton > 150, n_newton


def test_block_schedule_equals_the_serial_carried_solves(grooming_leg, locomotion):
    """The frame-parallel block schedule (csrc/seqik_block.cuh: speculate the closed-form warm step for 32 frames at once,
    accumulate the angle increments in frame order, verify warm_step's admission tests with the exact angles, replay the first
    frame that fails through the serial solver), emulated lane by lane on the host build, gives BIT-IDENTICAL angles, forward
    kinematics and evaluation counts to the serial carried solves -- on synthetic chains (limits active on the mid and hind
    legs), on both legs of the grooming trial (iterating solves, pitch angles parked on a limit, the singular episodes), on
    the locomotion recording (a fifth of its frames iterate), with ragged lengths and for a warm-started tail."""
    from seqikpy_b200 import data as D, synthetic as S
    from seqikpy_b200.kinematic_chain import KinematicChainSeq
    from seqikpy_b200.utils import calculate_body_size

    def same(pose, row, warm=None, carried=None):
        a1, f1, n1 = carried if carried is not None else H.run_carried_f32(pose, row, GN)
        a2, f2, n2, st = H.run_block_f32(pose, row, GN, warm=warm)
        assert np.array_equal(a1, a2) and np.array_equal(f1, f2) and np.array_equal(n1, n2)
        return (a1, f1, n1), st

    size, bounds, init = S.chain_constants()
    chain = KinematicChainSeq(bounds, list(S.LEGS), size)
    replays = frames = 0
    for tr in range(2):
        pose = S.make_trial(tr, 1000)
        for li, leg in enumerate(S.LEGS):
            _, st = same(pose[:, li], chain.pack_chain_params(leg, init[leg]))
            replays += int(st[2]); frames += 1000
    assert replays < 0.01 * frames                       # the schedule's premise on the benchmark workload: < 1 % of the frames replay
    for n in (1, 31, 32, 33, 100):
        same(S.make_trial(5, 128)[:n, 2], chain.pack_chain_params("RH", init["RH"]))
    chain_g = KinematicChainSeq(D.BOUNDS, ["RF", "LF"], None)
    for li, leg in enumerate(("RF", "LF")):
        row = chain_g.pack_chain_params(leg, D.INITIAL_ANGLES[leg])
        (a1, f1, n1), st = same(grooming_leg["pose"][li], row)
        assert st[2] < 0.06 * 6000
        # a frame range warm-started from the frame before it (on the 32-frame grid) continues the recording bit for bit
        a3, f3, n3, _ = H.run_block_f32(grooming_leg["pose"][li][3200:], row, GN, warm=a1[3199])
        assert np.array_equal(a3, a1[3200:]) and np.array_equal(f3, f1[3200:]) and np.array_equal(n3, n1[3200:])
    legs = list(locomotion["legs"])
    chain_l = KinematicChainSeq(D.BOUNDS_LOCOMOTION, legs, calculate_body_size(D.TEMPLATE_NMF_LOCOMOTION, legs))
    for i, leg in enumerate(legs):
        same(locomotion["aligned"][i], chain_l.pack_chain_params(leg, D.INITIAL_ANGLES_LOCOMOTION[leg]))
