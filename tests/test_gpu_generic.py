"""GPU parity tests of the generic 7-DOF leg IK (seqik_leg_solve_generic_f32/_f64 through LegInvKinGeneric and the
tensor API) against the oracle fixture tests/golden/generic_leg.npz.  See tests/test_generic_cpu.py for what parity
means for this under-determined problem: teacher-forced agreement solve by solve, claw residual / joint limits /
smoothness for a free-running recording."""
import pickle

import numpy as np
import pytest

import hostsim_build as H
from helpers import ANGLE_TOL, FK_TOL
from oracle import seqik_oracle as O
from test_generic_cpu import LEGS, consts, params_row, teacher_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    from seqikpy_b200 import _native, data, engine
    from seqikpy_b200.kinematic_chain import GENERIC_DOF_ORDER, KinematicChainGeneric
    from seqikpy_b200.leg_inverse_kinematics import LegInvKinGeneric
    _native.load_library()

    class A:
        pass
    a = A()
    a.torch, a.native, a.data, a.engine = torch, _native, data, engine
    a.Chain, a.Leg, a.ORDER = KinematicChainGeneric, LegInvKinGeneric, GENERIC_DOF_ORDER
    return a


@pytest.mark.parametrize("schedule", [2, 1])
@pytest.mark.parametrize("dtype,min_frac,med", [("float64", 0.96, 1e-7), ("float32", 0.84, 1e-5)])
def test_teacher_forced_single_solves(api, generic_gold, grooming_leg, dtype, min_frac, med, schedule):
    """Every solve of the fixture as its own one-frame chain, seeded with the oracle's previous answer; with the chain
    spread over eight lanes (schedule 2, the default) and with one lane per chain (schedule 1)."""
    torch = api.torch
    td = getattr(torch, dtype)
    n = int(generic_gold["n_frame"])
    for li, leg in enumerate(LEGS):
        _, seg, lb, ub, seed = consts(leg)
        pose = grooming_leg["pose"][li][:n]
        teacher = teacher_of(generic_gold, li, seed)
        rows = np.stack([params_row(seg, lb, ub, teacher[t]) for t in range(n)])
        d_pose = torch.tensor(pose[:, None], dtype=td, device="cuda")                 # (n chains, 1 frame, 5, 3)
        ang, fk, status, nfev = api.engine.leg_solve_generic(d_pose, torch.tensor(rows, dtype=td, device="cuda"), schedule=schedule)
        ang, fk = ang.cpu().numpy()[:, 0].astype(float), fk.cpu().numpy()[:, 0].astype(float)
        dev = np.abs(ang - generic_gold["oracle_angles"][li][:, 1:8]).max(axis=1)
        assert (dev <= ANGLE_TOL).mean() >= min_frac, (leg, (dev <= ANGLE_TOL).mean())
        assert np.median(dev) < med
        assert (status.cpu().numpy() == 1).all()
        r_ours = np.linalg.norm(fk[:, 8] - pose[:, 4], axis=1)
        r_ref = np.linalg.norm(generic_gold["oracle_fk_claw"][li] - pose[:, 4], axis=1)
        assert (r_ours - r_ref).max() < FK_TOL
        # same evaluation counts as the oracle on average (the same crawl)
        assert abs(nfev.cpu().numpy().mean() - generic_gold["oracle_stats"][li][:, 1].mean()) < 6
        # the device agrees with the host build of the same code about as well
        h_ang, _, _, _ = H.solve_generic(pose[:, [0, 4]], params_row(seg, lb, ub, seed), teacher[:, 1:8],
                                         dtype=np.float64 if dtype == "float64" else np.float32)
        assert (np.abs(h_ang - ang).max(axis=1) <= ANGLE_TOL).mean() >= min_frac


@pytest.mark.parametrize("precision", ["float64", "float32"])
def test_generic_class_free_running(api, generic_gold, grooming_leg, precision, tmp_path):
    """LegInvKinGeneric(...).run_ik_and_fk(): layout of the reference's dictionaries and pickles; every frame on the
    claw within the reference's residual + 1e-4 mm; joint limits; smoothness."""
    n = int(generic_gold["n_frame"])
    pose = {"RF_leg": grooming_leg["pose"][0][:n], "LF_leg": grooming_leg["pose"][1][:n], "R_head": np.zeros((n, 2, 3))}
    chain = api.Chain(api.data.BOUNDS, ["RF", "LF"])
    ik = api.Leg(pose, chain, api.data.INITIAL_ANGLES, log_level="ERROR", precision=precision)
    angles, fk = ik.run_ik_and_fk(export_path=tmp_path, hide_progress_bar=True)
    assert list(angles.keys()) == [f"Angle_{leg}_{d}" for leg in LEGS for d in api.ORDER]
    assert list(fk.keys()) == ["RF_leg", "LF_leg"]
    for v in angles.values():
        assert v.shape == (n,) and v.dtype == np.float64
    for li, leg in enumerate(LEGS):
        _, seg, lb, ub, _ = consts(leg)
        a7 = np.stack([angles[f"Angle_{leg}_{d}"] for d in api.ORDER], 1)
        f9 = fk[f"{leg}_leg"]
        assert f9.shape == (n, 9, 3) and f9.dtype == np.float64
        p5 = pose[f"{leg}_leg"]
        r_ours = np.linalg.norm(f9[:, 8] - p5[:, 4], axis=1)
        r_ref = np.linalg.norm(generic_gold["oracle_fk_claw"][li] - p5[:, 4], axis=1)
        assert (r_ours - r_ref).max() < FK_TOL
        assert np.all(a7 >= lb - 1e-6) and np.all(a7 <= ub + 1e-6)
        assert np.abs(O.fk_generic(a7, seg, p5[:, 0]) - f9).max() < (1e-9 if precision == "float64" else 5e-6)
        assert np.abs(f9[:, :4] - p5[:, :1]).max() < 1e-6 and np.array_equal(f9[:, 4], f9[:, 5])
        # smoothness: typical frame-to-frame motion like the oracle's (either path may jump between postures now and then)
        step_ours = np.abs(np.diff(a7, axis=0)).max(axis=1)
        step_ref = np.abs(np.diff(generic_gold["oracle_angles"][li][:, 1:8], axis=0)).max(axis=1)
        assert np.percentile(step_ours, 95) < 2 * np.percentile(step_ref, 95) and np.median(step_ours) < 2 * np.median(step_ref)
        assert step_ours.max() < np.pi
        assert ik.solver_stats[leg]["status"] == 1
        # frame 0 has no history: it must match the oracle's first solve
        assert np.abs(a7[0] - generic_gold["oracle_angles"][li][0, 1:8]).max() < (ANGLE_TOL if precision == "float64" else 5e-2)
    with open(tmp_path / "leg_joint_angles.pkl", "rb") as f:
        assert list(pickle.load(f).keys()) == list(angles.keys())
    with open(tmp_path / "forward_kinematics.pkl", "rb") as f:
        assert pickle.load(f)["LF_leg"].shape == (n, 9, 3)
    # calculate_ik_stage on one leg = the same numbers
    ik2 = api.Leg(pose, chain, api.data.INITIAL_ANGLES, log_level="ERROR", precision=precision)
    f1 = ik2.calculate_ik_stage(pose["RF_leg"][:, -1], pose["RF_leg"][:, 0], api.data.INITIAL_ANGLES["RF"]["stage_4"], "RF")
    assert np.array_equal(f1, fk["RF_leg"]) and np.array_equal(ik2.joint_angles_dict["Angle_RF_ThC_roll"], angles["Angle_RF_ThC_roll"])


def test_generic_schedule_invariance_and_errors(api, grooming_leg):
    """Chains per warp is scheduling only (bitwise identical results); bad arguments raise."""
    torch = api.torch
    _, seg, lb, ub, seed = consts("RF")
    rows = np.tile(params_row(seg, lb, ub, seed), (40, 1))
    pose = np.stack([grooming_leg["pose"][0][50 * k:50 * k + 60] for k in range(40)])      # 40 chains x 60 frames
    d_pose = torch.tensor(pose, dtype=torch.float32, device="cuda")
    d_rows = torch.tensor(rows, dtype=torch.float32, device="cuda")
    for schedule in (1, 2):
        ref = api.engine.leg_solve_generic(d_pose, d_rows, schedule=schedule)
        for cpw in (1, 3, 7, 32):
            out = api.engine.leg_solve_generic(d_pose, d_rows, chains_per_warp=cpw, schedule=schedule)
            assert torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1]) and torch.equal(out[3], ref[3])
    # the two mappings run the same iteration with different summation orders: same claw, same limits, nearly always the
    # same angles (the problem is under-determined: a few solves may settle on another point of the self-motion manifold)
    one = api.engine.leg_solve_generic(d_pose, d_rows, schedule=1)
    resid = lambda o: (o[1][:, :, 8] - d_pose[:, :, 4]).norm(dim=-1).max().item()
    assert resid(ref) < 1e-4 and resid(one) < 1e-4
    assert abs(float(ref[3].double().mean()) - float(one[3].double().mean())) < 0.1 * float(one[3].double().mean())
    # warm start replaces the seeds: continuing from frame 29 reproduces frames 30..59
    tail = api.engine.leg_solve_generic(d_pose[:, 30:].contiguous(), d_rows, warm=ref[0][:, 29].contiguous())
    assert torch.equal(tail[0], ref[0][:, 30:])
    # target_row: the claw may sit in any row
    alt = api.engine.leg_solve_generic(d_pose[:, :, [0, 4, 1]].contiguous(), d_rows, target_row=1)
    assert torch.equal(alt[0], ref[0])
    with pytest.raises(ValueError):
        api.engine.leg_solve_generic(d_pose, d_rows, target_row=0)
    with pytest.raises(ValueError):
        api.engine.leg_solve_generic(d_pose, d_rows[:3])
    # non-finite key points: the solve is skipped (status -1), like scipy's ValueError in the reference
    bad = d_pose.clone()
    bad[3, 10, 4, 0] = float("nan")
    st = api.engine.leg_solve_generic(bad, d_rows)[2].cpu().numpy()
    assert st[3] == -1 and (np.delete(st, 3) == 1).all()
    ik = api.Leg({"RF_leg": bad[3].cpu().numpy().astype(float)}, api.Chain(api.data.BOUNDS, ["RF"]), api.data.INITIAL_ANGLES,
                 log_level="ERROR")
    with pytest.raises(ValueError, match="not finite"):
        ik.run_ik_and_fk()
