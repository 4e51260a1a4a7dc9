"""Schedule 3 (frame-parallel blocks: a warp per chain, 32 frames per pass, csrc/seqik_block.cuh) against schedule 2 (the
stage pipeline): the two run the same per-frame arithmetic, so every output must agree BIT FOR BIT -- on the synthetic
workload, on the bundled grooming trial (iterating solves, pitch angles parked on their limits, the singular episodes), with
ragged frame counts, warm-started frame ranges, the alignment map applied on load, every FK layout and NaN key points."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    from seqikpy_b200 import _native as N, data as D, engine, synthetic as S
    from seqikpy_b200.batch import chain_param_table
    from seqikpy_b200.kinematic_chain import KinematicChainSeq
    N.load_library()
    size, bounds, init = S.chain_constants()
    chain = KinematicChainSeq(bounds, list(S.LEGS), size)

    class E:
        pass
    e = E()
    e.torch, e.N, e.D, e.engine, e.S = torch, N, D, engine, S
    e.pose = S.to_chains(torch.from_numpy(S.make_trials(range(8), 1000)).cuda())            # 48 chains x 1000 frames
    e.params = torch.from_numpy(chain_param_table(chain, init, S.LEGS, 8)).cuda()
    return e


def both(env, pose, params, **kw):
    """Schedule 2, then schedule 3 with the lean and with the robust kernel (which must agree with each other bit for bit:
    asserted here); returns (schedule 2, schedule 3)."""
    out = []
    for sched, variant in ((2, 0), (3, 1), (3, 2)):
        ang, fk, st, nf = env.engine.leg_solve(pose, params, schedule=sched, block_variant=variant, **kw)
        env.torch.cuda.synchronize()
        out.append((ang.cpu().numpy(), None if fk is None else fk.cpu().numpy(), st.cpu().numpy(), nf.cpu().numpy()))
    for x, y in zip(out[1], out[2]):
        assert (x is None and y is None) or np.array_equal(x, y, equal_nan=True)
    return out[:2]


def assert_same(a, b):
    for x, y, name in zip(a, b, ("angles", "fk", "status", "nfev")):
        if x is None:
            assert y is None
            continue
        assert x.shape == y.shape, name
        bad = np.argwhere(x != y)
        assert len(bad) == 0, (name, len(bad), bad[:5])


def test_block_schedule_is_the_default(env):
    """Automatic schedule = 3 for the default flags and all stages; 2 where schedule 3 does not apply."""
    pose = env.pose[:6, :96].contiguous()
    a0, f0, _, _ = env.engine.leg_solve(pose, env.params[:6])
    a3, f3, _, _ = env.engine.leg_solve(pose, env.params[:6], schedule=3)
    assert env.torch.equal(a0, a3) and env.torch.equal(f0, f3)
    with pytest.raises(Exception):
        env.engine.leg_solve(pose, env.params[:6], schedule=3, flags=env.N.FLAG_REFERENCE_ITERATES)


@pytest.mark.parametrize("n_frame", [1000, 999, 33, 32, 31, 5, 4, 1])
def test_synthetic_bitwise(env, n_frame):
    pose = env.pose[:, :n_frame].contiguous()
    a, b = both(env, pose, env.params)
    assert_same(a, b)
    assert (a[2] == 1).all()


@pytest.mark.parametrize("layout", ["full", "joints", None])
def test_fk_layouts_bitwise(env, layout):
    kw = dict(want_fk=layout is not None)
    if layout:
        kw["fk_layout"] = layout
    a, b = both(env, env.pose[:12, :200].contiguous(), env.params[:12], **kw)
    assert_same(a, b)


def test_resident_warps_do_not_change_results(env):
    ref = both(env, env.pose[:, :256].contiguous(), env.params)[1]
    for r in (1, 5, 13):
        ang, fk, st, nf = env.engine.leg_solve(env.pose[:, :256].contiguous(), env.params, schedule=3, chains_per_warp=r)
        assert np.array_equal(ang.cpu().numpy(), ref[0]) and np.array_equal(fk.cpu().numpy(), ref[1])


def test_warm_started_frame_ranges_bitwise(env):
    """Frame ranges solved in place, each warm-started from the frame before it: equal to one launch on the 32-frame grid."""
    torch = env.torch
    one = env.engine.leg_solve(env.pose, env.params, schedule=3)
    ang = torch.zeros_like(one[0]); fk = torch.zeros_like(one[1])
    for t0, t1 in ((0, 128), (128, 480), (480, 1000)):
        env.engine.leg_solve(env.pose, env.params, angles=ang, fk=fk, schedule=3, frames=(t0, t1))
    assert torch.equal(ang, one[0]) and torch.equal(fk, one[1])
    # an unaligned range (plain loads/stores instead of bulk copies; not on the 32-frame grid: equal to schedule 2 on the same range)
    for sched in (2, 3):
        a2 = one[0].clone(); f2 = one[1].clone()
        env.engine.leg_solve(env.pose, env.params, angles=a2, fk=f2, schedule=sched, frames=(333, 777))
        if sched == 2:
            ra, rf = a2, f2
    assert torch.equal(a2, ra) and torch.equal(f2, rf)


def test_alignment_map_on_load_bitwise(env):
    torch = env.torch
    n = 12
    aff = torch.tensor([[0.1, -0.2, 0.3, 1.05, 0.0, 0.0, 0.0, 0.0]], device="cuda").repeat(n, 1)
    aff[:, 4:7] = env.pose[:n, 0, 0]                                      # template coxa = the pose's own (constant) origin
    raw = (env.pose[:n, :300] - env.pose[:n, :1, :1]) / 1.05 + env.pose[:n, :1, :1] + 0.0
    a, b = both(env, raw.contiguous(), env.params[:n], affine=aff)
    assert_same(a, b)


def test_grooming_trial_bitwise(env, grooming_leg):
    """6000 frames x 2 legs of real data: ~3-4 % of the frames replay through the serial solver (iterating solves, pitch
    angles parked on a limit, the CTr_pitch = 0 episodes)."""
    from seqikpy_b200.kinematic_chain import KinematicChainSeq
    torch = env.torch
    chain = KinematicChainSeq(env.D.BOUNDS, ["RF", "LF"], None)
    params = torch.from_numpy(np.stack([chain.pack_chain_params(leg, env.D.INITIAL_ANGLES[leg]) for leg in ("RF", "LF")]).astype(np.float32)).cuda()
    pose = torch.from_numpy(np.ascontiguousarray(grooming_leg["pose"], dtype=np.float32)).cuda()
    a, b = both(env, pose, params)
    assert_same(a, b)
    assert (a[3] > 6000).any()                                             # some solves did iterate


def test_non_finite_key_points_bitwise(env):
    pose = env.pose[:6, :100].clone()
    pose[1, 40, 2, 1] = float("nan")
    pose[4, 0, 1, 0] = float("inf")
    a, b = both(env, pose, env.params[:6])
    for x, y in zip(a, b):
        assert np.array_equal(x, y, equal_nan=True)
    assert a[2][1] == -1 and a[2][4] == -1 and a[2][0] == 1


def test_config5_depth_windows_vs_oracle(env):
    """BASELINE config 5 is 100 000 serially warm-started frames.  Two synthetic chains are run to that depth on the GPU and the
    CPU oracle (memoryless between frames but for x0 = previous frame's angles, leg_inverse_kinematics.py:272) is run on
    three windows -- start, middle (across a 32-frame re-derivation boundary), end -- seeded with the GPU's own angles of the
    frame before each window: angles within 1e-3 rad, FK residual per joint not worse by more than 1e-4 mm."""
    from helpers import ANGLE_TOL, FK_TOL, F32_FK_NOISE, fk_residual
    from oracle import seqik_oracle as O
    from seqikpy_b200.kinematic_chain import DOF_ORDER, STAGE_ACTIVE_DOFS, STAGE_ACTIVE_SLOTS, KinematicChainSeq
    torch, S = env.torch, env.S
    n_frame = 100_000
    legs = ("RF", "RM")
    size, bounds, init = S.chain_constants()
    chain = KinematicChainSeq(bounds, list(S.LEGS), size)
    pose64 = S.make_trial(11, n_frame)                                                   # (F, 6, 5, 3) float64
    li = [S.LEGS.index(l) for l in legs]
    pose = torch.from_numpy(np.ascontiguousarray(pose64[:, li].transpose(1, 0, 2, 3), dtype=np.float32)).cuda()
    params = torch.from_numpy(np.stack([chain.pack_chain_params(l, init[l]) for l in legs]).astype(np.float32)).cuda()
    ang, fk, st, nf = env.engine.leg_solve(pose, params)
    ang, fk = ang.cpu().numpy().astype(np.float64), fk.cpu().numpy().astype(np.float64)
    assert (st.cpu().numpy() == 1).all()
    assert nf.cpu().numpy().sum() < 1.05 * 2 * 4 * n_frame                               # the depth run stays on the closed form
    for (t0, t1) in ((0, 300), (49_984, 50_300), (99_700, 100_000)):
        for ci, leg in enumerate(legs):
            seeds = {k: np.array(v, dtype=float) for k, v in init[leg].items()}
            if t0 > 0:
                for stage in (1, 2, 3, 4):
                    for slot, dof in zip(STAGE_ACTIVE_SLOTS[stage], STAGE_ACTIVE_DOFS[stage]):
                        seeds[f"stage_{stage}"][slot] = ang[ci, t0 - 1, DOF_ORDER.index(dof)]
            p = pose[ci, t0:t1].cpu().numpy().astype(np.float64)
            ref, ref_fk = O.run_ik_and_fk({f"{leg}_leg": p}, size, bounds, {leg: seeds})
            ref7 = np.stack([ref[f"Angle_{leg}_{d}"] for d in DOF_ORDER], 1)
            assert np.abs(ang[ci, t0:t1] - ref7).max() < ANGLE_TOL, (leg, t0, np.abs(ang[ci, t0:t1] - ref7).max())
            worse = fk_residual(fk[ci, t0:t1], p) - fk_residual(ref_fk[f"{leg}_leg"], p)
            assert worse.max() < FK_TOL + F32_FK_NOISE, (leg, t0, worse.max())


@pytest.mark.parametrize("flags_name", ["FLAG_DEFAULT", "FLAG_REFERENCE_ITERATES"])
def test_large_batch_on_the_automatic_schedule_is_bit_identical(env, flags_name):
    """19 200 chains (the 8 golden trials tiled 400 times) on the AUTOMATIC schedule -- schedule 3 for the default flags;
    schedule 2 with its large-batch packing and phase period (chains per warp 6, period 3 beyond ~16 000 chains) for the
    reference-iterates set -- give, chain for chain, the bits of the 48-chain run."""
    torch = env.torch
    flags = getattr(env.N, flags_name)
    n_frame, tiles = 512, 400
    pose = env.pose[:, :n_frame].contiguous()
    a0, f0, s0, n0 = env.engine.leg_solve(pose, env.params, flags=flags)
    big_pose = pose.repeat(tiles, 1, 1, 1)
    big_params = env.params.repeat(tiles, 1)
    a1, f1, s1, n1 = env.engine.leg_solve(big_pose, big_params, flags=flags)
    torch.cuda.synchronize()
    assert torch.equal(a1.view(tiles, 48, n_frame, 7), a0.expand(tiles, -1, -1, -1))
    assert torch.equal(f1.view(tiles, 48, n_frame, 9, 3), f0.expand(tiles, -1, -1, -1, -1))
    assert torch.equal(n1.view(tiles, 48, 4), n0.expand(tiles, -1, -1)) and torch.equal(s1.view(tiles, 48), s0.expand(tiles, -1))


def test_multi_gpu_driver_equals_one_device(env):
    """batch.MultiGpuLegIK: trial shards over every visible device, results in one pinned host tensor, bit-identical to a
    single-device session over the same trials.  With one visible device the driver still runs (one shard); with two or more
    (gpurun --gpus N) the shards really sit on different devices."""
    from seqikpy_b200.batch import BatchedLegIK, MultiGpuLegIK
    from seqikpy_b200.kinematic_chain import KinematicChainSeq
    torch, S = env.torch, env.S
    size, bounds, init = S.chain_constants()
    chain = KinematicChainSeq(bounds, list(S.LEGS), size)
    n_trial, n_frame = 8, 256
    host = env.pose.view(8, 6, 1000, 5, 3)[:, :, :n_frame].contiguous().cpu().pin_memory()
    one = BatchedLegIK(chain, init, S.LEGS, n_trial, n_frame, device="cuda:0")
    a1, f1 = one.solve_host(host)
    multi = MultiGpuLegIK(chain, init, S.LEGS, n_trial, n_frame)
    assert len(multi.devices) == torch.cuda.device_count() and sum(hi - lo for lo, hi in multi.shards) == n_trial
    a2, f2 = multi.solve_host(host)
    assert torch.equal(a1.view_as(a2), a2) and torch.equal(f1.view_as(f2), f2)
    # fewer trials than devices / uneven shards
    m3 = MultiGpuLegIK(chain, init, S.LEGS, 3, n_frame, devices=[f"cuda:{i}" for i in range(torch.cuda.device_count())] * 2)
    a3, f3 = m3.solve_host(host[:3].contiguous().pin_memory())
    assert torch.equal(a3, a2[:3]) and torch.equal(f3, f2[:3])


def test_batched_loader_feeds_the_device_paths(env):
    """loader.BatchedPoseLoader (raw anipose tables -> one pinned (n_trial, n_leg, n_frame, 5, 3) tensor) -> FusedPipeline
    (alignment statistics + align-on-load solve) against the dict API (AlignPose -> LegInvKinSeq) on the same recordings."""
    import conftest
    from seqikpy_b200.alignment import AlignPose
    from seqikpy_b200.batch import FusedPipeline
    from seqikpy_b200.kinematic_chain import DOF_ORDER, KinematicChainSeq
    from seqikpy_b200.leg_inverse_kinematics import LegInvKinSeq
    from seqikpy_b200.loader import BatchedPoseLoader, to_pose_dicts
    from seqikpy_b200.utils import calculate_body_size
    torch, D = env.torch, env.D
    ga = dict(np.load(conftest.GOLDEN / "grooming_align.npz"))
    legs = ["RF", "LF"]
    pts = {"RF_leg": [f"rf{i}" for i in range(5)], "LF_leg": [f"lf{i}" for i in range(5)]}

    def table(lo, hi):                                   # an anipose-style table cut from the bundled raw key points
        t = {}
        for seg, arr in (("RF_leg", ga["raw_full_RF"]), ("LF_leg", ga["raw_full_LF"])):
            for i, kp in enumerate(pts[seg]):
                for a, ax in enumerate("xyz"):
                    t[f"{kp}_{ax}"] = arr[lo:hi, i, a]
        return t
    batch = BatchedPoseLoader(legs, fmt="anipose", pts2align=pts).load([table(0, 640), table(2000, 2640), table(4000, 4640)])
    assert batch["legs"].is_pinned() and tuple(batch["legs"].shape) == (3, 2, 640, 5, 3)
    size = calculate_body_size(D.NMF_TEMPLATE, legs)
    chain = KinematicChainSeq(D.BOUNDS, legs, size)
    pipe = FusedPipeline(chain, D.INITIAL_ANGLES, legs, D.NMF_TEMPLATE, size, 3, 640, with_head=False)
    out = pipe.run(batch["legs"].cuda(non_blocking=True))
    torch.cuda.synchronize()
    for tr, raw in enumerate(to_pose_dicts(batch, legs)):
        al = AlignPose(raw, legs_list=legs, include_claw=False, body_template=D.NMF_TEMPLATE, log_level="ERROR").align_pose()
        ang, fk = LegInvKinSeq(al, chain, D.INITIAL_ANGLES, log_level="ERROR").run_ik_and_fk()
        for li, leg in enumerate(legs):
            ref = np.stack([ang[f"Angle_{leg}_{d}"] for d in DOF_ORDER], 1)
            ours = out["angles"][tr, li].cpu().numpy()
            close = np.abs(ours - ref).max(axis=1) < 2e-4                    # float32 pose rounding; singular episodes may differ
            assert close.mean() > 0.99, (tr, leg, close.mean())


def test_joints_wire_format_returns_the_reference_layout(env):
    """BatchedLegIK(wire="joints"): only the four joint rows cross the host link; the 9-row FK the caller receives (rows 0-3
    rebuilt from the host pose, row 4 from row 5, seqik_fk_expand_host_f32) is bit-identical to the plain call's."""
    from seqikpy_b200.batch import BatchedLegIK
    from seqikpy_b200.kinematic_chain import KinematicChainSeq
    torch, S = env.torch, env.S
    size, bounds, init = S.chain_constants()
    chain = KinematicChainSeq(bounds, list(S.LEGS), size)
    n_trial, n_frame = 8, 300
    host = env.pose.view(8, 6, 1000, 5, 3)[:, :, :n_frame].contiguous().cpu().pin_memory()
    a1, f1 = BatchedLegIK(chain, init, S.LEGS, n_trial, n_frame).solve_host(host)
    for threads in (1, 5):
        w = BatchedLegIK(chain, init, S.LEGS, n_trial, n_frame, wire="joints", expand_threads=threads)
        a2, f2 = w.solve_host(host, n_chunks=4)
        assert tuple(f2.shape) == (48, n_frame, 9, 3)
        assert torch.equal(a1, a2) and torch.equal(f1, f2)
        assert abs(w.mean_fk_error() - 0.0264) < 2e-3
