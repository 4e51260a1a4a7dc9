"""Secondary, driver-timed records of bench.py (rank 0 only): BASELINE config 5 through the fused pipeline, the HBM-bound
stream kernels against the measured copy bandwidth, and BASELINE config 2 through the reference-compatible dict API
(with the parity counts of that very run against the reference's shipped angles).  Every function returns a dict and is
called under try/except by bench.py: a failure here never costs the headline.  Nothing under oracle/ is used."""
import json
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def _events(torch):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _timeit(torch, fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = _events(torch)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def config5_fused(n_trial=100, n_frame=100_000, reps=3):
    """100 trials x 100 000 frames x 6 legs: alignment statistics + align-on-load leg IK/FK + head alignment + head angles in
    one device-resident pass (batch.FusedPipeline).  600 chains cannot fill a GPU: this is a latency figure."""
    import torch
    from seqikpy_b200 import data as D, synthetic as S
    from seqikpy_b200.batch import FusedPipeline
    from seqikpy_b200.kinematic_chain import KinematicChainSeq
    from seqikpy_b200.utils import calculate_body_size
    n_unique = min(n_trial, 4)
    _, bounds, init = S.chain_constants()
    tmpl = dict(D.TEMPLATE_NMF_LOCOMOTION)
    for k in ("R_Antenna_base", "L_Antenna_base", "R_Antenna_edge", "L_Antenna_edge", "Neck", "Thorax_mid", "R_wing", "L_wing"):
        tmpl[k] = D.NMF_TEMPLATE[k]
    size = calculate_body_size(tmpl, list(S.LEGS))
    chain = KinematicChainSeq(bounds, list(S.LEGS), size)
    t0 = time.perf_counter()
    legs = np.stack([S.to_raw(S.make_trial(tr, n_frame).astype(np.float32)).transpose(1, 0, 2, 3) for tr in range(n_unique)])
    heads = [S.make_head_trial(tr, n_frame, dtype=np.float32) for tr in range(n_unique)]
    rep = (n_trial + n_unique - 1) // n_unique

    def dev(a):
        return torch.from_numpy(np.ascontiguousarray(a)).cuda().repeat(rep, *([1] * (a.ndim - 1)))[:n_trial].contiguous()
    d_legs = dev(legs)
    d_r, d_l, d_th = (dev(S.to_raw(np.stack([h[i] for h in heads]))) for i in range(3))
    gen_s = time.perf_counter() - t0
    pipe = FusedPipeline(chain, init, S.LEGS, tmpl, size, n_trial, n_frame)
    out = {}

    def step():
        out.update(pipe.run(d_legs, d_r, d_l, d_th))
    ms = _timeit(torch, step, reps)
    sess = pipe.session
    nfev, status_ok = sess.nfev.double().sum(0), int((sess.status == 1).sum())
    ms_solver = _timeit(torch, lambda: sess.solve_device(d_legs, affine=out["leg_affine"].view(-1, 8), want_stats=False), reps, warm=1)
    lf = n_trial * 6 * n_frame
    res = {"workload": f"{n_trial} trials x {n_frame} frames x 6 legs (+ head), alignment + leg IK/FK + head IK fused on the device "
                       f"({n_unique} unique trials tiled)", "ms_per_pass": ms, "ms_solver_only": ms_solver, "leg_frames_per_s": lf / ms * 1e3,
           "chains": n_trial * 6, "us_per_frame_per_chain": ms * 1e3 / n_frame, "head_frames_per_s": n_trial * n_frame / ms * 1e3,
           "nfev_per_leg_frame_by_stage": (nfev / lf).tolist(), "chains_status_ok": status_ok,
           "mean_alignment_scale": float(out["leg_affine"][..., 3].mean()), "data_gen_s": gen_s, "timing": "CUDA events, inputs resident"}
    del pipe, d_legs, d_r, d_l, d_th, out
    torch.cuda.empty_cache()
    return res


def stream_kernels(hbm_peak, n_trial=1000, n_frame=1000):
    """FK, align-apply, alignment statistics, head angles, head-apply: achieved GB/s of algorithmic bytes / measured copy peak."""
    import torch
    from seqikpy_b200 import engine, synthetic as S
    from seqikpy_b200.batch import chain_param_table
    from seqikpy_b200.kinematic_chain import KinematicChainSeq
    size, bounds, init = S.chain_constants()
    chain = KinematicChainSeq(bounds, list(S.LEGS), size)
    n_chain = n_trial * 6
    g = torch.Generator(device="cuda").manual_seed(1)
    params = torch.from_numpy(chain_param_table(chain, init, S.LEGS, n_trial)).cuda()
    lb, ub = params[:, 4:11], params[:, 11:18]
    angles = (lb + (ub - lb).clamp(max=6.0) * torch.rand((n_chain, 7), device="cuda", generator=g))[:, None, :].expand(n_chain, n_frame, 7).contiguous()
    origin = torch.randn((n_chain, n_frame, 3), device="cuda", generator=g)
    pose = torch.randn((n_chain, n_frame, 5, 3), device="cuda", generator=g)
    aff = torch.rand((n_chain, 8), device="cuda", generator=g) + 0.5
    consts = torch.rand((n_chain, 4), device="cuda", generator=g) + 1.0
    lf = n_chain * n_frame
    nf = n_trial * n_frame * 6
    r = torch.randn((n_trial, 6 * n_frame, 2, 3), device="cuda", generator=g)
    l = torch.randn((n_trial, 6 * n_frame, 2, 3), device="cuda", generator=g)
    neck = torch.randn((n_trial, 3), device="cuda", generator=g)
    rest = torch.zeros((n_trial, 2), device="cuda")
    haff = torch.rand((n_trial, 8), device="cuda", generator=g) + 0.5
    x = torch.empty(1 << 28, device="cuda"); y = torch.empty_like(x)
    # key points as a recording has them (the synthetic workload's trials, tiled): every series lives in a narrow range
    base = S.to_chains(torch.from_numpy(S.make_trials(range(32), n_frame)).cuda())
    rec = base.repeat((n_chain + base.shape[0] - 1) // base.shape[0], 1, 1, 1)[:n_chain].contiguous()
    cases = [
        ("fk", lambda: engine.forward_kinematics(angles, origin, params), lf * (28 + 12 + 108)),
        ("align_apply", lambda: engine.align_apply(pose, aff), lf * 120),
        ("alignment_statistics (fused: key points -> affine rows)", lambda: engine.leg_affine(rec, consts), lf * 60 + n_chain * 32),
        ("alignment_statistics, normal-random key points (wide-range series: two radix passes)", lambda: engine.leg_affine(pose, consts), lf * 60 + n_chain * 32),
        ("alignment_statistics, three-kernel path (series written and re-read)", lambda: engine.leg_affine_unfused(rec, consts), lf * (60 + 28 + 28)),
        ("head_angles", lambda: engine.head_angles(r, l, neck, rest), nf * (48 + 28)),
        ("head_apply", lambda: engine.head_apply(r, haff), nf * 48),
        ("pchip_resample float32 (x10 upsampling of the angles tensor)", lambda: engine.pchip_resample(angles, 0.01, 0.001), lf * 7 * 4 * 11),
        ("torch copy_ of 1 GiB (reference point)", lambda: y.copy_(x), 2 * x.numel() * 4),
    ]
    out = {}
    for name, fn, nbytes in cases:
        ms = _timeit(torch, fn, 10, warm=3)
        out[name] = {"ms": ms, "GB/s": nbytes / ms / 1e6, "frac_of_measured_hbm": nbytes / ms / 1e6 / hbm_peak, "algorithmic_bytes": nbytes}
    out["units"] = (f"{n_chain} chains x {n_frame} frames; algorithmic bytes per unit in DESIGN.md 5.3 (alignment statistics: 60 B per "
                    f"leg-frame in, one 32-byte row per chain out; the three-kernel path also moves its series, 28 B out + 28 B in); "
                    f"peak = {hbm_peak} GB/s (MEASURED_PEAKS.json)")
    return out


def _fk_residual(fk9, pose5):
    return np.linalg.norm(np.asarray(fk9)[:, [5, 6, 7, 8]] - np.asarray(pose5)[:, 1:5], axis=2)


def config1_dict_api(reps=5):
    """BASELINE config 1 on the bundled df3d locomotion recording (6 legs x 100 frames): AlignPose -> LegInvKinSeq with the
    NeuroMechFly locomotion chain through the reference's class API (numpy dicts in, float64 dicts out).  Wall-clock per
    pipeline run and parity of the run against the CPU oracle's angles / FK residuals for the same input (tests/golden/
    locomotion.npz; the reference ships no outputs for this recording)."""
    import torch
    from seqikpy_b200 import data as D
    from seqikpy_b200.alignment import AlignPose
    from seqikpy_b200.kinematic_chain import DOF_ORDER, KinematicChainSeq
    from seqikpy_b200.leg_inverse_kinematics import LegInvKinSeq
    from seqikpy_b200.utils import calculate_body_size
    g = dict(np.load(ROOT / "tests" / "golden" / "locomotion.npz"))
    legs = [str(l) for l in g["legs"]]
    raw = {f"{leg}_leg": g["raw"][i] for i, leg in enumerate(legs)}
    chain = KinematicChainSeq(D.BOUNDS_LOCOMOTION, legs, calculate_body_size(D.TEMPLATE_NMF_LOCOMOTION, legs))

    def run():
        aligned = AlignPose(raw, legs_list=legs, include_claw=False, body_template=D.TEMPLATE_NMF_LOCOMOTION, log_level="ERROR").align_pose()
        angles, fk = LegInvKinSeq(aligned, chain, D.INITIAL_ANGLES_LOCOMOTION, log_level="ERROR").run_ik_and_fk(hide_progress_bar=True)
        return aligned, angles, fk
    run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        aligned, angles, fk = run()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps
    n = int(g["raw"].shape[1])
    worst_angle, worst_fk, err = 0.0, -1.0, []
    for i, leg in enumerate(legs):
        ours = np.stack([angles[f"Angle_{leg}_{d}"] for d in DOF_ORDER], 1)
        worst_angle = max(worst_angle, float(np.abs(ours - g["oracle_angles"][i]).max()))
        r_ours = _fk_residual(fk[f"{leg}_leg"], aligned[f"{leg}_leg"])
        r_ref = _fk_residual(g["oracle_fk"][i], g["aligned"][i])
        worst_fk = max(worst_fk, float((r_ours - r_ref).max()))
        err.append(float(r_ours.mean()))
    return {"workload": "bundled df3d_pose_result__210902_PR_Fly1 locomotion recording: AlignPose + LegInvKinSeq (6 legs x 100 frames, "
                        "NeuroMechFly locomotion chain) through the dict API, raw key points -> float64 dicts",
            "wall_ms_per_pipeline": wall * 1e3, "leg_frames_per_s": len(legs) * n / wall,
            "max_abs_angle_diff_vs_oracle_rad": worst_angle, "max_fk_residual_excess_vs_oracle_mm": worst_fk,
            "mean_fk_error_mm_by_leg": dict(zip(legs, err)),
            "aligned_max_abs_diff_vs_reference_class": float(max(np.abs(aligned[f"{leg}_leg"] - g["aligned"][i]).max() for i, leg in enumerate(legs)))}


def config2_dict_api(reps=3):
    """BASELINE config 2 on the bundled grooming trial (2 legs x 6000 frames + head): AlignPose -> HeadInverseKinematics ->
    LegInvKinSeq through the reference's class API (numpy dicts in, numpy dicts out; cf. the reference's
    examples/example_entire_pipeline.py:60-101, 'about 40 minutes').  Wall-clock per pipeline run, and the parity counts of
    the run against the reference's shipped leg_joint_angles.pkl / forward_kinematics residuals (tests/golden)."""
    import torch
    from seqikpy_b200 import data as D, synthetic as S
    from seqikpy_b200.alignment import AlignPose
    from seqikpy_b200.head_inverse_kinematics import HeadInverseKinematics
    from seqikpy_b200.kinematic_chain import DOF_ORDER, SEGMENTS, KinematicChainSeq
    from seqikpy_b200.leg_inverse_kinematics import LegInvKinSeq
    G = ROOT / "tests" / "golden"
    ga, gl, gh = (dict(np.load(G / n)) for n in ("grooming_align.npz", "grooming_leg.npz", "grooming_head.npz"))
    raw = {"RF_leg": ga["raw_full_RF"], "LF_leg": ga["raw_full_LF"]}
    chain = KinematicChainSeq(D.BOUNDS, ["RF", "LF"], body_size=None)

    def run():
        aligned = AlignPose(raw, legs_list=["RF", "LF"], include_claw=False, body_template=D.NMF_TEMPLATE, log_level="ERROR").align_pose()
        aligned["R_head"], aligned["L_head"] = gh["r_head"], gh["l_head"]
        head = HeadInverseKinematics(aligned, D.NMF_TEMPLATE, log_level="ERROR").compute_head_angles()
        ik = LegInvKinSeq(aligned, chain, D.INITIAL_ANGLES, log_level="ERROR")
        angles, fk = ik.run_ik_and_fk(hide_progress_bar=True)
        return aligned, head, angles, fk, ik
    run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        aligned, head, angles, fk, ik = run()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps
    n = 6000
    res = {"workload": "bundled anipose_220525_aJO_Fly001_001 grooming trial: AlignPose + head/antenna angles + LegInvKinSeq (2 legs x 6000 "
                       "frames) through the dict API, raw key points -> float64 dicts", "wall_ms_per_pipeline": wall * 1e3,
           "leg_frames_per_s": 2 * n / wall, "reference_statement": "about 40 minutes (reference examples/example_entire_pipeline.py:3)",
           "solver_evaluations_per_leg_frame": {leg: (ik.solver_stats[leg]["nfev"] / n).tolist() for leg in ("RF", "LF")}}
    # parity counts of this run against the reference's shipped outputs (windows = +-20 frames around the reference's own
    # CTr_pitch = 0 singular episodes, where its path is rounding noise: tests/helpers.singular_windows)
    size = chain.body_size
    for li, leg in enumerate(("RF", "LF")):
        ours = np.stack([angles[f"Angle_{leg}_{d}"] for d in DOF_ORDER], 1)
        ref = gl["ref_angles"][li]
        bad = np.where(np.abs(ours - ref).max(axis=1) > 1e-3)[0]
        sing = np.where(np.abs(ref[:, 3]) < 1e-6)[0]
        win = set()
        for t in sing:
            win.update(range(max(0, t - 20), t + 21))
        pose = gl["pose"][li]
        seg = np.array([size[f"{leg}_{s}"] for s in SEGMENTS])
        ref_fk = S.leg_key_points(ref, seg) + pose[:, :1]                                   # closed-form FK of the reference's angles
        r_ref = np.linalg.norm(ref_fk - pose[:, 1:5], axis=2)
        r_ours = _fk_residual(fk[f"{leg}_leg"], pose)
        worse = np.where(((r_ours - r_ref) > 1e-4 + 2e-6).any(axis=1))[0]
        res[f"{leg.lower()}_bad_frames_in_windows"] = int(sum(int(t) in win for t in bad))
        res[f"{leg.lower()}_bad_frames_outside_windows"] = int(sum(int(t) not in win for t in bad))
        res[f"{leg.lower()}_fk_worse_frames"] = int(len(worse))
        res[f"{leg.lower()}_fk_worse_frames_outside_windows"] = int(sum(int(t) not in win for t in worse))
        res[f"{leg.lower()}_max_abs_angle_diff_outside_windows"] = float(np.abs(ours - ref)[[t for t in range(n) if t not in win]].max())
        res[f"{leg.lower()}_mean_fk_error_mm"] = {"ours": float(r_ours.mean()), "reference": float(r_ref.mean())}
    href = gh.get("ref_angles")
    if href is not None:
        keys = list(head.keys())
        res["head_max_abs_angle_diff"] = float(max(np.abs(head[k] - href[i]).max() for i, k in enumerate(keys)))
    return res


if __name__ == "__main__":
    import sys
    sys.path.insert(0, str(ROOT))
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "config2"):
        print(json.dumps({"config2_dict_api": config2_dict_api()}))
    if which in ("all", "stream"):
        print(json.dumps({"stream_kernels": stream_kernels(float(peaks.get("hbm_gbs", 6650.0)))}))
    if which in ("all", "config5"):
        print(json.dumps({"config5_fused": config5_fused()}))
