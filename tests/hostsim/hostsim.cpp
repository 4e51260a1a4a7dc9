// Host build of csrc/seqik_core.cuh for CPU-side algorithm tests (tests/ only).
// TEST INFRASTRUCTURE: compiled by tests/hostsim_build.py with g++, loaded via ctypes by
// tests; the product package never loads it.
#include "seqik_core.cuh"
#include "seqik_generic.cuh"
#include <cstdint>

using namespace seqik;

template <typename R>
static void run_chain(const R* pose, int64_t n_frame, const R* seg, const R* lb, const R* ub,
                      const R* null_sq, const R* seed, R* angles, R* fk, int32_t* nfev, int32_t* status,
                      int stage_mask, int gn_mask) {
    ChainParams<R> P;
    for (int i = 0; i < 4; ++i) { P.seg[i] = seg[i]; P.null_sq[i] = null_sq[i]; }
    for (int i = 0; i < 7; ++i) { P.lb[i] = lb[i]; P.ub[i] = ub[i]; }
    R ang[7];
    for (int i = 0; i < 7; ++i) ang[i] = seed[i];
    for (int64_t t = 0; t < n_frame; ++t) {
        FrameStats fs;
        solve_frame<R>(P, pose + t * 15, ang, fk ? fk + t * 27 : nullptr, &fs, stage_mask, gn_mask);
        for (int i = 0; i < 7; ++i) angles[t * 7 + i] = ang[i];
        if (nfev) for (int s = 0; s < 4; ++s) { nfev[t * 4 + s] = fs.nfev[s]; status[t * 4 + s] = fs.status[s]; }
    }
}

// ---- the decoupled per-lane runner, driven serially (one step() at a time) ----
template <typename R> struct HostIO {
    const R* pose; const R* seg_; const R* lb_; const R* ub_; const R* nsq_; R* angles; R* fk;
    Vec3<R> kp(int64_t t, int row) const { const R* p = pose + t * 15 + row * 3; return {p[0], p[1], p[2]}; }
    void put_angles(int64_t t, const R* a, int i0, int i1) const { for (int i = i0; i < i1; ++i) angles[t * 7 + i] = a[i]; }
    R angle_in(int64_t t, int i) const { return angles[t * 7 + i]; }
    void put_fk(int64_t t, int row, const Vec3<R>& v) const { R* p = fk + t * 27 + row * 3; p[0] = v.x; p[1] = v.y; p[2] = v.z; }
    R seg(int i) const { return seg_[i]; }
    R lb(int i) const { return lb_[i]; }
    R ub(int i) const { return ub_[i]; }
    R null_sq(int i) const { return nsq_[i]; }
};
template <typename R>
static int64_t run_runner(const R* pose, int64_t n_frame, const R* seg, const R* lb, const R* ub, const R* null_sq,
                          const R* seed, R* angles, R* fk, uint32_t* nfev_sum, int stage_mask, int gn_mask) {
    HostIO<R> io{pose, seg, lb, ub, null_sq, angles, fk};
    ChainRunner<R, HostIO<R>> run;
    run.start(io, n_frame, seed, stage_mask, gn_mask);
    int64_t steps = 0;
    while (!run.finished()) { run.step(); ++steps; }
    nfev_sum[0] = run.nf0; nfev_sum[1] = run.nf1; nfev_sum[2] = run.nf2; nfev_sum[3] = run.nf3;
    return steps;
}

extern "C" {
int64_t hostsim_runner_f32(const float* pose, int64_t n_frame, const float* seg, const float* lb, const float* ub,
                           const float* null_sq, const float* seed, float* angles, float* fk, uint32_t* nfev_sum,
                           int stage_mask, int gn_mask) {
    return run_runner<float>(pose, n_frame, seg, lb, ub, null_sq, seed, angles, fk, nfev_sum, stage_mask, gn_mask);
}
void hostsim_chain_f32(const float* pose, int64_t n_frame, const float* seg, const float* lb, const float* ub,
                       const float* null_sq, const float* seed, float* angles, float* fk, int32_t* nfev,
                       int32_t* status, int stage_mask, int gn_mask) {
    run_chain<float>(pose, n_frame, seg, lb, ub, null_sq, seed, angles, fk, nfev, status, stage_mask, gn_mask);
}
void hostsim_chain_f64(const double* pose, int64_t n_frame, const double* seg, const double* lb, const double* ub,
                       const double* null_sq, const double* seed, double* angles, double* fk, int32_t* nfev,
                       int32_t* status, int stage_mask, int gn_mask) {
    run_chain<double>(pose, n_frame, seg, lb, ub, null_sq, seed, angles, fk, nfev, status, stage_mask, gn_mask);
}
}

// ---- the stage-pipeline kernel's per-lane arithmetic, run serially: the solve of (stage, frame) is carried from
// (stage, frame - 1) by StageSolve::restart and re-derived by init every SEQIK_RESYNC frames, exactly as
// leg_solve_pipe_kernel does (csrc/seqik_solver.cu); scheduling does not change the arithmetic.
static void run_carried(const float* pose, int64_t n_frame, const float* prm, float* angles, float* fk, int32_t* nfev,
                        int gn_mask) {
    typedef float R;
    StageSolve<R> S[4];
    bool carried[4] = {false, false, false, false};
    R xa[4], xb[4];
    for (int s = 0; s < 4; ++s) { xa[s] = (s == 3) ? 0.f : prm[18 + 2 * s]; xb[s] = prm[18 + ((s == 3) ? 6 : 2 * s + 1)]; }
    const R inf = Num<R>::inf();
    const bool esc = (gn_mask >> 4) & 1;
    for (int64_t t = 0; t < n_frame; ++t) {
        const R* kp = pose + t * 15;
        const Vec3<R> o = {kp[0], kp[1], kp[2]};
        Mat3<R> A = {{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}};
        Vec3<R> piv = {0.f, 0.f, 0.f};
        R* f9 = fk + t * 27;
        for (int s = 0; s < 4; ++s) {
            const int kind = (s == 0) ? KIND_XY : KIND_ZY;
            const int ia = 2 * s, ib = (s == 3) ? 6 : 2 * s + 1;
            const R lb0 = (s == 3) ? -inf : prm[4 + ia], ub0 = (s == 3) ? inf : prm[11 + ia];
            const R lb1 = prm[4 + ib], ub1 = prm[11 + ib];
            const int n_full = (s == 0) ? 4 : (s == 1) ? 6 : (s == 2) ? 8 : 9;
            const int gn = stage_mode(gn_mask, s);
            const Vec3<R> k = {kp[3 * (s + 1)], kp[3 * (s + 1) + 1], kp[3 * (s + 1) + 2]};
            const Vec3<R> rel = {(k.x - o.x) - piv.x, (k.y - o.y) - piv.y, (k.z - o.z) - piv.z};
            const Vec3<R> q3 = mulT(A, rel);
            const bool fresh = !(carried[s] && (t & (SEQIK_RESYNC - 1)) != 0);
            if (fresh) {
                const bool had = S[s].have_bt && carried[s];
                const float l0 = S[s].sl0, l1 = S[s].cl0, u0 = S[s].su0, u1 = S[s].cu0;
                S[s].set_problem(kind, prm[s], (s == 3) ? 0.f : 1.f, prm[25 + s], n_full, gn);
                if (!carried[s]) S[s].set_limit_trig(lb0, ub0);          // once per (chain, stage), like the kernel
                else { S[s].have_bt = had; S[s].sl0 = l0; S[s].cl0 = l1; S[s].su0 = u0; S[s].cu0 = u1; }
                S[s].set_iterate(xa[s], xb[s]); carried[s] = true;
            }
            S[s].restart(q3, lb0, ub0, lb1, ub1, fresh, t > 0);
            for (;;) {
                while (!S[s].done()) S[s].trip();
                if (!(esc && S[s].escape())) break;
            }
            xa[s] = S[s].x0; xb[s] = S[s].angle_b();
            if (s != 3) angles[t * 7 + ia] = xa[s];
            angles[t * 7 + ib] = xb[s];
            nfev[t * 4 + s] = S[s].nfev;
            const Vec3<R> Af = mul(A, S[s].res());
            const Vec3<R> np_ = {(piv.x + rel.x) + Af.x, (piv.y + rel.y) + Af.y, (piv.z + rel.z) + Af.z};
            const Vec3<R> jw = {np_.x + o.x, np_.y + o.y, np_.z + o.z};
            f9[3 * s] = o.x; f9[3 * s + 1] = o.y; f9[3 * s + 2] = o.z;
            f9[15 + 3 * s] = jw.x; f9[16 + 3 * s] = jw.y; f9[17 + 3 * s] = jw.z;
            if (s == 0) { f9[12] = jw.x; f9[13] = jw.y; f9[14] = jw.z; }
            A = rotate_frame(A, kind, S[s].sa, S[s].ca, S[s].sin_b(), S[s].cos_b());
            piv = np_;
        }
    }
}
extern "C" void hostsim_carried_f32(const float* pose, int64_t n_frame, const float* prm, float* angles, float* fk,
                                    int32_t* nfev, int gn_mask) {
    run_carried(pose, n_frame, prm, angles, fk, nfev, gn_mask);
}

// ---- generic 7-DOF solve (csrc/seqik_generic.cuh): pose (N,2,3) = ThC origin + claw per frame; prm = 32-float row
// (seg 0..3, lb 4..10, ub 11..17, seed 18..24, null_sq 25; joints in generic chain order).  teacher: NULL, or (N,7)
// seeds that replace the warm start of every frame (teacher-forced single solves).
template <typename R>
static void run_generic(const R* pose, int64_t n_frame, const R* prm, const R* teacher, R* angles, R* fk,
                        int32_t* nfev, int32_t* status) {
    GenericSolve<R> S;
    for (int i = 0; i < 7; ++i) S.x[i] = prm[18 + i];
    for (int64_t t = 0; t < n_frame; ++t) {
        const R* p = pose + t * 6;
        if (teacher) for (int i = 0; i < 7; ++i) S.x[i] = teacher[t * 7 + i];
        S.start(prm, Vec3<R>{p[3] - p[0], p[4] - p[1], p[5] - p[2]}, prm[25]);
        while (!S.done()) S.trip();
        for (int i = 0; i < 7; ++i) angles[t * 7 + i] = S.x[i];
        nfev[t] = S.nfev; status[t] = S.status;
        if (fk) {
            Vec3<R> org[3], claw;
            S.joints(org, &claw);
            R* o = fk + t * 27;
            for (int r = 0; r < 4; ++r) { o[3 * r] = p[0]; o[3 * r + 1] = p[1]; o[3 * r + 2] = p[2]; }
            const Vec3<R> rows[5] = {org[0], org[0], org[1], org[2], claw};
            for (int r = 0; r < 5; ++r) { o[12 + 3 * r] = rows[r].x + p[0]; o[13 + 3 * r] = rows[r].y + p[1]; o[14 + 3 * r] = rows[r].z + p[2]; }
        }
    }
}

extern "C" {
void hostsim_generic_f32(const float* pose, int64_t n_frame, const float* prm, const float* teacher, float* angles,
                         float* fk, int32_t* nfev, int32_t* status) {
    run_generic<float>(pose, n_frame, prm, teacher, angles, fk, nfev, status);
}
void hostsim_generic_f64(const double* pose, int64_t n_frame, const double* prm, const double* teacher, double* angles,
                         double* fk, int32_t* nfev, int32_t* status) {
    run_generic<double>(pose, n_frame, prm, teacher, angles, fk, nfev, status);
}
}

// ---- self-checks of refactors that are otherwise only visible on the GPU: the select-based frame rotation equals the
// branching one bit for bit, and restart() equals its three pieces.  Returns the number of mismatches.
extern "C" int hostsim_selfcheck(const float* rnd, int n) {
    int bad = 0;
    for (int i = 0; i + 16 <= n; i += 16) {
        const float* r = rnd + i;
        const Mat3<float> A = {{r[0], r[1], r[2]}, {r[3], r[4], r[5]}, {r[6], r[7], r[8]}};
        for (int kind = 0; kind < 2; ++kind) {
            const Mat3<float> X = rotate_frame(A, kind, r[9], r[10], r[11], r[12]);
            const Mat3<float> Y = rotate_frame_sel(A, kind, r[9], r[10], r[11], r[12]);
            const float* x = &X.c0.x; const float* y = &Y.c0.x;
            for (int k = 0; k < 9; ++k) bad += !(x[k] == y[k]);
        }
        for (int fresh = 0; fresh < 2; ++fresh) {
            StageSolve<float> S1, S2;
            const Vec3<float> q0 = {0.3f * r[0], 0.3f * r[1], -0.4f + 0.1f * r[2]};
            S1.init(KIND_ZY, 0.4f, 1.f, q0, 0.2f * r[3], -0.8f + 0.2f * r[4], -1.f, 1.f, -2.f, 0.f, 1.5f, 6, 15);
            while (!S1.done()) S1.trip();
            S2 = S1;
            const Vec3<float> q1 = {q0.x + 0.02f * r[5], q0.y + 0.02f * r[6], q0.z + 0.02f * r[7]};
            S1.restart(q1, -1.f, 1.f, -2.f, 0.f, fresh != 0, true);
            const Vec3<float> q = S2.restart_a(q1, -1.f, 1.f, -2.f, 0.f, fresh != 0);
            S2.warm_step(q, -1.f, 1.f, -2.f - S2.shift, 0.f - S2.shift, S2.closed_form && S2.gn_mode);
            S2.restart_b(q, -1.f, 1.f, -2.f, 0.f);
            bad += !(S1.x0 == S2.x0 && S1.x1 == S2.x1 && S1.status == S2.status && S1.cost == S2.cost && S1.sa == S2.sa && S1.cb == S2.cb);
        }
    }
    return bad;
}

// ---- the closed-form warm step on its own: a Rz(a) Ry(b) stage (two variables, or one with has_a = 0) carried at
// (a, b) meets target q; out = (a', b', seeded, seed_at, cost at the new point).
extern "C" void hostsim_warm_step_f64(double L, double has_a, const double* ab, const double* q, const double* lbub, double* out) {
    StageSolve<double> S;
    S.set_problem(KIND_ZY, L, has_a, 1.0, 6, 15);
    S.set_limit_trig(lbub[0], lbub[1]);
    S.set_iterate(ab[0], ab[1]);
    const Vec3<double> qq = {q[0], q[1], q[2]};
    S.restart(qq, lbub[0], lbub[1], lbub[2], lbub[3], true, true);
    out[0] = S.x0; out[1] = S.angle_b(); out[2] = S.seeded ? 1.0 : 0.0; out[3] = S.seed_at; out[4] = S.cost;
}

// ---- one solve step on its own: init at (a, b) against q, then `n_trip` evaluations; out = (a', b', cost, nfev, status)
extern "C" void hostsim_trips_f64(double L, const double* ab, const double* q, const double* lbub, int mode, int n_trip, double* out) {
    StageSolve<double> S;
    const Vec3<double> qq = {q[0], q[1], q[2]};
    S.init(KIND_ZY, L, 1.0, qq, ab[0], ab[1], lbub[0], lbub[1], lbub[2], lbub[3], 1.0, 6, mode);
    for (int i = 0; i < n_trip && !S.done(); ++i) S.trip();
    out[0] = S.x0; out[1] = S.angle_b(); out[2] = S.cost; out[3] = S.nfev; out[4] = S.status;
}
