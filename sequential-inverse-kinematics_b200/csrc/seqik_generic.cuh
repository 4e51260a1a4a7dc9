// seqik_generic.cuh -- per-lane solver core of the GENERIC (single-target, 7-DOF) leg IK.
//
// Reference: LegInvKinGeneric.calculate_ik_stage (seqikpy/leg_inverse_kinematics.py:473-547) on the chain of
// KinematicChainGeneric.create_leg_chain (seqikpy/kinematic_chain.py:444-532): Base link, ThC_roll (Z), ThC_yaw (X),
// ThC_pitch (Y), CTr_pitch (Y, after the coxa), CTr_roll (Z), FTi_pitch (Y, after the femur), TiTa_pitch (Y, after the
// tibia), Claw (after the tarsus).  One target -- the claw -- per frame; the solve is ikpy Chain.inverse_kinematics ->
// scipy.optimize.least_squares (method "trf", 2-point Jacobian, ftol = xtol = gtol = 1e-8) over all 9 chain slots, of
// which Base and Claw have identically-zero Jacobian columns.
//
// This header restates scipy's bounded Trust-Region-Reflective iteration (scipy/optimize/_lsq/trf.py:206-413,
// common.py) on the 7 joints:
//   * the Jacobian is analytic: column i = (world axis of joint i) x (claw - origin of joint i);
//   * the two inert slots enter through `null_sq` (their squared norm: initial trust radius trf.py:236, xtol test
//     common.py:705-718) and max_nfev = 100 * 9;
//   * m = 3 < n, so solve_lsq_trust_region (common.py:57-168) always takes its rank-deficient branch: up to 10
//     safeguarded Newton iterations on the Levenberg parameter alpha, the step rescaled to the trust radius, alpha
//     carried between iterations.  scipy evaluates p(alpha) = -V (s uf) / (s^2 + alpha) from one SVD of the augmented
//     matrix [J_h; diag(sqrt(C))], i.e. p(alpha) = -(J_h^T J_h + C + alpha I)^-1 g_h.  With D = C + alpha I (diagonal,
//     positive) that is  p = -D^-1 J_h^T (I + J_h D^-1 J_h^T)^-1 f: ONE SYMMETRIC 3x3 SOLVE instead of an SVD.  D^-1 is
//     normalised by its largest entry (W = eps / D, eps = min D), which also gives d p / d alpha without cancellation:
//         d W / d alpha = W (1 - W) / eps =: U / eps,   M = eps I + J_h W J_h^T,   y = M^-1 f,   p = -W J_h^T y,
//         d p / d alpha = -(U / eps) J_h^T y + W J_h^T M^-1 (y + (J_h U J_h^T) y / eps),
//     every term of the size of the result, so float32 is enough (tests/model_generic.py is the executable
//     specification: in float64 it reproduces scipy as closely as scipy reproduces itself under a 1e-12 input
//     perturbation -- the problem is under-determined).
//
// One trip() = one function evaluation: the FIRST trip of a solve evaluates the (strictly feasible) seed and sets the
// initial trust radius, every later one solves the trust-region subproblem, evaluates the chosen step, accepts or
// rejects it and applies scipy's termination tests.  A warp of lanes at different positions therefore runs one
// straight-line block per iteration.
//
// Compiled by nvcc for the kernel (seqik_generic.cu) and by g++ for the host-side test harness (tests/hostsim).
#pragma once
#include "seqik_core.cuh"

namespace seqik {

constexpr int GEN_DOF = 7;
constexpr int GEN_N_FULL = 9;             // chain slots seen by scipy (Base + 7 joints + Claw)

template <typename R> struct GenNum;
template <> struct GenNum<float> {
    static SK_HD float floor_() { return 1e-30f; }            // keeps 1/(C + alpha) finite
    static SK_HD float next_(float x, float to) { return nextafterf(x, to); }
    static SK_HD void sincos_(float x, float* s, float* c) { float v; Num<float>::sincosv_(x, s, c, &v); }
};
template <> struct GenNum<double> {
    static SK_HD double floor_() { return 1e-300; }
    static SK_HD double next_(double x, double to) { return nextafter(x, to); }
    static SK_HD void sincos_(double x, double* s, double* c) {
#if defined(__CUDA_ARCH__)
        sincos(x, s, c);
#else
        *s = sin(x); *c = cos(x);
#endif
    }
};

template <typename R> SK_HD Vec3<R> cross(const Vec3<R>& a, const Vec3<R>& b) {
    typedef Num<R> N;
    return {N::fma_(a.y, b.z, -(a.z * b.y)), N::fma_(a.z, b.x, -(a.x * b.z)), N::fma_(a.x, b.y, -(a.y * b.x))};
}
template <typename R> SK_HD Vec3<R> sub(const Vec3<R>& a, const Vec3<R>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
// a - s b
template <typename R> SK_HD Vec3<R> nmadd(const Vec3<R>& a, R s, const Vec3<R>& b) {
    typedef Num<R> N;
    return {N::fma_(-s, b.x, a.x), N::fma_(-s, b.y, a.y), N::fma_(-s, b.z, a.z)};
}
template <typename R> SK_HD void rot_x(Mat3<R>& F, R s, R c) { const Vec3<R> a = lin(F.c1, c, F.c2, s); F.c2 = lin(F.c2, c, F.c1, -s); F.c1 = a; }
template <typename R> SK_HD void rot_y(Mat3<R>& F, R s, R c) { const Vec3<R> a = lin(F.c0, c, F.c2, -s); F.c2 = lin(F.c0, s, F.c2, c); F.c0 = a; }
template <typename R> SK_HD void rot_z(Mat3<R>& F, R s, R c) { const Vec3<R> a = lin(F.c0, c, F.c1, s); F.c1 = lin(F.c1, c, F.c0, -s); F.c0 = a; }

// Geometry of the generic chain with the ThC joint at 0: origins of CTr (org[0]), FTi (org[1]), TiTa (org[2]), the claw,
// and (AX) the world rotation axis of every joint.  sn/cs: sin/cos of the 7 joint angles.
template <typename R, bool AX>
SK_HD void generic_chain(const R* sn, const R* cs, R coxa, R femur, R tibia, R tarsus, Vec3<R>* org, Vec3<R>* claw, Vec3<R>* ax) {
    Mat3<R> F = {{R(1), R(0), R(0)}, {R(0), R(1), R(0)}, {R(0), R(0), R(1)}};
    if (AX) ax[0] = F.c2;
    rot_z(F, sn[0], cs[0]);                       // ThC_roll
    if (AX) ax[1] = F.c0;
    rot_x(F, sn[1], cs[1]);                       // ThC_yaw
    if (AX) ax[2] = F.c1;
    rot_y(F, sn[2], cs[2]);                       // ThC_pitch
    Vec3<R> o = {-coxa * F.c2.x, -coxa * F.c2.y, -coxa * F.c2.z};
    org[0] = o;
    if (AX) ax[3] = F.c1;
    rot_y(F, sn[3], cs[3]);                       // CTr_pitch
    if (AX) ax[4] = F.c2;
    rot_z(F, sn[4], cs[4]);                       // CTr_roll
    o = nmadd(o, femur, F.c2);
    org[1] = o;
    if (AX) ax[5] = F.c1;
    rot_y(F, sn[5], cs[5]);                       // FTi_pitch
    o = nmadd(o, tibia, F.c2);
    org[2] = o;
    if (AX) ax[6] = F.c1;
    rot_y(F, sn[6], cs[6]);                       // TiTa_pitch
    *claw = nmadd(o, tarsus, F.c2);
}

// inverse of a symmetric 3x3 (m00 m01 m02 m11 m12 m22) by the adjugate of the matrix normalised by its largest diagonal entry
template <typename R> struct Sym3 { R a, b, c, d, e, f; };
template <typename R> SK_HD Sym3<R> inv_sym3(const Sym3<R>& M) {
    typedef Num<R> N;
    const R sc = N::rcp_(N::max_(N::max_(N::abs_(M.a), N::abs_(M.d)), N::max_(N::abs_(M.f), GenNum<R>::floor_())));
    const R a = M.a * sc, b = M.b * sc, c = M.c * sc, d = M.d * sc, e = M.e * sc, f = M.f * sc;
    const R c00 = N::fma_(d, f, -(e * e)), c01 = N::fma_(c, e, -(b * f)), c02 = N::fma_(b, e, -(c * d));
    R det = N::fma_(a, c00, N::fma_(b, c01, c * c02));
    if (N::abs_(det) < GenNum<R>::floor_()) det = N::copysign_(GenNum<R>::floor_(), det);
    const R r = N::rcp_(det) * sc;
    return {c00 * r, c01 * r, c02 * r, N::fma_(a, f, -(c * c)) * r, N::fma_(b, c, -(a * e)) * r, N::fma_(a, d, -(b * b)) * r};
}
template <typename R> SK_HD Vec3<R> mul(const Sym3<R>& M, const Vec3<R>& v) {
    typedef Num<R> N;
    return {N::fma_(M.c, v.z, N::fma_(M.b, v.y, M.a * v.x)), N::fma_(M.e, v.z, N::fma_(M.d, v.y, M.b * v.x)),
            N::fma_(M.f, v.z, N::fma_(M.e, v.y, M.c * v.x))};
}
template <typename R> SK_HD void add_outer(Sym3<R>& M, R w, const Vec3<R>& j) {
    typedef Num<R> N;
    const R wx = w * j.x, wy = w * j.y, wz = w * j.z;
    M.a = N::fma_(wx, j.x, M.a); M.b = N::fma_(wx, j.y, M.b); M.c = N::fma_(wx, j.z, M.c);
    M.d = N::fma_(wy, j.y, M.d); M.e = N::fma_(wy, j.z, M.e); M.f = N::fma_(wz, j.z, M.f);
}
template <typename R> SK_HD R dot7(const R* a, const R* b) {
    R s = a[0] * b[0];
#pragma unroll
    for (int i = 1; i < GEN_DOF; ++i) s = Num<R>::fma_(a[i], b[i], s);
    return s;
}

template <typename R> SK_HD R gen_quad(const Vec3<R>* Jh, const R* C, const R* gh, const R* s) {   // evaluate_quadratic
    typedef Num<R> N;
    Vec3<R> js = {R(0), R(0), R(0)};
    R q = R(0);
#pragma unroll
    for (int i = 0; i < GEN_DOF; ++i) {
        js.x = N::fma_(Jh[i].x, s[i], js.x); js.y = N::fma_(Jh[i].y, s[i], js.y); js.z = N::fma_(Jh[i].z, s[i], js.z);
        q = N::fma_(s[i] * C[i], s[i], q);
    }
    return N::fma_(R(0.5), dot(js, js) + q, dot7(s, gh));
}
template <typename R> SK_HD Vec3<R> gen_jdot(const Vec3<R>* Jh, const R* s) {
    typedef Num<R> N;
    Vec3<R> js = {R(0), R(0), R(0)};
#pragma unroll
    for (int i = 0; i < GEN_DOF; ++i) { js.x = N::fma_(Jh[i].x, s[i], js.x); js.y = N::fma_(Jh[i].y, s[i], js.y); js.z = N::fma_(Jh[i].z, s[i], js.z); }
    return js;
}
template <typename R> SK_HD R gen_cdot(const R* a, const R* C, const R* b) {     // a^T diag(C) b
    R s = R(0);
#pragma unroll
    for (int i = 0; i < GEN_DOF; ++i) s = Num<R>::fma_(a[i] * C[i], b[i], s);
    return s;
}
// common.py step_size_to_bound: per-variable steps (inf where s = 0) and their minimum
template <typename R> SK_HD R gen_to_bound(const R* x, const R* s, const R* lb, const R* ub, R* steps) {
    typedef Num<R> N;
    R best = N::inf();
#pragma unroll
    for (int i = 0; i < GEN_DOF; ++i) {
        const R rs = R(1) / s[i];
        R st = N::max_((lb[i] - x[i]) * rs, (ub[i] - x[i]) * rs);
        if (s[i] == R(0)) st = N::inf();
        steps[i] = st;
        best = N::min_(best, st);
    }
    return best;
}
// common.py minimize_quadratic_1d on [lo, hi]
template <typename R> SK_HD void gen_minq(R a, R b, R lo, R hi, R c, R* t_out, R* y_out) {
    R bt = lo, by = lo * (a * lo + b) + c;
    const R yh = hi * (a * hi + b) + c;
    if (yh < by) { bt = hi; by = yh; }
    if (a != R(0)) {
        const R ext = R(-0.5) * b / a;
        if (lo < ext && ext < hi) { const R ye = ext * (a * ext + b) + c; if (ye < by) { bt = ext; by = ye; } }
    }
    *t_out = bt; *y_out = by;
}

// trf.py select_step when x + p leaves the box: the trust-region step cut at the first bound, its reflection there, or
// the scaled anti-gradient -- whichever the quadratic model likes best.  NOT a rare path here: the rescale of the
// rank-deficient branch makes |p_h| = Delta, and Delta starts at |x0| ~ 3 rad, so most steps of a generic solve are
// chosen by this function (typically the 1-D minimiser along the anti-gradient: the reference crawls linearly).
// p, ph: in = trust-region step; out = chosen step.
template <typename R>
SK_HD R gen_select_general(const R* x, const R* lb, const R* ub, const R* d, const R* C, const R* gh, const Vec3<R>* Jh,
                           R Delta, R theta, R* p, R* ph) {
    typedef Num<R> N;
    R steps[GEN_DOF];
    const R p_stride = gen_to_bound(x, p, lb, ub, steps);
    R r_h[GEN_DOF], r[GEN_DOF], x_on[GEN_DOF];
#pragma unroll
    for (int i = 0; i < GEN_DOF; ++i) {
        const bool hit = steps[i] == p_stride && p[i] != R(0);
        r_h[i] = hit ? -ph[i] : ph[i];
        r[i] = d[i] * r_h[i];
        p[i] *= p_stride; ph[i] *= p_stride;
        x_on[i] = x[i] + p[i];
    }
    // intersect_trust_region(p_h, r_h, Delta): the positive root
    const R a = dot7(r_h, r_h), b = dot7(ph, r_h);
    const R c = N::min_(dot7(ph, ph) - Delta * Delta, R(0));
    const R disc = N::sqrt_(N::max_(N::fma_(b, b, -(a * c)), R(0)));
    const R qq = -(b + N::copysign_(disc, b));
    R t1 = R(0), t2 = R(0);
    if (qq != R(0)) { t1 = qq / a; t2 = c / qq; }
    const R to_tr = N::max_(t1, t2);
    const R to_bd = gen_to_bound(x_on, r, lb, ub, steps);
    const R r_stride = N::min_(to_bd, to_tr);
    R r_l = R(0), r_u = R(-1);
    if (r_stride > R(0)) {
        r_l = (R(1) - theta) * p_stride / r_stride;
        r_u = (r_stride == to_bd) ? theta * to_bd : to_tr;
    }
    R r_value = N::inf();
    if (r_l <= r_u) {
        // build_quadratic_1d(J_h, g_h, r_h, s0 = p_h, diag = diag_h)
        const Vec3<R> v = gen_jdot(Jh, r_h), u = gen_jdot(Jh, ph);
        const R aa = R(0.5) * (dot(v, v) + gen_cdot(r_h, C, r_h));
        const R bb = dot7(gh, r_h) + dot(u, v) + gen_cdot(ph, C, r_h);
        const R cc = R(0.5) * dot(u, u) + dot7(gh, ph) + R(0.5) * gen_cdot(ph, C, ph);
        R rs;
        gen_minq(aa, bb, r_l, r_u, cc, &rs, &r_value);
#pragma unroll
        for (int i = 0; i < GEN_DOF; ++i) { r_h[i] = N::fma_(r_h[i], rs, ph[i]); r[i] = r_h[i] * d[i]; }
    }
#pragma unroll
    for (int i = 0; i < GEN_DOF; ++i) { p[i] *= theta; ph[i] *= theta; }
    const R p_value = gen_quad(Jh, C, gh, ph);
    R ag_h[GEN_DOF], ag[GEN_DOF];
#pragma unroll
    for (int i = 0; i < GEN_DOF; ++i) { ag_h[i] = -gh[i]; ag[i] = d[i] * ag_h[i]; }
    const R to_tr2 = Delta * N::rsqrt_(dot7(ag_h, ag_h));
    const R to_bd2 = gen_to_bound(x, ag, lb, ub, steps);
    const R ag_hi = (to_bd2 < to_tr2) ? theta * to_bd2 : to_tr2;
    const Vec3<R> vg = gen_jdot(Jh, ag_h);
    R ags, ag_value;
    gen_minq(R(0.5) * (dot(vg, vg) + gen_cdot(ag_h, C, ag_h)), dot7(gh, ag_h), R(0), ag_hi, R(0), &ags, &ag_value);
    if (p_value < r_value && p_value < ag_value) return -p_value;
    if (r_value < p_value && r_value < ag_value) {
#pragma unroll
        for (int i = 0; i < GEN_DOF; ++i) { p[i] = r[i]; ph[i] = r_h[i]; }
        return -r_value;
    }
#pragma unroll
    for (int i = 0; i < GEN_DOF; ++i) { p[i] = ag[i] * ags; ph[i] = ag_h[i] * ags; }
    return -ag_value;
}

// ---------------------------------------------------------------------------------
// One generic solve: the iterate + one evaluation per trip()
// ---------------------------------------------------------------------------------
template <typename R>
struct GenericSolve {
    typedef Num<R> N;
    typedef GenNum<R> G;
    // problem: constants row (seg 0..3, lb 4..10, ub 11..17, joints in chain order), target, inert slots
    const R* prm; Vec3<R> q; R null_sq;
    // iterate
    R x[GEN_DOF]; Vec3<R> J[GEN_DOF]; R g[GEN_DOF]; Vec3<R> f;
    R cost, Delta, alpha; int nfev, status;

    SK_HD R ldc(int i) const {
#if defined(__CUDA_ARCH__)
        return __ldg(prm + i);
#else
        return prm[i];
#endif
    }
    SK_HD R lb(int i) const { return ldc(4 + i); }
    SK_HD R ub(int i) const { return ldc(11 + i); }
    SK_HD bool done() const { return status != ST_RUNNING; }

    // `x` must already hold the seed (the previous frame's angles: leg_inverse_kinematics.py:524)
    SK_HD void start(const R* prm_, const Vec3<R>& target, R null_sq_) {
        prm = prm_; q = target; null_sq = null_sq_; nfev = 0; status = ST_RUNNING; alpha = R(0); Delta = R(1); cost = R(0);
    }

    // Coleman-Li scaling (common.py CL_scaling_vector); every joint has finite bounds.  Returns |g v|_inf.
    SK_HD R scaling(R* v, R* dv) const {
        R gn = R(0);
#pragma unroll
        for (int i = 0; i < GEN_DOF; ++i) {
            v[i] = R(1); dv[i] = R(0);
            if (g[i] < R(0)) { v[i] = ub(i) - x[i]; dv[i] = R(-1); }
            else if (g[i] > R(0)) { v[i] = x[i] - lb(i); dv[i] = R(1); }
            gn = N::max_(gn, N::abs_(g[i] * v[i]));
        }
        return gn;
    }

    // p(alpha) = -(J_h^T J_h + C + alpha I)^-1 g_h and (DP) its alpha-derivative, see the header comment
    template <bool DP>
    SK_HD void tr_point(const Vec3<R>* Jh, const R* C, R al, R* p, R* dp) const {
        // alpha may be NEGATIVE here: scipy's last Newton update is not safeguarded (common.py:156-161), so D can have
        // entries of either sign.  The identities above hold for any non-zero normaliser; eps = the entry of smallest
        // magnitude keeps |W| <= 1.
        R D[GEN_DOF], W[GEN_DOF], U[GEN_DOF];
        R eps = N::inf();
#pragma unroll
        for (int i = 0; i < GEN_DOF; ++i) {
            R Di = C[i] + al;
            if (N::abs_(Di) < G::floor_()) Di = N::copysign_(G::floor_(), Di);
            D[i] = Di;
            if (N::abs_(Di) < N::abs_(eps)) eps = Di;
        }
        Sym3<R> M = {eps, R(0), R(0), eps, R(0), eps};
        Sym3<R> K = {R(0), R(0), R(0), R(0), R(0), R(0)};
#pragma unroll
        for (int i = 0; i < GEN_DOF; ++i) {
            const R rd = N::rcp_(D[i]);
            W[i] = eps * rd;
            add_outer(M, W[i], Jh[i]);
            if (DP) { U[i] = W[i] * ((D[i] - eps) * rd); add_outer(K, U[i], Jh[i]); }
        }
        const Sym3<R> Mi = inv_sym3(M);
        const Vec3<R> y = mul(Mi, f);
        R jy[GEN_DOF];
#pragma unroll
        for (int i = 0; i < GEN_DOF; ++i) { jy[i] = dot(Jh[i], y); p[i] = -(W[i] * jy[i]); }
        if (DP) {
            const R re = N::rcp_(eps);
            const Vec3<R> ky = mul(K, y);
            const Vec3<R> z = mul(Mi, Vec3<R>{N::fma_(ky.x, re, y.x), N::fma_(ky.y, re, y.y), N::fma_(ky.z, re, y.z)});
#pragma unroll
            for (int i = 0; i < GEN_DOF; ++i) dp[i] = N::fma_(W[i], dot(Jh[i], z), -((U[i] * re) * jy[i]));
        }
    }

    // the step to evaluate next (trf.py:296-335: scaling, trust-region subproblem, select_step); xt = x + step made strictly feasible
    SK_HD void propose(R* xt, R* step_h_norm, R* step_norm, R* pred) {
        R v[GEN_DOF], dv[GEN_DOF], d[GEN_DOF], C[GEN_DOF], gh[GEN_DOF];
        Vec3<R> Jh[GEN_DOF];
        const R g_norm = scaling(v, dv);
#pragma unroll
        for (int i = 0; i < GEN_DOF; ++i) {
            d[i] = N::sqrt_(v[i]); C[i] = g[i] * dv[i]; gh[i] = d[i] * g[i];
            Jh[i] = {J[i].x * d[i], J[i].y * d[i], J[i].z * d[i]};
        }
        const R theta = N::max_(R(0.995), R(1) - g_norm);
        const R gh_norm = N::sqrt_(dot7(gh, gh));
        // ---- solve_lsq_trust_region, rank-deficient branch (common.py:132-166)
        const R rD = N::rcp_(Delta);
        R a_up = gh_norm * rD, a_lo = R(0);
        if (alpha == R(0)) alpha = R(0.001) * a_up;
        bool brk = false;
        R p[GEN_DOF], dp[GEN_DOF];
#pragma unroll 1
        for (int it = 0; it < 10; ++it) {
            if (!brk) {
                if (alpha < a_lo || alpha > a_up) alpha = N::max_(R(0.001) * a_up, N::sqrt_(a_lo * a_up));
                tr_point<true>(Jh, C, alpha, p, dp);
                const R pn = N::sqrt_(dot7(p, p));
                const R phi = pn - Delta;
                const R dphi = dot7(p, dp) * N::rcp_(pn);
                if (phi < R(0)) a_up = alpha;
                const R ratio = phi * N::rcp_(dphi);
                a_lo = N::max_(a_lo, alpha - ratio);
                alpha -= (phi + Delta) * ratio * rD;
                brk = N::abs_(phi) < R(0.01) * Delta;
            }
        }
        R ph[GEN_DOF];
        tr_point<false>(Jh, C, alpha, ph, (R*)nullptr);
        const R scl = Delta * N::rsqrt_(dot7(ph, ph));
        bool inb = true;
#pragma unroll
        for (int i = 0; i < GEN_DOF; ++i) {
            ph[i] *= scl; p[i] = d[i] * ph[i];
            const R xn = x[i] + p[i];
            inb = inb && xn >= lb(i) && xn <= ub(i);
        }
        R lbv[GEN_DOF], ubv[GEN_DOF];
#pragma unroll
        for (int i = 0; i < GEN_DOF; ++i) { lbv[i] = lb(i); ubv[i] = ub(i); }
        if (inb) *pred = -gen_quad(Jh, C, gh, ph);
        else *pred = gen_select_general(x, lbv, ubv, d, C, gh, Jh, Delta, theta, p, ph);
        *step_h_norm = N::sqrt_(dot7(ph, ph));
        const R* step = p;
        *step_norm = N::sqrt_(dot7(step, step));
        // make_strictly_feasible(x + step, lb, ub, rstep = 0)
#pragma unroll
        for (int i = 0; i < GEN_DOF; ++i) {
            R xn = x[i] + step[i];
            const R l = lbv[i], u = ubv[i];
            if (xn <= l) xn = G::next_(l, u);
            if (xn >= u) xn = G::next_(u, l);
            xt[i] = xn;
        }
    }

    // least_squares.py: x0 = make_strictly_feasible(x0, lb, ub) (rstep = 1e-10; in float32 the nudge is one ulp)
    SK_HD void feasible_seed(R* xt) const {
#pragma unroll
        for (int i = 0; i < GEN_DOF; ++i) {
            const R l = lb(i), u = ub(i);
            R xv = x[i];
            const R dl = xv - l, du = u - xv;
            if (dl <= N::min_(du, R(1e-10) * N::max_(R(1), N::abs_(l)))) { xv = l + R(1e-10) * N::max_(R(1), N::abs_(l)); if (xv <= l) xv = G::next_(l, u); }
            else if (du <= N::min_(dl, R(1e-10) * N::max_(R(1), N::abs_(u)))) { xv = u - R(1e-10) * N::max_(R(1), N::abs_(u)); if (xv >= u) xv = G::next_(u, l); }
            if (xv < l || xv > u) xv = R(0.5) * (l + u);
            xt[i] = xv;
        }
    }

    // one function evaluation
    SK_HD void trip() {
        const bool first = nfev == 0;
        R xt[GEN_DOF];
        R step_h_norm = R(0), step_norm = R(0), pred = R(0);
        if (first) feasible_seed(xt); else propose(xt, &step_h_norm, &step_norm, &pred);
        R sn[GEN_DOF], cs[GEN_DOF];
#pragma unroll
        for (int i = 0; i < GEN_DOF; ++i) G::sincos_(xt[i], &sn[i], &cs[i]);
        Vec3<R> org[3], claw, ax[GEN_DOF];
        generic_chain<R, true>(sn, cs, ldc(0), ldc(1), ldc(2), ldc(3), org, &claw, ax);
        const Vec3<R> fn = sub(claw, q);
        const R cost_new = R(0.5) * dot(fn, fn);
        ++nfev;
        bool accept = first;
        if (first) {
            if (!(cost_new < N::inf())) { status = ST_NONFINITE; return; }      // scipy raises: the solve is skipped
        } else if (!(cost_new < N::inf())) {
            Delta = R(0.25) * step_h_norm;                                    // trf.py:344-346
        } else {
            const R actual = cost - cost_new;
            // update_tr_radius (common.py)
            R ratio = R(0);
            if (pred > R(0)) ratio = actual / pred; else if (pred == R(0) && actual == R(0)) ratio = R(1);
            R Delta_new = Delta;
            if (ratio < R(0.25)) Delta_new = R(0.25) * step_h_norm;
            else if (ratio > R(0.75) && step_h_norm > R(0.95) * Delta) Delta_new = R(2) * Delta;
            // check_termination (common.py)
            const R x_norm = N::sqrt_(null_sq + dot7(x, x));
            const bool ft = actual < R(1e-8) * cost && ratio > R(0.25);
            const bool xt_ = step_norm < R(1e-8) * (R(1e-8) + x_norm);
            if (ft && xt_) status = ST_BOTH; else if (ft) status = ST_FTOL; else if (xt_) status = ST_XTOL;
            if (status == ST_RUNNING) { alpha *= Delta / Delta_new; Delta = Delta_new; }
            accept = actual > R(0);
        }
        if (accept) {
            cost = cost_new; f = fn;
#pragma unroll
            for (int i = 0; i < GEN_DOF; ++i) x[i] = xt[i];
            // Jacobian column i = axis_i x (claw - origin_i); gradient g = J^T f
            const Vec3<R> l1 = sub(claw, org[0]), l2 = sub(claw, org[1]), l3 = sub(claw, org[2]);
            J[0] = cross(ax[0], claw); J[1] = cross(ax[1], claw); J[2] = cross(ax[2], claw);
            J[3] = cross(ax[3], l1); J[4] = cross(ax[4], l1); J[5] = cross(ax[5], l2); J[6] = cross(ax[6], l3);
#pragma unroll
            for (int i = 0; i < GEN_DOF; ++i) g[i] = dot(J[i], f);
        }
        if (first) {                 // trf.py:234-238: Delta0 = |x0 / sqrt(v)| over ALL chain slots
            R v[GEN_DOF], dv[GEN_DOF];
            scaling(v, dv);
            R s = null_sq;
#pragma unroll
            for (int i = 0; i < GEN_DOF; ++i) s = N::fma_(x[i] * x[i], N::rcp_(v[i]), s);
            Delta = N::sqrt_(s);
            if (Delta == R(0)) Delta = R(1);
        }
        // head of scipy's outer loop (trf.py:262-273): gtol test (it overrides ftol/xtol), evaluation limit
        if (accept || status != ST_RUNNING) {
            R v[GEN_DOF], dv[GEN_DOF];
            if (scaling(v, dv) < R(1e-8)) status = ST_GTOL;
        }
        if (status == ST_RUNNING && nfev >= 100 * GEN_N_FULL) status = ST_MAXFEV;
    }

    // 9 rows of forward_kinematics(full_kinematics=True) relative to the ThC: rows 0-3 zero, 4-5 CTr, 6 FTi, 7 TiTa, 8 claw
    SK_HD void joints(Vec3<R>* org, Vec3<R>* claw) const {
        R sn[GEN_DOF], cs[GEN_DOF];
#pragma unroll
        for (int i = 0; i < GEN_DOF; ++i) G::sincos_(x[i], &sn[i], &cs[i]);
        generic_chain<R, false>(sn, cs, ldc(0), ldc(1), ldc(2), ldc(3), org, claw, (Vec3<R>*)nullptr);
    }
};

}  // namespace seqik
