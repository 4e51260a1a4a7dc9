// seqik_generic_group.cuh -- the generic (single-target, 7-DOF) leg-IK solve with ONE CHAIN SPREAD OVER EIGHT LANES.
//
// Same iteration as GenericSolve (seqik_generic.cuh: scipy's bounded Trust-Region-Reflective loop on the chain of
// KinematicChainGeneric, seqikpy/kinematic_chain.py:444-532, driven by LegInvKinGeneric.calculate_ik_stage,
// seqikpy/leg_inverse_kinematics.py:473-547), but lane g of an aligned group of eight owns joint g (lane 7 is a joint that
// does not move: zero Jacobian column, zero step) and everything a solve sums over the joints -- the 3x3 matrix of the
// Levenberg subproblem, norms, dot products, the minima of select_step -- is a butterfly reduction over the group
// (shfl.xor 1, 2, 4: every lane ends with the same bits because the additions pair up symmetrically).  The per-joint work
// (reciprocals, weights, outer products, bound tests) happens once instead of seven times and a lane holds a quarter of the
// state: 115 registers in float32, 183 and NO spills in float64 (the one-lane form: 194, and 255 with 480 B spilled).
// Measured: the evaluation remains a latency chain -- a butterfly sum costs what the seven-term sum did --
// so a single chain is only 13 % faster (12.4 against 14.2 us per evaluation); the gain is in float64 (6 000 chains 282
// against 410 ms) and in batches that fit the GPU at once (float32 6 000 chains: 71 against 82 ms).  Larger batches stay on
// the one-lane kernel, which keeps five times as many chains in flight per SM.  The scalars of a solve (cost, radius,
// alpha, counters, status) are replicated in the eight lanes and stay identical because they are computed from reduced
// values only.
//
// Sums are taken in butterfly order, not joint by joint, so results differ from the one-lane form in the last bits; for
// this under-determined problem that is the same as another BLAS under scipy: parity is the teacher-forced
// agreement with the oracle, which both forms meet.  Device only (the one-lane form stays the host-buildable specification).
#pragma once
#include "seqik_generic.cuh"

namespace seqik {

#if defined(__CUDACC__)

template <typename R>
struct GroupSolve {
    typedef Num<R> N;
    typedef GenNum<R> G;
    // lane-own: joint `gl` of the chain
    int gl, gbase; unsigned gmask; bool real;
    R x, lb, ub, g; Vec3<R> J;
    // replicated over the group
    R seg0, seg1, seg2, seg3, null_sq;
    Vec3<R> q, f;
    R cost, Delta, alpha; int nfev, status;

    __device__ __forceinline__ R xsh(R v, int m) const { return __shfl_xor_sync(gmask, v, m); }
    __device__ __forceinline__ R gsum(R v) const { v += xsh(v, 1); v += xsh(v, 2); v += xsh(v, 4); return v; }
    __device__ __forceinline__ R gmin(R v) const { v = N::min_(v, xsh(v, 1)); v = N::min_(v, xsh(v, 2)); v = N::min_(v, xsh(v, 4)); return v; }
    __device__ __forceinline__ R gmax(R v) const { v = N::max_(v, xsh(v, 1)); v = N::max_(v, xsh(v, 2)); v = N::max_(v, xsh(v, 4)); return v; }
    __device__ __forceinline__ Vec3<R> gsum3(Vec3<R> v) const {
#pragma unroll
        for (int m = 1; m < 8; m <<= 1) { const R a = xsh(v.x, m), b = xsh(v.y, m), c = xsh(v.z, m); v.x += a; v.y += b; v.z += c; }
        return v;
    }
    __device__ __forceinline__ bool gall(bool p) const { return (__ballot_sync(gmask, p) & gmask) == gmask; }
    __device__ __forceinline__ bool done() const { return status != ST_RUNNING; }

    __device__ __forceinline__ void init(const R* prm, int lane) {
        gl = lane & 7; gbase = lane & 24; gmask = 0xffu << gbase; real = gl < GEN_DOF;
        seg0 = __ldg(prm); seg1 = __ldg(prm + 1); seg2 = __ldg(prm + 2); seg3 = __ldg(prm + 3); null_sq = __ldg(prm + 25);
        lb = real ? __ldg(prm + 4 + gl) : R(-1); ub = real ? __ldg(prm + 11 + gl) : R(1);
        x = R(0); g = R(0); J = {R(0), R(0), R(0)};
        status = ST_GTOL; nfev = 0; alpha = R(0); Delta = R(1); cost = R(0); f = {R(0), R(0), R(0)}; q = f;
    }
    __device__ __forceinline__ void start(const Vec3<R>& target) {
        q = target; nfev = 0; status = ST_RUNNING; alpha = R(0); Delta = R(1); cost = R(0);
    }

    // Coleman-Li scaling of the own joint (common.py CL_scaling_vector)
    __device__ __forceinline__ void scaling(R& v, R& dv) const {
        v = R(1); dv = R(0);
        if (g < R(0)) { v = ub - x; dv = R(-1); }
        else if (g > R(0)) { v = x - lb; dv = R(1); }
    }

    // own component of p(alpha) = -(J_h^T J_h + C + alpha I)^-1 g_h and (DP) of its alpha-derivative (seqik_generic.cuh header)
    template <bool DP>
    __device__ __forceinline__ void tr_point(const Vec3<R>& Jh, R C, R al, R& p, R& dp) const {
        R Di = C + al;
        if (N::abs_(Di) < G::floor_()) Di = N::copysign_(G::floor_(), Di);
        // eps = the entry of smallest magnitude over the real joints (ties: the larger value), the same bits in every lane
        R ea = real ? N::abs_(Di) : N::inf(), ev = Di;
#pragma unroll
        for (int m = 1; m < 8; m <<= 1) {
            const R oa = xsh(ea, m), ov = xsh(ev, m);
            const bool take = oa < ea || (oa == ea && ov > ev);
            ea = take ? oa : ea; ev = take ? ov : ev;
        }
        const R eps = ev;
        const R rd = N::rcp_(Di);
        const R W = real ? eps * rd : R(0);
        const R U = real ? W * ((Di - eps) * rd) : R(0);
        const R wx = W * Jh.x, wy = W * Jh.y, wz = W * Jh.z;
        R m0 = wx * Jh.x, m1 = wx * Jh.y, m2 = wx * Jh.z, m3 = wy * Jh.y, m4 = wy * Jh.z, m5 = wz * Jh.z;
#pragma unroll
        for (int m = 1; m < 8; m <<= 1) {
            const R a0 = xsh(m0, m), a1 = xsh(m1, m), a2 = xsh(m2, m), a3 = xsh(m3, m), a4 = xsh(m4, m), a5 = xsh(m5, m);
            m0 += a0; m1 += a1; m2 += a2; m3 += a3; m4 += a4; m5 += a5;
        }
        const Sym3<R> M = {m0 + eps, m1, m2, m3 + eps, m4, m5 + eps};
        const Sym3<R> Mi = inv_sym3(M);
        const Vec3<R> y = mul(Mi, f);
        const R jy = dot(Jh, y);
        p = -(W * jy);
        if (DP) {
            // K y = sum_i U_i J_h,i (J_h,i . y): three sums instead of the six entries of K
            const R uj = U * jy;
            const Vec3<R> ky = gsum3(Vec3<R>{uj * Jh.x, uj * Jh.y, uj * Jh.z});
            const R re = N::rcp_(eps);
            const Vec3<R> z = mul(Mi, Vec3<R>{N::fma_(ky.x, re, y.x), N::fma_(ky.y, re, y.y), N::fma_(ky.z, re, y.z)});
            dp = N::fma_(W, dot(Jh, z), -((U * re) * jy));
        }
    }

    // common.py step_size_to_bound, own joint (inf where the step component is 0)
    __device__ __forceinline__ R to_bound(R xv, R s) const {
        const R rs = R(1) / s;
        R st = N::max_((lb - xv) * rs, (ub - xv) * rs);
        if (s == R(0) || !real) st = N::inf();
        return st;
    }
    // evaluate_quadratic of a step given by its own component
    __device__ __forceinline__ R quad(const Vec3<R>& Jh, R C, R gh, R s) const {
        const Vec3<R> js = gsum3(Vec3<R>{Jh.x * s, Jh.y * s, Jh.z * s});
        const R qc = gsum((s * C) * s), sg = gsum(s * gh);
        return N::fma_(R(0.5), dot(js, js) + qc, sg);
    }

    // trf.py select_step when x + p leaves the box (gen_select_general of the one-lane form)
    __device__ __forceinline__ R select_general(R d, R C, R gh, const Vec3<R>& Jh, R theta, R& p, R& ph) const {
        const R steps = to_bound(x, p);
        const R p_stride = gmin(steps);
        const bool hit = steps == p_stride && p != R(0);
        R r_h = hit ? -ph : ph;
        R r = d * r_h;
        p *= p_stride; ph *= p_stride;
        const R x_on = x + p;
        const R a = gsum(r_h * r_h), b = gsum(ph * r_h), pp = gsum(ph * ph);
        const R c = N::min_(pp - Delta * Delta, R(0));
        const R disc = N::sqrt_(N::max_(N::fma_(b, b, -(a * c)), R(0)));
        const R qq = -(b + N::copysign_(disc, b));
        R t1 = R(0), t2 = R(0);
        if (qq != R(0)) { t1 = qq / a; t2 = c / qq; }
        const R to_tr = N::max_(t1, t2);
        const R to_bd = gmin(to_bound(x_on, r));
        const R r_stride = N::min_(to_bd, to_tr);
        R r_l = R(0), r_u = R(-1);
        if (r_stride > R(0)) {
            r_l = (R(1) - theta) * p_stride / r_stride;
            r_u = (r_stride == to_bd) ? theta * to_bd : to_tr;
        }
        R r_value = N::inf();
        if (r_l <= r_u) {                                           // (group-uniform)
            const Vec3<R> v = gsum3(Vec3<R>{Jh.x * r_h, Jh.y * r_h, Jh.z * r_h});
            const Vec3<R> u = gsum3(Vec3<R>{Jh.x * ph, Jh.y * ph, Jh.z * ph});
            const Vec3<R> s3 = gsum3(Vec3<R>{(r_h * C) * r_h, (ph * C) * r_h, (ph * C) * ph});
            const R g_r = gsum(gh * r_h), g_p = gsum(gh * ph);
            const R aa = R(0.5) * (dot(v, v) + s3.x);
            const R bb = g_r + dot(u, v) + s3.y;
            const R cc = R(0.5) * dot(u, u) + g_p + R(0.5) * s3.z;
            R rs;
            gen_minq(aa, bb, r_l, r_u, cc, &rs, &r_value);
            r_h = N::fma_(r_h, rs, ph); r = r_h * d;
        }
        p *= theta; ph *= theta;
        const R p_value = quad(Jh, C, gh, ph);
        const R ag_h = -gh, ag = d * ag_h;
        const R to_tr2 = Delta * N::rsqrt_(gsum(ag_h * ag_h));
        const R to_bd2 = gmin(to_bound(x, ag));
        const R ag_hi = (to_bd2 < to_tr2) ? theta * to_bd2 : to_tr2;
        const Vec3<R> vg = gsum3(Vec3<R>{Jh.x * ag_h, Jh.y * ag_h, Jh.z * ag_h});
        const R sa = gsum((ag_h * C) * ag_h), sb = gsum(gh * ag_h);
        R ags, ag_value;
        gen_minq(R(0.5) * (dot(vg, vg) + sa), sb, R(0), ag_hi, R(0), &ags, &ag_value);
        if (p_value < r_value && p_value < ag_value) return -p_value;
        if (r_value < p_value && r_value < ag_value) { p = r; ph = r_h; return -r_value; }
        p = ag * ags; ph = ag_h * ags;
        return -ag_value;
    }

    // the step to evaluate next (trf.py:296-335); xt = own component of x + step, strictly feasible
    __device__ __forceinline__ void propose(R& xt, R& step_h_norm, R& step_norm, R& pred) {
        R v, dv;
        scaling(v, dv);
        const R g_norm = gmax(N::abs_(g * v));
        const R d = N::sqrt_(v), C = g * dv, gh = d * g;
        const Vec3<R> Jh = {J.x * d, J.y * d, J.z * d};
        const R theta = N::max_(R(0.995), R(1) - g_norm);
        const R gh_norm = N::sqrt_(gsum(gh * gh));
        // solve_lsq_trust_region, rank-deficient branch (common.py:132-166)
        const R rD = N::rcp_(Delta);
        R a_up = gh_norm * rD, a_lo = R(0);
        if (alpha == R(0)) alpha = R(0.001) * a_up;
        bool brk = false;
        R p = R(0), dp = R(0);
#pragma unroll 1
        for (int it = 0; it < 10; ++it) {
            if (!brk) {                                             // (group-uniform)
                if (alpha < a_lo || alpha > a_up) alpha = N::max_(R(0.001) * a_up, N::sqrt_(a_lo * a_up));
                tr_point<true>(Jh, C, alpha, p, dp);
                R s0 = p * p, s1 = p * dp;
#pragma unroll
                for (int m = 1; m < 8; m <<= 1) { const R b0 = xsh(s0, m), b1 = xsh(s1, m); s0 += b0; s1 += b1; }
                const R pn = N::sqrt_(s0);
                const R phi = pn - Delta;
                const R dphi = s1 * N::rcp_(pn);
                if (phi < R(0)) a_up = alpha;
                const R ratio = phi * N::rcp_(dphi);
                a_lo = N::max_(a_lo, alpha - ratio);
                alpha -= (phi + Delta) * ratio * rD;
                brk = N::abs_(phi) < R(0.01) * Delta;
            }
        }
        R ph, unused;
        tr_point<false>(Jh, C, alpha, ph, unused);
        const R scl = Delta * N::rsqrt_(gsum(ph * ph));
        ph *= scl; p = d * ph;
        const R xn = x + p;
        const bool inb = gall(!real || (xn >= lb && xn <= ub));
        if (inb) pred = -quad(Jh, C, gh, ph);
        else pred = select_general(d, C, gh, Jh, theta, p, ph);
        R n0 = ph * ph, n1 = p * p;
#pragma unroll
        for (int m = 1; m < 8; m <<= 1) { const R b0 = xsh(n0, m), b1 = xsh(n1, m); n0 += b0; n1 += b1; }
        step_h_norm = N::sqrt_(n0); step_norm = N::sqrt_(n1);
        // make_strictly_feasible(x + step, lb, ub, rstep = 0)
        R xv = x + p;
        if (xv <= lb) xv = G::next_(lb, ub);
        if (xv >= ub) xv = G::next_(ub, lb);
        xt = real ? xv : R(0);
    }

    // least_squares.py: x0 = make_strictly_feasible(x0, lb, ub) (rstep = 1e-10)
    __device__ __forceinline__ R feasible_seed() const {
        R xv = x;
        const R dl = xv - lb, du = ub - xv;
        if (dl <= N::min_(du, R(1e-10) * N::max_(R(1), N::abs_(lb)))) { xv = lb + R(1e-10) * N::max_(R(1), N::abs_(lb)); if (xv <= lb) xv = G::next_(lb, ub); }
        else if (du <= N::min_(dl, R(1e-10) * N::max_(R(1), N::abs_(ub)))) { xv = ub - R(1e-10) * N::max_(R(1), N::abs_(ub)); if (xv >= ub) xv = G::next_(ub, lb); }
        if (xv < lb || xv > ub) xv = R(0.5) * (lb + ub);
        return real ? xv : R(0);
    }

    // the chain at the angles whose own component is xa: every lane gathers the seven sin/cos and runs the (short) chain
    template <bool AX>
    __device__ __forceinline__ void chain(R xa, Vec3<R>* org, Vec3<R>* claw, Vec3<R>* ax) const {
        R s_own, c_own;
        G::sincos_(xa, &s_own, &c_own);
        R sn[GEN_DOF], cs[GEN_DOF];
#pragma unroll
        for (int i = 0; i < GEN_DOF; ++i) { sn[i] = __shfl_sync(gmask, s_own, gbase | i); cs[i] = __shfl_sync(gmask, c_own, gbase | i); }
        generic_chain<R, AX>(sn, cs, seg0, seg1, seg2, seg3, org, claw, ax);
    }

    // one function evaluation
    __device__ __forceinline__ void trip() {
        const bool first = nfev == 0;
        R xt, step_h_norm = R(0), step_norm = R(0), pred = R(0);
        if (first) xt = feasible_seed(); else propose(xt, step_h_norm, step_norm, pred);
        Vec3<R> org[3], claw, ax[GEN_DOF];
        chain<true>(xt, org, &claw, ax);
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
            R ratio = R(0);
            if (pred > R(0)) ratio = actual / pred; else if (pred == R(0) && actual == R(0)) ratio = R(1);
            R Delta_new = Delta;
            if (ratio < R(0.25)) Delta_new = R(0.25) * step_h_norm;
            else if (ratio > R(0.75) && step_h_norm > R(0.95) * Delta) Delta_new = R(2) * Delta;
            const R x_norm = N::sqrt_(null_sq + gsum(x * x));
            const bool ft = actual < R(1e-8) * cost && ratio > R(0.25);
            const bool xt_ = step_norm < R(1e-8) * (R(1e-8) + x_norm);
            if (ft && xt_) status = ST_BOTH; else if (ft) status = ST_FTOL; else if (xt_) status = ST_XTOL;
            if (status == ST_RUNNING) { alpha *= Delta / Delta_new; Delta = Delta_new; }
            accept = actual > R(0);
        }
        if (accept) {                                               // (group-uniform)
            cost = cost_new; f = fn; x = xt;
            // Jacobian column of the own joint = axis x (claw - origin of the joint); gradient component g = J . f
            Vec3<R> a = ax[0], o = {R(0), R(0), R(0)};
            if (gl == 1) a = ax[1];
            if (gl == 2) a = ax[2];
            if (gl == 3) { a = ax[3]; o = org[0]; }
            if (gl == 4) { a = ax[4]; o = org[0]; }
            if (gl == 5) { a = ax[5]; o = org[1]; }
            if (gl == 6) { a = ax[6]; o = org[2]; }
            const Vec3<R> lever = sub(claw, o);
            J = cross(a, lever);
            if (!real) J = {R(0), R(0), R(0)};
            g = dot(J, f);
        }
        R v, dv;
        scaling(v, dv);
        if (first) {                 // trf.py:234-238: Delta0 = |x0 / sqrt(v)| over ALL chain slots
            Delta = N::sqrt_(null_sq + gsum(real ? (x * x) * N::rcp_(v) : R(0)));
            if (Delta == R(0)) Delta = R(1);
        }
        // head of scipy's outer loop (trf.py:262-273): gtol test (it overrides ftol/xtol), evaluation limit
        if (accept || status != ST_RUNNING) {
            if (gmax(N::abs_(g * v)) < R(1e-8)) status = ST_GTOL;
        }
        if (status == ST_RUNNING && nfev >= 100 * GEN_N_FULL) status = ST_MAXFEV;
    }
};

#endif  // __CUDACC__

}  // namespace seqik
