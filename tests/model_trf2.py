"""Scalar float64 Python model of the device solver ("TRF2").  TEST/DEV TOOL.

It is the executable specification of what ``csrc/seqik_core.cuh`` computes:
scipy's bounded Trust-Region-Reflective iteration
(scipy/optimize/_lsq/trf.py:206-413, common.py) restricted to the <=2 ACTIVE
variables of a SeqIKPy stage, written in the pivot frame of the stage
(SURVEY.md 3.4), with the frozen/null chain slots folded into two scalars
(their squared norm enters the initial trust radius and the xtol test).

It is used by CPU tests to check the algorithm against the oracle without a
GPU; the product never imports it.
"""
import math

EPS = 2.220446049250313e-16


def stage_w(kind, L, a, b):
    """End point of the solved segment in the pivot frame + its 2 Jacobian columns.

    kind 0: Rx(a) Ry(b) (0,0,-L)   (stage 1: yaw, pitch)
    kind 1: Rz(a) Ry(b) (0,0,-L)   (stages 2, 3: roll, pitch)
    kind 2:       Ry(b) (0,0,-L)   (stage 4: pitch only; column a is zero)
    """
    sa, ca, sb, cb = math.sin(a), math.cos(a), math.sin(b), math.cos(b)
    if kind == 0:
        w = (-L * sb, L * cb * sa, -L * cb * ca)
        ja = (0.0, L * cb * ca, L * cb * sa)
        jb = (-L * cb, -L * sb * sa, L * sb * ca)
    elif kind == 1:
        w = (-L * sb * ca, -L * sb * sa, -L * cb)
        ja = (L * sb * sa, -L * sb * ca, 0.0)
        jb = (-L * cb * ca, -L * cb * sa, L * sb)
    else:
        w = (-L * sb, 0.0, -L * cb)
        ja = (0.0, 0.0, 0.0)
        jb = (-L * cb, 0.0, L * sb)
    return w, ja, jb


def _nextafter(x, toward):
    return math.nextafter(x, toward)


def trf2(kind, L, q, x0, lb, ub, null_sq, n_full, ftol=1e-8, xtol=1e-8, gtol=1e-8, trace=None):
    """Bounded TRF on the active pair x=(a,b).  For kind 2 the slot `a` is inert
    (lb=ub=x0 is not required: its Jacobian column is zero, it behaves like a null slot,
    so callers pass its value through null_sq instead and set na=1)."""
    na = 1 if kind == 2 else 2
    idx = (1,) if kind == 2 else (0, 1)
    x = list(x0)
    # least_squares: make_strictly_feasible(x0, lb, ub)  (rstep=1e-10)
    for i in idx:
        if x[i] <= lb[i]:  # on lower bound
            x[i] = lb[i] + 1e-10 * max(1.0, abs(lb[i]))
        if x[i] >= ub[i]:
            x[i] = ub[i] - 1e-10 * max(1.0, abs(ub[i]))
        if not (lb[i] < x[i] < ub[i]):
            x[i] = 0.5 * (lb[i] + ub[i])

    def fun(xx):
        w, ja, jb = stage_w(kind, L, xx[0], xx[1])
        f = (w[0] - q[0], w[1] - q[1], w[2] - q[2])
        return f, (ja, jb)

    def dot3(u, v):
        return u[0] * v[0] + u[1] * v[1] + u[2] * v[2]

    f, J = fun(x)
    nfev = 1
    cost = 0.5 * dot3(f, f)
    g = [dot3(J[0], f), dot3(J[1], f)]
    if kind == 2:
        g[0] = 0.0

    def cl(xx, gg):
        v = [1.0, 1.0]
        dv = [0.0, 0.0]
        for i in idx:
            if gg[i] < 0 and math.isfinite(ub[i]):
                v[i] = ub[i] - xx[i]
                dv[i] = -1.0
            elif gg[i] > 0 and math.isfinite(lb[i]):
                v[i] = xx[i] - lb[i]
                dv[i] = 1.0
        return v, dv

    v, dv = cl(x, g)
    Delta = math.sqrt(null_sq + sum(x[i] * x[i] / v[i] for i in idx))
    if Delta == 0:
        Delta = 1.0
    max_nfev = 100 * n_full
    alpha = 0.0
    status = None
    nit = 0
    while True:
        v, dv = cl(x, g)
        g_norm = max(abs(g[i] * v[i]) for i in idx)
        if g_norm < gtol:
            status = 1
        if status is not None or nfev == max_nfev:
            break
        d = [math.sqrt(v[0]), math.sqrt(v[1])]
        diag_h = [g[0] * dv[0], g[1] * dv[1]]
        g_h = [d[0] * g[0], d[1] * g[1]]
        Jh = ([J[0][k] * d[0] for k in range(3)], [J[1][k] * d[1] for k in range(3)])
        if kind == 2:
            Jh = ([0.0, 0.0, 0.0], Jh[1])
            g_h[0] = 0.0
        # B = Jh^T Jh + diag_h  (2x2 symmetric);  eigen-decomposition = SVD of the augmented matrix
        b00 = dot3(Jh[0], Jh[0]) + diag_h[0]
        b01 = dot3(Jh[0], Jh[1])
        b11 = dot3(Jh[1], Jh[1]) + diag_h[1]
        if kind == 2:
            lam = [b11, 0.0]
            V = ((0.0, 1.0), (1.0, 0.0))  # columns: V[:,0]=(0,1), V[:,1]=(1,0)
        else:
            tr = 0.5 * (b00 + b11)
            df = 0.5 * (b00 - b11)
            rad = math.hypot(df, b01)
            l1 = tr + rad
            l2 = tr - rad
            # stable smaller eigenvalue
            if l1 != 0:
                l2 = (b00 * b11 - b01 * b01) / l1
            lam = [l1, max(l2, 0.0)]
            # eigenvector for l1
            if df >= 0:
                ex, ey = df + rad, b01
            else:
                ex, ey = b01, rad - df
            nrm = math.hypot(ex, ey)
            if nrm == 0:
                ex, ey = 1.0, 0.0
            else:
                ex, ey = ex / nrm, ey / nrm
            V = ((ex, -ey), (ey, ex))  # V[row][col]
        # suf = s * uf = V^T g_h
        suf = [V[0][0] * g_h[0] + V[1][0] * g_h[1], V[0][1] * g_h[0] + V[1][1] * g_h[1]]
        theta = max(0.995, 1 - g_norm)
        actual = -1.0
        while actual <= 0 and nfev < max_nfev:
            # ---- solve_lsq_trust_region, rank-deficient branch (m=3 < n_full always)
            def phi_dphi(al):
                t0 = suf[0] / (lam[0] + al) if (lam[0] + al) != 0 else 0.0
                t1 = suf[1] / (lam[1] + al) if (lam[1] + al) != 0 else 0.0
                pn = math.hypot(t0, t1)
                phi = pn - Delta
                dd = 0.0
                if (lam[0] + al) != 0:
                    dd += suf[0] ** 2 / (lam[0] + al) ** 3
                if (lam[1] + al) != 0:
                    dd += suf[1] ** 2 / (lam[1] + al) ** 3
                return phi, (-dd / pn if pn != 0 else -math.inf)
            a_up = math.hypot(suf[0], suf[1]) / Delta
            a_lo = 0.0
            if alpha == 0:
                alpha = max(0.001 * a_up, math.sqrt(a_lo * a_up))
            n_it = 0
            for it in range(10):
                if alpha < a_lo or alpha > a_up:
                    alpha = max(0.001 * a_up, math.sqrt(a_lo * a_up))
                phi, dphi = phi_dphi(alpha)
                if phi < 0:
                    a_up = alpha
                ratio = phi / dphi
                a_lo = max(a_lo, alpha - ratio)
                alpha -= (phi + Delta) * ratio / Delta
                n_it += 1
                if abs(phi) < 0.01 * Delta:
                    break
            t0 = suf[0] / (lam[0] + alpha) if (lam[0] + alpha) != 0 else 0.0
            t1 = suf[1] / (lam[1] + alpha) if (lam[1] + alpha) != 0 else 0.0
            ph = [-(V[0][0] * t0 + V[0][1] * t1), -(V[1][0] * t0 + V[1][1] * t1)]
            pn = math.hypot(ph[0], ph[1])
            ph = [ph[0] * Delta / pn, ph[1] * Delta / pn]
            p = [d[0] * ph[0], d[1] * ph[1]]
            if kind == 2:
                ph[0] = 0.0
                p[0] = 0.0
            step, step_h, pred = _select_step(x, Jh, diag_h, g_h, p, ph, d, Delta, lb, ub, theta, idx)
            x_new = list(x)
            for i in idx:
                xi = x[i] + step[i]
                if xi <= lb[i]:
                    xi = _nextafter(lb[i], ub[i])
                if xi >= ub[i]:
                    xi = _nextafter(ub[i], lb[i])
                x_new[i] = xi
            f_new, J_new = fun(x_new)
            nfev += 1
            step_h_norm = math.hypot(step_h[0], step_h[1])
            cost_new = 0.5 * dot3(f_new, f_new)
            actual = cost - cost_new
            # update_tr_radius
            if pred > 0:
                ratio = actual / pred
            elif pred == actual == 0:
                ratio = 1.0
            else:
                ratio = 0.0
            Delta_new = Delta
            if ratio < 0.25:
                Delta_new = 0.25 * step_h_norm
            elif ratio > 0.75 and step_h_norm > 0.95 * Delta:
                Delta_new = 2.0 * Delta
            step_norm = math.hypot(step[0], step[1])
            x_norm = math.sqrt(null_sq + sum(x[i] * x[i] for i in idx))
            ft = actual < ftol * cost and ratio > 0.25
            xt = step_norm < xtol * (xtol + x_norm)
            if trace is not None:
                trace.append(dict(nit=nit, nfev=nfev, x=tuple(x_new), cost=cost, cost_new=cost_new,
                                  Delta=Delta, alpha=alpha, ratio=ratio, pred=pred, n_it=n_it))
            if ft and xt:
                status = 4
            elif ft:
                status = 2
            elif xt:
                status = 3
            if status is not None:
                break
            alpha *= Delta / Delta_new
            Delta = Delta_new
        if actual > 0:
            x, f, J, cost = x_new, f_new, J_new, cost_new
            g = [dot3(J[0], f), dot3(J[1], f)]
            if kind == 2:
                g[0] = 0.0
        nit += 1
    if status is None:
        status = 0
    return x, status, nfev, cost


def _select_step(x, Jh, diag_h, g_h, p, ph, d, Delta, lb, ub, theta, idx):
    def dot3(u, v):
        return u[0] * v[0] + u[1] * v[1] + u[2] * v[2]

    def Jdot(s):
        return [Jh[0][k] * s[0] + Jh[1][k] * s[1] for k in range(3)]

    def evalq(s):
        js = Jdot(s)
        return 0.5 * (dot3(js, js) + s[0] * diag_h[0] * s[0] + s[1] * diag_h[1] * s[1]) + s[0] * g_h[0] + s[1] * g_h[1]

    def to_bound(xx, pp):
        best = math.inf
        hits = [0, 0]
        steps = [math.inf, math.inf]
        for i in idx:
            if pp[i] != 0:
                steps[i] = max((lb[i] - xx[i]) / pp[i], (ub[i] - xx[i]) / pp[i])
        best = min(steps)
        for i in idx:
            if steps[i] == best and pp[i] != 0:
                hits[i] = 1 if pp[i] > 0 else -1
        return best, hits

    inb = all(lb[i] <= x[i] + p[i] <= ub[i] for i in idx)
    if inb:
        return list(p), list(ph), -evalq(ph)
    p_stride, hits = to_bound(x, p)
    r_h = [(-ph[i] if hits[i] else ph[i]) for i in range(2)]
    r = [d[0] * r_h[0], d[1] * r_h[1]]
    p = [p[0] * p_stride, p[1] * p_stride]
    ph = [ph[0] * p_stride, ph[1] * p_stride]
    x_on = [x[0] + p[0], x[1] + p[1]]
    # intersect_trust_region(ph, r_h, Delta) -> positive root
    a = r_h[0] ** 2 + r_h[1] ** 2
    b = ph[0] * r_h[0] + ph[1] * r_h[1]
    c = ph[0] ** 2 + ph[1] ** 2 - Delta ** 2
    c = min(c, 0.0)
    disc = math.sqrt(max(b * b - a * c, 0.0))
    qq = -(b + math.copysign(disc, b))
    if qq != 0:
        t1, t2 = qq / a, c / qq
    else:
        t1, t2 = 0.0, 0.0
    to_tr = max(t1, t2)
    to_bd, _ = to_bound(x_on, r)
    r_stride = min(to_bd, to_tr)
    if r_stride > 0:
        r_l = (1 - theta) * p_stride / r_stride
        r_u = theta * to_bd if r_stride == to_bd else to_tr
    else:
        r_l, r_u = 0.0, -1.0
    if r_l <= r_u:
        # build_quadratic_1d(Jh, g_h, r_h, s0=ph, diag=diag_h)
        vv = Jdot(r_h)
        aa = 0.5 * (dot3(vv, vv) + r_h[0] * diag_h[0] * r_h[0] + r_h[1] * diag_h[1] * r_h[1])
        bb = g_h[0] * r_h[0] + g_h[1] * r_h[1]
        uu = Jdot(ph)
        bb += dot3(uu, vv)
        cc = 0.5 * dot3(uu, uu) + g_h[0] * ph[0] + g_h[1] * ph[1]
        bb += ph[0] * diag_h[0] * r_h[0] + ph[1] * diag_h[1] * r_h[1]
        cc += 0.5 * (ph[0] * diag_h[0] * ph[0] + ph[1] * diag_h[1] * ph[1])
        rs, r_value = _minq(aa, bb, r_l, r_u, cc)
        r_h = [r_h[0] * rs + ph[0], r_h[1] * rs + ph[1]]
        r = [r_h[0] * d[0], r_h[1] * d[1]]
    else:
        r_value = math.inf
    p = [p[0] * theta, p[1] * theta]
    ph = [ph[0] * theta, ph[1] * theta]
    p_value = evalq(ph)
    ag_h = [-g_h[0], -g_h[1]]
    ag = [d[0] * ag_h[0], d[1] * ag_h[1]]
    to_tr = Delta / math.hypot(ag_h[0], ag_h[1])
    to_bd, _ = to_bound(x, ag)
    ag_stride = theta * to_bd if to_bd < to_tr else to_tr
    vv = Jdot(ag_h)
    aa = 0.5 * (dot3(vv, vv) + ag_h[0] * diag_h[0] * ag_h[0] + ag_h[1] * diag_h[1] * ag_h[1])
    bb = g_h[0] * ag_h[0] + g_h[1] * ag_h[1]
    ags, ag_value = _minq(aa, bb, 0.0, ag_stride, 0.0)
    ag_h = [ag_h[0] * ags, ag_h[1] * ags]
    ag = [ag[0] * ags, ag[1] * ags]
    if p_value < r_value and p_value < ag_value:
        return p, ph, -p_value
    if r_value < p_value and r_value < ag_value:
        return r, r_h, -r_value
    return ag, ag_h, -ag_value


def _minq(a, b, lo, hi, c):
    ts = [lo, hi]
    if a != 0:
        ext = -0.5 * b / a
        if lo < ext < hi:
            ts.append(ext)
    best_t, best_y = None, None
    for t in ts:
        y = t * (a * t + b) + c
        if best_y is None or y < best_y:
            best_t, best_y = t, y
    return best_t, best_y


# ----------------------------------------------------------------------------
# chain driver: the 4 stages of one leg over N frames, frame-major
# ----------------------------------------------------------------------------
def _rot(axis, t):
    c, s = math.cos(t), math.sin(t)
    if axis == 0:
        return ((1, 0, 0), (0, c, -s), (0, s, c))
    if axis == 1:
        return ((c, 0, s), (0, 1, 0), (-s, 0, c))
    return ((c, -s, 0), (s, c, 0), (0, 0, 1))


def _mm(A, B):
    return tuple(tuple(sum(A[i][k] * B[k][j] for k in range(3)) for j in range(3)) for i in range(3))


def _col2(A):
    return (A[0][2], A[1][2], A[2][2])


def solve_leg(pose, seg_len, lb7, ub7, seed7, null_sq4, solver=trf2, stats=None):
    """pose (N,5,3) -> angles (N,7), fk (N,9,3).  DOF order = oracle.DOF_ORDER."""
    import numpy as np
    n = pose.shape[0]
    ang = np.zeros((n, 7))
    fk = np.zeros((n, 9, 3))
    prev = list(seed7)
    cx, fe, ti, ta = seg_len
    nfull = (4, 6, 8, 9)
    for t in range(n):
        o = pose[t, 0]
        tgt = [pose[t, s] - o for s in range(1, 5)]
        # stage 1: A = I, pivot = 0
        (y, p), st, nf, _ = solver(0, cx, tuple(tgt[0]), (prev[0], prev[1]), lb7[0:2], ub7[0:2], null_sq4[0], nfull[0])
        if stats is not None:
            stats.append((1, t, st, nf))
        A = _mm(_rot(0, y), _rot(1, p))
        z = _col2(A)
        piv = (-cx * z[0], -cx * z[1], -cx * z[2])
        q = tuple(sum(A[k][i] * (tgt[1][k] - piv[k]) for k in range(3)) for i in range(3))
        (r, cp), st, nf, _ = solver(1, fe, q, (prev[2], prev[3]), lb7[2:4], ub7[2:4], null_sq4[1], nfull[1])
        if stats is not None:
            stats.append((2, t, st, nf))
        A = _mm(_mm(A, _rot(2, r)), _rot(1, cp))
        z = _col2(A)
        piv2 = (piv[0] - fe * z[0], piv[1] - fe * z[1], piv[2] - fe * z[2])
        q = tuple(sum(A[k][i] * (tgt[2][k] - piv2[k]) for k in range(3)) for i in range(3))
        (cr, fp), st, nf, _ = solver(1, ti, q, (prev[4], prev[5]), lb7[4:6], ub7[4:6], null_sq4[2], nfull[2])
        if stats is not None:
            stats.append((3, t, st, nf))
        A = _mm(_mm(A, _rot(2, cr)), _rot(1, fp))
        z = _col2(A)
        piv3 = (piv2[0] - ti * z[0], piv2[1] - ti * z[1], piv2[2] - ti * z[2])
        q = tuple(sum(A[k][i] * (tgt[3][k] - piv3[k]) for k in range(3)) for i in range(3))
        (_, tp), st, nf, _ = solver(2, ta, q, (0.0, prev[6]), (0.0, lb7[6]), (0.0, ub7[6]), null_sq4[3], nfull[3])
        if stats is not None:
            stats.append((4, t, st, nf))
        A = _mm(A, _rot(1, tp))
        z = _col2(A)
        piv4 = (piv3[0] - ta * z[0], piv3[1] - ta * z[1], piv3[2] - ta * z[2])
        prev = [y, p, r, cp, cr, fp, tp]
        ang[t] = prev
        fk[t, 4] = fk[t, 5] = piv
        fk[t, 6] = piv2
        fk[t, 7] = piv3
        fk[t, 8] = piv4
        fk[t] += o
    return ang, fk


def null_sq_from_seeds(init):
    """Squared norm of the frozen/null slots of the 4 stage seed vectors
    (chain order): everything except the stage's active slots."""
    act = {1: (1, 2), 2: (3, 4), 3: (5, 6), 4: (7,)}
    out = []
    for s in (1, 2, 3, 4):
        v = init[f"stage_{s}"]
        out.append(float(sum(float(v[i]) ** 2 for i in range(len(v)) if i not in act[s])))
    return out


def seeds7(init):
    return [float(init["stage_1"][1]), float(init["stage_1"][2]), float(init["stage_2"][3]),
            float(init["stage_2"][4]), float(init["stage_3"][5]), float(init["stage_3"][6]),
            float(init["stage_4"][7])]
