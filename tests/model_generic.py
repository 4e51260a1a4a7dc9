"""numpy model of the device solver of the GENERIC 7-DOF leg IK ("TRF7").  TEST/DEV TOOL.

It is the executable specification of what ``csrc/seqik_generic.cuh`` computes: scipy's bounded
Trust-Region-Reflective iteration (scipy/optimize/_lsq/trf.py:206-413, common.py) on the chain of the reference's
``KinematicChainGeneric`` (seqikpy/kinematic_chain.py:424-532: Base, ThC_roll Z, ThC_yaw X, ThC_pitch Y, CTr_pitch Y,
CTr_roll Z, FTi_pitch Y, TiTa_pitch Y, Claw) with ONE target, the claw (leg_inverse_kinematics.py:406-613).

Differences from a literal transcription, none of which changes the iteration in exact arithmetic:
  * the Jacobian is analytic (axis x lever) instead of scipy's 2-point finite difference;
  * only the 7 joints are carried; the inert chain slots (Base link, Claw: zero Jacobian columns) enter through
    ``null_sq`` (their squared norm: initial trust radius trf.py:236, xtol test common.py:705-718) and max_nfev = 100*9;
  * the trust-region subproblem is solved without an SVD.  scipy takes the SVD of the (3+n) x n augmented matrix
    [J_h; diag(sqrt(C))] and evaluates p(alpha) = -V (s uf)/(s^2 + alpha), i.e. p(alpha) = -(J_h^T J_h + C + alpha I)^-1 g_h.
    With D = C + alpha I diagonal and positive this is  p = -D^-1 J_h^T (I + J_h D^-1 J_h^T)^-1 f  (a 3x3 solve), and
    its alpha-derivative has an equally cancellation-free form once D^-1 is normalised by its largest entry
    (``tr_point``), so float32 is enough on the device.  m = 3 < n always, hence solve_lsq_trust_region
    (common.py:57-168) always takes its rank-deficient branch: at most 10 safeguarded Newton iterations on alpha,
    the step rescaled to the trust radius, alpha carried between iterations.

``dtype=np.float32`` runs the same arithmetic in single precision (what the device does, up to fma contraction
and SFU reciprocals).  The product never imports this file.
"""
import numpy as np

N_DOF = 7
GENERIC_DOF_ORDER = ("ThC_roll", "ThC_yaw", "ThC_pitch", "CTr_pitch", "CTr_roll", "FTi_pitch", "TiTa_pitch")
_AXES = (2, 0, 1, 1, 2, 1, 1)            # rotation axis of each joint: Z X Y Y Z Y Y
_SEG_BEFORE = (-1, -1, -1, 0, -1, 1, 2)  # segment whose length offsets the joint along -z of the parent frame


def _rot(axis, t, dt):
    c, s = np.cos(t, dtype=dt), np.sin(t, dtype=dt)
    o, z = dt(1), dt(0)
    if axis == 0:
        return np.array([[o, z, z], [z, c, -s], [z, s, c]], dtype=dt)
    if axis == 1:
        return np.array([[c, z, s], [z, o, z], [-s, z, c]], dtype=dt)
    return np.array([[c, -s, z], [s, c, z], [z, z, o]], dtype=dt)


def chain_points(x, seg, dt=np.float64):
    """Joint origins (7,3) + claw (3,) + world joint axes (7,3) of the generic chain, ThC at 0."""
    R = np.eye(3, dtype=dt)
    o = np.zeros(3, dtype=dt)
    origins, axes = [], []
    for i in range(N_DOF):
        if _SEG_BEFORE[i] >= 0:
            o = o - R[:, 2] * dt(seg[_SEG_BEFORE[i]])
        origins.append(o.copy())
        axes.append(R[:, _AXES[i]].copy())
        R = (R @ _rot(_AXES[i], x[i], dt)).astype(dt)
    claw = o - R[:, 2] * dt(seg[3])
    return np.array(origins, dtype=dt), claw.astype(dt), np.array(axes, dtype=dt)


def residual_jacobian(x, seg, target, dt=np.float64):
    origins, claw, axes = chain_points(x, seg, dt)
    J = np.cross(axes, claw[None, :] - origins).T.astype(dt)          # (3, 7)
    return (claw - target).astype(dt), J


def fk_rows(x, seg, dt=np.float64):
    """The 9 rows ikpy's forward_kinematics(full_kinematics=True) yields for the generic chain (ThC at 0)."""
    origins, claw, _ = chain_points(x, seg, dt)
    return np.vstack([np.zeros((1, 3), dtype=dt), origins, claw[None]])


# ----------------------------------------------------------------------------------------------------------
def tr_point(Jh, f, C, alpha, dt):
    """p(alpha) and d p / d alpha of the regularised step, float32-safe (see the module docstring)."""
    D = C + alpha                                 # alpha (hence D) may be negative: scipy's last Newton update is not safeguarded
    eps = D[np.argmin(np.abs(D))]
    W = eps / D                                   # |W| <= 1, exactly 1 for the entry of smallest magnitude
    U = W * ((D - eps) / D)                       # W (1 - W) without cancellation
    Mh = (Jh * W) @ Jh.T + eps * np.eye(3, dtype=dt)
    Minv = np.linalg.inv(Mh.astype(np.float64)).astype(dt) if dt is np.float64 else _inv3(Mh, dt)
    y = Minv @ f
    jy = Jh.T @ y
    p = -W * jy
    K = (Jh * U) @ Jh.T
    z = Minv @ (y + (K @ y) / eps)
    dp = -(U / eps) * jy + W * (Jh.T @ z)
    return p.astype(dt), dp.astype(dt)


def _inv3(M, dt):
    """Symmetric 3x3 inverse by the adjugate (what the device does)."""
    a, b, c, d, e, f = M[0, 0], M[0, 1], M[0, 2], M[1, 1], M[1, 2], M[2, 2]
    c00, c01, c02 = d * f - e * e, c * e - b * f, b * e - c * d
    det = a * c00 + b * c01 + c * c02
    r = dt(1) / det
    c11, c12, c22 = a * f - c * c, b * c - a * e, a * d - b * b
    return (np.array([[c00, c01, c02], [c01, c11, c12], [c02, c12, c22]], dtype=dt) * r).astype(dt)


def _norm(v, dt):
    return np.sqrt(np.dot(v, v), dtype=dt)


def _to_bound(x, s, lb, ub):
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        steps = np.where(s != 0, np.maximum((lb - x) / s, (ub - x) / s), np.inf)
    m = steps.min()
    return m, (steps == m) * np.sign(s)


def _minq(a, b, lo, hi, c=0.0):
    ts = [lo, hi]
    if a != 0:
        ext = -0.5 * b / a
        if lo < ext < hi:
            ts.append(ext)
    ys = [t * (a * t + b) + c for t in ts]
    k = int(np.argmin(ys))
    return ts[k], ys[k]


def _quad(Jh, diag_h, g_h, s):
    js = Jh @ s
    return 0.5 * (np.dot(js, js) + np.dot(s * diag_h, s)) + np.dot(s, g_h)


def select_step(x, Jh, diag_h, g_h, p, p_h, d, Delta, lb, ub, theta, dt):
    """trf.py select_step: trust-region step, its reflection at the first bound hit, or the scaled gradient."""
    if np.all((x + p >= lb) & (x + p <= ub)):
        return p, p_h, -_quad(Jh, diag_h, g_h, p_h)
    p_stride, hits = _to_bound(x, p, lb, ub)
    r_h = np.where(hits != 0, -p_h, p_h)
    r = d * r_h
    p = p * p_stride
    p_h = p_h * p_stride
    x_on = x + p
    a = np.dot(r_h, r_h)
    b = np.dot(p_h, r_h)
    c = min(np.dot(p_h, p_h) - Delta * Delta, dt(0))
    disc = np.sqrt(max(b * b - a * c, dt(0)))
    q = -(b + np.copysign(disc, b))
    t1, t2 = (q / a, c / q) if q != 0 else (dt(0), dt(0))
    to_tr = max(t1, t2)
    to_bd, _ = _to_bound(x_on, r, lb, ub)
    r_stride = min(to_bd, to_tr)
    if r_stride > 0:
        r_l = (1 - theta) * p_stride / r_stride
        r_u = theta * to_bd if r_stride == to_bd else to_tr
    else:
        r_l, r_u = 0.0, -1.0
    if r_l <= r_u:
        v = Jh @ r_h
        u = Jh @ p_h
        aa = 0.5 * (np.dot(v, v) + np.dot(r_h * diag_h, r_h))
        bb = np.dot(g_h, r_h) + np.dot(u, v) + np.dot(p_h * diag_h, r_h)
        cc = 0.5 * np.dot(u, u) + np.dot(g_h, p_h) + 0.5 * np.dot(p_h * diag_h, p_h)
        rs, r_value = _minq(aa, bb, r_l, r_u, cc)
        r_h = r_h * rs + p_h
        r = r_h * d
    else:
        r_value = np.inf
    p = p * theta
    p_h = p_h * theta
    p_value = _quad(Jh, diag_h, g_h, p_h)
    ag_h = -g_h
    ag = d * ag_h
    to_tr = Delta / _norm(ag_h, dt)
    to_bd, _ = _to_bound(x, ag, lb, ub)
    ag_stride = theta * to_bd if to_bd < to_tr else to_tr
    v = Jh @ ag_h
    aa = 0.5 * (np.dot(v, v) + np.dot(ag_h * diag_h, ag_h))
    bb = np.dot(g_h, ag_h)
    ags, ag_value = _minq(aa, bb, 0.0, ag_stride)
    if p_value < r_value and p_value < ag_value:
        return p, p_h, -p_value
    if r_value < p_value and r_value < ag_value:
        return r, r_h, -r_value
    return ag * ags, ag_h * ags, -ag_value


def trf7(seg, target, x0, lb, ub, null_sq=0.0, n_full=9, ftol=1e-8, xtol=1e-8, gtol=1e-8, dtype=np.float64,
         trace=None):
    """One bounded solve.  x0, lb, ub: the 7 joints in GENERIC_DOF_ORDER.  Returns (x, status, nfev, cost)."""
    dt = dtype
    lb = np.asarray(lb, dtype=dt)
    ub = np.asarray(ub, dtype=dt)
    x = np.asarray(x0, dtype=dt).copy()
    target = np.asarray(target, dtype=dt)
    null_sq = dt(null_sq)
    # least_squares.py: x0 = make_strictly_feasible(x0, lb, ub)   (rstep = 1e-10)
    lo = (x - lb) <= np.minimum(ub - x, dt(1e-10) * np.maximum(1, np.abs(lb)))
    hi = (ub - x) <= np.minimum(x - lb, dt(1e-10) * np.maximum(1, np.abs(ub)))
    x[lo] = (lb + dt(1e-10) * np.maximum(1, np.abs(lb)))[lo]
    x[hi] = (ub - dt(1e-10) * np.maximum(1, np.abs(ub)))[hi]
    tight = (x < lb) | (x > ub)
    x[tight] = (0.5 * (lb + ub))[tight]

    def cl(x, g):
        v = np.ones(N_DOF, dtype=dt)
        dv = np.zeros(N_DOF, dtype=dt)
        m = g < 0
        v[m] = (ub - x)[m]
        dv[m] = -1
        m = g > 0
        v[m] = (x - lb)[m]
        dv[m] = 1
        return v, dv

    f, J = residual_jacobian(x, seg, target, dt)
    nfev = 1
    cost = dt(0.5) * np.dot(f, f)
    g = J.T @ f
    v, dv = cl(x, g)
    Delta = np.sqrt(null_sq + np.sum(x * x / v), dtype=dt)
    if Delta == 0:
        Delta = dt(1)
    max_nfev = 100 * n_full
    alpha = dt(0)
    status = None
    while True:
        v, dv = cl(x, g)
        g_norm = np.max(np.abs(g * v))
        if g_norm < gtol:
            status = 1
        if status is not None or nfev == max_nfev:
            break
        d = np.sqrt(v)
        diag_h = g * dv                       # C >= 0
        g_h = d * g
        Jh = J * d
        theta = max(dt(0.995), dt(1) - g_norm)
        gh_norm = _norm(g_h, dt)
        actual = dt(-1)
        while actual <= 0 and nfev < max_nfev:
            # ---- solve_lsq_trust_region, rank-deficient branch
            a_up = gh_norm / Delta
            a_lo = dt(0)
            if alpha == 0:
                alpha = dt(0.001) * a_up
            n_it = 0
            for _ in range(10):
                if alpha < a_lo or alpha > a_up:
                    alpha = max(dt(0.001) * a_up, np.sqrt(a_lo * a_up, dtype=dt))
                p, dp = tr_point(Jh, f, diag_h, alpha, dt)
                pn = _norm(p, dt)
                phi = pn - Delta
                dphi = np.dot(p, dp) / pn
                if phi < 0:
                    a_up = alpha
                ratio = phi / dphi
                a_lo = max(a_lo, alpha - ratio)
                alpha = alpha - (phi + Delta) * ratio / Delta
                n_it += 1
                if abs(phi) < dt(0.01) * Delta:
                    break
            p_h, _ = tr_point(Jh, f, diag_h, alpha, dt)
            p_h = p_h * (Delta / _norm(p_h, dt))
            p = d * p_h
            step, step_h, pred = select_step(x, Jh, diag_h, g_h, p, p_h, d, Delta, lb, ub, theta, dt)
            x_new = (x + step).astype(dt)
            m = x_new <= lb
            x_new[m] = np.nextafter(lb, ub)[m]
            m = x_new >= ub
            x_new[m] = np.nextafter(ub, lb)[m]
            f_new, J_new = residual_jacobian(x_new, seg, target, dt)
            nfev += 1
            step_h_norm = _norm(step_h, dt)
            cost_new = dt(0.5) * np.dot(f_new, f_new)
            actual = cost - cost_new
            if pred > 0:
                ratio = actual / pred
            elif pred == actual == 0:
                ratio = dt(1)
            else:
                ratio = dt(0)
            Delta_new = Delta
            if ratio < 0.25:
                Delta_new = dt(0.25) * step_h_norm
            elif ratio > 0.75 and step_h_norm > dt(0.95) * Delta:
                Delta_new = dt(2) * Delta
            step_norm = _norm(step, dt)
            x_norm = np.sqrt(null_sq + np.dot(x, x), dtype=dt)
            ft = actual < ftol * cost and ratio > 0.25
            xt = step_norm < xtol * (xtol + x_norm)
            if trace is not None:
                trace.append(dict(nfev=nfev, cost=float(cost), cost_new=float(cost_new), Delta=float(Delta),
                                  alpha=float(alpha), ratio=float(ratio), n_it=n_it, step_norm=float(step_norm)))
            if ft and xt:
                status = 4
            elif ft:
                status = 2
            elif xt:
                status = 3
            if status is not None:
                break
            alpha = alpha * (Delta / Delta_new)
            Delta = Delta_new
        if actual > 0:
            x, f, J, cost = x_new, f_new, J_new, cost_new
            g = J.T @ f
    if status is None:
        status = 0
    return x, status, nfev, float(cost)


def solve_leg_generic(pose, seg, lb, ub, seed, null_sq=0.0, dtype=np.float64, stats=None):
    """Warm-started frame loop (leg_inverse_kinematics.py:521-533): pose (N,>=2,3) with the ThC at row 0 and the
    claw at row -1 -> angles (N,7) in GENERIC_DOF_ORDER, fk (N,9,3)."""
    n = pose.shape[0]
    ang = np.zeros((n, N_DOF))
    fk = np.zeros((n, 9, 3))
    x = np.asarray(seed, dtype=float)
    for t in range(n):
        x, st, nf, cost = trf7(seg, pose[t, -1] - pose[t, 0], x, lb, ub, null_sq, dtype=dtype)
        ang[t] = x
        fk[t] = fk_rows(x, seg) + pose[t, 0]
        if stats is not None:
            stats.append((t, st, nf, cost))
    return ang, fk
