"""ORACLE (test infrastructure, NOT product code) - exact float64 solver for the MPC QP.

    min 1/2 u'Hu + g'u   s.t.  0 <= C u <= ub          (drake branch, mpc.py:277-286)

with C = kron(I_4H, pyramid(mu)) and ub = (+inf x4, gait*fz_max) per foot-step
(mpc.py:237-260).  The reference delegates this to Drake 1.15.0 `Solve()` (OSQP with
polish; requirements.txt:23,51) which is not available offline.  H >= 2R > 0 makes
the optimum unique, so any exact method is a valid oracle.  This one is deliberately
generic and independent of the CUDA solver's structure-exploiting tricks:

  1. swing feet (ub_fz <= 0) are pinned to zero (0 <= fz <= 0 and |fx|,|fy| <= mu fz);
  2. a dense Mehrotra predictor-corrector interior-point method on A u <= b to ~1e-12;
  3. active rows are read off the complementarity split (slack < multiplier), the
     equality-constrained QP on that set is solved through a null-space basis from an
     SVD (rank-deficient active sets - the pyramid apex - are fine), and the KKT
     conditions are CHECKED: primal feasibility plus a non-negative least-squares fit
     of the multipliers; rows are added / dropped until the check passes.

`QPSolution.verified` is only True when max KKT violation <= `kkt_tol`.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
from scipy.linalg import cho_factor, cho_solve, null_space
from scipy.optimize import nnls


@dataclass
class QPSolution:
    u: np.ndarray                    # [12H] optimum (swing entries exactly 0)
    y: np.ndarray                    # [20H] multipliers on the reference rows: Hu+g+C'y=0, y<=0 @lower, y>=0 @upper
    active_lower: np.ndarray         # [20H] bool, (Cu)_r == 0 enforced
    active_upper: np.ndarray         # [20H] bool, (Cu)_r == ub_r enforced (fz rows only)
    kkt_stationarity: float
    kkt_primal: float
    ipm_iterations: int
    polish_rounds: int
    verified: bool
    info: dict = field(default_factory=dict)


def _stance_rows(mu: float, ub_fz: np.ndarray):
    """Inequalities A u_s <= b on the stance variables; 6 rows per stance foot-step.

    Row order per foot: fx>=-mu fz, fx<=mu fz, fy>=-mu fz, fy<=mu fz, fz>=0, fz<=fmax,
    i.e. the reference's 5 two-sided rows split into one-sided ones (rows 0-4 lower,
    row 5 = upper of reference row 4)."""
    stance = np.flatnonzero(ub_fz > 0)
    ns = len(stance)
    A = np.zeros((6 * ns, 3 * ns))
    b = np.zeros(6 * ns)
    for p in range(ns):
        c = 3 * p
        A[6 * p + 0, [c, c + 2]] = (-1, -mu)
        A[6 * p + 1, [c, c + 2]] = (1, -mu)
        A[6 * p + 2, [c + 1, c + 2]] = (-1, -mu)
        A[6 * p + 3, [c + 1, c + 2]] = (1, -mu)
        A[6 * p + 4, c + 2] = -1
        A[6 * p + 5, c + 2] = 1
        b[6 * p + 5] = ub_fz[stance[p]]
    return stance, A, b


def ipm_qp(H, g, A, b, u0, tol=1e-10, max_iter=100):
    """Feasible-start Mehrotra predictor-corrector for min 1/2u'Hu+g'u, Au<=b (float64)."""
    u = u0.copy()
    s = b - A @ u
    assert np.all(s > 0), "interior start required"
    m = len(b)
    lam = np.maximum(1.0 / s, 1e-3)
    scale = 1.0 + np.abs(g).max()
    it = 0
    for it in range(1, max_iter + 1):
        grad = H @ u + g
        rd = grad + A.T @ lam
        mu_c = s @ lam / m
        if mu_c <= tol * scale and np.abs(rd).max() <= tol * scale * 10:
            break
        d = lam / s
        K = H + (A.T * d) @ A
        try:
            cf = cho_factor(K)
        except np.linalg.LinAlgError:                      # barrier terms swamped H: hand over to the polish
            break
        # predictor
        du = cho_solve(cf, -grad)
        ds = -A @ du
        dlam = -lam - d * ds
        a_p = _max_step(s, ds)
        a_d = _max_step(lam, dlam)
        mu_aff = (s + a_p * ds) @ (lam + a_d * dlam) / m
        sigma = (mu_aff / mu_c) ** 3
        # corrector
        corr = (sigma * mu_c - ds * dlam) / s
        du = cho_solve(cf, -grad - A.T @ corr)
        ds = -A @ du
        dlam = corr - lam - d * ds
        a_p = min(1.0, 0.995 * _max_step(s, ds, cap=np.inf))
        a_d = min(1.0, 0.995 * _max_step(lam, dlam, cap=np.inf))
        u = u + a_p * du
        s = b - A @ u
        lam = lam + a_d * dlam
    return u, s, lam, it


def _max_step(v, dv, cap=1.0):
    neg = dv < 0
    if not np.any(neg):
        return cap
    return min(cap, float(np.min(-v[neg] / dv[neg])))


def _polish(H, g, A, b, act):
    """Minimise on {A_act u = b_act} via an SVD null-space basis (rank-deficient safe)."""
    n = H.shape[0]
    if not np.any(act):
        return np.linalg.solve(H, -g)
    Aa, ba = A[act], b[act]
    u_p = np.linalg.lstsq(Aa, ba, rcond=None)[0]
    Z = null_space(Aa, rcond=1e-12)
    if Z.shape[1] == 0:
        return u_p
    w = np.linalg.solve(Z.T @ H @ Z, -Z.T @ (g + H @ u_p))
    return u_p + Z @ w


def solve_qp_exact(H, g, mu, ub_fz, kkt_tol=1e-8, feas_tol=1e-9, max_rounds=60) -> QPSolution:
    mu = float(np.float32(mu))                               # the reference's C is float32 (mpc.py:239-245)
    H = np.asarray(H, dtype=np.float64)
    g = np.asarray(g, dtype=np.float64)
    ub_fz = np.asarray(ub_fz, dtype=np.float64)
    nfeet = len(ub_fz)
    n, m = 3 * nfeet, 5 * nfeet
    u_full = np.zeros(n)
    y_full = np.zeros(m)
    act_lo = np.zeros(m, dtype=bool)
    act_up = np.zeros(m, dtype=bool)
    stance, A, b = _stance_rows(mu, ub_fz)
    swing = np.setdiff1d(np.arange(nfeet), stance)
    act_lo.reshape(nfeet, 5)[swing, :] = True           # swing: every row sits on its (zero) bound
    act_up.reshape(nfeet, 5)[swing, 4] = True
    ns = len(stance)
    if ns == 0:
        return QPSolution(u_full, _swing_multipliers(y_full, g, H, u_full, mu, swing), act_lo, act_up,
                          0.0, 0.0, 0, 0, True)
    var = (3 * stance[:, None] + np.arange(3)[None, :]).ravel()
    Hs, gs = H[np.ix_(var, var)], g[var]

    u0 = np.zeros(3 * ns)
    u0[2::3] = np.minimum(0.5 * ub_fz[stance], 10.0)
    u, s, lam, iters = ipm_qp(Hs, gs, A, b, u0)
    act = s < lam

    verified, rounds = False, 0
    stat = prim = np.inf
    yA = np.zeros(len(b))
    seen = set()
    for rounds in range(1, max_rounds + 1):
        u = _polish(Hs, gs, A, b, act)
        slack = b - A @ u
        grad = Hs @ u + gs
        viol = slack < -feas_tol
        yA[:] = 0.0
        if np.any(act):
            ya, _ = nnls(A[act].T, -grad, maxiter=50 * int(act.sum()) + 100)
            yA[act] = ya
        stat = float(np.abs(grad + A.T @ yA).max())
        prim = float(max(0.0, -slack.min()))
        if not np.any(viol) and stat <= kkt_tol:
            verified = True
            break
        key = act.tobytes()
        if key in seen:                                    # cycling guard: fall back to one-at-a-time
            pass
        seen.add(key)
        if np.any(viol):
            act = act.copy()
            act[np.argmin(slack)] = True
        else:
            yls = np.linalg.lstsq(A[act].T, -grad, rcond=None)[0]
            rows = np.flatnonzero(act)
            act = act.copy()
            act[rows[np.argmin(yls)]] = False

    u_full[var] = u
    # map one-sided stance rows back to the reference's 5 two-sided rows per foot
    slack = b - A @ u
    tight = np.abs(slack) <= 1e-9 * (1.0 + np.abs(b))
    for p, k in enumerate(stance):
        act_lo[5 * k:5 * k + 5] = tight[6 * p:6 * p + 5]
        act_up[5 * k + 4] = tight[6 * p + 5]
        y_full[5 * k:5 * k + 5] = -yA[6 * p:6 * p + 5]     # lower-bound multipliers are <= 0
        y_full[5 * k + 4] += yA[6 * p + 5]
    y_full = _swing_multipliers(y_full, g, H, u_full, mu, swing)
    return QPSolution(u_full, y_full, act_lo, act_up, stat, prim, iters, rounds, verified,
                      info=dict(n_stance=ns, n_active=int(act.sum())))


def _swing_multipliers(y, g, H, u, mu, swing):
    """Any multipliers that balance the gradient on the pinned swing feet (for KKT reporting).

    At a swing foot all 5 rows are tight at 0 (and row 4 also at its upper bound 0), so
    the normal cone is all of R^3; pick y on rows 0/2 (lower) and the fz row."""
    if len(swing) == 0:
        return y
    grad = H @ u + g
    for k in swing:
        gx, gy, gz = grad[3 * k:3 * k + 3]
        # C'y = -grad with rows (1,0,mu),(−1,0,mu),(0,1,mu),(0,−1,mu),(0,0,1)
        y[5 * k + 0] = min(-gx, 0.0)
        y[5 * k + 1] = min(gx, 0.0)
        y[5 * k + 2] = min(-gy, 0.0)
        y[5 * k + 3] = min(gy, 0.0)
        y[5 * k + 4] = -gz - mu * (y[5 * k] + y[5 * k + 1] + y[5 * k + 2] + y[5 * k + 3])
    return y


def kkt_report(H, g, mu, ub_fz, u, tol_active=1e-6):
    """Independent KKT check of a candidate u against (H, g, constraints): returns
    (stationarity residual after a per-foot NNLS multiplier fit, primal violation,
    active_lower[20H], active_upper[20H]) with activity decided on primal slack."""
    mu = float(np.float32(mu))                               # the reference's C is float32 (mpc.py:239-245)
    H = np.asarray(H, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    nfeet = len(ub_fz)
    grad = H @ u + np.asarray(g, dtype=np.float64)
    pyr = np.array([[1, 0, mu], [-1, 0, mu], [0, 1, mu], [0, -1, mu], [0, 0, 1.0]])
    act_lo = np.zeros(5 * nfeet, dtype=bool)
    act_up = np.zeros(5 * nfeet, dtype=bool)
    stat = prim = 0.0
    for k in range(nfeet):
        f = u[3 * k:3 * k + 3]
        cu = pyr @ f
        ub = max(float(ub_fz[k]), 0.0)
        scale = 1.0 + np.abs(f).max()
        lo = cu <= tol_active * scale
        up = np.zeros(5, dtype=bool)
        up[4] = cu[4] >= ub - tol_active * scale
        prim = max(prim, float(max(0.0, (-cu).max(), cu[4] - ub)))
        # normals of the tight one-sided rows (outward): lower rows -> -pyr[r], upper -> +pyr[4]
        normals = [-pyr[r] for r in range(5) if lo[r]] + ([pyr[4]] if up[4] else [])
        gk = grad[3 * k:3 * k + 3]
        if normals:
            N = np.array(normals).T
            _, res = nnls(N, -gk)
            stat = max(stat, float(res))
        else:
            stat = max(stat, float(np.abs(gk).max()))
        act_lo[5 * k:5 * k + 5] = lo
        act_up[5 * k:5 * k + 5] = up
    return stat, prim, act_lo, act_up
