"""Pin the oracle's exact QP solver (oracle/qp_exact.py).  The reference delegates the solve to
Drake/OSQP (absent offline); the QP is strictly convex so the optimum is unique and any exact
method must agree.  Checks: KKT conditions, an independent solver (scipy trust-constr / SLSQP),
and hand-solvable cases."""
import numpy as np
import pytest
from scipy.optimize import LinearConstraint, minimize

from helpers import make_batch
from oracle.qp_exact import kkt_report, solve_qp_exact
from pympc_quadruped_b200 import A1Config, Gait

MU = 0.7
PYR = np.array([[1, 0, MU], [-1, 0, MU], [0, 1, MU], [0, -1, MU], [0, 0, 1.0]])


def test_hand_solvable_single_foot():
    # min 1/2 |f - t|^2 over the pyramid: targets inside, outside a side face, below the apex, above the cap
    H = np.eye(3)
    for target, expect in [((1.0, -2.0, 10.0), (1.0, -2.0, 10.0)),
                           ((0.0, 0.0, -5.0), (0.0, 0.0, 0.0)),
                           ((0.0, 0.0, 600.0), (0.0, 0.0, 500.0))]:
        sol = solve_qp_exact(H, -np.array(target), MU, np.array([500.0]))
        assert sol.verified and np.allclose(sol.u, expect, atol=1e-9)
    # outside the +x face: projection onto the plane fx = mu fz
    t = np.array([10.0, 0.0, 5.0])
    nrm = np.array([1.0, 0.0, -MU])
    expect = t - nrm * (nrm @ t) / (nrm @ nrm)
    sol = solve_qp_exact(H, -t, MU, np.array([500.0]))
    assert sol.verified and np.allclose(sol.u, expect, atol=1e-9)
    assert sol.active_lower[1] and not sol.active_lower[0]
    # swing foot is pinned to zero and reports every row tight
    sol = solve_qp_exact(H, -t, MU, np.array([0.0]))
    assert sol.verified and np.all(sol.u == 0) and sol.active_lower.all() and sol.active_upper[4]


def test_kkt_conditions_and_independent_solver():
    batch = make_batch(A1Config, 3, 6, "aggressive", (Gait.TROTTING10, Gait.STANDING), 77)
    for b in range(batch["B"]):
        H, g, ub = batch["qps"][b]
        sol = batch["sols"][b]
        n = H.shape[0]
        # KKT with the solver's own multipliers: H u + g + C' y = 0, sign and complementarity
        C = np.kron(np.eye(n // 3), PYR)
        cu = C @ sol.u
        assert np.abs(H @ sol.u + g + C.T @ sol.y).max() <= 1e-8
        assert cu.min() >= -1e-9 and np.all(cu[4::5] <= ub[4::5] + 1e-9)
        stat, prim, lo, up = kkt_report(H, g, MU, ub[4::5], sol.u)
        assert stat <= 1e-7 and prim <= 1e-9
        assert np.array_equal(lo, sol.active_lower) and np.array_equal(up, sol.active_upper)
        # independent: scipy SLSQP from a different start on the stance variables
        ubf = np.where(np.isfinite(ub), ub, 1e9).astype(np.float64)
        fun = lambda u: 0.5 * u @ H @ u + g @ u
        jac = lambda u: H @ u + g
        cons = [{"type": "ineq", "fun": lambda u: C @ u, "jac": lambda u: C},
                {"type": "ineq", "fun": lambda u: ubf - C @ u, "jac": lambda u: -C}]
        r = minimize(fun, np.zeros(n), jac=jac, constraints=cons, method="SLSQP", options=dict(maxiter=500, ftol=1e-14))
        assert fun(sol.u) <= fun(r.x) + 1e-7 * (1 + abs(fun(r.x)))
        assert np.abs(r.x - sol.u).max() <= 5e-3 * (1 + np.abs(sol.u).max())


def test_unverified_is_reported_not_hidden():
    # a non-convex "H" cannot be verified; the oracle must say so instead of returning garbage silently
    H = np.diag([1.0, 1.0, -1.0])
    try:
        sol = solve_qp_exact(H, np.array([0.0, 0.0, -1.0]), MU, np.array([500.0]))
        ok = sol.verified and abs(sol.u[2] - 500.0) < 1e-6     # the only KKT point that is a minimiser is the cap
        assert ok or not sol.verified
    except (np.linalg.LinAlgError, AssertionError):
        pass
