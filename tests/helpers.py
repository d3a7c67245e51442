"""Shared test helpers: seeded QP input batches and their oracle answers."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.mpc_oracle import OracleMPC, RobotState            # noqa: E402
from oracle.qp_exact import solve_qp_exact                      # noqa: E402
from pympc_quadruped_b200.configs import with_horizon           # noqa: E402
from pympc_quadruped_b200.synth import synth_states, synth_gait_tables  # noqa: E402


def make_batch(robot_cfg, horizon, B, regime, gaits, seed, solve=True):
    """Inputs of `mpcq_solve` for B seeded synthetic robots plus, per env, the oracle's
    (H, g, ub) from the reference construction and its exact optimum."""
    st = synth_states(B, robot_cfg, regime, seed=seed)
    tabs = synth_gait_tables(B, horizon, gaits, seed=seed)
    cfg = with_horizon(horizon)
    x0 = np.zeros((B, 13), np.float32)
    yaw = np.zeros(B)
    xref = np.zeros((B, 13 * horizon), np.float32)
    sols, qps = [], []
    for b in range(B):
        m = OracleMPC(cfg, robot_cfg)
        rd = RobotState(st["quat_base"][b], st["pos_base"][b], st["ang_vel_base"][b], st["lin_vel_base"][b],
                        st["pos_base_feet"][b])
        m.update_robot_state(rd)
        m.is_first_run = False
        m.xpos_base_desired = float(m.current_state[3])
        m.ypos_base_desired = float(m.current_state[4])
        m.yaw_desired = m.yaw
        vel = rd.R_base @ st["vel_cmd_body"][b]
        xr = m.reference_trajectory(vel, float(st["yaw_rate_cmd"][b]))
        x0[b], yaw[b], xref[b] = m.current_state, m.yaw, xr
        if solve:
            Hm, g, Cm, lb, ub = m.build_qp(xr, tabs[b])
            sol = solve_qp_exact(Hm, g, m.mu, ub[4::5])
            assert sol.verified
            sols.append(sol)
            qps.append((Hm, g, ub))
    return dict(x0=x0, yaw=yaw, feet=st["pos_base_feet"].reshape(B, 12).copy(), gait=tabs.astype(np.float32),
                xref=xref, sols=sols, qps=qps, cfg=cfg, B=B, horizon=horizon)


def oracle_chunk(args):
    """Worker of `oracle_pool_solve` (picklable): the reference's QP construction restated + the exact solve, per env."""
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:
        pass
    horizon, robot_name, x0, yaw, feet, xref, gait = args
    from oracle import mpc_oracle as mo
    from pympc_quadruped_b200 import configs
    robot = getattr(configs, robot_name)
    cfg = with_horizon(horizon)
    Qbar = np.kron(np.identity(horizon), cfg.Q)
    Rbar = np.kron(np.identity(horizon), cfg.R)
    out = np.zeros((x0.shape[0], 12 * horizon))
    for b in range(x0.shape[0]):
        Ac, Bc = mo.state_space_model(float(yaw[b]), feet[b].reshape(4, 3), robot.base_inertia_base, robot.mass_base)
        Ad, Bd = mo.discretize(Ac, Bc, 0.05)
        Hm, g = mo.qp_cost(Ad, Bd, x0[b].astype(np.float32), xref[b].astype(np.float32), Qbar, Rbar, horizon)
        _, _, ub = mo.qp_constraints(gait[b], cfg.friction_coef, robot.fz_max, horizon)
        sol = solve_qp_exact(Hm, g, cfg.friction_coef, ub[4::5])
        assert sol.verified
        out[b] = sol.u
    return out


def oracle_pool_solve(batch, robot_name, idx=None, workers=None):
    """Oracle optimum u [len(idx), 12H] of the envs `idx` of a make_batch(..., solve=False) batch, on all host cores."""
    import multiprocessing as mp
    idx = np.arange(batch["B"]) if idx is None else np.asarray(idx)
    workers = workers or min(os.cpu_count() or 1, max(1, len(idx) // 8))
    bounds = np.linspace(0, len(idx), workers + 1).astype(int)
    jobs = []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        if hi > lo:
            s = idx[lo:hi]
            jobs.append((batch["horizon"], robot_name, batch["x0"][s], batch["yaw"][s], batch["feet"][s], batch["xref"][s], batch["gait"][s]))
    with mp.get_context("spawn").Pool(workers) as pool:
        outs = pool.map(oracle_chunk, jobs)
    return np.concatenate(outs)
