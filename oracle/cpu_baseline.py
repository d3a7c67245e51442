"""ORACLE (test infrastructure, NOT product code) - the reference's per-robot CPU loop, timed.

Used only by bench.py (`cpu_baseline` leg and `--impl reference`) and tests.  One call of
`solve_chunk` does, per environment, exactly what the reference's control tick does on the hot
path (scripts/isaacgym_a1.py:141-143): update_robot_state -> generate_reference_trajectory ->
_generate_state_space_model -> _discretize_continuous_model (twice, as the reference does) ->
_generate_QP_cost -> _generate_QP_constraints -> solve.  The construction is the numpy
restatement pinned bit-for-bit against the reference (tests/test_oracle_golden.py); the solve is
oracle.qp_exact because Drake/OSQP (the reference's solver) is not installable offline.
"""
from __future__ import annotations

import os
import time

import numpy as np


def solve_chunk(args):
    """args = (mpc_config_kwargs, robot_name, states dict of arrays, gait tables) -> (forces, t_build, t_solve)"""
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:
        pass
    horizon, robot_name, st, tabs = args
    from oracle.mpc_oracle import OracleMPC, RobotState
    from oracle.qp_exact import solve_qp_exact
    from pympc_quadruped_b200 import configs
    robot = getattr(configs, robot_name)
    cfg = configs.with_horizon(horizon)
    B = st["quat_base"].shape[0]
    forces = np.zeros((B, 12))
    t_build = t_solve = 0.0
    for b in range(B):
        t0 = time.perf_counter()
        m = OracleMPC(cfg, robot)
        rd = RobotState(st["quat_base"][b], st["pos_base"][b], st["ang_vel_base"][b], st["lin_vel_base"][b],
                        st["pos_base_feet"][b])
        m.update_robot_state(rd)
        m.is_first_run = False
        m.xpos_base_desired = float(m.current_state[3])
        m.ypos_base_desired = float(m.current_state[4])
        # a regular (non-first) control tick whose integrated desired xy equals the current xy;
        # yaw_desired follows mpc.py:92
        m.yaw_desired = m.yaw + m.dt_control * float(st["yaw_rate_cmd"][b])
        xr = m.reference_trajectory(rd.R_base @ st["vel_cmd_body"][b], float(st["yaw_rate_cmd"][b]))
        H, g, C, lb, ub = m.build_qp(xr, tabs[b])
        t1 = time.perf_counter()
        sol = solve_qp_exact(H, g, m.mu, ub[4::5])
        t2 = time.perf_counter()
        forces[b] = sol.u[:12]
        t_build += t1 - t0
        t_solve += t2 - t1
    return forces, t_build, t_solve


def run_parallel(pool, horizon, robot_name, st, tabs, n_workers):
    """Split the sample over the pool; returns (forces, wall seconds, cpu build s, cpu solve s)."""
    B = st["quat_base"].shape[0]
    bounds = np.linspace(0, B, n_workers + 1).astype(int)
    keys = ("quat_base", "pos_base", "ang_vel_base", "lin_vel_base", "pos_base_feet", "vel_cmd_body", "yaw_rate_cmd")
    jobs = [(horizon, robot_name, {k: st[k][lo:hi] for k in keys}, tabs[lo:hi])
            for lo, hi in zip(bounds[:-1], bounds[1:]) if hi > lo]
    t0 = time.perf_counter()
    outs = pool.map(solve_chunk, jobs)
    wall = time.perf_counter() - t0
    return np.concatenate([o[0] for o in outs]), wall, sum(o[1] for o in outs), sum(o[2] for o in outs)
