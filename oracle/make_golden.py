"""ORACLE tooling - generate tests/golden/*.npz from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  The reference's
linear_mpc/mpc.py cannot be imported as-is (matplotlib, pydrake, qpsolvers and
pinocchio are absent), so empty stub modules are injected for those names; the
QP-CONSTRUCTION path (mpc.py:55-260) then runs unmodified.  The solve
(mpc.py:277-290, Drake/OSQP) cannot run; the stored `f_star` / `u_star` come from
oracle.qp_exact on the reference's own (H, g, C, lb, ub) and are marked as such.

Usage:  python -m oracle.make_golden            (one subprocess per horizon, because the
        reference's Gait enum reads LinearMpcConfig.horizon at import time, gait.py:47-50)
"""
from __future__ import annotations

import os
import subprocess
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def import_reference(horizon: int):
    for name in ("matplotlib", "matplotlib.pyplot", "pydrake", "pydrake.all", "qpsolvers", "pinocchio"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pydrake.all"].MathematicalProgram = None
    sys.modules["pydrake.all"].Solve = None
    sys.modules["qpsolvers"].solve_qp = None
    for sub in ("config", "utils", "linear_mpc"):
        sys.path.insert(0, os.path.join(REF, sub))
    import linear_mpc_configs
    linear_mpc_configs.LinearMpcConfig.horizon = horizon     # before `gait` is imported
    import robot_configs
    import gait as ref_gait
    import mpc as ref_mpc
    import kinematics as ref_kin
    return linear_mpc_configs.LinearMpcConfig, robot_configs, ref_gait, ref_mpc, ref_kin


class _RD:
    pass


def _robot_data(st, b, ref_kin):
    rd = _RD()
    rd.quat_base = st["quat_base"][b]
    rd.pos_base = st["pos_base"][b]
    rd.ang_vel_base = st["ang_vel_base"][b]
    rd.lin_vel_base = st["lin_vel_base"][b]
    rd.pos_base_feet = [st["pos_base_feet"][b, i] for i in range(4)]
    rd.R_base = ref_kin.quat2matrix(rd.quat_base)
    return rd


def generate(horizon: int) -> None:
    sys.path.insert(0, ROOT)
    MpcCfg, robot_configs, ref_gait, ref_mpc, ref_kin = import_reference(horizon)
    from pympc_quadruped_b200.synth import synth_states, SEED_BASE
    from oracle.qp_exact import solve_qp_exact
    import io
    import contextlib

    gaits = {"TROTTING10": ref_gait.Gait.TROTTING10, "PACING10": ref_gait.Gait.PACING10,
             "STANDING": ref_gait.Gait.STANDING, "TROTTING16": ref_gait.Gait.TROTTING16,
             "JUMPING16": ref_gait.Gait.JUMPING16}
    out = {}

    # ---- gait tables: every pattern at every iteration of one period -------------------------
    for name, gt in gaits.items():
        tabs = []
        for it in range(0, 20 * gt.num_segment, 20):
            gt.set_iteration(20, it)
            tabs.append(gt.get_gait_table().copy())
        out[f"gait/{name}"] = np.stack(tabs)

    # ---- survey known-answer state + seeded synthetic states ----------------
    for robot_name in ("A1", "Aliengo"):
        rcfg = getattr(robot_configs, robot_name + "Config")
        cases = []
        h = rcfg.base_height_des
        kat = dict(quat_base=np.array([[1., 0, 0, 0]]), pos_base=np.array([[0, 0, h]]),
                   lin_vel_base=np.array([[0.5, 0, 0]]), ang_vel_base=np.zeros((1, 3)),
                   pos_base_feet=np.array([[[0.18, 0.13, -h], [0.18, -0.13, -h],
                                            [-0.18, 0.13, -h], [-0.18, -0.13, -h]]]),
                   vel_cmd_body=np.array([[0.5, 0, 0]]), yaw_rate_cmd=np.array([0.0]))
        cases.append(("kat", kat, 0, "TROTTING10", 0))
        n_syn = 6 if horizon == 10 else 2
        for regime in ("nominal", "aggressive"):
            st = synth_states(n_syn, rcfg, regime, seed=SEED_BASE + 100 + horizon)
            for b in range(n_syn):
                gname = ("TROTTING10", "PACING10", "STANDING", "TROTTING16", "JUMPING16", "TROTTING10")[b % 6]
                cases.append((f"{regime}{b}", st, b, gname, 20 * ((3 * b + 1) % 10)))
        for tag, st, b, gname, git in cases:
            ctrl = ref_mpc.ModelPredictiveController(MpcCfg, rcfg)
            rd = _robot_data(st, b, ref_kin)
            gt = gaits[gname]
            gt.set_iteration(20, git)
            table = gt.get_gait_table().copy()
            ctrl.update_robot_state(rd)
            vel_des = rd.R_base @ st["vel_cmd_body"][b]
            yaw_rate = float(st["yaw_rate_cmd"][b])
            # state as after the first `update_mpc_if_needed` preamble (mpc.py:84-88) except
            # that desired xy = current xy as SURVEY 8d specifies (kat: 0)
            ctrl.xpos_base_desired = 0.0 if tag == "kat" else float(ctrl.current_state[3])
            ctrl.ypos_base_desired = 0.0 if tag == "kat" else float(ctrl.current_state[4])
            ctrl.yaw_desired = ctrl.yaw
            ctrl.is_first_run = False
            x_ref = ctrl.generate_reference_trajectory(vel_des, yaw_rate)
            Ac, Bc = ctrl._generate_state_space_model()
            Ad, Bd = ctrl._discretize_continuous_model(Ac, Bc)
            H, g = ctrl._generate_QP_cost(Ad, Bd, ctrl.current_state, x_ref)
            C, lb, ub = ctrl._generate_QP_constraints(table)
            sol = solve_qp_exact(H, g, ctrl.mu, ub[4::5])
            assert sol.verified, (robot_name, tag)
            k = f"{robot_name}/{tag}/"
            out[k + "quat_base"] = rd.quat_base
            out[k + "pos_base"] = rd.pos_base
            out[k + "ang_vel_base"] = rd.ang_vel_base
            out[k + "lin_vel_base"] = rd.lin_vel_base
            out[k + "pos_base_feet"] = np.stack(rd.pos_base_feet)
            out[k + "R_base"] = rd.R_base
            out[k + "vel_cmd_body"] = st["vel_cmd_body"][b]
            out[k + "yaw_rate_cmd"] = np.float64(yaw_rate)
            out[k + "gait_table"] = table
            out[k + "xy_des"] = np.array([ctrl.xpos_base_desired, ctrl.ypos_base_desired])
            out[k + "current_state"] = ctrl.current_state.copy()
            out[k + "yaw"] = np.float64(ctrl.yaw)
            out[k + "x_ref"] = x_ref
            out[k + "roll_pitch_init"] = np.array([ctrl.roll_init, ctrl.pitch_init], dtype=np.float64)
            out[k + "Ac"], out[k + "Bc"], out[k + "Ad"], out[k + "Bd"] = Ac, Bc, Ad, Bd
            if horizon <= 16 and tag in ("kat", "nominal0", "aggressive0", "aggressive2"):
                out[k + "H"] = H
            else:                                           # keep the fixture small: digest + a few rows
                n = H.shape[0]
                out[k + "H_rows"] = H[[0, 1, 2, n // 3 - 1, n // 2, n - 1]]
                out[k + "H_digest"] = np.array([np.trace(H), H.sum(), np.abs(H).sum(), (H * H).sum()])
            out[k + "g"] = g
            out[k + "lb"], out[k + "ub"] = lb, ub
            if tag == "kat" and robot_name == "A1":
                out["C"] = C
            out[k + "u_star__oracle"] = sol.u              # NOT from the reference (Drake absent)
            out[k + "active_lower__oracle"] = sol.active_lower
            out[k + "active_upper__oracle"] = sol.active_upper

        # ---- stateful sequence: 61 control ticks through the public API (mpc.py:81-108) ------
        if horizon == 10:
            ctrl = ref_mpc.ModelPredictiveController(MpcCfg, rcfg)
            # the reference solves through Drake; substitute the oracle's exact solver for that one call
            ctrl._solve_mpc = (lambda self: lambda ref, gait_table, solver='drake', debug=False:
                               _solve_via_oracle(self, ref, gait_table, solve_qp_exact))(ctrl)
            st = synth_states(4, rcfg, "nominal", seed=SEED_BASE + 300)
            gt = gaits["TROTTING10"]
            forces, refs, des = [], [], []
            for tick in range(61):
                b = (tick // 20) % 4                         # a new "measured" state every MPC period
                rd = _robot_data(st, b, ref_kin)
                gt.set_iteration(20, tick)
                table = gt.get_gait_table()
                ctrl.update_robot_state(rd)
                with contextlib.redirect_stdout(io.StringIO()):
                    f = ctrl.update_mpc_if_needed(tick, st["vel_cmd_body"][b], float(st["yaw_rate_cmd"][b]),
                                                  table, solver='drake')
                forces.append(np.array(f))
                refs.append(ctrl.ref_traj.copy())
                des.append([ctrl.xpos_base_desired, ctrl.ypos_base_desired, ctrl.yaw_desired,
                            ctrl.roll_init, ctrl.pitch_init])
            k = f"{robot_name}/seq/"
            for name in ("quat_base", "pos_base", "ang_vel_base", "lin_vel_base", "pos_base_feet",
                         "vel_cmd_body", "yaw_rate_cmd"):
                out[k + name] = st[name]
            out[k + "forces__oracle_solver"] = np.stack(forces)
            out[k + "ref_traj"] = np.stack(refs)
            out[k + "desired"] = np.array(des, dtype=np.float64)

    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, f"reference_h{horizon}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


def _solve_via_oracle(ctrl, ref_traj, gait_table, solve_qp_exact):
    Ac, Bc = ctrl._generate_state_space_model()
    Ad, Bd = ctrl._discretize_continuous_model(Ac, Bc)
    H, g = ctrl._generate_QP_cost(Ad, Bd, ctrl.current_state, ref_traj)
    C, lb, ub = ctrl._generate_QP_constraints(gait_table)
    sol = solve_qp_exact(H, g, ctrl.mu, ub[4::5])
    assert sol.verified
    return sol.u


if __name__ == "__main__":
    if len(sys.argv) > 1:
        generate(int(sys.argv[1]))
    else:
        env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1")
        for h in (10, 16, 30):
            subprocess.run([sys.executable, "-m", "oracle.make_golden", str(h)], check=True, cwd=ROOT, env=env)
