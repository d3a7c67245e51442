"""ORACLE tooling - BASELINE configs[0]: 1 000 consecutive MPC updates of ONE A1 robot through the UNMODIFIED reference
class (public API, scripts/mujoco_aliengo.py:204-207 call sequence), over a synthetic "recorded" trajectory.

Runs only in the build container (needs /root/reference).  State t of the trajectory is sample t of the seeded nominal
generator, the gait is TROTTING10 at control tick 20 t (iteration = t mod 10), every call is an MPC update
(iter_counter = 20 t), and the controller object is ONE instance for the whole sequence, so its integrators (desired x / y /
yaw, roll / pitch compensation, mpc.py:84-93,121-152) carry over like in the simulator loop.  The reference's Drake call is
replaced by oracle.qp_exact on the reference's own (H, g, ub) - Drake is not installable offline - and marked as such.

Writes tests/golden/reference_cfg0_seq.npz.   Usage:  python -m oracle.make_golden_cfg0
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
N_UPDATES = 1000


def main():
    sys.path.insert(0, ROOT)
    from oracle.make_golden import import_reference, _robot_data, _solve_via_oracle, OUT
    MpcCfg, robot_configs, ref_gait, ref_mpc, ref_kin = import_reference(10)
    from oracle.qp_exact import solve_qp_exact
    from pympc_quadruped_b200.synth import SEED_BASE, synth_states
    rcfg = robot_configs.A1Config
    st = synth_states(N_UPDATES, rcfg, "nominal", seed=SEED_BASE + 0)
    ctrl = ref_mpc.ModelPredictiveController(MpcCfg, rcfg)
    ctrl._solve_mpc = (lambda self: lambda ref, gait_table, solver='drake', debug=False:
                       _solve_via_oracle(self, ref, gait_table, solve_qp_exact))(ctrl)
    gt = ref_gait.Gait.TROTTING10
    forces, des, refsum = [], [], []
    for t in range(N_UPDATES):
        rd = _robot_data(st, t, ref_kin)
        gt.set_iteration(20, 20 * t)
        ctrl.update_robot_state(rd)
        with contextlib.redirect_stdout(io.StringIO()):
            f = ctrl.update_mpc_if_needed(20 * t, st["vel_cmd_body"][t], float(st["yaw_rate_cmd"][t]), gt.get_gait_table(), solver='drake')
        forces.append(np.array(f, dtype=np.float64))
        des.append([ctrl.xpos_base_desired, ctrl.ypos_base_desired, ctrl.yaw_desired, ctrl.roll_init, ctrl.pitch_init])
        r = ctrl.ref_traj.astype(np.float64)
        refsum.append([r.sum(), np.abs(r).sum(), (r * np.arange(1, r.size + 1)).sum()])
    out = {name: st[name] for name in ("quat_base", "pos_base", "ang_vel_base", "lin_vel_base", "pos_base_feet", "vel_cmd_body", "yaw_rate_cmd")}
    out["forces__oracle_solver"] = np.stack(forces)
    out["desired"] = np.array(des, dtype=np.float64)
    out["ref_traj_checksums"] = np.array(refsum, dtype=np.float64)
    path = os.path.join(OUT, "reference_cfg0_seq.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
