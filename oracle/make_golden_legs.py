"""ORACLE tooling - generate tests/golden/reference_legs.npz from the UNMODIFIED reference leg layer.

Runs only in the build container (needs /root/reference).  `linear_mpc/swing_foot_trajectory_generator.py`,
`linear_mpc/leg_controller.py` and `linear_mpc/gait.py` are imported as they are, with empty stub modules for matplotlib and
pinocchio.  pydrake is absent, so `pydrake.all.PiecewisePolynomial` is a stand-in whose `CubicHermite` is the restated Drake
construction of `oracle/leg_oracle.py` - the ONE function on this path that is not the reference's own code; the state machine,
the frame changes, the break points and the torque map around it run unmodified.  Robot data is synthetic (pinocchio is absent):
random but smooth base motion, random foot Jacobians, the attributes the two classes read (`utils/robot_data.py:70-108`).

Usage:  python -m oracle.make_golden_legs
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REF = "/root/reference"
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
OUT = os.path.join(ROOT, "tests", "golden", "reference_legs.npz")


class _Poly:
    def __init__(self, breaks, samples, order=0):
        self.breaks, self.samples, self.order = breaks, samples, order

    def value(self, t):
        from oracle.leg_oracle import cubic_hermite_zero_velocity
        return cubic_hermite_zero_velocity(self.breaks, self.samples, t)[self.order].reshape(3, 1)

    def derivative(self, n):
        assert n == 1
        return _Poly(self.breaks, self.samples, 1)


class _PiecewisePolynomial:
    @staticmethod
    def CubicHermite(breaks, samples, sample_dots):
        assert not np.any(sample_dots)
        return _Poly(breaks, samples)


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "pydrake", "pydrake.all", "pinocchio", "qpsolvers"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pydrake.all"].PiecewisePolynomial = _PiecewisePolynomial
    for sub in ("config", "utils", "linear_mpc"):
        sys.path.insert(0, os.path.join(REF, sub))
    import gait as ref_gait
    import leg_controller as ref_leg
    import robot_configs
    import swing_foot_trajectory_generator as ref_swing
    return ref_gait, ref_leg, ref_swing, robot_configs


def synth_leg_data(num_robots: int, ticks: int, seed: int):
    """Per robot, per control tick: the RobotData attributes the leg layer reads, plus MPC forces and commands."""
    rng = np.random.default_rng(seed)
    R, T = num_robots, ticks
    t = np.arange(T)[None, :, None] * 1e-3
    yaw = rng.uniform(-np.pi, np.pi, (R, 1)) + 0.3 * t[..., 0]
    roll, pitch = 0.05 * np.sin(7 * t[..., 0] + rng.uniform(0, 6, (R, 1))), 0.05 * np.cos(5 * t[..., 0] + rng.uniform(0, 6, (R, 1)))
    cr, sr, cp, sp, cy, sy = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    Rb = np.empty((R, T, 3, 3))
    Rb[..., 0, 0], Rb[..., 0, 1], Rb[..., 0, 2] = cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr
    Rb[..., 1, 0], Rb[..., 1, 1], Rb[..., 1, 2] = sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr
    Rb[..., 2, 0], Rb[..., 2, 1], Rb[..., 2, 2] = -sp, cp * sr, cp * cr
    vel = rng.uniform(-0.3, 1.0, (R, 1, 3)) + 0.1 * np.sin(11 * t + rng.uniform(0, 6, (R, 1, 3)))
    vel[..., 2] *= 0.1
    pos = rng.uniform(-0.5, 0.5, (R, 1, 3)) + np.cumsum(vel, axis=1) * 1e-3
    pos[..., 2] = 0.3 + 0.02 * np.sin(9 * t[..., 0])
    hips = np.array([[0.183, 0.13, 0.0], [0.183, -0.13, 0.0], [-0.183, 0.13, 0.0], [-0.183, -0.13, 0.0]])
    thighs = hips[None, None] + rng.uniform(-0.01, 0.01, (R, T, 4, 3))
    base_feet = hips[None, None] + np.array([0.0, 0.0, -0.3]) + rng.uniform(-0.08, 0.08, (R, T, 4, 3))
    pos_feet = pos[:, :, None, :] + np.einsum("rtij,rtlj->rtli", Rb, base_feet)
    base_vel_feet = rng.uniform(-1.0, 1.0, (R, T, 4, 3))
    Jv_blocks = rng.uniform(-0.4, 0.4, (R, T, 4, 3, 3))
    forces = rng.uniform(-20.0, 20.0, (R, T, 12))
    forces[..., 2::3] = rng.uniform(0.0, 120.0, (R, T, 4))
    v_des = np.concatenate([rng.uniform(0.0, 1.4, (R, 1)), rng.uniform(-0.2, 0.2, (R, 1)), np.zeros((R, 1))], axis=1)
    yaw_rate = rng.uniform(-0.5, 0.5, R)
    d = dict(R_base=Rb, pos_base=pos, lin_vel_base=vel, base_pos_base_thighs=thighs, base_pos_base_feet=base_feet,
             pos_feet=pos_feet, base_vel_base_feet=base_vel_feet, Jv_blocks=Jv_blocks, contact_forces=forces, v_des=v_des, yaw_rate=yaw_rate)
    return {k: v.astype(np.float32) for k, v in d.items()}      # float32-exact inputs keep the committed fixture small


def robot_data_at(d, r, t):
    from oracle.leg_oracle import expand_jacobians
    rd = types.SimpleNamespace()
    rd.R_base = d["R_base"][r, t]
    rd.pos_base = d["pos_base"][r, t]
    rd.lin_vel_base = d["lin_vel_base"][r, t]
    rd.base_pos_base_thighs = [d["base_pos_base_thighs"][r, t, i] for i in range(4)]
    rd.base_pos_base_feet = [d["base_pos_base_feet"][r, t, i] for i in range(4)]
    rd.base_vel_base_feet = [d["base_vel_base_feet"][r, t, i] for i in range(4)]
    rd.pos_feet = [d["pos_feet"][r, t, i] for i in range(4)]
    rd.Jv_feet = list(expand_jacobians(d["Jv_blocks"][r, t]))
    return rd


GAITS = ("TROTTING10", "PACING10", "TROTTING16", "JUMPING16")
NUM_ROBOTS, TICKS, SEED, START = 8, 340, 20261022, (0, 37, 101, 160, 3, 250, 77, 199)


def main() -> None:
    sys.path.insert(0, ROOT)
    ref_gait, ref_leg, ref_swing, robot_configs = import_reference()
    out = synth_leg_data(NUM_ROBOTS, TICKS, SEED)
    d = {k: v.astype(np.float64) for k, v in out.items()}
    ibm = 20
    pos_t = np.zeros((NUM_ROBOTS, TICKS, 4, 3))
    vel_t = np.zeros_like(pos_t)
    tau = np.zeros((NUM_ROBOTS, TICKS, 12), dtype=np.float32)
    swing = np.zeros((NUM_ROBOTS, TICKS, 4))
    times = np.zeros((NUM_ROBOTS, 2))
    gait_params = np.zeros((NUM_ROBOTS, 9), dtype=np.int32)
    for r in range(NUM_ROBOTS):
        cfg = robot_configs.A1Config if r % 2 == 0 else robot_configs.AliengoConfig
        gait = getattr(ref_gait.Gait, GAITS[r % len(GAITS)])
        gens = [ref_swing.SwingFootTrajectoryGenerator(leg) for leg in range(4)]
        ctrl = ref_leg.LegController(cfg.Kp_swing, cfg.Kd_swing)
        times[r] = gait.swing_time, gait.stance_time
        gait_params[r] = [*gait.stance_offsets, *gait.stance_durations, gait.num_segment]
        for t in range(TICKS):
            rd = robot_data_at(d, r, t)
            gait.set_iteration(ibm, START[r] + t)                          # scripts/isaacgym_a1.py:137-138
            ss = gait.get_swing_state()
            swing[r, t] = ss
            pt, vt = np.zeros((4, 3)), np.zeros((4, 3))
            for leg in range(4):                                           # scripts/isaacgym_a1.py:146-162
                if ss[leg] > 0:
                    gens[leg].set_foot_placement(rd, gait, d["v_des"][r], float(d["yaw_rate"][r]))
                    pt[leg], vt[leg] = gens[leg].compute_traj_swingfoot(rd, gait)
                torque = ctrl.update(rd, d["contact_forces"][r, t], ss, pt, vt)
            pos_t[r, t], vel_t[r, t], tau[r, t] = pt, vt, torque
    out.update(pos_targets=pos_t, vel_targets=vel_t, torque_cmds=tau, swing_state=swing, swing_stance_time=times,
               gait_params=gait_params, start_iteration=np.array(START), iterations_between_mpc=np.array(ibm),
               robot=np.array([r % 2 for r in range(NUM_ROBOTS)]))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
