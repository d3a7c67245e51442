"""ORACLE (test infrastructure, never shipped or measured): CPU restatement of the reference's per-leg layer that consumes the
MPC forces - SURVEY.md 8f row 4.

  * `OracleSwingFootTrajectoryGenerator` follows `linear_mpc/swing_foot_trajectory_generator.py:15-129` (foot placement state
    machine `:82-129`, swing target in the base frame `:65-80`, three-point trajectory `:38-63`), with the same dtype flow
    (float32 base position / velocity, float64 everything else, float32 break points).
  * `cubic_hermite_zero_velocity` restates the one third-party call on that path, Drake 1.15.0
    `PiecewisePolynomial.CubicHermite(breaks, samples, sample_dots)` + `.value(t)` / `.derivative(1).value(t)`
    (`requirements.txt:23`; call site `swing_foot_trajectory_generator.py:53-57`).  pydrake is absent from this image; the
    published construction is: on a segment [t_a, t_b] of length h with end values y_a, y_b and end slopes d_a, d_b,
        p(s) = y_a + d_a s + c2 s^2 + c3 s^3,  s = t - t_a,
        c2 = (3 (y_b - y_a)/h - 2 d_a - d_b) / h,   c3 = (d_a + d_b - 2 (y_b - y_a)/h) / h^2,
    evaluation clamps t to [breaks[0], breaks[-1]].  The reference passes zero slopes.  "Parity unpinned" for this one
    function in the strict sense (no Drake output exists to compare with); everything around it is pinned against the
    unmodified reference code (`oracle/make_golden_legs.py`, `tests/golden/reference_legs.npz`).
  * `OracleLegController` follows `linear_mpc/leg_controller.py:10-91`.

Arithmetic follows numpy >= 2 promotion (the fixtures are generated under numpy 2.3): `np.float64 scalar * float32 array` is
float64.  Under the reference's pinned numpy 1.24 a few of those products are rounded to float32 first; the difference is at
most one float32 ulp of a centimetre-sized term (1e-9 m) and is covered by the tolerance of the device tests.
"""
from __future__ import annotations

import numpy as np

FOOT_Z_FINAL = -0.0255          # swing_foot_trajectory_generator.py:117


def cubic_hermite_zero_velocity(breaks, samples, t):
    """Value and first derivative at t of the piecewise cubic through `samples[:, k]` at `breaks[k]` with zero end slopes."""
    br = np.asarray(breaks, dtype=np.float64).reshape(-1)
    y = np.asarray(samples, dtype=np.float64)
    t = min(max(float(t), br[0]), br[-1])
    k = len(br) - 2
    for i in range(len(br) - 1):
        if t < br[i + 1]:
            k = i
            break
    h = br[k + 1] - br[k]
    s = t - br[k]
    dy = y[:, k + 1] - y[:, k]
    c2 = (3.0 * dy / h) / h
    c3 = (-2.0 * dy / h) / (h * h)
    pos = y[:, k] + s * s * (c2 + c3 * s)
    vel = s * (2.0 * c2 + 3.0 * c3 * s)
    return pos, vel


def expand_jacobians(blocks):
    """[..., 4, 3, 3] per-leg joint blocks -> [..., 4, 3, 18] foot Jacobians in the reference layout (6 floating-base columns,
    then 3 per leg, `utils/robot_data.py:119-133`); every column outside a leg's own block is filled with 0.25 so that a wrong
    slice in `leg_controller.py:84,88` would show."""
    blocks = np.asarray(blocks, dtype=np.float64)
    J = np.full(blocks.shape[:-1] + (18,), 0.25)
    for leg in range(4):
        J[..., leg, :, 6 + 3 * leg:9 + 3 * leg] = blocks[..., leg, :, :]
    return J


class OracleSwingFootTrajectoryGenerator:
    def __init__(self, leg_id, dt_control=0.001, swing_height=0.1, gravity=9.81):
        self.dt_control, self.swing_height, self.gravity = dt_control, swing_height, gravity   # :31-34
        self.is_first_swing = True
        self.remaining_swing_time = 0.0
        self.leg_id = leg_id
        self.footpos_init = np.zeros(3)
        self.footpos_final = np.zeros(3)

    def generate_swing_foot_trajectory(self, total_swing_time, cur_swing_time):                 # :38-63
        breaks = np.array([[0.0], [total_swing_time / 2.0], [total_swing_time]], dtype=np.float32)
        mid = (self.footpos_init + self.footpos_final) / 2
        mid[2] = self.swing_height
        pts = np.hstack((np.reshape(self.footpos_init, (3, 1)), mid.reshape(3, 1), np.reshape(self.footpos_final, (3, 1))))
        return cubic_hermite_zero_velocity(breaks, pts, cur_swing_time)

    def compute_traj_swingfoot(self, robot_data, swing_time):                                   # :65-80
        pos_base = np.array(robot_data.pos_base, dtype=np.float32)
        vel_base = np.array(robot_data.lin_vel_base, dtype=np.float32)
        R_base = robot_data.R_base
        cur = swing_time - self.remaining_swing_time
        pos, vel = self.generate_swing_foot_trajectory(swing_time, cur)
        return R_base.T @ (pos - pos_base), R_base.T @ (vel - vel_base)

    def set_foot_placement(self, robot_data, swing_time, stance_time, swing_state, base_vel_base_des, yaw_turn_rate_des):   # :82-129
        pos_base = np.array(robot_data.pos_base, dtype=np.float32)
        vel_base = np.array(robot_data.lin_vel_base, dtype=np.float32)
        R_base = robot_data.R_base
        thigh = robot_data.base_pos_base_thighs[self.leg_id]
        vel_base_des = R_base @ base_vel_base_des
        if self.is_first_swing:
            self.remaining_swing_time = swing_time
        else:
            self.remaining_swing_time -= self.dt_control
        th = yaw_turn_rate_des * 0.5 * stance_time
        RotZ = np.array([[np.cos(th), -np.sin(th), 0.0], [np.sin(th), np.cos(th), 0.0], [0.0, 0.0, 1.0]])
        final = pos_base + R_base @ (RotZ @ thigh + base_vel_base_des * self.remaining_swing_time) \
            + 0.5 * stance_time * vel_base + 0.03 * (vel_base - vel_base_des)
        final[0] += (0.5 * pos_base[2] / self.gravity) * (vel_base[1] * yaw_turn_rate_des)
        final[1] += (0.5 * pos_base[2] / self.gravity) * (-vel_base[0] * yaw_turn_rate_des)
        final[2] = FOOT_Z_FINAL
        self.footpos_final = final
        if self.is_first_swing:
            self.is_first_swing = False
            self.footpos_init = robot_data.pos_feet[self.leg_id]
        if swing_state >= 1:
            self.is_first_swing = True


class OracleLegController:
    def __init__(self, Kp_swing, Kd_swing):                                                     # leg_controller.py:29-32
        self.Kp, self.Kd = Kp_swing, Kd_swing
        self.torque_cmds = np.zeros(12, dtype=np.float32)

    def update(self, robot_data, contact_forces, swing_states, pos_targets, vel_targets):       # :38-91
        R = robot_data.R_base
        for leg in range(4):
            Jv = robot_data.Jv_feet[leg]
            if swing_states[leg]:
                err = self.Kp @ (R @ pos_targets[leg] - R @ robot_data.base_pos_base_feet[leg]) \
                    + self.Kd @ (R @ vel_targets[leg] - R @ robot_data.base_vel_base_feet[leg])
                tau = Jv.T @ err
            else:
                tau = Jv.T @ -contact_forces[3 * leg:3 * leg + 3]
            self.torque_cmds[3 * leg:3 * leg + 3] = tau[6 + 3 * leg:6 + 3 * leg + 3]
        return self.torque_cmds


def leg_layer_tick(gens, ctrl, robot_data, swing_time, stance_time, swing_states, contact_forces, v_des, yaw_rate):
    """One robot, one control tick of the loop body `scripts/isaacgym_a1.py:146-162` after the MPC call."""
    pos_t, vel_t = np.zeros((4, 3)), np.zeros((4, 3))
    for leg in range(4):
        if swing_states[leg] > 0:
            gens[leg].set_foot_placement(robot_data, swing_time, stance_time, swing_states[leg], v_des, yaw_rate)
            pos_t[leg], vel_t[leg] = gens[leg].compute_traj_swingfoot(robot_data, swing_time)
    tau = ctrl.update(robot_data, contact_forces, swing_states, pos_t, vel_t)
    return pos_t, vel_t, np.array(tau)
