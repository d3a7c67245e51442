"""Batched per-leg layer behind the MPC: the batch counterparts of the reference's
`SwingFootTrajectoryGenerator` (linear_mpc/swing_foot_trajectory_generator.py:15-129, one object per leg per robot) and
`LegController` (linear_mpc/leg_controller.py:10-91, one per robot), as driven by scripts/isaacgym_a1.py:143-162.

Both are thin holders of device state around one elementwise kernel each (`mpcq_swing_targets`, `mpcq_leg_torques`); the
arithmetic lives in libmpcq.so and there is no host path.  Kinematics (foot Jacobians, foot / thigh positions and velocities -
pinocchio in the reference, utils/robot_data.py:97-108) stay the caller's: they arrive as batched tensors.
"""
from __future__ import annotations

import torch

from . import _capi
from .configs import AliengoConfig, LinearMpcConfig


class BatchedLegKinematics:
    """The RobotData fields the leg layer reads (utils/robot_data.py:97-108), batched, float64, on the engine's device."""

    def __init__(self, pos_base, lin_vel_base, R_base, base_pos_base_thighs, pos_feet, base_pos_base_feet, base_vel_base_feet,
                 Jv_feet):
        self.pos_base = pos_base                            # [B,3] world
        self.lin_vel_base = lin_vel_base                    # [B,3] world
        self.R_base = R_base                                # [B,3,3]
        self.base_pos_base_thighs = base_pos_base_thighs    # [B,4,3] base frame
        self.pos_feet = pos_feet                            # [B,4,3] world
        self.base_pos_base_feet = base_pos_base_feet        # [B,4,3] base frame
        self.base_vel_base_feet = base_vel_base_feet        # [B,4,3] base frame
        self.Jv_feet = Jv_feet                              # [B,4,3,18] (reference layout) or [B,4,3,3] (joint blocks)


def _leg_params(mpc_config, robot_config):
    # the reference's generator reads AliengoConfig.swing_height whatever the robot (swing_foot_trajectory_generator.py:33)
    return _capi.make_leg_params(robot_config.Kp_swing, robot_config.Kd_swing, AliengoConfig.swing_height,
                                 mpc_config.dt_control, mpc_config.gravity)


class BatchedSwingFootTrajectoryGenerator:
    """4 x num_envs generators: `set_foot_placement` + `compute_traj_swingfoot` for every swinging leg in one launch."""

    def __init__(self, engine, num_envs: int, mpc_config=LinearMpcConfig, robot_config=AliengoConfig):
        self.engine, self.num_envs = engine, int(num_envs)
        self.params = _leg_params(mpc_config, robot_config)
        dev, f64 = engine.device, torch.float64
        B = self.num_envs
        self.swing_active = torch.zeros((B, 4), dtype=torch.uint8, device=dev)        # 0 <=> the reference's is_first_swing
        self.remaining_swing_time = torch.zeros((B, 4), dtype=f64, device=dev)
        self.footpos_init = torch.zeros((B, 4, 3), dtype=f64, device=dev)
        self.footpos_final = torch.zeros((B, 4, 3), dtype=f64, device=dev)
        self.pos_targets = torch.zeros((B, 4, 3), dtype=f64, device=dev)
        self.vel_targets = torch.zeros((B, 4, 3), dtype=f64, device=dev)

    def update(self, robot_data, gait, base_vel_base_des, yaw_turn_rate_des):
        """One control tick: `gait` is a `BatchedGaitSchedule` after `set_iteration` (swing states and swing / stance times),
        `base_vel_base_des` [B,3] and `yaw_turn_rate_des` [B] the commands.  Returns (pos_targets, vel_targets) [B,4,3]: the
        swing-foot target relative to the base in the base frame, zero rows for stance legs."""
        return self.engine.swing_targets(
            self.params, robot_data.pos_base, robot_data.lin_vel_base, robot_data.R_base, robot_data.base_pos_base_thighs,
            robot_data.pos_feet, gait.get_swing_state(), base_vel_base_des, yaw_turn_rate_des, gait.swing_time, gait.stance_time,
            (self.swing_active, self.remaining_swing_time, self.footpos_init, self.footpos_final),
            pos_targets=self.pos_targets, vel_targets=self.vel_targets)


class BatchedLegController:
    """`LegController.update` for num_envs robots: float32 joint torques [B,12]."""

    def __init__(self, engine, num_envs: int, Kp_swing, Kd_swing, mpc_config=LinearMpcConfig):
        self.engine, self.num_envs = engine, int(num_envs)
        self.params = _capi.make_leg_params(Kp_swing, Kd_swing, AliengoConfig.swing_height, mpc_config.dt_control, mpc_config.gravity)
        self.torque_cmds = torch.zeros((self.num_envs, 12), dtype=torch.float32, device=engine.device)

    def update(self, robot_data, contact_forces, swing_states, pos_targets_swingfeet, vel_targets_swingfeet):
        return self.engine.leg_torques(self.params, robot_data.Jv_feet, robot_data.R_base, robot_data.base_pos_base_feet,
                                       robot_data.base_vel_base_feet, contact_forces, swing_states, pos_targets_swingfeet,
                                       vel_targets_swingfeet, torque_cmds=self.torque_cmds)
