"""Constructor inputs of the MPC path: controller and robot constants.

Mirrors the *attribute names and values* of the reference's class-attribute
configs so that either the reference's own classes or these can be handed to
the controller (the reference passes the classes themselves, not instances:
``ModelPredictiveController(LinearMpcConfig, A1Config)``,
scripts/isaacgym_a1.py:93).

  LinearMpcConfig  <- config/linear_mpc_configs.py:4-24
  RobotConfig / AliengoConfig / A1Config  <- config/robot_configs.py:9-56
  com_inertia()    <- utils/dynamics.py:3-18 (float32 symmetric 3x3)

Only the attributes the MPC path reads are required by the engine
(`extract_mpc_constants`); swing gains etc. are carried for completeness.
"""
from __future__ import annotations

import numpy as np


def com_inertia(ixx, ixy, ixz, iyy, iyz, izz) -> np.ndarray:
    """Symmetric body inertia, float32 like the reference (utils/dynamics.py:14-18)."""
    return np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]], dtype=np.float32)


class LinearMpcConfig:
    dt_control: float = 0.001
    iteration_between_mpc: int = 20
    dt_mpc: float = 0.05
    horizon: int = 16
    gravity: float = 9.81
    friction_coef: float = 0.7
    # r, p, y, x, y, z, wx, wy, wz, vx, vy, vz, g
    Q: np.ndarray = np.diag([5., 5., 10., 10., 10., 50., 0.01, 0.01, 0.2, 0.2, 0.2, 0.2, 0.])
    R: np.ndarray = np.diag([1e-5] * 12)
    cmd_xvel: float = 0.
    cmd_yvel: float = 0.
    cmd_yaw_turn_rate: float = 0.


def with_horizon(horizon: int, base=LinearMpcConfig):
    """A LinearMpcConfig subclass with a different horizon (BASELINE configs use 10 / 30)."""
    return type(f"{base.__name__}H{horizon}", (base,), {"horizon": int(horizon)})


class RobotConfig:
    mass_base: float
    base_height_des: float
    base_inertia_base: np.ndarray
    fz_max: float
    swing_height: float
    Kp_swing: np.ndarray
    Kd_swing: np.ndarray


class AliengoConfig(RobotConfig):
    mass_base: float = 9.042
    base_height_des: float = 0.38
    base_inertia_base = com_inertia(0.033260231, -0.000451628, 0.000487603,
                                    0.16117211, 4.8356e-05, 0.17460442)
    fz_max = 500.
    swing_height = 0.1
    Kp_swing = np.diag([200., 200., 200.])
    Kd_swing = np.diag([20., 20., 20.])


class A1Config(RobotConfig):
    mass_base: float = 4.713
    base_height_des: float = 0.42
    # the reference scales the URDF trunk inertia by 10 (config/robot_configs.py:50)
    base_inertia_base = com_inertia(0.01683993, 8.3902e-05, 0.000597679,
                                    0.056579028, 2.5134e-05, 0.064713601) * 10
    fz_max = 500.
    swing_height = 0.1
    Kp_swing = np.diag([700., 700., 700.])
    Kd_swing = np.diag([20., 20., 20.])


# The reference hard-codes the MPC step to 0.05 s and ignores dt_mpc
# (linear_mpc/mpc.py:38).  Kept as a named constant so the quirk is visible.
REFERENCE_MPC_DT = 0.05


def extract_mpc_constants(mpc_config, robot_config) -> dict:
    """Read exactly what `_load_parameters` reads (linear_mpc/mpc.py:35-52).

    Accepts classes or instances, the reference's or ours.  Q and R must be
    diagonal (they are in the reference); the engine keeps only the diagonals.
    """
    Q = np.asarray(mpc_config.Q, dtype=np.float64)
    R = np.asarray(mpc_config.R, dtype=np.float64)
    if Q.shape != (13, 13) or R.shape != (12, 12):
        raise ValueError(f"Q must be 13x13 and R 12x12, got {Q.shape} and {R.shape}")
    if np.any(Q - np.diag(np.diag(Q))) or np.any(R - np.diag(np.diag(R))):
        raise ValueError("the batched engine supports diagonal Q and R only (as in the reference config)")
    if np.any(np.diag(R) <= 0):
        raise ValueError("R must be positive (it makes the QP strictly convex)")
    inertia = np.asarray(robot_config.base_inertia_base, dtype=np.float32)
    if inertia.shape != (3, 3):
        raise ValueError("base_inertia_base must be 3x3")
    horizon = int(mpc_config.horizon)
    if horizon < 1:
        raise ValueError("horizon must be >= 1")
    return dict(
        dt_control=float(mpc_config.dt_control),
        iterations_between_mpc=int(mpc_config.iteration_between_mpc),
        dt=REFERENCE_MPC_DT,
        horizon=horizon,
        mu=float(mpc_config.friction_coef),
        fz_max=float(robot_config.fz_max),
        gravity=float(mpc_config.gravity),
        inertia=inertia,
        mass=float(robot_config.mass_base),
        com_height_des=float(robot_config.base_height_des),
        q_diag=np.diag(Q).copy(),
        r_diag=np.diag(R).copy(),
    )
