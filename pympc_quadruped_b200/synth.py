"""Seeded synthetic robot states for parity tests and the benchmark.

There is no simulator in this environment, so the "recorded states" of the
BASELINE configs are drawn from the generator SURVEY.md section 8(d) specifies
(`numpy.random.default_rng(seed)`):

  nominal     small roll/pitch, feet under the hips, gentle velocities: the
              unconstrained optimum is usually feasible.
  aggressive  large roll/pitch/rates, low body, scattered feet: 3..28 stance
              rows of the friction pyramid end up active (solver stress).

The fields are the five `RobotData` fields the MPC path consumes
(utils/robot_data.py:70-76,144-149): quat_base (w,x,y,z), pos_base,
ang_vel_base, lin_vel_base (world frame), pos_base_feet (world-frame
base->foot, legs FL,FR,RL,RR), plus R_base and the command.
"""
from __future__ import annotations

import numpy as np

from .gait import Gait, gait_tables

SEED_BASE = 20261018

_HIPS_A1 = np.array([[0.183, 0.13, 0.0], [0.183, -0.13, 0.0],
                     [-0.183, 0.13, 0.0], [-0.183, -0.13, 0.0]])


def zyx_to_matrix(roll, pitch, yaw) -> np.ndarray:
    """R = Rz(yaw) Ry(pitch) Rx(roll), batched -> [B,3,3]."""
    cr, sr = np.cos(roll), np.sin(roll)
    cp, sp = np.cos(pitch), np.sin(pitch)
    cy, sy = np.cos(yaw), np.sin(yaw)
    R = np.empty(np.shape(roll) + (3, 3))
    R[..., 0, 0] = cy * cp
    R[..., 0, 1] = cy * sp * sr - sy * cr
    R[..., 0, 2] = cy * sp * cr + sy * sr
    R[..., 1, 0] = sy * cp
    R[..., 1, 1] = sy * sp * sr + cy * cr
    R[..., 1, 2] = sy * sp * cr - cy * sr
    R[..., 2, 0] = -sp
    R[..., 2, 1] = cp * sr
    R[..., 2, 2] = cp * cr
    return R


def zyx_to_quat(roll, pitch, yaw) -> np.ndarray:
    """Unit quaternion (w,x,y,z) of Rz(yaw) Ry(pitch) Rx(roll), batched -> [B,4]."""
    cr, sr = np.cos(roll / 2), np.sin(roll / 2)
    cp, sp = np.cos(pitch / 2), np.sin(pitch / 2)
    cy, sy = np.cos(yaw / 2), np.sin(yaw / 2)
    return np.stack([cy * cp * cr + sy * sp * sr,
                     cy * cp * sr - sy * sp * cr,
                     cy * sp * cr + sy * cp * sr,
                     sy * cp * cr - cy * sp * sr], axis=-1)


def synth_states(num_envs: int, robot_config, regime: str = "mixed", seed: int = SEED_BASE,
                 hip_scale: float | None = None) -> dict:
    """Draw `num_envs` robot states + commands.  regime: nominal | aggressive | mixed (50/50)."""
    if regime not in ("nominal", "aggressive", "mixed"):
        raise ValueError(regime)
    rng = np.random.default_rng(seed)
    B = int(num_envs)
    if regime == "mixed":
        aggressive = rng.random(B) < 0.5
    else:
        aggressive = np.full(B, regime == "aggressive")
    if hip_scale is None:
        hip_scale = 1.3 if float(robot_config.mass_base) > 6.0 else 1.0
    h_des = float(robot_config.base_height_des)

    def pick(nom, agg):
        return np.where(aggressive.reshape((B,) + (1,) * (nom.ndim - 1)), agg, nom)

    u = lambda lo, hi, shape: rng.uniform(lo, hi, size=shape)
    roll = pick(u(-0.05, 0.05, B), u(-0.15, 0.15, B))
    pitch = pick(u(-0.05, 0.05, B), u(-0.15, 0.15, B))
    yaw = u(-np.pi, np.pi, B)
    pos = np.stack([u(-0.05, 0.05, B), u(-0.05, 0.05, B),
                    pick(h_des + u(-0.03, 0.03, B), u(0.25, 0.36, B))], axis=-1)
    v_body = pick(np.stack([u(0, 1.0, B), u(-0.1, 0.1, B), u(-0.05, 0.05, B)], -1),
                  np.stack([u(-1.5, 1.5, B), u(-0.5, 0.5, B), u(-0.3, 0.3, B)], -1))
    omega = pick(u(-0.1, 0.1, (B, 3)), u(-1.0, 1.0, (B, 3)))
    foot_off = pick(np.stack([u(-0.05, 0.05, (B, 4)), u(-0.02, 0.02, (B, 4))], -1),
                    np.stack([u(-0.1, 0.1, (B, 4)), u(-0.05, 0.05, (B, 4))], -1))
    cmd_vx = u(0.0, 1.4, B)
    yaw_rate = pick(u(-0.3, 0.3, B), u(-1.0, 1.0, B))

    R = zyx_to_matrix(roll, pitch, yaw)
    feet_body = np.empty((B, 4, 3))
    feet_body[:, :, 0:2] = hip_scale * _HIPS_A1[None, :, 0:2] + foot_off
    feet_body[:, :, 2] = -pos[:, 2:3]
    feet = np.einsum('bij,bkj->bki', R, feet_body)
    return dict(
        quat_base=zyx_to_quat(roll, pitch, yaw),
        pos_base=pos,
        ang_vel_base=omega,
        lin_vel_base=np.einsum('bij,bj->bi', R, v_body),
        pos_base_feet=feet,
        R_base=R,
        vel_cmd_body=np.stack([cmd_vx, np.zeros(B), np.zeros(B)], -1),
        yaw_rate_cmd=yaw_rate,
        aggressive=aggressive,
    )


GAIT_MIX = (Gait.TROTTING10, Gait.PACING10, Gait.BOUNDING10)


def synth_gait_params(num_envs: int, gaits=(Gait.TROTTING10,), seed: int = SEED_BASE):
    """Per-env random pattern from `gaits` and random phase: (stance_offsets [B,4], stance_durations [B,4],
    num_segment [B], iteration [B]) - the inputs of `gait_tables` / of the device kernel `mpcq_gait_tables`."""
    rng = np.random.default_rng(seed + 7919)
    which = rng.integers(0, len(gaits), size=num_envs)
    offs = np.stack([gaits[k].stance_offsets for k in which])
    durs = np.stack([gaits[k].stance_durations for k in which])
    segs = np.array([gaits[k].num_segment for k in which])
    iteration = rng.integers(0, segs)
    return offs, durs, segs, iteration


def synth_gait_tables(num_envs: int, horizon: int, gaits=(Gait.TROTTING10,),
                      seed: int = SEED_BASE) -> np.ndarray:
    """Per-env random pattern from `gaits` and random phase -> float32 [B, 4*horizon]."""
    offs, durs, segs, iteration = synth_gait_params(num_envs, gaits, seed)
    return gait_tables(offs, durs, segs, iteration, horizon)
