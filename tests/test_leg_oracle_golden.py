"""The leg-layer oracle (oracle/leg_oracle.py) against the sequence recorded from the unmodified reference
classes (tests/golden/reference_legs.npz, written by oracle/make_golden_legs.py)."""
import os
import types

import numpy as np
import pytest

from oracle.leg_oracle import (OracleLegController, OracleSwingFootTrajectoryGenerator, cubic_hermite_zero_velocity,
                               expand_jacobians, leg_layer_tick)

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_legs.npz")
KP = {0: np.diag([700.0] * 3), 1: np.diag([200.0] * 3)}       # robot 0 = A1Config, 1 = AliengoConfig (config/robot_configs.py:55,36)
KD = np.diag([20.0] * 3)


def load_gold():
    z = np.load(GOLD)
    return {k: (z[k].astype(np.float64) if z[k].dtype == np.float32 and k != "torque_cmds" else z[k]) for k in z.files}


def robot_data_at(d, r, t):
    rd = types.SimpleNamespace(R_base=d["R_base"][r, t], pos_base=d["pos_base"][r, t], lin_vel_base=d["lin_vel_base"][r, t])
    for k in ("base_pos_base_thighs", "base_pos_base_feet", "base_vel_base_feet", "pos_feet"):
        setattr(rd, k, list(d[k][r, t]))
    rd.Jv_feet = list(expand_jacobians(d["Jv_blocks"][r, t]))
    return rd


def test_oracle_reproduces_the_reference_sequence_bitwise():
    d = load_gold()
    R, T = d["pos_base"].shape[:2]
    swings = 0
    for r in range(R):
        gens = [OracleSwingFootTrajectoryGenerator(leg) for leg in range(4)]
        ctrl = OracleLegController(KP[int(d["robot"][r])], KD)
        sw, st = np.float64(d["swing_stance_time"][r, 0]), np.float64(d["swing_stance_time"][r, 1])
        for t in range(T):
            ss = d["swing_state"][r, t]
            pt, vt, tau = leg_layer_tick(gens, ctrl, robot_data_at(d, r, t), sw, st, ss, d["contact_forces"][r, t],
                                         d["v_des"][r], float(d["yaw_rate"][r]))
            assert np.array_equal(pt, d["pos_targets"][r, t]) and np.array_equal(vt, d["vel_targets"][r, t]), (r, t)
            assert np.array_equal(tau, d["torque_cmds"][r, t]), (r, t)
            swings += int((ss > 0).sum())
    assert swings > 1000        # the fixture really exercises swing phases (and stance ones)
    assert (d["swing_state"] == 0).sum() > 1000


def test_hermite_interpolates_and_clamps():
    br = np.array([0.0, 0.05, 0.1], dtype=np.float32)
    pts = np.array([[0.0, 0.5, 1.0], [1.0, 1.0, 1.0], [-0.02, 0.1, -0.02]])
    for k in range(3):
        p, v = cubic_hermite_zero_velocity(br, pts, float(br[k]))
        assert np.allclose(p, pts[:, k], atol=1e-15) and np.allclose(v, 0.0, atol=1e-12)
    p, v = cubic_hermite_zero_velocity(br, pts, 0.025)
    assert np.allclose(p, 0.5 * (pts[:, 0] + pts[:, 1])) and np.allclose(v[0], 1.5 * 0.5 / float(br[1]))
    assert np.array_equal(cubic_hermite_zero_velocity(br, pts, -1.0)[0], pts[:, 0])
    assert np.allclose(cubic_hermite_zero_velocity(br, pts, 7.0)[0], pts[:, 2], atol=1e-15)
    # derivative is the derivative (central difference)
    h = 1e-7
    for t in (0.013, 0.049, 0.0777):
        pa, _ = cubic_hermite_zero_velocity(br, pts, t - h)
        pb, _ = cubic_hermite_zero_velocity(br, pts, t + h)
        assert np.allclose((pb - pa) / (2 * h), cubic_hermite_zero_velocity(br, pts, t)[1], atol=1e-6)
