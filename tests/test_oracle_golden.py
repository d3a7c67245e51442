"""Pin the oracle: the numpy restatement (oracle/mpc_oracle.py) and the closed form the CUDA
engine uses (oracle/structured.py) against fixtures generated from the UNMODIFIED reference
(oracle/make_golden.py -> tests/golden/reference_h{10,16,30}.npz).

The reference has no tests or golden vectors of its own, so these fixtures -
outputs of the reference's own functions run in the build container - are the pin.
"""
import os

import numpy as np
import pytest

from oracle import mpc_oracle as mo
from oracle.qp_exact import solve_qp_exact
from oracle.structured import dense_hessian, structured_qp
from pympc_quadruped_b200 import A1Config, AliengoConfig, LinearMpcConfig, with_horizon
from pympc_quadruped_b200.gait import Gait

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ROBOTS = {"A1": A1Config, "Aliengo": AliengoConfig}


def _load(h):
    return np.load(os.path.join(GOLD, f"reference_h{h}.npz"))


def _tags(z, robot):
    return sorted({k.split("/")[1] for k in z.keys() if k.startswith(robot + "/")} - {"seq"})


def _oracle_for(z, robot, tag, horizon):
    k = f"{robot}/{tag}/"
    m = mo.OracleMPC(with_horizon(horizon), ROBOTS[robot])
    rd = mo.RobotState(z[k + "quat_base"], z[k + "pos_base"], z[k + "ang_vel_base"], z[k + "lin_vel_base"],
                       z[k + "pos_base_feet"], z[k + "R_base"])
    m.update_robot_state(rd)
    m.is_first_run = False
    m.xpos_base_desired, m.ypos_base_desired = (0.0, 0.0) if tag == "kat" else (float(m.current_state[3]), float(m.current_state[4]))
    m.yaw_desired = m.yaw
    return m, rd, k


@pytest.mark.parametrize("horizon", [10, 16, 30])
def test_restatement_reproduces_reference_bitwise(horizon):
    z = _load(horizon)
    for robot in ROBOTS:
        for tag in _tags(z, robot):
            m, rd, k = _oracle_for(z, robot, tag, horizon)
            assert np.array_equal(m.current_state, z[k + "current_state"])
            assert m.yaw == float(z[k + "yaw"])
            x_ref = m.reference_trajectory(rd.R_base @ z[k + "vel_cmd_body"], float(z[k + "yaw_rate_cmd"]))
            assert np.array_equal(x_ref, z[k + "x_ref"])
            assert np.array_equal([m.xpos_base_desired, m.ypos_base_desired], z[k + "xy_des"])
            H, g, C, lb, ub, mid = m.build_qp(x_ref, z[k + "gait_table"], with_intermediates=True)
            for name in ("Ac", "Bc", "Ad", "Bd"):
                assert np.array_equal(mid[name], z[k + name]), (robot, tag, name)
            assert np.array_equal(g, z[k + "g"])
            assert np.array_equal(lb, z[k + "lb"]) and np.array_equal(ub, z[k + "ub"])
            if k + "H" in z:
                assert np.array_equal(H, z[k + "H"])
            else:
                n = H.shape[0]
                assert np.array_equal(H[[0, 1, 2, n // 3 - 1, n // 2, n - 1]], z[k + "H_rows"])
                assert np.allclose([np.trace(H), H.sum(), np.abs(H).sum(), (H * H).sum()], z[k + "H_digest"], rtol=1e-13)
    assert np.array_equal(mo.qp_constraints(np.ones(40, np.float32), 0.7, 500.0, 10)[0], _load(10)["C"])


def test_survey_known_answers():
    """SURVEY.md 8c.3 digits (computed during the survey from the reference's functions)."""
    z = _load(10)
    H, g = z["A1/kat/H"], z["A1/kat/g"]
    assert np.allclose(H[0, 0:3], [0.016145367, -0.002008632, 0.005294291], atol=5e-10)
    assert np.isclose(H[119, 119], 1.1809069461538672e-4, rtol=1e-12)
    assert np.isclose(np.trace(H), 2.2332838019847143, rtol=1e-12)
    assert np.isclose(np.linalg.norm(g), 6.962314063224849, rtol=1e-12)
    f = z["A1/kat/u_star__oracle"][:12]
    assert np.allclose(f, [-3.686563554, -1.788557749, 29.038091815, 0, 0, 0, 0, 0, 0,
                           -3.678219826, -1.800118919, 17.198448912], atol=2e-6)
    f = z["Aliengo/kat/u_star__oracle"][:12]
    assert np.allclose(f, [-3.975663848, -2.795317523, 52.527432073, 0, 0, 0, 0, 0, 0,
                           -3.975513047, -2.795532211, 36.18154613], atol=2e-6)


@pytest.mark.parametrize("horizon", [10, 16, 30])
def test_closed_form_matches_reference_construction(horizon):
    """The Kronecker closed form (what the kernels compute) vs the reference's expm/Su/Sx build:
    H, g to 2e-6 relative (float32 rounding of the reference's Su), optimum well inside 1e-3 N."""
    z = _load(horizon)
    q, r = np.diag(LinearMpcConfig.Q), np.diag(LinearMpcConfig.R)
    for robot, cfg in ROBOTS.items():
        for tag in _tags(z, robot):
            k = f"{robot}/{tag}/"
            M00, M11, g, _ = structured_qp(float(z[k + "yaw"]), z[k + "pos_base_feet"], cfg.base_inertia_base,
                                           cfg.mass_base, 0.05, q, r, z[k + "current_state"], z[k + "x_ref"], horizon)
            Hd = dense_hessian(M00, M11, r, horizon)
            assert np.abs(g - z[k + "g"]).max() <= 2e-6 * np.abs(z[k + "g"]).max()
            if k + "H" in z:
                assert np.abs(Hd - z[k + "H"]).max() <= 2e-6 * np.abs(z[k + "H"]).max()
            else:
                n = Hd.shape[0]
                rows = Hd[[0, 1, 2, n // 3 - 1, n // 2, n - 1]]
                assert np.abs(rows - z[k + "H_rows"]).max() <= 2e-6 * np.abs(z[k + "H_rows"]).max()
            if horizon == 10 or tag == "kat":
                sol = solve_qp_exact(Hd, g, 0.7, z[k + "ub"][4::5])
                assert sol.verified
                assert np.abs(sol.u - z[k + "u_star__oracle"]).max() <= 5e-4
                assert np.array_equal(sol.active_lower, z[k + "active_lower__oracle"])


def test_gait_tables_match_reference():
    z = _load(10)
    for name in ("TROTTING10", "PACING10", "STANDING", "TROTTING16", "JUMPING16"):
        sched = getattr(Gait, name).with_horizon(10)
        ref = z[f"gait/{name}"]
        for i in range(ref.shape[0]):
            sched.set_iteration(20, 20 * i)
            assert np.array_equal(sched.get_gait_table(), ref[i]), (name, i)
            assert sched.get_gait_table().dtype == np.float32
