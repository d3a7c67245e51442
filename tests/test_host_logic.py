"""Host-side logic on CPU: the controller mirror (state assembly, integrators, decimation, reference
trajectory) against the reference's own outputs, the C-ABI library surface, gait schedules."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from fake_engine import OracleEngine
from oracle import mpc_oracle as mo
from pympc_quadruped_b200 import A1Config, AliengoConfig, Gait, LinearMpcConfig, _capi, with_horizon
from pympc_quadruped_b200.controller import (BatchedModelPredictiveController, BatchedRobotData,
                                             ModelPredictiveController)
from pympc_quadruped_b200.gait import gait_tables
from pympc_quadruped_b200.synth import synth_states

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_c_abi_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "mpcq.h")).read()
    declared = set(re.findall(r"^\s*(?:const\s+char\*|int|void)\s+(mpcq_\w+)\s*\(", header, flags=re.M))
    assert declared == set(_capi.EXPORTS), declared ^ set(_capi.EXPORTS)
    from conftest import load_module
    lib = ctypes.CDLL(load_module("mpcq_lib_build", os.path.join(ROOT, "pympc_quadruped_b200", "csrc", "build.py")).build())
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mpcq_version() == 100
    # struct mirror has the C layout's size: 4 ints + 5 doubles + 34 doubles + 4 ints + 4 doubles
    assert ctypes.sizeof(_capi.MpcqConfig) == 16 + 8 * 5 + 8 * 34 + 16 + 56


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_no_cpu_fallback():
    lib = _capi.load_library()
    cfg = _capi.make_config(__import__("pympc_quadruped_b200.configs", fromlist=["x"]).extract_mpc_constants(with_horizon(10), A1Config))
    h = ctypes.c_void_p()
    assert lib.mpcq_create(ctypes.byref(cfg), ctypes.byref(h)) == -3          # MPCQ_ERR_NO_DEVICE
    assert b"no CPU path" in lib.mpcq_last_error(None)
    with pytest.raises(RuntimeError):
        BatchedModelPredictiveController(with_horizon(10), A1Config, 2, device="cuda:0")
    with pytest.raises(RuntimeError):
        from pympc_quadruped_b200.engine import MpcqEngine
        MpcqEngine(with_horizon(10), A1Config, device="cpu")


class _RD:
    pass


@pytest.mark.parametrize("robot_name,robot", [("A1", A1Config), ("Aliengo", AliengoConfig)])
def test_single_robot_adapter_follows_reference_sequence(robot_name, robot):
    """61 control ticks through the reference's public API (fixture generated from the unmodified
    reference class, its Drake call replaced by the oracle solver): integrators, decimation, cached
    forces and reference trajectories must agree."""
    z = np.load(os.path.join(GOLD, "reference_h10.npz"))
    k = f"{robot_name}/seq/"
    cfg = with_horizon(10)
    ctrl = ModelPredictiveController(cfg, robot, device="cpu", dtype=torch.float64, engine=OracleEngine(cfg, robot, torch.float64))
    gt = Gait.TROTTING10.with_horizon(10)
    for tick in range(61):
        b = (tick // 20) % 4
        rd = _RD()
        rd.quat_base, rd.pos_base = z[k + "quat_base"][b], z[k + "pos_base"][b]
        rd.ang_vel_base, rd.lin_vel_base = z[k + "ang_vel_base"][b], z[k + "lin_vel_base"][b]
        rd.pos_base_feet = [z[k + "pos_base_feet"][b, i] for i in range(4)]
        rd.R_base = mo.quat_to_matrix(rd.quat_base)
        gt.set_iteration(20, tick)
        ctrl.update_robot_state(rd)
        f = ctrl.update_mpc_if_needed(tick, z[k + "vel_cmd_body"][b], float(z[k + "yaw_rate_cmd"][b]), gt.get_gait_table())
        assert f.shape == (12,) and f.dtype == np.float64
        assert np.abs(f - z[k + "forces__oracle_solver"][tick]).max() <= 1e-4
        # float32 storage: equal up to one float32 ulp.  (The fixture was generated under numpy 2, whose NEP-50
        # promotion keeps `python_float * np.float32` in float32; the reference pins numpy 1.24 where those
        # scalar ops run in float64, which is what the controller does - see DESIGN.md "rounding points".)
        ref = z[k + "ref_traj"][tick]
        assert np.abs(ctrl.ref_traj - ref).max() <= 1.2e-7 * max(1.0, np.abs(ref).max()), tick
        des = z[k + "desired"][tick]
        got = [float(ctrl.xpos_base_desired[0]), float(ctrl.ypos_base_desired[0]), float(ctrl.yaw_desired[0]),
               float(ctrl.roll_init[0]), float(ctrl.pitch_init[0])]
        assert np.allclose(got[:3], des[:3], rtol=0, atol=1e-15), tick
        assert np.allclose(got[3:], des[3:], rtol=5e-7, atol=1e-12), tick
    assert ctrl.engine.calls == 4                       # MPC runs every 20th tick only (mpc.py:95)
    assert ctrl.iterations_between_mpc == 20 and ctrl.horizon == 10 and ctrl.dt == 0.05


def test_batched_controller_equals_independent_oracle_controllers():
    B, H = 6, 10
    cfg = with_horizon(H)
    st = synth_states(B, A1Config, "mixed", seed=99)
    eng = OracleEngine(cfg, A1Config, torch.float64)
    ctrl = BatchedModelPredictiveController(cfg, A1Config, B, device="cpu", dtype=torch.float64, engine=eng)
    singles = [mo.OracleMPC(cfg, A1Config) for _ in range(B)]
    gt = Gait.PACING10.with_horizon(H)
    rng = np.random.default_rng(0)
    for tick in (0, 1, 2, 19, 20, 21, 40):
        scale = 1.0 + 0.01 * tick
        gt.set_iteration(20, tick)
        table = gt.get_gait_table()
        ctrl.update_robot_state(BatchedRobotData(st["quat_base"], st["pos_base"] * scale, st["ang_vel_base"], st["lin_vel_base"] * scale,
                                                 st["pos_base_feet"]))
        fb = ctrl.update_mpc_if_needed(tick, st["vel_cmd_body"], st["yaw_rate_cmd"], np.tile(table, (B, 1)))
        for b in range(B):
            m = singles[b]
            m.update_robot_state(mo.RobotState(st["quat_base"][b], st["pos_base"][b] * scale, st["ang_vel_base"][b],
                                               st["lin_vel_base"][b] * scale, st["pos_base_feet"][b]))
            fs = m.update_mpc_if_needed(tick, st["vel_cmd_body"][b], float(st["yaw_rate_cmd"][b]), table)
            assert np.array_equal(ctrl.current_state[b].numpy(), m.current_state)
            assert np.abs(ctrl.ref_traj[b].numpy() - m.ref_traj).max() <= 1.2e-7 * max(1.0, np.abs(m.ref_traj).max()), (tick, b)
            assert abs(float(ctrl.xpos_base_desired[b]) - m.xpos_base_desired) <= 1e-15
            assert abs(float(ctrl.pitch_init[b]) - float(m.pitch_init)) <= 5e-7 * abs(float(m.pitch_init)) + 1e-12
            assert np.abs(fb[b].numpy() - fs).max() <= 1e-4
    with pytest.raises(NotImplementedError):
        ctrl.update_mpc_if_needed(0, st["vel_cmd_body"], st["yaw_rate_cmd"], np.tile(table, (B, 1)), solver="qpsolvers")


def test_vectorised_gait_tables_and_phase_states():
    for sched in (Gait.TROTTING10, Gait.PACING10, Gait.BOUNDING10, Gait.STANDING, Gait.JUMPING16, Gait.TROTTING16):
        for H in (10, 30):
            s = sched.with_horizon(H)
            its = np.arange(0, 40)
            tabs = gait_tables(np.tile(s.stance_offsets, (40, 1)), np.tile(s.stance_durations, (40, 1)),
                               np.full(40, s.num_segment), its % s.num_segment, H)
            for it in its:
                s.set_iteration(20, 20 * it)
                ref = np.array([[1.0 if ((i + 1 + s.iteration - s.stance_offsets[j]) % s.num_segment) < s.stance_durations[j] else 0.0
                                 for j in range(4)] for i in range(H)], dtype=np.float32).reshape(-1)
                assert np.array_equal(s.get_gait_table(), ref)
                assert np.array_equal(tabs[it], ref)
    t = Gait.TROTTING10.with_horizon(10)
    t.set_iteration(20, 50)
    assert np.allclose(t.get_stance_state() + t.get_swing_state() >= 0, True)
    assert t.swing_time == pytest.approx(0.02 * 5) and t.stance_time == pytest.approx(0.02 * 5)


def test_config_extraction_validates():
    from pympc_quadruped_b200.configs import extract_mpc_constants
    c = extract_mpc_constants(LinearMpcConfig, A1Config)
    assert c["horizon"] == 16 and c["dt"] == 0.05 and c["inertia"].dtype == np.float32
    assert np.allclose(c["inertia"], np.array(A1Config.base_inertia_base))

    class BadQ(LinearMpcConfig):
        Q = np.ones((13, 13))
    with pytest.raises(ValueError):
        extract_mpc_constants(BadQ, A1Config)

    class BadR(LinearMpcConfig):
        R = np.zeros((12, 12))
    with pytest.raises(ValueError):
        extract_mpc_constants(BadR, A1Config)


def test_isaacgym_root_state_glue_reorders_like_the_reference_script():
    """scripts/isaacgym_a1.py:119-133: pos 0:3, quaternion (x,y,z,w) -> (w,x,y,z), lin vel 7:10, ang vel 10:13; base->foot =
    foot - base (utils/robot_data.py:144-149).  Fake tensors: the simulator is not in the image."""
    import torch
    from pympc_quadruped_b200.controller import BatchedRobotData
    rng = np.random.default_rng(5)
    rs = torch.as_tensor(rng.normal(size=(6, 13)).astype(np.float32))
    feet_w = torch.as_tensor(rng.normal(size=(6, 4, 3)).astype(np.float32))
    rd = BatchedRobotData.from_isaacgym(rs, foot_positions_world=feet_w)
    for b in range(6):                                           # the reference's per-robot statements
        q_imre = rs[b, 3:7].numpy()
        q_reim = np.array([q_imre[3], q_imre[0], q_imre[1], q_imre[2]], dtype=np.float32)
        assert np.array_equal(rd.quat_base[b].numpy(), q_reim.astype(np.float64))
        assert np.array_equal(rd.pos_base[b].numpy(), rs[b, 0:3].numpy().astype(np.float64))
        assert np.array_equal(rd.lin_vel_base[b].numpy(), rs[b, 7:10].numpy().astype(np.float64))
        assert np.array_equal(rd.ang_vel_base[b].numpy(), rs[b, 10:13].numpy().astype(np.float64))
        assert np.array_equal(rd.pos_base_feet[b].numpy(), feet_w[b].numpy().astype(np.float64) - rs[b, 0:3].numpy().astype(np.float64))
    ids = torch.tensor([4, 1])
    sub = BatchedRobotData.from_isaacgym(rs, pos_base_feet=rd.pos_base_feet[ids], env_ids=ids)
    assert torch.equal(sub.quat_base, rd.quat_base[ids]) and torch.equal(sub.pos_base_feet, rd.pos_base_feet[ids])
    with pytest.raises(ValueError):
        BatchedRobotData.from_isaacgym(rs)
    with pytest.raises(ValueError):
        BatchedRobotData.from_isaacgym(rs[:, :12], pos_base_feet=rd.pos_base_feet)
