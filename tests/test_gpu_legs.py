"""GPU parity of the per-leg layer through the C ABI: `mpcq_swing_targets` + `mpcq_leg_torques` (driven by the
device contact schedule `mpcq_gait_tables`) against the sequence recorded from the UNMODIFIED reference classes
(tests/golden/reference_legs.npz) and against the oracle restatement on random batches.

Tolerances (floating point, stated here): swing targets are float64 on both sides but differ in summation order, in the
device's sincos and in two products the reference rounds differently under numpy 1.24 / 2.x (oracle/leg_oracle.py header):
|d pos| <= 1e-8 m, |d vel| <= 1e-7 m/s (values 0.1-1 m, up to 15 m/s).  Torques are float32 on both sides: 2 float32 ulp of
the largest term, written as |d tau| <= 4e-7 * (1 + sum |J| |e|)."""
import os
import types

import numpy as np
import pytest
import torch

from oracle.leg_oracle import (OracleLegController, OracleSwingFootTrajectoryGenerator, expand_jacobians, leg_layer_tick)
from pympc_quadruped_b200 import A1Config, AliengoConfig, BatchedGaitSchedule, Gait, GaitSchedule
from pympc_quadruped_b200.configs import with_horizon

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_legs.npz")
POS_TOL, VEL_TOL = 1e-8, 1e-7
DEV = "cuda:0"


def _engine(robot, dtype):
    from pympc_quadruped_b200.engine import MpcqEngine
    return MpcqEngine(with_horizon(10), robot, dtype=dtype, device=DEV)


def _dev(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a)).to(device=DEV, dtype=dtype)


def _kin(d, sel, t):
    from pympc_quadruped_b200 import BatchedLegKinematics
    g = lambda k: d[k][sel, t].contiguous()
    return BatchedLegKinematics(g("pos_base"), g("lin_vel_base"), g("R_base"), g("base_pos_base_thighs"), g("pos_feet"),
                                g("base_pos_base_feet"), g("base_vel_base_feet"), g("Jv_feet"))


def _tau_tol(J, e):
    return 4e-7 * (1.0 + np.einsum("...rc,...r->...c", np.abs(J), np.abs(e)))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
def test_leg_layer_reproduces_the_reference_sequence(dtype):
    from pympc_quadruped_b200 import BatchedLegController, BatchedSwingFootTrajectoryGenerator
    z = np.load(GOLD)
    R, T = z["pos_base"].shape[:2]
    ibm = int(z["iterations_between_mpc"])
    d = {k: _dev(z[k]) for k in ("pos_base", "lin_vel_base", "R_base", "base_pos_base_thighs", "pos_feet", "base_pos_base_feet",
                                 "base_vel_base_feet", "v_des", "yaw_rate")}
    d["Jv_feet"] = _dev(expand_jacobians(z["Jv_blocks"]))
    forces = _dev(z["contact_forces"], dtype)
    worst = dict(pos=0.0, vel=0.0, tau=0.0)
    for robot_id, robot in ((0, A1Config), (1, AliengoConfig)):
        sel = np.nonzero(z["robot"] == robot_id)[0]
        eng = _engine(robot, dtype)
        gp = z["gait_params"][sel]
        gait = BatchedGaitSchedule(eng, [GaitSchedule("fixture", int(p[8]), p[0:4], p[4:8], horizon=10) for p in gp])
        assert np.array_equal(gait.swing_time.cpu().numpy(), z["swing_stance_time"][sel, 0])        # np.float64 products, exact
        assert np.array_equal(gait.stance_time.cpu().numpy(), z["swing_stance_time"][sel, 1])
        swing = BatchedSwingFootTrajectoryGenerator(eng, len(sel), robot_config=robot)
        legs = BatchedLegController(eng, len(sel), robot.Kp_swing, robot.Kd_swing)
        start = _dev(z["start_iteration"][sel], torch.int32)
        tsel = torch.as_tensor(sel, device=DEV)
        got_p, got_v, got_tau, got_ss = [], [], [], []
        for t in range(T):
            kin = _kin(d, tsel, t)
            gait.set_iteration(ibm, start + t)
            pt, vt = swing.update(kin, gait, d["v_des"][tsel].contiguous(), d["yaw_rate"][tsel].contiguous())
            tau = legs.update(kin, forces[tsel, t].contiguous(), gait.get_swing_state(), pt, vt)
            got_p.append(pt.clone()); got_v.append(vt.clone()); got_tau.append(tau.clone()); got_ss.append(gait.get_swing_state().clone())
        torch.cuda.synchronize()
        P, V = torch.stack(got_p, 1).cpu().numpy(), torch.stack(got_v, 1).cpu().numpy()
        TAU, SS = torch.stack(got_tau, 1).cpu().numpy(), torch.stack(got_ss, 1).cpu().numpy()
        assert np.array_equal(SS, z["swing_state"][sel])                       # device schedule == reference Gait, bit for bit
        assert TAU.dtype == np.float32
        dp, dv = np.abs(P - z["pos_targets"][sel]).max(), np.abs(V - z["vel_targets"][sel]).max()
        assert dp <= POS_TOL and dv <= VEL_TOL, (dp, dv)
        assert np.all(P[SS <= 0] == 0.0) and np.all(V[SS <= 0] == 0.0)         # stance legs: zero targets, exactly
        # torque tolerance from the terms of the map itself
        Jb = z["Jv_blocks"][sel].astype(np.float64)
        Rb = z["R_base"][sel].astype(np.float64)
        kp, kd = np.diag(robot.Kp_swing), np.diag(robot.Kd_swing)
        e_sw = kp * np.abs(np.einsum("rtij,rtlj->rtli", Rb, z["pos_targets"][sel] - z["base_pos_base_feet"][sel])) \
            + kd * np.abs(np.einsum("rtij,rtlj->rtli", Rb, z["vel_targets"][sel] - z["base_vel_base_feet"][sel])) + 50.0
        e_st = np.abs(z["contact_forces"][sel].astype(np.float64)).reshape(len(sel), T, 4, 3)
        e = np.where((SS != 0)[..., None], e_sw, e_st)
        tol = _tau_tol(Jb, e).reshape(len(sel), T, 12)
        dt = np.abs(TAU.astype(np.float64) - z["torque_cmds"][sel].astype(np.float64))
        assert np.all(dt <= tol), float((dt / tol).max())
        worst = dict(pos=max(worst["pos"], dp), vel=max(worst["vel"], dv), tau=max(worst["tau"], float(dt.max())))
    print("leg layer vs reference sequence:", worst)


def test_random_batch_matches_oracle_and_edge_cases():
    """One tick on 4 096 robots from arbitrary generator states (mid-swing, swing start, swing end, stance) against the
    oracle on a subsample; properties on the whole batch; argument errors."""
    from pympc_quadruped_b200 import _capi
    B, rng = 4096, np.random.default_rng(20261023)
    eng = _engine(A1Config, torch.float32)
    lp = _capi.make_leg_params(A1Config.Kp_swing, A1Config.Kd_swing, 0.1, 0.001, 9.81)
    f32x = lambda a: a.astype(np.float32).astype(np.float64)
    yaw = rng.uniform(-3, 3, B)
    Rb = np.zeros((B, 3, 3)); Rb[:, 0, 0] = np.cos(yaw); Rb[:, 0, 1] = -np.sin(yaw); Rb[:, 1, 0] = np.sin(yaw); Rb[:, 1, 1] = np.cos(yaw); Rb[:, 2, 2] = 1
    pos, vel = rng.uniform(-1, 1, (B, 3)), rng.uniform(-1, 1, (B, 3))
    thighs, feet_w = rng.uniform(-0.2, 0.2, (B, 4, 3)), rng.uniform(-1, 1, (B, 4, 3))
    ss = rng.choice([0.0, 0.0, 0.3, 0.77, 1.0, 1.0 + 1e-9], size=(B, 4)) * rng.choice([1.0, 0.5], size=(B, 4))
    ss[0] = [np.nan, 0.0, 0.5, 1.0]
    vdes, yr = rng.uniform(-1, 1.4, (B, 3)), rng.uniform(-1, 1, B)
    tsw, tst = rng.choice([0.1, 0.16, 0.24], B), rng.choice([0.1, 0.16, 0.08], B)
    active0 = rng.integers(0, 2, (B, 4)).astype(np.uint8)
    rem0, init0 = rng.uniform(0.0, 0.1, (B, 4)), rng.uniform(-1, 1, (B, 4, 3))
    fin0 = rng.uniform(-1, 1, (B, 4, 3))
    state = (_dev(active0, torch.uint8), _dev(rem0), _dev(init0), _dev(fin0))
    bpf, bvf = rng.uniform(-0.4, 0.4, (B, 4, 3)), rng.uniform(-2, 2, (B, 4, 3))
    Jb = rng.uniform(-0.4, 0.4, (B, 4, 3, 3))
    f = f32x(rng.uniform(-50, 120, (B, 12)))
    pt, vt = eng.swing_targets(lp, _dev(pos), _dev(vel), _dev(Rb), _dev(thighs), _dev(feet_w), _dev(ss), _dev(vdes), _dev(yr),
                               _dev(tsw), _dev(tst), state)
    tau18 = eng.leg_torques(lp, _dev(expand_jacobians(Jb)), _dev(Rb), _dev(bpf), _dev(bvf), _dev(f, torch.float32), _dev(ss), pt, vt)
    tau3 = eng.leg_torques(lp, _dev(Jb), _dev(Rb), _dev(bpf), _dev(bvf), _dev(f, torch.float32), _dev(ss), pt, vt)
    torch.cuda.synchronize()
    assert torch.equal(tau18, tau3)                                 # joint blocks alone == reference 18-column layout
    P, V, TAU = pt.cpu().numpy(), vt.cpu().numpy(), tau18.cpu().numpy()
    act1, rem1, init1, fin1 = (s.cpu().numpy() for s in state)
    sw = ss > 0                                                     # NaN: not swinging for the generator ...
    assert np.all(P[~sw] == 0) and np.all(V[~sw] == 0)
    assert np.array_equal(act1[~sw], active0[~sw]) and np.array_equal(rem1[~sw], rem0[~sw])          # untouched state
    assert np.array_equal(init1[~sw], init0[~sw]) and np.array_equal(fin1[~sw], fin0[~sw])
    assert np.array_equal(act1[sw], (ss[sw] < 1.0).astype(np.uint8))
    assert np.all(fin1[sw][:, 2] == -0.0255)
    assert np.all(np.isfinite(P)) and np.all(np.isfinite(V))
    # ... but truthy for the controller (`if swing_states[leg]:`): leg 0 of robot 0 takes the PD branch with zero targets
    e = -(A1Config.Kp_swing @ (Rb[0] @ bpf[0, 0]) + A1Config.Kd_swing @ (Rb[0] @ bvf[0, 0]))
    assert np.allclose(TAU[0, 0:3], Jb[0, 0].T @ e, rtol=1e-5, atol=1e-4)
    for b in list(range(48)) + list(rng.integers(0, B, 80)):
        if np.isnan(ss[b]).any():
            continue
        rd = types.SimpleNamespace(R_base=Rb[b], pos_base=pos[b], lin_vel_base=vel[b], base_pos_base_thighs=list(thighs[b]),
                                   pos_feet=list(feet_w[b]), base_pos_base_feet=list(bpf[b]), base_vel_base_feet=list(bvf[b]),
                                   Jv_feet=list(expand_jacobians(Jb[b])))
        gens = [OracleSwingFootTrajectoryGenerator(leg) for leg in range(4)]
        for leg, g in enumerate(gens):
            g.is_first_swing, g.remaining_swing_time = not active0[b, leg], rem0[b, leg]
            g.footpos_init, g.footpos_final = init0[b, leg].copy(), fin0[b, leg].copy()
        ctrl = OracleLegController(A1Config.Kp_swing, A1Config.Kd_swing)
        op, ov, otau = leg_layer_tick(gens, ctrl, rd, np.float64(tsw[b]), np.float64(tst[b]), ss[b], f[b], vdes[b], float(yr[b]))
        scale = 1.0 + np.abs(ov).max()
        assert np.abs(P[b] - op).max() <= POS_TOL and np.abs(V[b] - ov).max() <= VEL_TOL * scale, (b, np.abs(V[b] - ov).max())
        assert np.allclose(TAU[b], otau, rtol=2e-6, atol=2e-4), (b, np.abs(TAU[b] - otau).max())
        for leg, g in enumerate(gens):
            if ss[b, leg] > 0:
                assert abs(rem1[b, leg] - g.remaining_swing_time) <= 1e-15 and np.abs(fin1[b, leg] - g.footpos_final).max() <= POS_TOL
                assert np.array_equal(init1[b, leg], g.footpos_init)
    # empty batch and argument errors
    e0 = lambda *s, dt=torch.float64: torch.empty(s, dtype=dt, device=DEV)
    eng.swing_targets(lp, e0(0, 3), e0(0, 3), e0(0, 9), e0(0, 4, 3), e0(0, 4, 3), e0(0, 4), e0(0, 3), e0(0), e0(0), e0(0),
                      (e0(0, 4, dt=torch.uint8), e0(0, 4), e0(0, 4, 3), e0(0, 4, 3)))
    assert eng.leg_torques(lp, e0(0, 4, 3, 3), e0(0, 9), e0(0, 4, 3), e0(0, 4, 3), e0(0, 12, dt=torch.float32), e0(0, 4), e0(0, 4, 3),
                           e0(0, 4, 3)).shape == (0, 12)
    with pytest.raises(ValueError):
        eng.leg_torques(lp, e0(4, 4, 3, 12), e0(4, 9), e0(4, 4, 3), e0(4, 4, 3), e0(4, 12, dt=torch.float32), e0(4, 4), e0(4, 4, 3), e0(4, 4, 3))
    with pytest.raises(TypeError):
        eng.leg_torques(lp, e0(4, 4, 3, 3), e0(4, 9), e0(4, 4, 3), e0(4, 4, 3), e0(4, 12), e0(4, 4), e0(4, 4, 3), e0(4, 4, 3))
    assert eng.lib.mpcq_leg_torques(eng._h, 4, None, None, 7, None, None, None, None, None, None, None, None, None) == -1
    assert b"ncol" in eng.lib.mpcq_last_error(eng._h)
