"""GPU parity: the sm_100a path (through the C ABI) against the oracle on identical seeded inputs.

Tolerance (BASELINE.json north_star): |df| <= 1e-3 N or 1e-4 relative, KKT residual <= 1e-6 in
fp64 mode, constraint activity identical.  The oracle is the reference's own QP construction
(restated, pinned in test_oracle_golden.py) + an exact fp64 solve.
"""
import numpy as np
import pytest
import torch

from helpers import make_batch
from oracle.qp_exact import kkt_report
from pympc_quadruped_b200 import A1Config, AliengoConfig, Gait
from pympc_quadruped_b200 import _capi
from pympc_quadruped_b200.synth import GAIT_MIX

pytestmark = pytest.mark.gpu

ABS_TOL, REL_TOL = 1e-3, 1e-4

CASES = [
    # name, robot, horizon, B, regime, gaits, dtype, seed
    ("a1_trot_h10_f32", A1Config, 10, 96, "mixed", (Gait.TROTTING10,), torch.float32, 11),
    ("aliengo_mix_h10_f64", AliengoConfig, 10, 96, "mixed", GAIT_MIX, torch.float64, 12),
    ("aliengo_mix_h10_f32_aggr", AliengoConfig, 10, 64, "aggressive", GAIT_MIX, torch.float32, 13),
    ("a1_stand_h10_f32", A1Config, 10, 48, "mixed", (Gait.STANDING,), torch.float32, 14),
    ("a1_h16_gaits_f32", A1Config, 16, 32, "mixed", (Gait.TROTTING16, Gait.STANDING, Gait.JUMPING16, Gait.PACING16), torch.float32, 15),
    ("a1_trot_h30_f32", A1Config, 30, 24, "mixed", (Gait.TROTTING10,), torch.float32, 16),
    ("a1_trot_h30_f64", A1Config, 30, 16, "mixed", (Gait.TROTTING10,), torch.float64, 17),
    ("a1_stand_h30_f64", A1Config, 30, 6, "mixed", (Gait.STANDING,), torch.float64, 18),
]


def _engine(batch, robot, dtype):
    from pympc_quadruped_b200.engine import MpcqEngine
    return MpcqEngine(batch["cfg"], robot, dtype=dtype, device="cuda:0")


def _to_dev(batch, dtype):
    dev = "cuda:0"
    t = lambda a, dt: torch.as_tensor(a).to(device=dev, dtype=dt)
    return (t(batch["x0"], dtype), t(batch["feet"], dtype), t(batch["gait"], torch.float32), t(batch["xref"], dtype),
            t(batch["yaw"], dtype))


def _slacks(u, ub_fz, H, mu=0.7):
    """(slack to the lower bound, slack to the upper bound) of the reference's 20H rows C u in [0, ub]"""
    f = np.asarray(u, dtype=np.float64).reshape(4 * H, 3)
    cu = np.stack([f[:, 0] + mu * f[:, 2], -f[:, 0] + mu * f[:, 2], f[:, 1] + mu * f[:, 2], -f[:, 1] + mu * f[:, 2], f[:, 2]], axis=1)
    up = np.full_like(cu, np.inf)
    up[:, 4] = ub_fz - cu[:, 4]
    return cu.reshape(-1), up.reshape(-1)


def _activity(active_row, H):
    lo = ((active_row[:, None] >> np.arange(5)[None, :]) & 1).astype(bool).reshape(-1)
    up = np.zeros(20 * H, dtype=bool)
    up[4::5] = (active_row >> 5) & 1
    return lo, up


@pytest.mark.parametrize("name,robot,H,B,regime,gaits,dtype,seed", CASES, ids=[c[0] for c in CASES])
def test_solve_matches_oracle(name, robot, H, B, regime, gaits, dtype, seed):
    batch = make_batch(robot, H, B, regime, gaits, seed)
    eng = _engine(batch, robot, dtype)
    x0, feet, gait, xref, yaw = _to_dev(batch, dtype)
    res = eng.solve(x0, feet, gait, xref, yaw=yaw)
    torch.cuda.synchronize()
    u = res.u.double().cpu().numpy()
    f = res.forces.double().cpu().numpy()
    status = res.status.cpu().numpy()
    active = res.active.cpu().numpy()
    resid = res.resid.cpu().numpy()
    if dtype == torch.float64:
        Hdev, gdev, _ = eng.build_qp(x0, feet, gait, xref, yaw=yaw)
        Hdev, gdev = Hdev.cpu().numpy(), gdev.cpu().numpy()
    assert np.all(status & _capi.ST_VERIFIED), f"unverified envs: {np.flatnonzero(~(status & 1).astype(bool))} status {status}"
    assert not np.any(status & (_capi.ST_NUMERIC | _capi.ST_MAXITER))
    worst, n_rows, n_weak = 0.0, 0, 0
    ub_fz = batch["gait"].astype(np.float64) * 500.0
    for b in range(B):
        sol = batch["sols"][b]
        assert np.array_equal(f[b], u[b, :12])
        tol = np.maximum(ABS_TOL, REL_TOL * np.abs(sol.u).max())
        err = np.abs(u[b] - sol.u).max()
        worst = max(worst, err / tol)
        assert err <= tol, f"env {b}: |du| = {err:.3e} > {tol:.3e}"
        lo, up = _activity(active[b], H)
        n_rows += lo.size
        if not (np.array_equal(lo, sol.active_lower) and np.array_equal(up, sol.active_upper)):
            # The two QPs differ by the reference's float32 rounding of Su (~5e-7 relative), so a WEAKLY active
            # row (tight with a ~zero multiplier) may be tight in one optimum and a hair inside in the other.
            # Any row whose flag differs must be that case: (nearly) tight in BOTH solutions.
            slack_o = _slacks(sol.u, ub_fz[b], H)
            slack_g = _slacks(u[b], ub_fz[b], H)
            diff = np.flatnonzero((lo != sol.active_lower) | (up != sol.active_upper))
            scale = 1.0 + np.abs(sol.u).max()
            assert np.all(np.minimum(slack_o[0][diff], slack_o[1][diff]) <= 1e-4 * scale), f"env {b}: activity differs on a strongly inactive row"
            assert np.all(np.minimum(slack_g[0][diff], slack_g[1][diff]) <= 1e-4 * scale), f"env {b}: activity differs on a strongly inactive row"
            n_weak += diff.size
        # independent KKT check of the GPU point against the REFERENCE-constructed (H, g): the reference builds
        # Su in float32, so its (H, g) differ from the exact-arithmetic closed form by ~5e-7 relative, which
        # shows up as a ~1e-5 stationarity residual on ~50 N forces
        Hm, g, ub = batch["qps"][b]
        stat, prim, _, _ = kkt_report(Hm, g, 0.7, ub[4::5], u[b])
        assert prim <= (1e-9 if dtype == torch.float64 else 2e-7 * (1.0 + np.abs(u[b]).max()))   # f32: output rounding (float ulp of the forces)
        assert stat <= 5e-5 * (1.0 + np.abs(g).max())
        if dtype == torch.float64:                                     # KKT <= 1e-6 on the engine's own QP data
            stat, prim, _, _ = kkt_report(Hdev[b], gdev[b], 0.7, ub[4::5], u[b])
            assert stat <= 1e-6 and prim <= 1e-9
    assert n_weak <= max(2, n_rows // 2000), f"{n_weak} weakly-active rows differ out of {n_rows}"
    if dtype == torch.float64:
        assert resid[:, 0].max() <= 1e-6 and resid[:, 1].max() <= 1e-9
    print(f"{name}: worst err/tol = {worst:.3f}, factorisations mean {res.iters[:, 0].float().mean():.2f} "
          f"max {int(res.iters[:, 0].max())}, fallback envs {int((status & 2).astype(bool).sum())}")


def test_build_qp_matches_reference_construction():
    batch = make_batch(A1Config, 10, 16, "mixed", (Gait.TROTTING10, Gait.STANDING), 21)
    for dtype in (torch.float32, torch.float64):
        eng = _engine(batch, A1Config, dtype)
        x0, feet, gait, xref, yaw = _to_dev(batch, dtype)
        Hm, g, ub = eng.build_qp(x0, feet, gait, xref, yaw=yaw)
        torch.cuda.synchronize()
        Hm, g, ub = Hm.cpu().numpy(), g.cpu().numpy(), ub.cpu().numpy()
        for b in range(batch["B"]):
            Href, gref, ubref = batch["qps"][b]
            assert np.abs(Hm[b] - Href).max() <= 2e-6 * np.abs(Href).max()
            assert np.abs(g[b] - gref).max() <= 2e-6 * np.abs(gref).max()
            assert np.array_equal(ub[b], ubref.astype(np.float64))
            assert np.abs(Hm[b] - Hm[b].T).max() <= 1e-15 * np.abs(Hm[b]).max()


def test_golden_known_answers():
    """SURVEY 8c.3 known-answer states stored in tests/golden (generated from the unmodified reference)."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_h10.npz"))
    from pympc_quadruped_b200 import with_horizon
    from pympc_quadruped_b200.engine import MpcqEngine
    for robot_name, robot in (("A1", A1Config), ("Aliengo", AliengoConfig)):
        tags = sorted({k.split("/")[1] for k in z.keys() if k.startswith(robot_name + "/")} - {"seq"})
        for dtype in (torch.float32, torch.float64):
            eng = MpcqEngine(with_horizon(10), robot, dtype=dtype)
            t = lambda a, dt=dtype: torch.as_tensor(np.stack(a)).to(device="cuda:0", dtype=dt)
            x0 = t([z[f"{robot_name}/{g}/current_state"] for g in tags])
            feet = t([z[f"{robot_name}/{g}/pos_base_feet"].reshape(12) for g in tags])
            gait = t([z[f"{robot_name}/{g}/gait_table"][:40] for g in tags], torch.float32)
            xref = t([z[f"{robot_name}/{g}/x_ref"] for g in tags])
            yaw = t([z[f"{robot_name}/{g}/yaw"] for g in tags])
            res = eng.solve(x0, feet, gait, xref, yaw=yaw)
            u = res.u.double().cpu().numpy()
            for i, g in enumerate(tags):
                ustar = z[f"{robot_name}/{g}/u_star__oracle"]
                tol = max(ABS_TOL, REL_TOL * np.abs(ustar).max())
                assert np.abs(u[i] - ustar).max() <= tol, (robot_name, g)


def test_host_entry_point_equals_device_entry_point():
    batch = make_batch(A1Config, 10, 33, "mixed", (Gait.TROTTING10,), 31, solve=False)
    eng = _engine(batch, A1Config, torch.float32)
    x0, feet, gait, xref, yaw = _to_dev(batch, torch.float32)
    res = eng.solve(x0, feet, gait, xref, yaw=yaw)
    host = eng.solve_host(batch["x0"], batch["feet"], batch["gait"], batch["xref"], yaw=batch["yaw"],
                          want=("u", "iters", "resid", "status", "active"))
    assert np.array_equal(host["forces"], res.forces.cpu().numpy())
    assert np.array_equal(host["u"], res.u.cpu().numpy())
    assert np.array_equal(host["status"], res.status.cpu().numpy())
    assert np.array_equal(host["active"], res.active.cpu().numpy())
    assert np.array_equal(host["iters"], res.iters.cpu().numpy())


def test_host_entry_point_chunked_pinned_and_pageable(monkeypatch):
    """B large enough for the chunked multi-stream pipeline of mpcq_solve_host; pinned buffers are used for DMA
    directly, pageable ones are staged - all three routes must agree bit for bit."""
    B = 2500
    batch = make_batch(A1Config, 10, B, "mixed", (Gait.TROTTING10, Gait.STANDING), 32, solve=False)
    monkeypatch.setenv("MPCQ_HOST_CHUNKS", "2")                 # batches below 8 192 robots take one chunk by default
    eng = _engine(batch, A1Config, torch.float32)
    x0, feet, gait, xref, yaw = _to_dev(batch, torch.float32)
    res = eng.solve(x0, feet, gait, xref, yaw=yaw)
    host = eng.solve_host(batch["x0"], batch["feet"], batch["gait"], batch["xref"], yaw=batch["yaw"], want=("u", "status", "iters"))
    pin = lambda a, dt: torch.empty(a.shape, dtype=dt, pin_memory=True).copy_(torch.as_tensor(a).to(dt)).numpy()
    out = {"forces": torch.empty((B, 12), dtype=torch.float32, pin_memory=True).numpy(),
           "u": torch.empty((B, 120), dtype=torch.float32, pin_memory=True).numpy()}
    pinned = eng.solve_host(pin(batch["x0"], torch.float32), pin(batch["feet"], torch.float32), pin(batch["gait"], torch.float32),
                            pin(batch["xref"], torch.float32), yaw=pin(batch["yaw"], torch.float32), want=("u", "status"), out=out)
    assert pinned["forces"] is out["forces"]
    for r in (host, pinned):
        assert np.array_equal(r["forces"], res.forces.cpu().numpy())
        assert np.array_equal(r["u"], res.u.cpu().numpy())
        assert np.array_equal(r["status"], res.status.cpu().numpy())
    assert np.array_equal(host["iters"], res.iters.cpu().numpy())
    assert eng.last_launch_count == 4 * 2                       # (2 schedule launches + 2 size classes) x 2 chunks


def test_edge_cases_empty_batch_all_swing_and_errors():
    from pympc_quadruped_b200 import with_horizon
    from pympc_quadruped_b200.engine import MpcqEngine
    eng = MpcqEngine(with_horizon(10), A1Config)
    dev = "cuda:0"
    z = lambda *s: torch.zeros(s, device=dev)
    res = eng.solve(z(0, 13), z(0, 12), z(0, 40), z(0, 130))
    assert res.forces.shape == (0, 12)
    # all-swing contact table: u = 0, flagged
    batch = make_batch(A1Config, 10, 4, "nominal", (Gait.TROTTING10,), 41, solve=False)
    x0, feet, gait, xref, yaw = _to_dev(batch, torch.float32)
    res = eng.solve(x0, feet, torch.zeros_like(gait), xref, yaw=yaw)
    assert torch.all(res.forces == 0) and torch.all(res.status & _capi.ST_NO_STANCE)
    assert torch.all(res.active == 0x3F)
    with pytest.raises(ValueError):
        eng.solve(x0, feet, gait[:, :20], xref)
    with pytest.raises(TypeError):
        eng.solve(x0.double(), feet, gait, xref)
    with pytest.raises(ValueError):
        MpcqEngine(with_horizon(40), A1Config)


def test_schedule_does_not_change_results():
    """The expected-work-first launch order only reorders CTAs: results are bitwise those of the natural order."""
    batch = make_batch(A1Config, 10, 1500, "mixed", (Gait.TROTTING10, Gait.STANDING), 81, solve=False)
    from pympc_quadruped_b200.engine import MpcqEngine
    x0, feet, gait, xref, yaw = _to_dev(batch, torch.float32)
    a = MpcqEngine(batch["cfg"], A1Config).solve(x0, feet, gait, xref, yaw=yaw)
    eng = MpcqEngine(batch["cfg"], A1Config, schedule=-1)
    b = eng.solve(x0, feet, gait, xref, yaw=yaw)
    assert eng.last_launch_count == 2
    for k in ("forces", "u", "iters", "status", "active"):
        assert torch.equal(getattr(a, k), getattr(b, k)), k


def test_non_finite_inputs_are_flagged_not_propagated():
    """A robot with NaN / inf state must come back flagged MPCQ_ST_NUMERIC (never VERIFIED) without disturbing its neighbours."""
    batch = make_batch(A1Config, 10, 16, "mixed", (Gait.TROTTING10,), 71, solve=False)
    eng = _engine(batch, A1Config, torch.float32)
    x0, feet, gait, xref, yaw = _to_dev(batch, torch.float32)
    good = eng.solve(x0, feet, gait, xref, yaw=yaw).u.clone()
    x0b, feetb, xrefb = x0.clone(), feet.clone(), xref.clone()
    x0b[3, 9] = float("nan")
    feetb[7, 2] = float("inf")
    xrefb[11, 5] = float("nan")
    res = eng.solve(x0b, feetb, gait, xrefb, yaw=yaw)
    st = res.status.cpu().numpy()
    for b in (3, 7, 11):
        assert st[b] & _capi.ST_NUMERIC and not (st[b] & _capi.ST_VERIFIED), (b, st[b])
    keep = [b for b in range(16) if b not in (3, 7, 11)]
    assert np.all(st[keep] & _capi.ST_VERIFIED)
    assert torch.equal(res.u[keep], good[keep])


def test_internal_overflow_terminates_and_is_flagged():
    """Finite but absurd inputs that overflow inside the kernel must terminate (bounded loops, uniform control flow)."""
    batch = make_batch(A1Config, 10, 8, "mixed", (Gait.TROTTING10,), 72, solve=False)
    for dtype in (torch.float32, torch.float64):
        eng = _engine(batch, A1Config, dtype)
        x0, feet, gait, xref, yaw = _to_dev(batch, dtype)
        feet = feet.clone(); x0 = x0.clone()
        feet[1, :] = 1e25
        x0[2, 3:6] = 1e30
        res = eng.solve(x0, feet, gait, xref, yaw=yaw)
        st = res.status.cpu().numpy()
        u = res.u.double().cpu().numpy()
        for b in (1, 2):
            assert not (st[b] & _capi.ST_VERIFIED) or np.all(np.isfinite(u[b]))
        keep = [0, 3, 4, 5, 6, 7]
        assert np.all(st[keep] & _capi.ST_VERIFIED)


def test_batch_permutation_and_split_invariance():
    """Environments are independent: any split / order of the batch gives bitwise-equal results
    (this is what makes the multi-GPU sharding exact)."""
    batch = make_batch(A1Config, 10, 64, "mixed", (Gait.TROTTING10, Gait.STANDING), 51, solve=False)
    eng = _engine(batch, A1Config, torch.float32)
    x0, feet, gait, xref, yaw = _to_dev(batch, torch.float32)
    full = eng.solve(x0, feet, gait, xref, yaw=yaw).u.clone()
    perm = torch.randperm(64, device="cuda:0", generator=torch.Generator(device="cuda:0").manual_seed(0))
    p = eng.solve(x0[perm], feet[perm], gait[perm], xref[perm], yaw=yaw[perm]).u
    assert torch.equal(p, full[perm])
    lo = eng.solve(x0[:20], feet[:20], gait[:20], xref[:20], yaw=yaw[:20]).u.clone()
    hi = eng.solve(x0[20:], feet[20:], gait[20:], xref[20:], yaw=yaw[20:]).u
    assert torch.equal(torch.cat([lo, hi]), full)


def test_fused_assembly_equals_torch_statement_and_reference_sequence():
    """mpcq_assemble (state assembly + integrators + reference trajectory in one kernel) against the torch statement of the
    same arithmetic on the device, over a multi-tick sequence with MPC decimation; and the B = 1 adapter against the
    fixture recorded from the reference class."""
    import os
    from pympc_quadruped_b200 import with_horizon
    from pympc_quadruped_b200.controller import BatchedModelPredictiveController, BatchedRobotData, ModelPredictiveController
    from pympc_quadruped_b200.synth import synth_states, synth_gait_tables
    B, H = 257, 10
    cfg = with_horizon(H)
    st = synth_states(B, A1Config, "mixed", seed=61)
    for dtype in (torch.float32, torch.float64):
        from torch_statement import StatementEngine
        from pympc_quadruped_b200.configs import extract_mpc_constants
        fused = BatchedModelPredictiveController(cfg, A1Config, B, dtype=dtype)
        plain = BatchedModelPredictiveController(cfg, A1Config, B, dtype=dtype,
                                                 engine=StatementEngine(fused.engine, extract_mpc_constants(cfg, A1Config)))
        for tick in (0, 1, 2, 20, 21, 40):
            s = 1.0 + 0.002 * tick
            rd = BatchedRobotData(st["quat_base"], st["pos_base"] * s, st["ang_vel_base"], st["lin_vel_base"] * s, st["pos_base_feet"],
                                  st["R_base"] if tick % 2 else None)
            tabs = synth_gait_tables(B, H, (Gait.TROTTING10,), seed=tick)
            ff, fp = None, None
            for c in (fused, plain):
                c.update_robot_state(rd)
                f = c.update_mpc_if_needed(tick, st["vel_cmd_body"], st["yaw_rate_cmd"], tabs)
                ff, fp = (f, fp) if c is fused else (ff, f)
            assert torch.equal(fused.current_state.float(), plain.current_state.float())
            assert torch.equal(fused.ref_traj.float(), plain.ref_traj.float()), tick
            assert torch.allclose(fused._xy_des, plain._xy_des, rtol=0, atol=1e-15)
            assert torch.allclose(fused.yaw_desired, plain.yaw_desired, rtol=0, atol=1e-15)
            assert torch.allclose(fused._rp_init, plain._rp_init, rtol=0, atol=1e-15)
            assert torch.equal(ff, fp)
    # single-robot adapter on the real engine against the recorded reference sequence
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_h10.npz"))
    k = "A1/seq/"
    ctrl = ModelPredictiveController(cfg, A1Config, dtype=torch.float64)
    gt = Gait.TROTTING10.with_horizon(10)

    class RD:
        pass
    from oracle.mpc_oracle import quat_to_matrix
    for tick in range(61):
        b = (tick // 20) % 4
        rd = RD()
        rd.quat_base, rd.pos_base = z[k + "quat_base"][b], z[k + "pos_base"][b]
        rd.ang_vel_base, rd.lin_vel_base = z[k + "ang_vel_base"][b], z[k + "lin_vel_base"][b]
        rd.pos_base_feet = [z[k + "pos_base_feet"][b, i] for i in range(4)]
        rd.R_base = quat_to_matrix(rd.quat_base)
        gt.set_iteration(20, tick)
        ctrl.update_robot_state(rd)
        f = ctrl.update_mpc_if_needed(tick, z[k + "vel_cmd_body"][b], float(z[k + "yaw_rate_cmd"][b]), gt.get_gait_table())
        ref = z[k + "forces__oracle_solver"][tick]
        assert np.abs(f - ref).max() <= max(1e-3, 1e-4 * np.abs(ref).max()), tick
        assert np.abs(ctrl.ref_traj - z[k + "ref_traj"][tick]).max() <= 1.2e-7 * max(1.0, np.abs(z[k + "ref_traj"][tick]).max())


FULL_SIZE = [
    # BASELINE.json configs[1], [2], [3] at their full sizes: name, robot, H, B, gaits, dtype, seed, oracle sample
    ("cfg2_a1_trot_4096_f32", A1Config, 10, 4096, (Gait.TROTTING10,), torch.float32, 21, 48),
    ("cfg3_aliengo_mix_16384_f64", AliengoConfig, 10, 16384, GAIT_MIX, torch.float64, 22, 48),
    ("cfg4_a1_h30_4096_f32", A1Config, 30, 4096, (Gait.TROTTING10,), torch.float32, 23, 12),
]


@pytest.mark.parametrize("name,robot,H,B,gaits,dtype,seed,nsample", FULL_SIZE, ids=[c[0] for c in FULL_SIZE])
def test_full_size_properties(name, robot, H, B, gaits, dtype, seed, nsample):
    """BASELINE sizes: size-independent properties on every environment (verified status, fp64 KKT residuals, feasibility
    of the returned forces recomputed here from u, swing forces exactly zero, f = u[:12]), agreement of the fp32 and fp64
    modes within the parity tolerance, and the oracle on a random subsample."""
    batch = make_batch(robot, H, B, "mixed", gaits, seed, solve=False)
    eng = _engine(batch, robot, dtype)
    x0, feet, gait, xref, yaw = _to_dev(batch, dtype)
    res = eng.solve(x0, feet, gait, xref, yaw=yaw)
    torch.cuda.synchronize()
    u = res.u.double().cpu().numpy()
    status = res.status.cpu().numpy()
    resid = res.resid.cpu().numpy()
    assert np.all(status & _capi.ST_VERIFIED) and not np.any(status & (_capi.ST_NUMERIC | _capi.ST_MAXITER))
    assert np.array_equal(res.forces.double().cpu().numpy(), u[:, :12])
    mu, fz_max = float(eng.consts["mu"]), float(eng.consts["fz_max"])
    f = u.reshape(B, 4 * H, 3)
    stance = batch["gait"].reshape(B, 4 * H) > 0
    assert np.all(f[~stance] == 0.0)                                   # swing foot-steps carry exactly no force
    out_tol = 1e-9 if dtype == torch.float64 else 2e-5                 # f32 mode: rounding of ~100 N outputs to float
    assert np.all(f[..., 2] >= -out_tol) and np.all(f[..., 2] <= fz_max + out_tol)
    assert np.all(np.abs(f[..., 0]) <= mu * f[..., 2] + out_tol) and np.all(np.abs(f[..., 1]) <= mu * f[..., 2] + out_tol)
    gscale = 1.0 + np.abs(u).max()
    assert resid[:, 1].max() <= 1e-9 * gscale                          # primal violation measured on the fp64 iterate
    assert resid[:, 0].max() <= (1e-6 if dtype == torch.float64 else 1e-4)   # reduced gradient (f32: 2 min(R) * 2e-4 N cap)
    # the other arithmetic mode solves the same QPs: same optimum within the parity tolerance
    other = torch.float64 if dtype == torch.float32 else torch.float32
    eng2 = _engine(batch, robot, other)
    a2 = _to_dev(batch, other)
    u2 = eng2.solve(a2[0], a2[1], a2[2], a2[3], yaw=a2[4]).u.double().cpu().numpy()
    tol = np.maximum(ABS_TOL, REL_TOL * np.abs(u).max(axis=1))
    assert np.all(np.abs(u - u2).max(axis=1) <= tol)
    # oracle on a random subsample
    from oracle.mpc_oracle import OracleMPC, RobotState
    from oracle.qp_exact import solve_qp_exact
    from pympc_quadruped_b200.synth import synth_states
    st = synth_states(B, robot, "mixed", seed=seed)
    rng = np.random.default_rng(seed)
    worst = 0.0
    for b in rng.choice(B, size=nsample, replace=False):
        m = OracleMPC(batch["cfg"], robot)
        m.update_robot_state(RobotState(st["quat_base"][b], st["pos_base"][b], st["ang_vel_base"][b], st["lin_vel_base"][b],
                                        st["pos_base_feet"][b]))
        Hm, g, _, _, ub = m.build_qp(batch["xref"][b], batch["gait"][b])
        sol = solve_qp_exact(Hm, g, m.mu, ub[4::5])
        assert sol.verified
        t = max(ABS_TOL, REL_TOL * np.abs(sol.u).max())
        err = np.abs(u[b] - sol.u).max()
        assert err <= t, f"env {b}: |du| = {err:.3e} > {t:.3e}"
        worst = max(worst, err / t)
    print(f"{name}: all {B} verified, worst err/tol on {nsample} oracle samples {worst:.3f}, "
          f"factorisations mean {res.iters[:, 0].float().mean():.2f} max {int(res.iters[:, 0].max())}")


def test_measure_peaks_reports_plausible_numbers():
    import ctypes
    lib = _capi.load_library()
    out = (ctypes.c_double * 4)()
    assert lib.mpcq_measure_peaks(0, out) == 0
    fp32, fp64, smem, sms = out[0], out[1], out[2], out[3]
    assert sms >= 100 and 20.0 < fp32 < 200.0 and 5.0 < fp64 < 100.0 and 5e3 < smem < 1e5, (fp32, fp64, smem, sms)
    assert lib.mpcq_measure_peaks(-1, out) == -1 and lib.mpcq_measure_peaks(0, None) == -1


def test_device_gait_tables_equal_the_reference_schedule():
    """mpcq_gait_tables against the per-robot schedule objects (pinned to the reference's Gait in
    test_oracle_golden.py): contact table bit for bit, swing / stance phase states exactly (same float32 phase, float64
    arithmetic), robots out of phase, every named pattern."""
    from pympc_quadruped_b200 import BatchedGaitSchedule
    from pympc_quadruped_b200.engine import MpcqEngine
    from pympc_quadruped_b200.configs import with_horizon
    names = ["STANDING", "TROTTING16", "TROTTING10", "JUMPING16", "PACING16", "PACING10", "BOUNDING10"]
    for H in (10, 16):
        eng = MpcqEngine(with_horizon(H), A1Config, dtype=torch.float32, device="cuda:0")
        rng = np.random.default_rng(H)
        scheds = [getattr(Gait, names[i % len(names)]).with_horizon(H) for i in range(70)]
        bg = BatchedGaitSchedule(eng, scheds)
        for ibm in (20, 7):
            cur = rng.integers(0, 5000, size=70)
            bg.set_iteration(ibm, torch.as_tensor(cur, device="cuda:0"))
            torch.cuda.synchronize()
            tab = bg.get_gait_table().cpu().numpy()
            sw, stn = bg.get_swing_state().cpu().numpy(), bg.get_stance_state().cpu().numpy()
            for b, s in enumerate(scheds):
                s.set_iteration(ibm, int(cur[b]))
                assert np.array_equal(tab[b], s.get_gait_table()), (H, ibm, b)
                with np.errstate(all="ignore"):
                    assert np.array_equal(sw[b], np.asarray(s.get_swing_state(), dtype=np.float64), equal_nan=True), (H, ibm, b)
                    assert np.array_equal(stn[b], np.asarray(s.get_stance_state(), dtype=np.float64), equal_nan=True), (H, ibm, b)
        bg.set_iteration(20, 40)                                   # one tick for all robots
        s = scheds[2]; s.set_iteration(20, 40)
        assert np.array_equal(bg.get_gait_table()[2].cpu().numpy(), s.get_gait_table())
    # the table feeds the solver directly
    batch = make_batch(A1Config, 10, 8, "mixed", (Gait.TROTTING10,), 31)
    eng = _engine(batch, A1Config, torch.float32)
    bg = BatchedGaitSchedule(eng, [Gait.TROTTING10.with_horizon(10)] * 8)
    bg.set_iteration(20, 60)
    x0, feet, _, xref, yaw = _to_dev(batch, torch.float32)
    res = eng.solve(x0, feet, bg.get_gait_table(), xref, yaw=yaw)
    assert np.all(res.status.cpu().numpy() & _capi.ST_VERIFIED)


def test_warm_start_same_optimum_one_factorisation():
    """mpcq_set_warm_start on the device: a solve's own faces are verified with ONE factorisation; arbitrary guesses end at
    the same (unique) optimum; zeroed faces are the cold start bit for bit."""
    batch = make_batch(A1Config, 10, 256, "mixed", (Gait.TROTTING10,), 51, solve=False)
    eng = _engine(batch, A1Config, torch.float32)
    x0, feet, gait, xref, yaw = _to_dev(batch, torch.float32)
    cold = eng.solve(x0, feet, gait, xref, yaw=yaw)
    faces = torch.full((256, 40), 0xEE, dtype=torch.uint8, device="cuda:0")
    zero = torch.zeros_like(faces)
    c2 = eng.solve(x0, feet, gait, xref, yaw=yaw, faces_in=zero, faces_out=faces)
    assert torch.equal(c2.u, cold.u) and torch.equal(c2.iters, cold.iters)
    assert torch.all(faces[gait == 0] == 0) and torch.all((faces & 0xC0) == 0)
    out2 = torch.empty_like(faces)
    warm = eng.solve(x0, feet, gait, xref, yaw=yaw, faces_in=faces, faces_out=out2)
    assert torch.all(warm.status & _capi.ST_VERIFIED) and torch.all(warm.iters[:, 0] == 1)
    assert torch.equal(out2, faces)
    assert (warm.u.double() - cold.u.double()).abs().max() <= 2e-4
    assert float(cold.iters[:, 0].float().mean()) > 3.0
    guess = torch.randint(0, 256, faces.shape, dtype=torch.uint8, device="cuda:0")
    rnd = eng.solve(x0, feet, gait, xref, yaw=yaw, faces_in=guess)
    assert torch.all(rnd.status & _capi.ST_VERIFIED)
    tol = torch.clamp(REL_TOL * cold.u.double().abs().amax(dim=1), min=ABS_TOL)
    assert torch.all((rnd.u.double() - cold.u.double()).abs().amax(dim=1) <= tol)
    # the handle is back to cold starts afterwards
    again = eng.solve(x0, feet, gait, xref, yaw=yaw)
    assert torch.equal(again.u, cold.u)


def test_controller_warm_start_tracks_the_cold_controller():
    """BatchedModelPredictiveController(warm_start=True) over a drifting state sequence with an advancing gait: same forces
    as the cold controller within the parity tolerance, fewer factorisations."""
    from pympc_quadruped_b200.controller import BatchedModelPredictiveController, BatchedRobotData
    from pympc_quadruped_b200.configs import with_horizon
    from pympc_quadruped_b200.synth import synth_states
    from pympc_quadruped_b200 import BatchedGaitSchedule
    B, H = 512, 10
    st = synth_states(B, A1Config, "mixed", seed=61)
    rng = np.random.default_rng(61)
    drift = {k: rng.normal(size=st[k].shape) for k in ("pos_base", "ang_vel_base", "lin_vel_base")}
    ctrls = [BatchedModelPredictiveController(with_horizon(H), A1Config, B, warm_start=w) for w in (False, True)]
    gaits = BatchedGaitSchedule(ctrls[0].engine, [Gait.TROTTING10.with_horizon(H)] * B)
    dv = lambda a: torch.as_tensor(a, device="cuda:0")
    nfac = [[], []]
    for t in range(8):
        rd = BatchedRobotData(dv(st["quat_base"]), dv(st["pos_base"] + 0.001 * t * drift["pos_base"]),
                              dv(st["ang_vel_base"] + 0.01 * t * drift["ang_vel_base"]),
                              dv(st["lin_vel_base"] + 0.01 * t * drift["lin_vel_base"]), dv(st["pos_base_feet"]), dv(st["R_base"]))
        it = t * ctrls[0].iterations_between_mpc
        gaits.set_iteration(ctrls[0].iterations_between_mpc, it)
        forces = []
        for i, c in enumerate(ctrls):
            c.update_robot_state(rd)
            forces.append(c.update_mpc_if_needed(it, dv(st["vel_cmd_body"]), dv(st["yaw_rate_cmd"]), gaits.get_gait_table()).double())
            assert torch.all(c.last_result.status & _capi.ST_VERIFIED)
            nfac[i].append(float(c.last_result.iters[:, 0].float().mean()))
        tol = torch.clamp(REL_TOL * forces[0].abs().amax(dim=1), min=ABS_TOL)
        assert torch.all((forces[0] - forces[1]).abs().amax(dim=1) <= tol), t
    # (the states here do not follow the model's prediction, so the shifted faces are only a rough guess)
    assert np.mean(nfac[1][1:]) <= np.mean(nfac[0][1:]) + 0.05, nfac
    print(f"factorisations per update: cold {np.mean(nfac[0][1:]):.2f}, warm {np.mean(nfac[1][1:]):.2f}")


def test_warm_start_on_a_model_consistent_rollout():
    """Closed loop in the MPC's own model: the state advances one horizon step under the first-step forces (Ad, Bd of the
    reference construction), the contact table and the reference trajectory shift by one step.  By the principle of
    optimality the shifted previous faces are then (nearly) the new optimum's faces: the warm-started solves must give the
    cold solves' forces with far fewer factorisations."""
    from oracle.mpc_oracle import state_space_model, discretize
    from pympc_quadruped_b200.gait import gait_tables
    from pympc_quadruped_b200.synth import synth_gait_params
    B, H, T = 96, 10, 6
    batch = make_batch(A1Config, H, B, "mixed", (Gait.TROTTING10,), 71, solve=False)
    eng = _engine(batch, A1Config, torch.float32)
    off, dur, seg, it0 = synth_gait_params(B, (Gait.TROTTING10,), 71)
    x0 = batch["x0"].astype(np.float64); feet = batch["feet"].astype(np.float64).copy(); xref = batch["xref"].astype(np.float64).reshape(B, H, 13).copy()
    yaw = batch["yaw"].astype(np.float64).copy()
    inertia = np.asarray(A1Config.base_inertia_base, dtype=np.float32)
    faces = torch.zeros((B, 4 * H), dtype=torch.uint8, device="cuda:0")
    fin = torch.zeros_like(faces)
    t32 = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda:0")
    nf_cold, nf_warm = [], []
    for t in range(T):
        gait = gait_tables(off, dur, seg, it0 + t, H)
        args = (t32(x0), t32(feet), t32(gait), t32(xref.reshape(B, 13 * H)))
        cold = eng.solve(*args, yaw=t32(yaw))
        fin[:, :4 * (H - 1)] = faces[:, 4:]
        fin[:, 4 * (H - 1):] = faces[:, 4 * (H - 1):]
        warm = eng.solve(*args, yaw=t32(yaw), faces_in=fin if t > 0 else None, faces_out=faces)
        assert torch.all(cold.status & _capi.ST_VERIFIED) and torch.all(warm.status & _capi.ST_VERIFIED)
        tol = torch.clamp(REL_TOL * cold.u.double().abs().amax(dim=1), min=ABS_TOL)
        assert torch.all((warm.u.double() - cold.u.double()).abs().amax(dim=1) <= tol), t
        nf_cold.append(float(cold.iters[:, 0].float().mean())); nf_warm.append(float(warm.iters[:, 0].float().mean()))
        u0 = cold.forces.double().cpu().numpy()
        for b in range(B):                                    # one model step under the applied forces
            Ac, Bc = state_space_model(yaw[b], feet[b].reshape(4, 3), inertia, float(A1Config.mass_base))
            Ad, Bd = discretize(Ac, Bc, 0.05)
            xn = Ad.astype(np.float64) @ x0[b] + Bd.astype(np.float64) @ u0[b]
            feet[b] -= np.tile(xn[3:6] - x0[b, 3:6], 4)       # feet stay where they are in the world
            x0[b] = xn
            yaw[b] = xn[2]
        xref[:, :-1] = xref[:, 1:].copy()
        xref[:, -1, 2:5] += xref[:, -1, 2:5] - xref[:, -3, 2:5] if H > 2 else 0.0
    print(f"model-consistent rollout: factorisations per update cold {np.mean(nf_cold[1:]):.2f}, warm {np.mean(nf_warm[1:]):.2f}")
    assert np.mean(nf_warm[1:]) < 0.75 * np.mean(nf_cold[1:]), (nf_cold, nf_warm)


def test_solve_is_graph_capturable():
    """include/mpcq.h promises that mpcq_solve stays ONE ordered, capturable operation on the caller's stream although it forks
    its size classes onto private streams: capture it in a CUDA graph, replay on new inputs, compare bit for bit with eager calls."""
    B = 1024
    sets = [make_batch(A1Config, 10, B, "mixed", (Gait.TROTTING10, Gait.STANDING), 41 + i, solve=False) for i in range(2)]
    eng = _engine(sets[0], A1Config, torch.float32)
    dev = [_to_dev(b, torch.float32) for b in sets]
    eager = []
    for x0, feet, gait, xref, yaw in dev:
        r = eng.solve(x0, feet, gait, xref, yaw=yaw, want=("u", "status"))
        eager.append((r.forces.clone(), r.u.clone(), r.status.clone()))
    static = [t.clone() for t in dev[0]]
    out = eng.solve(*static[:4], yaw=static[4], want=("u", "status"))          # warm-up: buffers of the handle are allocated
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        eng.solve(*static[:4], yaw=static[4], want=("u", "status"), out=out)
    for k in (1, 0, 1):
        for dst, src in zip(static, dev[k]):
            dst.copy_(src)
        out.forces.zero_(); out.u.zero_(); out.status.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out.forces, eager[k][0]) and torch.equal(out.u, eager[k][1]) and torch.equal(out.status, eager[k][2]), k
    # replay timing next to eager launches (information only)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    res = {}
    for name, fn in (("eager", lambda: eng.solve(*static[:4], yaw=static[4], want=("u", "status"), out=out)), ("graph", g.replay)):
        for _ in range(5):
            fn()
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 50
    print("mpcq_solve, 1024 robots, ms per call:", res)
