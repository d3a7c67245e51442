"""CPU tests of the DEVICE code: pympc_quadruped_b200/csrc/mpcq_core.cuh compiled by g++ with the
warp emulated as 32 lock-step coroutines (tests/emu).  Same source, same arithmetic as the sm_100a
kernels apart from the lane scheduling, so these tests pin the kernel logic (model assembly,
panel Cholesky, triangular solves, face updates, fallback) against the oracle without a GPU.
The emulator is test infrastructure: the package never loads it."""
import ctypes as C

import numpy as np
import pytest

from helpers import make_batch
from pympc_quadruped_b200 import A1Config, AliengoConfig, Gait, _capi
from pympc_quadruped_b200.configs import extract_mpc_constants
from pympc_quadruped_b200.synth import GAIT_MIX


def emu_solve(lib, batch, robot, f64, **knobs):
    B, H = batch["B"], batch["horizon"]
    cfg = _capi.make_config(extract_mpc_constants(batch["cfg"], robot), _capi.MPCQ_F64 if f64 else _capi.MPCQ_F32, **knobs)
    rt = np.float64 if f64 else np.float32
    x0, yaw, feet, xref = (batch[k].astype(rt) for k in ("x0", "yaw", "feet", "xref"))
    gait = np.ascontiguousarray(batch["gait"], dtype=np.float32)
    out = dict(f=np.zeros((B, 12), rt), u=np.zeros((B, 12 * H), rt), iters=np.zeros((B, 2), np.int32),
               resid=np.zeros((B, 2)), status=np.zeros(B, np.int32), active=np.zeros((B, 4 * H), np.uint8))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    fn = lib.mpcq_emu_solve_f64 if f64 else lib.mpcq_emu_solve_f32
    rc = fn(C.byref(cfg), B, p(x0), p(yaw), p(feet), p(gait), p(xref), p(out["f"]), p(out["u"]), p(out["iters"]),
            p(out["resid"]), p(out["status"]), p(out["active"]))
    assert rc == 0
    return out


def check_against_oracle(out, batch, f64):
    H = batch["horizon"]
    assert np.all(out["status"] & _capi.ST_VERIFIED)
    for b, sol in enumerate(batch["sols"]):
        tol = max(1e-3, 1e-4 * np.abs(sol.u).max())
        assert np.abs(out["u"][b].astype(np.float64) - sol.u).max() <= tol, b
        assert np.array_equal(out["f"][b], out["u"][b, :12])
        a = out["active"][b]
        lo = ((a[:, None] >> np.arange(5)[None, :]) & 1).astype(bool).reshape(-1)
        up = np.zeros(20 * H, dtype=bool)
        up[4::5] = (a >> 5) & 1
        assert np.array_equal(lo, sol.active_lower) and np.array_equal(up, sol.active_upper), b
    assert out["resid"][:, 1].max() <= 1e-9
    assert out["resid"][:, 0].max() <= (1e-9 if f64 else 1e-6)


CASES = [
    ("a1_trot_h10_f32", A1Config, 10, 24, "mixed", (Gait.TROTTING10,), False, 101),
    ("aliengo_mix_h10_f64", AliengoConfig, 10, 24, "mixed", GAIT_MIX, True, 102),
    ("aliengo_mix_h10_f32_aggressive", AliengoConfig, 10, 16, "aggressive", GAIT_MIX, False, 103),
    ("a1_stand_h10_f32", A1Config, 10, 8, "mixed", (Gait.STANDING,), False, 104),
    ("a1_h16_gaits_f64", A1Config, 16, 8, "mixed", (Gait.TROTTING16, Gait.JUMPING16, Gait.PACING16), True, 105),
    ("a1_trot_h30_f64", A1Config, 30, 3, "mixed", (Gait.TROTTING10,), True, 106),
]


@pytest.mark.parametrize("name,robot,H,B,regime,gaits,f64,seed", CASES, ids=[c[0] for c in CASES])
def test_device_code_matches_oracle(emu_lib, name, robot, H, B, regime, gaits, f64, seed):
    batch = make_batch(robot, H, B, regime, gaits, seed)
    out = emu_solve(emu_lib, batch, robot, f64)
    check_against_oracle(out, batch, f64)


def test_fallback_path_is_exact(emu_lib):
    """Force the primal active-set fallback (no primal-dual rounds allowed): same optimum."""
    batch = make_batch(A1Config, 10, 8, "aggressive", (Gait.TROTTING10,), 111)
    out = emu_solve(emu_lib, batch, A1Config, True, max_pdas_rounds=1)
    check_against_oracle(out, batch, True)
    assert np.any(out["status"] & _capi.ST_FALLBACK)


def test_all_swing_and_single_stance(emu_lib):
    batch = make_batch(A1Config, 10, 3, "nominal", (Gait.TROTTING10,), 112, solve=False)
    batch["gait"][0] = 0.0                      # no stance at all
    batch["gait"][1] = 0.0
    batch["gait"][1, 4 * 3 + 2] = 1.0           # one single stance foot-step
    out = emu_solve(emu_lib, batch, A1Config, False)
    assert out["status"][0] & _capi.ST_NO_STANCE and np.all(out["u"][0] == 0) and np.all(out["active"][0] == 0x3F)
    assert out["status"][1] & _capi.ST_VERIFIED
    nz = np.flatnonzero(out["u"][1])
    assert set(nz) <= {3 * 14, 3 * 14 + 1, 3 * 14 + 2}


def test_iteration_cap_is_reported(emu_lib):
    batch = make_batch(A1Config, 10, 6, "aggressive", (Gait.STANDING,), 113, solve=False)
    out = emu_solve(emu_lib, batch, A1Config, False, max_pdas_rounds=1, max_as_iter=1)
    bad = ~(out["status"] & _capi.ST_VERIFIED).astype(bool)
    assert bad.any() and np.all(out["status"][bad] & _capi.ST_MAXITER)
    assert out["resid"][:, 1].max() <= 1e-9      # the returned point is still feasible


def test_non_finite_inputs_exit_early_and_uniformly(emu_lib):
    """NaN / inf inputs: flagged MPCQ_ST_NUMERIC, zero forces, and - checked by the emulator on every run - no lane
    may diverge around a collective (on hardware that is a hang, which is how this case was found)."""
    batch = make_batch(A1Config, 10, 6, "mixed", (Gait.TROTTING10,), 114, solve=False)
    batch["x0"][1, 9] = np.nan
    batch["feet"][2, 2] = np.inf
    batch["xref"][4, 5] = np.nan
    out = emu_solve(emu_lib, batch, A1Config, False)
    for b in (1, 2, 4):
        assert out["status"][b] == _capi.ST_NUMERIC and np.all(out["u"][b] == 0)
    for b in (0, 3, 5):
        assert out["status"][b] & _capi.ST_VERIFIED


def test_internal_overflow_is_flagged(emu_lib):
    """Finite but absurd inputs that overflow inside (foot lever 1e25 m): must end flagged, never verified, without
    divergent lanes (the emulator aborts on divergence)."""
    batch = make_batch(A1Config, 10, 4, "mixed", (Gait.TROTTING10,), 115, solve=False)
    batch["feet"][1, :] = 1e25
    batch["x0"][2, 3:6] = 1e30
    for f64 in (False, True):
        out = emu_solve(emu_lib, batch, A1Config, f64)
        for b in (1, 2):
            assert not (out["status"][b] & _capi.ST_VERIFIED) or np.all(np.isfinite(out["u"][b]))
        for b in (0, 3):
            assert out["status"][b] & _capi.ST_VERIFIED


def emu_solve_warm(lib, batch, robot, faces_in, **knobs):
    B, H = batch["B"], batch["horizon"]
    cfg = _capi.make_config(extract_mpc_constants(batch["cfg"], robot), _capi.MPCQ_F32, **knobs)
    rt = np.float32
    x0, yaw, feet, xref = (batch[k].astype(rt) for k in ("x0", "yaw", "feet", "xref"))
    gait = np.ascontiguousarray(batch["gait"], dtype=np.float32)
    out = dict(f=np.zeros((B, 12), rt), u=np.zeros((B, 12 * H), rt), iters=np.zeros((B, 2), np.int32),
               resid=np.zeros((B, 2)), status=np.zeros(B, np.int32), active=np.zeros((B, 4 * H), np.uint8),
               faces=np.full((B, 4 * H), 0xEE, np.uint8))
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    rc = lib.mpcq_emu_solve_warm_f32(C.byref(cfg), B, p(x0), p(yaw), p(feet), p(gait), p(xref), p(out["f"]), p(out["u"]),
                                     p(out["iters"]), p(out["resid"]), p(out["status"]), p(out["active"]), p(faces_in), p(out["faces"]))
    assert rc == 0
    return out


def test_warm_start_same_optimum_fewer_rounds(emu_lib):
    """mpcq_set_warm_start: the faces a solve returns, handed back, are verified with ONE factorisation; a wrong guess
    (all apex / random codes) still ends at the same (unique) optimum; swing foot-steps report code 0."""
    batch = make_batch(A1Config, 10, 10, "aggressive", (Gait.TROTTING10,), 41)
    cold = emu_solve_warm(emu_lib, batch, A1Config, None)
    check_against_oracle(cold, batch, False)
    assert np.array_equal(cold["u"], emu_solve(emu_lib, batch, A1Config, False)["u"])        # NULL faces_in = the cold path
    assert cold["iters"][:, 0].max() > 2
    stance = batch["gait"] > 0
    assert np.all(cold["faces"][~stance] == 0) and np.all((cold["faces"] & 0xC0) == 0)
    warm = emu_solve_warm(emu_lib, batch, A1Config, cold["faces"])
    check_against_oracle(warm, batch, False)
    assert np.all(warm["iters"][:, 0] == 1), warm["iters"][:, 0]
    assert np.array_equal(warm["faces"], cold["faces"])
    assert np.abs(warm["u"].astype(np.float64) - cold["u"]).max() <= 2e-4
    rng = np.random.default_rng(3)
    for guess in (np.full_like(cold["faces"], 0x30), rng.integers(0, 256, size=cold["faces"].shape).astype(np.uint8)):
        out = emu_solve_warm(emu_lib, batch, A1Config, guess)
        check_against_oracle(out, batch, False)                  # same optimum, same constraint activity
        # the face codes may differ only on weakly active rows (tight with a zero multiplier: in or out of the face)
        assert (out["faces"] != cold["faces"]).sum() <= 2


def test_leg_layer_device_code_reproduces_the_reference_sequence(emu_lib):
    """The device bodies of `mpcq_swing_targets` / `mpcq_leg_torques` (csrc/mpcq_legs.cuh, compiled for the host) over the
    340-tick x 8-robot sequence recorded from the unmodified reference classes (tests/golden/reference_legs.npz).
    Tolerances as in tests/test_gpu_legs.py: 1e-8 m, 1e-7 m/s, 2 float32 ulp of the largest torque term."""
    import ctypes as C
    import os
    from oracle.leg_oracle import expand_jacobians
    from pympc_quadruped_b200 import _capi
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_legs.npz"))
    R, T = z["pos_base"].shape[:2]
    d = {k: np.ascontiguousarray(z[k].astype(np.float64)) for k in z.files if z[k].dtype == np.float32 and k != "torque_cmds"}
    Jv = np.ascontiguousarray(expand_jacobians(d["Jv_blocks"]))
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    emu_lib.mpcq_emu_swing_targets.argtypes = [C.c_int, C.POINTER(_capi.MpcqLegParams)] + [C.c_void_p] * 16
    emu_lib.mpcq_emu_leg_torques.argtypes = [C.c_int, C.POINTER(_capi.MpcqLegParams), C.c_void_p, C.c_int] + [C.c_void_p] * 4 + \
        [C.c_int] + [C.c_void_p] * 4
    emu_lib.mpcq_emu_swing_targets.restype = emu_lib.mpcq_emu_leg_torques.restype = None
    kp = {0: np.diag([700.0] * 3), 1: np.diag([200.0] * 3)}                # robot 0 = A1Config, 1 = AliengoConfig
    worst = [0.0, 0.0, 0.0]
    for r in range(R):                                                     # one robot per "batch" (per-robot gains)
        lp = _capi.make_leg_params(kp[int(z["robot"][r])], np.diag([20.0] * 3), 0.1, 0.001, 9.81)
        active, rem = np.zeros((1, 4), np.uint8), np.zeros((1, 4))
        init, fin = np.zeros((1, 4, 3)), np.zeros((1, 4, 3))
        tsw, tst = np.array([z["swing_stance_time"][r, 0]]), np.array([z["swing_stance_time"][r, 1]])
        vdes, yr = np.ascontiguousarray(d["v_des"][r:r + 1]), np.ascontiguousarray(d["yaw_rate"][r:r + 1])
        for t in range(T):
            g = lambda k: np.ascontiguousarray(d[k][r, t])
            ss = np.ascontiguousarray(z["swing_state"][r, t])
            pt, vt, tau = np.empty((1, 4, 3)), np.empty((1, 4, 3)), np.empty((1, 12), np.float32)
            Rb, bpf, bvf = g("R_base"), g("base_pos_base_feet"), g("base_vel_base_feet")
            ins = [g("pos_base"), g("lin_vel_base"), Rb, g("base_pos_base_thighs"), g("pos_feet"), ss, vdes, yr, tsw, tst]
            emu_lib.mpcq_emu_swing_targets(1, C.byref(lp), *[P(a) for a in ins], P(active), P(rem), P(init), P(fin), P(pt), P(vt))
            for f64 in (0, 1):
                f = np.ascontiguousarray(d["contact_forces"][r, t].astype(np.float64 if f64 else np.float32))
                Jt = np.ascontiguousarray(Jv[r, t])
                emu_lib.mpcq_emu_leg_torques(1, C.byref(lp), P(Jt), 18, P(Rb), P(bpf), P(bvf), P(f), f64, P(ss), P(pt), P(vt), P(tau))
                ref = z["torque_cmds"][r, t].astype(np.float64)
                dt = np.abs(tau[0].astype(np.float64) - ref)
                assert np.all(dt <= 4e-7 * (1.0 + 60.0 * np.abs(ref).max() + 1e3)), (r, t, dt.max())
                worst[2] = max(worst[2], float((dt / (1.0 + np.abs(ref))).max()))
            dp, dv = np.abs(pt[0] - z["pos_targets"][r, t]).max(), np.abs(vt[0] - z["vel_targets"][r, t]).max()
            assert dp <= 1e-8 and dv <= 1e-7, (r, t, dp, dv)
            worst[0], worst[1] = max(worst[0], dp), max(worst[1], dv)
    assert worst[2] <= 2e-6, worst            # relative torque error: float32 rounding level


@pytest.mark.parametrize("nw", [1, 2, 3, 12])
def test_team_argmin_lets_a_nan_win_in_any_warp(emu_lib, nw):
    """team::reduce_argmin (the ratio test of the monotone fallback): the minimum with ties towards the smaller tag, the same
    answer in every thread, and a NaN held by ANY warp of the team wins - it must be noticed (ST_NUMERIC), not skipped."""
    import ctypes as C
    nt = 32 * nw
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    rng = np.random.default_rng(5 + nw)
    def run(vals):
        out_v, out_tag = np.zeros(nt), np.zeros(nt, np.int32)
        emu_lib.mpcq_emu_team_argmin(nw, p(vals, C.c_double), p(out_v, C.c_double), p(out_tag, C.c_int))
        assert np.all(out_tag == out_tag[0])
        assert np.all(out_v == out_v[0]) or np.all(np.isnan(out_v))
        return out_v[0], int(out_tag[0])
    vals = rng.uniform(1.0, 2.0, nt)
    k = int(rng.integers(nt))
    vals[k] = 0.5
    assert run(vals) == (0.5, k)
    vals[(k + 37) % nt] = 0.5                                        # a tie: the smaller tag
    assert run(vals) == (0.5, min(k, (k + 37) % nt))
    for where in sorted({0, nt - 1, nt // 2, 32 * (nw - 1) + 7}):    # a NaN in the first, the last and a middle warp
        v2 = vals.copy()
        v2[where] = np.nan
        v, tag = run(v2)
        assert np.isnan(v) and tag == where
