"""Device-side engine handle: torch tensors in, torch tensors out, through the C ABI.

`MpcqEngine.solve` is the batched replacement of `ModelPredictiveController._solve_mpc`
(reference linear_mpc/mpc.py:262-290).  torch is used for device memory and streams only;
all arithmetic happens inside libmpcq.so (sm_100a kernels).  No CPU path exists.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _capi
from .configs import extract_mpc_constants


@dataclass
class SolveResult:
    forces: torch.Tensor          # [B,12]  first-step GRFs  (= _solve_mpc(...)[0:12])
    u: torch.Tensor | None        # [B,12H] whole optimum
    iters: torch.Tensor | None    # [B,2]   factorisations, fallback iterations
    resid: torch.Tensor | None    # [B,2]   reduced gradient, primal violation (fp64)
    status: torch.Tensor | None   # [B]     MPCQ_ST_* bits
    active: torch.Tensor | None   # [B,4H]  constraint-activity bit masks


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class MpcqEngine:
    """One handle per (config, robot, dtype, device)."""

    def __init__(self, mpc_config, robot_config, dtype=torch.float32, device="cuda:0", **knobs):
        if dtype not in (torch.float32, torch.float64):
            raise ValueError("dtype must be torch.float32 or torch.float64")
        self.lib = _capi.load_library()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("MpcqEngine needs a CUDA device: the engine has no CPU fallback")
        self.dtype = dtype
        self.consts = extract_mpc_constants(mpc_config, robot_config)
        self.horizon = self.consts["horizon"]
        cfg = _capi.make_config(self.consts, _capi.MPCQ_F64 if dtype == torch.float64 else _capi.MPCQ_F32,
                                self.device.index or 0, **knobs)
        self._h = C.c_void_p()
        self._tick_refs = {}                                  # buffers of asynchronous ticks in flight, per slot
        rc = self.lib.mpcq_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            msg = f"mpcq_create failed ({rc}): {self.lib.mpcq_last_error(None).decode()}"
            raise ValueError(msg) if rc in (-1, -4) else RuntimeError(msg)     # invalid / unsupported configuration

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.mpcq_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------------------------------
    def _check(self, name, t, shape, dtype):
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} must be a torch.Tensor")
        if t.device != self.device:
            raise ValueError(f"{name} is on {t.device}, engine is on {self.device}")
        if t.dtype != dtype:
            raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(t.shape)}")
        return t.contiguous()

    def _err(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.lib.mpcq_last_error(self._h).decode()}")

    def solve(self, x0, r_feet, gait, x_ref, yaw=None, want=("u", "iters", "resid", "status", "active"), out=None,
              faces_in=None, faces_out=None):
        """x0 [B,13], r_feet [B,12] or [B,4,3], gait float32 [B,4H], x_ref [B,13H], yaw [B] optional.
        `faces_in` / `faces_out`: uint8 [B,4H] warm-start face codes (`mpcq_set_warm_start`); zeros = cold start."""
        B, H = x0.shape[0], self.horizon
        x0 = self._check("x0", x0, (B, 13), self.dtype)
        r_feet = self._check("r_feet", r_feet.reshape(B, 12), (B, 12), self.dtype)
        gait = self._check("gait", gait, (B, 4 * H), torch.float32)
        x_ref = self._check("x_ref", x_ref, (B, 13 * H), self.dtype)
        if yaw is not None:
            yaw = self._check("yaw", yaw, (B,), self.dtype)
        if out is None:
            mk = lambda shape, dt: torch.empty(shape, dtype=dt, device=self.device)
            out = SolveResult(
                forces=mk((B, 12), self.dtype),
                u=mk((B, 12 * H), self.dtype) if "u" in want else None,
                iters=mk((B, 2), torch.int32) if "iters" in want else None,
                resid=mk((B, 2), torch.float64) if "resid" in want else None,
                status=mk((B,), torch.int32) if "status" in want else None,
                active=mk((B, 4 * H), torch.uint8) if "active" in want else None)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        warm = faces_in is not None or faces_out is not None
        if warm:
            for name, t in (("faces_in", faces_in), ("faces_out", faces_out)):
                if t is not None and self._check(name, t, (B, 4 * H), torch.uint8) is not t:
                    raise ValueError(f"{name} must be a contiguous uint8 tensor on the engine's device")
            self._err(self.lib.mpcq_set_warm_start(self._h, _ptr(faces_in), _ptr(faces_out)), "mpcq_set_warm_start")
        try:
            rc = self.lib.mpcq_solve(self._h, B, _ptr(x0), _ptr(yaw), _ptr(r_feet), _ptr(gait), _ptr(x_ref),
                                     _ptr(out.forces), _ptr(out.u), _ptr(out.iters), _ptr(out.resid), _ptr(out.status),
                                     _ptr(out.active), C.c_void_p(stream))
        finally:
            if warm:
                self.lib.mpcq_set_warm_start(self._h, None, None)
        self._err(rc, "mpcq_solve")
        return out

    def solve_host(self, x0, r_feet, gait, x_ref, yaw=None, want=("status",), out=None):
        """numpy in / numpy out through `mpcq_solve_host` (H2D + solve + D2H inside the call).
        Page-locked arrays (e.g. `torch.empty(..., pin_memory=True).numpy()`) are used for DMA directly;
        `out` may hold preallocated result arrays (keys forces/u/iters/resid/status/active)."""
        rt = np.float64 if self.dtype == torch.float64 else np.float32
        B, H = x0.shape[0], self.horizon
        x0 = np.ascontiguousarray(x0, dtype=rt).reshape(B, 13)
        r_feet = np.ascontiguousarray(r_feet, dtype=rt).reshape(B, 12)
        gait = np.ascontiguousarray(gait, dtype=np.float32).reshape(B, 4 * H)
        x_ref = np.ascontiguousarray(x_ref, dtype=rt).reshape(B, 13 * H)
        yaw = None if yaw is None else np.ascontiguousarray(yaw, dtype=rt).reshape(B)
        spec = dict(forces=((B, 12), rt), u=((B, 12 * H), rt), iters=((B, 2), np.int32), resid=((B, 2), np.float64),
                    status=((B,), np.int32), active=((B, 4 * H), np.uint8))
        res = {}
        for key, (shape, dt) in spec.items():
            if key != "forces" and key not in want:
                continue
            if out is not None and key in out:
                a = out[key]
                if a.shape != shape or a.dtype != dt or not a.flags["C_CONTIGUOUS"]:
                    raise ValueError(f"out[{key!r}] must be a C-contiguous {dt} array of shape {shape}")
                res[key] = a
            else:
                res[key] = np.empty(shape, dt)
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        rc = self.lib.mpcq_solve_host(self._h, B, p(x0), p(yaw), p(r_feet), p(gait), p(x_ref), p(res["forces"]),
                                      p(res.get("u")), p(res.get("iters")), p(res.get("resid")), p(res.get("status")),
                                      p(res.get("active")))
        self._err(rc, "mpcq_solve_host")
        return res

    def tick_host(self, state_cmd, gait_params, iterations_between_mpc: int, first_run=False, out=None, validate=True):
        """One MPC update of B robots from HOST state through `mpcq_tick_host`: `state_cmd` [B,29] float64
        (quat 4 | pos 3 | omega 3 | vel 3 | pos_base_feet 12 | v_des_body 3 | yaw_rate), `gait_params` [B,10] int32
        (stance_offsets 4 | stance_durations 4 | num_segment | cur_iteration).  Gait table, state assembly, reference
        trajectory and the solve run on the device; the controller's integrator state lives in the handle.
        `first_run`: False / True (mpc.py:84-88) / 2 (respawn: desired pose = current pose).  `validate=False` skips the
        host-side pass over gait_params (num_segment >= 1) in hot loops.
        Returns dict(forces [B,12], status [B]) as numpy arrays (`out` may hold preallocated, e.g. page-locked, ones)."""
        rt = np.float64 if self.dtype == torch.float64 else np.float32
        state_cmd = np.ascontiguousarray(state_cmd, dtype=np.float64)
        B = state_cmd.shape[0]
        if state_cmd.shape != (B, 29):
            raise ValueError(f"state_cmd must be [B,29], got {state_cmd.shape}")
        gait_params = np.ascontiguousarray(gait_params, dtype=np.int32)
        if gait_params.shape != (B, 10):
            raise ValueError(f"gait_params must be [B,10], got {gait_params.shape}")
        if validate and np.any(gait_params[:, 8] < 1):
            raise ValueError("num_segment must be >= 1")
        res = {}
        for key, shape, dt in (("forces", (B, 12), rt), ("status", (B,), np.int32)):
            if out is not None and key in out:
                a = out[key]
                if a.shape != shape or a.dtype != dt or not a.flags["C_CONTIGUOUS"]:
                    raise ValueError(f"out[{key!r}] must be a C-contiguous {dt} array of shape {shape}")
                res[key] = a
            else:
                res[key] = np.empty(shape, dt)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = self.lib.mpcq_tick_host(self._h, B, p(state_cmd), p(gait_params), int(iterations_between_mpc), int(first_run),
                                     p(res["forces"]), p(res["status"]))
        self._err(rc, "mpcq_tick_host")
        return res

    def tick_submit(self, slot: int, state_cmd, gait_params, iterations_between_mpc: int, first_run, out):
        """Asynchronous `tick_host` on pipeline `slot` (0 / 1): page-locked numpy arrays only (`out` = dict(forces, status)),
        returns once the work is queued; `tick_wait(slot)` returns when `out` holds the results.  The arrays are kept alive
        by the engine until then."""
        rt = np.float64 if self.dtype == torch.float64 else np.float32
        B = state_cmd.shape[0]
        for name, a, shape, dt in (("state_cmd", state_cmd, (B, 29), np.float64), ("gait_params", gait_params, (B, 10), np.int32),
                                   ("out['forces']", out["forces"], (B, 12), rt), ("out['status']", out["status"], (B,), np.int32)):
            if not isinstance(a, np.ndarray) or a.shape != shape or a.dtype != dt or not a.flags.c_contiguous:
                raise ValueError(f"{name}: expected a C-contiguous {np.dtype(dt).name} array of shape {shape}")
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        self._err(self.lib.mpcq_tick_host_submit(self._h, int(slot), B, p(state_cmd), p(gait_params), int(iterations_between_mpc),
                                                 int(first_run), p(out["forces"]), p(out["status"])), "mpcq_tick_host_submit")
        self._tick_refs[int(slot)] = (state_cmd, gait_params, out["forces"], out["status"])

    def tick_wait(self, slot: int):
        self._err(self.lib.mpcq_tick_host_wait(self._h, int(slot)), "mpcq_tick_host_wait")
        self._tick_refs.pop(int(slot), None)

    def tick_reset(self):
        self._err(self.lib.mpcq_tick_reset(self._h), "mpcq_tick_reset")

    def build_qp(self, x0, r_feet, gait, x_ref, yaw=None):
        """Dense (H [B,n,n], g [B,n], ub [B,20H]) in float64 - the data the reference hands to its solver."""
        B, H = x0.shape[0], self.horizon
        n = 12 * H
        x0 = self._check("x0", x0, (B, 13), self.dtype)
        r_feet = self._check("r_feet", r_feet.reshape(B, 12), (B, 12), self.dtype)
        gait = self._check("gait", gait, (B, 4 * H), torch.float32)
        x_ref = self._check("x_ref", x_ref, (B, 13 * H), self.dtype)
        if yaw is not None:
            yaw = self._check("yaw", yaw, (B,), self.dtype)
        Hm = torch.empty((B, n, n), dtype=torch.float64, device=self.device)
        g = torch.empty((B, n), dtype=torch.float64, device=self.device)
        ub = torch.empty((B, 20 * H), dtype=torch.float64, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.lib.mpcq_build_qp(self._h, B, _ptr(x0), _ptr(yaw), _ptr(r_feet), _ptr(gait), _ptr(x_ref),
                                    _ptr(Hm), _ptr(g), _ptr(ub), C.c_void_p(stream))
        self._err(rc, "mpcq_build_qp")
        return Hm, g, ub

    @property
    def last_launch_count(self) -> int:
        return int(self.lib.mpcq_last_launch_count(self._h))

    def assemble(self, quat, pos, omega, vel, v_des_body, yaw_rate_des, xy_des, yaw_des, rp_init, first_run, do_mpc,
                 x0, yaw, x_ref, R_base=None):
        """Fused state assembly + command integration + reference trajectory (`mpcq_assemble`): float64 device
        tensors in, controller state updated in place, (x0, yaw, x_ref) written in the engine's dtype."""
        B, H = quat.shape[0], self.horizon
        f64 = torch.float64
        quat = self._check("quat", quat, (B, 4), f64)
        pos = self._check("pos", pos, (B, 3), f64)
        omega = self._check("omega", omega, (B, 3), f64)
        vel = self._check("vel", vel, (B, 3), f64)
        v_des_body = self._check("v_des_body", v_des_body, (B, 3), f64)
        yaw_rate_des = self._check("yaw_rate_des", yaw_rate_des, (B,), f64)
        if R_base is not None:
            R_base = self._check("R_base", R_base.reshape(B, 9), (B, 9), f64)
        for name, t, shape, dt in (("xy_des", xy_des, (B, 2), f64), ("yaw_des", yaw_des, (B,), f64), ("rp_init", rp_init, (B, 2), f64),
                                   ("x0", x0, (B, 13), self.dtype), ("yaw", yaw, (B,), self.dtype),
                                   ("x_ref", x_ref, (B, 13 * H), self.dtype)):
            if self._check(name, t, shape, dt) is not t:
                raise ValueError(f"{name} must be contiguous (it is written in place)")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.lib.mpcq_assemble(self._h, B, _ptr(quat), _ptr(pos), _ptr(omega), _ptr(vel), _ptr(R_base), _ptr(v_des_body),
                                    _ptr(yaw_rate_des), _ptr(xy_des), _ptr(yaw_des), _ptr(rp_init), int(first_run),
                                    int(bool(do_mpc)), _ptr(x0), _ptr(yaw), _ptr(x_ref), C.c_void_p(stream))
        self._err(rc, "mpcq_assemble")

    def gait_tables(self, stance_offsets, stance_durations, num_segment, cur_iteration, iterations_between_mpc,
                    table=None, swing_state=None, stance_state=None, want_states=False):
        """Contact tables (and swing / stance phase states) for B robots on the device (`mpcq_gait_tables`): int32 device
        tensors in, float32 [B, 4H] table out - the `gait` argument of `solve`."""
        B, H = stance_offsets.shape[0], self.horizon
        i32 = torch.int32
        stance_offsets = self._check("stance_offsets", stance_offsets, (B, 4), i32)
        stance_durations = self._check("stance_durations", stance_durations, (B, 4), i32)
        num_segment = self._check("num_segment", num_segment, (B,), i32)
        cur_iteration = self._check("cur_iteration", cur_iteration, (B,), i32)
        if table is None:
            table = torch.empty((B, 4 * H), dtype=torch.float32, device=self.device)
        elif self._check("table", table, (B, 4 * H), torch.float32) is not table:
            raise ValueError("table must be contiguous (it is written in place)")
        if want_states:
            swing_state = torch.empty((B, 4), dtype=torch.float64, device=self.device) if swing_state is None else swing_state
            stance_state = torch.empty((B, 4), dtype=torch.float64, device=self.device) if stance_state is None else stance_state
        for name, t in (("swing_state", swing_state), ("stance_state", stance_state)):
            if t is not None and self._check(name, t, (B, 4), torch.float64) is not t:
                raise ValueError(f"{name} must be contiguous (it is written in place)")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.lib.mpcq_gait_tables(self._h, B, _ptr(stance_offsets), _ptr(stance_durations), _ptr(num_segment),
                                       _ptr(cur_iteration), int(iterations_between_mpc), _ptr(table), _ptr(swing_state),
                                       _ptr(stance_state), C.c_void_p(stream))
        self._err(rc, "mpcq_gait_tables")
        return table, swing_state, stance_state

    def swing_targets(self, leg_params, pos_base, lin_vel_base, R_base, base_pos_base_thighs, pos_feet, swing_state, v_des_body,
                      yaw_rate_des, swing_time, stance_time, state, pos_targets=None, vel_targets=None):
        """Swing-foot targets of B robots x 4 legs (`mpcq_swing_targets`); `state` = (swing_active uint8 [B,4], remaining [B,4],
        footpos_init [B,4,3], footpos_final [B,4,3]) is advanced in place."""
        B, f64 = pos_base.shape[0], torch.float64
        ins = [self._check(n, t, sh, f64) for n, t, sh in (
            ("pos_base", pos_base, (B, 3)), ("lin_vel_base", lin_vel_base, (B, 3)), ("R_base", R_base.reshape(B, 9), (B, 9)),
            ("base_pos_base_thighs", base_pos_base_thighs, (B, 4, 3)), ("pos_feet", pos_feet, (B, 4, 3)),
            ("swing_state", swing_state, (B, 4)), ("v_des_body", v_des_body, (B, 3)), ("yaw_rate_des", yaw_rate_des, (B,)),
            ("swing_time", swing_time, (B,)), ("stance_time", stance_time, (B,)))]
        active, remaining, init, final = state
        if pos_targets is None:
            pos_targets = torch.empty((B, 4, 3), dtype=f64, device=self.device)
        if vel_targets is None:
            vel_targets = torch.empty((B, 4, 3), dtype=f64, device=self.device)
        for n, t, sh, dt in (("swing_active", active, (B, 4), torch.uint8), ("remaining_swing_time", remaining, (B, 4), f64),
                             ("footpos_init", init, (B, 4, 3), f64), ("footpos_final", final, (B, 4, 3), f64),
                             ("pos_targets", pos_targets, (B, 4, 3), f64), ("vel_targets", vel_targets, (B, 4, 3), f64)):
            if self._check(n, t, sh, dt) is not t:
                raise ValueError(f"{n} must be contiguous (it is written in place)")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.lib.mpcq_swing_targets(self._h, B, C.byref(leg_params), *[_ptr(t) for t in ins], _ptr(active), _ptr(remaining),
                                         _ptr(init), _ptr(final), _ptr(pos_targets), _ptr(vel_targets), C.c_void_p(stream))
        self._err(rc, "mpcq_swing_targets")
        return pos_targets, vel_targets

    def leg_torques(self, leg_params, Jv_feet, R_base, base_pos_base_feet, base_vel_base_feet, contact_forces, swing_state,
                    pos_targets, vel_targets, torque_cmds=None):
        """Joint torques [B,12] float32 of B robots (`mpcq_leg_torques`); Jv_feet is [B,4,3,18] (reference layout) or [B,4,3,3]."""
        B, f64 = R_base.shape[0], torch.float64
        if Jv_feet.dim() != 4 or Jv_feet.shape[-1] not in (3, 18):
            raise ValueError("Jv_feet must have shape [B,4,3,18] or [B,4,3,3]")
        ncol = int(Jv_feet.shape[-1])
        Jv_feet = self._check("Jv_feet", Jv_feet, (B, 4, 3, ncol), f64)
        R_base = self._check("R_base", R_base.reshape(B, 9), (B, 9), f64)
        ins = [self._check(n, t, sh, dt) for n, t, sh, dt in (
            ("base_pos_base_feet", base_pos_base_feet, (B, 4, 3), f64), ("base_vel_base_feet", base_vel_base_feet, (B, 4, 3), f64),
            ("contact_forces", contact_forces, (B, 12), self.dtype), ("swing_state", swing_state, (B, 4), f64),
            ("pos_targets", pos_targets, (B, 4, 3), f64), ("vel_targets", vel_targets, (B, 4, 3), f64))]
        if torque_cmds is None:
            torque_cmds = torch.empty((B, 12), dtype=torch.float32, device=self.device)
        elif self._check("torque_cmds", torque_cmds, (B, 12), torch.float32) is not torque_cmds:
            raise ValueError("torque_cmds must be contiguous (it is written in place)")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.lib.mpcq_leg_torques(self._h, B, C.byref(leg_params), _ptr(Jv_feet), ncol, _ptr(R_base), *[_ptr(t) for t in ins],
                                       _ptr(torque_cmds), C.c_void_p(stream))
        self._err(rc, "mpcq_leg_torques")
        return torque_cmds

    def set_profiling(self, enable: bool) -> None:
        self._err(self.lib.mpcq_set_profiling(self._h, int(bool(enable))), "mpcq_set_profiling")

    def last_kernel_ms(self):
        """Per-launch durations (ms) of the last `solve`, one per size class (needs set_profiling(True))."""
        buf = (C.c_float * 8)()
        n = self.lib.mpcq_last_kernel_ms(self._h, buf, 8)
        if n < 0:
            self._err(n, "mpcq_last_kernel_ms")
        return [float(buf[i]) for i in range(min(n, 8))]
