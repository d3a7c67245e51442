"""ctypes binding of include/mpcq.h (libmpcq.so).

There is no CPU path: if the CUDA library is not built or no device is present, loading /
`mpcq_create` raises.  Build with ``python -c "import __graft_entry__ as g; g.build()"`` or
``python pympc_quadruped_b200/csrc/build.py``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

MPCQ_F32, MPCQ_F64 = 0, 1
ST_VERIFIED, ST_FALLBACK, ST_MAXITER, ST_NUMERIC, ST_NO_STANCE = 1, 2, 4, 8, 32

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "_build", "libmpcq.so")


class MpcqConfig(C.Structure):
    """mirror of `mpcq_config` (include/mpcq.h)"""
    _fields_ = [
        ("horizon", C.c_int32), ("dtype", C.c_int32), ("device", C.c_int32), ("reserved0", C.c_int32),
        ("dt", C.c_double), ("mu", C.c_double), ("fz_max", C.c_double), ("mass", C.c_double),
        ("gravity", C.c_double),
        ("inertia", C.c_double * 9), ("q_diag", C.c_double * 13), ("r_diag", C.c_double * 12),
        ("max_pdas_rounds", C.c_int32), ("max_as_iter", C.c_int32), ("max_refine", C.c_int32),
        ("schedule", C.c_int32),
        ("tol_primal", C.c_double), ("tol_dual", C.c_double), ("tol_residual", C.c_double),
        ("tol_active", C.c_double), ("tol_residual_loose", C.c_double),
        ("dt_control", C.c_double), ("com_height_des", C.c_double),
    ]


def make_config(consts: dict, dtype: int = MPCQ_F32, device: int = 0, **knobs) -> MpcqConfig:
    """`consts` is the dict of configs.extract_mpc_constants (what mpc.py:35-52 reads)."""
    cfg = MpcqConfig()
    cfg.horizon = int(consts["horizon"])
    cfg.dtype = int(dtype)
    cfg.device = int(device)
    cfg.dt = float(consts["dt"])
    cfg.mu = float(consts["mu"])
    cfg.fz_max = float(consts["fz_max"])
    cfg.mass = float(consts["mass"])
    cfg.gravity = float(consts["gravity"])
    cfg.dt_control = float(consts.get("dt_control", 0.001))
    cfg.com_height_des = float(consts.get("com_height_des", 0.0))
    cfg.inertia[:] = [float(v) for v in np.asarray(consts["inertia"], dtype=np.float32).reshape(9)]
    cfg.q_diag[:] = [float(v) for v in consts["q_diag"]]
    cfg.r_diag[:] = [float(v) for v in consts["r_diag"]]
    for k, v in knobs.items():
        if not hasattr(cfg, k):
            raise TypeError(f"unknown solver knob {k!r}")
        setattr(cfg, k, v)
    return cfg


_SOLVE_ARGS = [C.c_void_p] * 5 + [C.c_void_p] * 6          # x0,yaw,feet,gait,xref | f,u,iters,resid,status,active


class MpcqLegParams(C.Structure):
    """mirror of `mpcq_leg_params` (include/mpcq.h)"""
    _fields_ = [("kp_swing", C.c_double * 9), ("kd_swing", C.c_double * 9), ("swing_height", C.c_double),
                ("dt_control", C.c_double), ("gravity", C.c_double), ("foot_z_final", C.c_double)]


def make_leg_params(kp_swing, kd_swing, swing_height, dt_control, gravity, foot_z_final=-0.0255) -> MpcqLegParams:
    import numpy as np
    lp = MpcqLegParams()
    lp.kp_swing[:] = [float(v) for v in np.asarray(kp_swing, dtype=np.float64).reshape(9)]
    lp.kd_swing[:] = [float(v) for v in np.asarray(kd_swing, dtype=np.float64).reshape(9)]
    lp.swing_height, lp.dt_control, lp.gravity, lp.foot_z_final = float(swing_height), float(dt_control), float(gravity), float(foot_z_final)
    return lp


def bind(lib: C.CDLL) -> C.CDLL:
    lib.mpcq_version.restype = C.c_int
    lib.mpcq_create.argtypes = [C.POINTER(MpcqConfig), C.POINTER(C.c_void_p)]
    lib.mpcq_create.restype = C.c_int
    lib.mpcq_destroy.argtypes = [C.c_void_p]
    lib.mpcq_destroy.restype = None
    lib.mpcq_last_error.argtypes = [C.c_void_p]
    lib.mpcq_last_error.restype = C.c_char_p
    lib.mpcq_solve.argtypes = [C.c_void_p, C.c_int32] + _SOLVE_ARGS + [C.c_void_p]
    lib.mpcq_solve.restype = C.c_int
    lib.mpcq_solve_host.argtypes = [C.c_void_p, C.c_int32] + _SOLVE_ARGS
    lib.mpcq_solve_host.restype = C.c_int
    lib.mpcq_tick_host.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.mpcq_tick_host.restype = C.c_int
    lib.mpcq_tick_host_submit.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.mpcq_tick_host_submit.restype = C.c_int
    lib.mpcq_tick_host_wait.argtypes = [C.c_void_p, C.c_int32]
    lib.mpcq_tick_host_wait.restype = C.c_int
    lib.mpcq_tick_reset.argtypes = [C.c_void_p]
    lib.mpcq_tick_reset.restype = C.c_int
    lib.mpcq_build_qp.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 5 + [C.c_void_p] * 3 + [C.c_void_p]
    lib.mpcq_build_qp.restype = C.c_int
    lib.mpcq_assemble.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 10 + [C.c_int32, C.c_int32] + [C.c_void_p] * 4
    lib.mpcq_assemble.restype = C.c_int
    lib.mpcq_last_launch_count.argtypes = [C.c_void_p]
    lib.mpcq_last_launch_count.restype = C.c_int
    lib.mpcq_set_profiling.argtypes = [C.c_void_p, C.c_int32]
    lib.mpcq_set_profiling.restype = C.c_int
    lib.mpcq_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int32]
    lib.mpcq_last_kernel_ms.restype = C.c_int
    lib.mpcq_gait_tables.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 4 + [C.c_int32] + [C.c_void_p] * 4
    lib.mpcq_gait_tables.restype = C.c_int
    lib.mpcq_set_warm_start.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mpcq_set_warm_start.restype = C.c_int
    lib.mpcq_swing_targets.argtypes = [C.c_void_p, C.c_int32, C.POINTER(MpcqLegParams)] + [C.c_void_p] * 16 + [C.c_void_p]
    lib.mpcq_swing_targets.restype = C.c_int
    lib.mpcq_leg_torques.argtypes = [C.c_void_p, C.c_int32, C.POINTER(MpcqLegParams), C.c_void_p, C.c_int32] + [C.c_void_p] * 8 + [C.c_void_p]
    lib.mpcq_leg_torques.restype = C.c_int
    lib.mpcq_measure_peaks.argtypes = [C.c_int32, C.POINTER(C.c_double)]
    lib.mpcq_measure_peaks.restype = C.c_int
    return lib


EXPORTS = ("mpcq_version", "mpcq_create", "mpcq_destroy", "mpcq_last_error", "mpcq_solve",
           "mpcq_solve_host", "mpcq_build_qp", "mpcq_assemble", "mpcq_last_launch_count", "mpcq_set_profiling",
           "mpcq_last_kernel_ms", "mpcq_measure_peaks", "mpcq_gait_tables", "mpcq_set_warm_start",
           "mpcq_swing_targets", "mpcq_leg_torques", "mpcq_tick_host", "mpcq_tick_reset", "mpcq_tick_host_submit", "mpcq_tick_host_wait")

_lib = None


def load_library() -> C.CDLL:
    """Load libmpcq.so or fail loudly (the engine has no fallback path)."""
    global _lib
    if _lib is None:
        path = os.environ.get("MPCQ_LIB_PATH", LIB_PATH)       # development: another build of the same library (A/B experiments)
        if path != LIB_PATH:
            _lib = bind(C.CDLL(path))
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA engine is not built (run __graft_entry__.build()). "
                "pympc_quadruped_b200 has no CPU fallback.")
        _lib = bind(C.CDLL(LIB_PATH))
    return _lib
