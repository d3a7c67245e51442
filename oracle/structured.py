"""ORACLE (test infrastructure, NOT product code) - closed-form restatement of the QP data.

The CUDA engine never forms Su / Sx or runs expm; it uses the algebraic structure of
the reference's model (verified against the reference's arrays):

  Ac^3 = 0, Ac^2 Bc = 0  (mpc.py:184-186)  =>  A^m Bd = B0 + (m + 1/2) B1,
      B0 = dt Bc          (rows 6:9 = dt G_a,       rows 9:12 = dt/m I)
      B1 = dt^2 Ac Bc     (rows 0:3 = dt^2 Rz' G_a, rows 3:6  = dt^2/m I)
  with G_a = inv(Rz I_b Rz') [r_a]x  (mpc.py:187-190).  B0 and B1 have disjoint row support
  and Q is diagonal, so  B0' Q B1 = 0  and the condensed Hessian (mpc.py:232) is

      H = 2 (N (x) M00 + S (x) M11 + I (x) R),        N_ij = H - max(i,j),
      M00 = B0' Q B0,   M11 = B1' Q B1   (12x12),     S_ij = sum_{k>=max(i,j)} (k-i+1/2)(k-j+1/2)

  and the linear term (mpc.py:233) is an adjoint recursion over the free-response error
  e_k = A^(k+1) x0 - xref_k  (closed form: constant-velocity + gravity).

This module states that math in numpy float64 so tests can pin it against the golden
(H, g) the unmodified reference produced (tests/golden/*.npz).  It mirrors what
pympc_quadruped_b200/csrc/mpcq_kernels.cu computes per environment.
"""
from __future__ import annotations

import numpy as np


def horizon_tables(horizon: int):
    """N_ij and S_ij, the only horizon-dependent constants of the Hessian."""
    idx = np.arange(horizon)
    N = (horizon - np.maximum.outer(idx, idx)).astype(np.float64)
    S = np.zeros((horizon, horizon))
    for i in range(horizon):
        for j in range(horizon):
            k = np.arange(max(i, j), horizon)
            S[i, j] = np.sum((k - i + 0.5) * (k - j + 0.5))
    return N, S


def foot_maps(yaw: float, feet, inertia_f32, mass: float):
    """Rz (float32-rounded like the reference) and G_a = inv(world_I) [r_a]x rounded to
    float32 on store (mpc.py:177-190).  Returned as float64 arrays holding float32 values."""
    c, s = np.float32(np.cos(yaw)), np.float32(np.sin(yaw))
    Rz = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float64)
    I_b = np.asarray(inertia_f32, dtype=np.float64)
    world_I = (Rz @ I_b @ Rz.T).astype(np.float32).astype(np.float64)
    inv_I = np.linalg.inv(world_I)
    G = np.zeros((4, 3, 3))
    for a in range(4):
        r = np.asarray(feet[a], dtype=np.float64)
        sk = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
        G[a] = (inv_I @ sk).astype(np.float32)
    return Rz, G


def structured_qp(yaw, feet, inertia_f32, mass, dt, q_diag, r_diag, x0, x_ref, horizon):
    """Closed-form (M00, M11, g) of one environment, float64."""
    q = np.asarray(q_diag, dtype=np.float64)
    Rz, G = foot_maps(yaw, feet, inertia_f32, mass)
    inv_m = np.float64(np.float32(1.0 / mass))             # Bc[9:12] = I/m stored float32
    B0 = np.zeros((13, 12))
    B1 = np.zeros((13, 12))
    for a in range(4):
        B0[6:9, 3 * a:3 * a + 3] = dt * G[a]
        B0[9:12, 3 * a:3 * a + 3] = dt * inv_m * np.eye(3)
        B1[0:3, 3 * a:3 * a + 3] = dt * dt * (Rz.T @ G[a])
        B1[3:6, 3 * a:3 * a + 3] = dt * dt * inv_m * np.eye(3)
    M00 = B0.T @ (q[:, None] * B0)
    M11 = B1.T @ (q[:, None] * B1)
    # free response A^(k+1) x0 (Ac x0 = [Rz' w, v, 0, (0,0,x0[12])], Ac^2 x0 = [0, (0,0,x0[12]), 0, 0])
    x0 = np.asarray(x0, dtype=np.float64)
    xr = np.asarray(x_ref, dtype=np.float64).reshape(horizon, 13)
    acx = np.zeros(13)
    acx[0:3] = Rz.T @ x0[6:9]
    acx[3:6] = x0[9:12]
    acx[11] = x0[12]
    ac2x = np.zeros(13)
    ac2x[5] = x0[12]
    g = np.zeros((horizon, 12))
    E0 = np.zeros(13)
    E1 = np.zeros(13)
    for j in range(horizon - 1, -1, -1):
        t = (j + 1) * dt
        qe = q * (x0 + t * acx + 0.5 * t * t * ac2x - xr[j])
        E1 = E1 + E0 + 0.5 * qe
        E0 = E0 + qe
        g[j] = 2.0 * (B0.T @ E0 + B1.T @ E1)
    return M00, M11, g.reshape(-1), (B0, B1)


def dense_hessian(M00, M11, r_diag, horizon):
    N, S = horizon_tables(horizon)
    return 2.0 * (np.kron(N, M00) + np.kron(S, M11) + np.kron(np.eye(horizon), np.diag(r_diag)))
