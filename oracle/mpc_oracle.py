"""ORACLE (test infrastructure, NOT product code) - CPU restatement of the MPC hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
leg may import this package.  The product path (pympc_quadruped_b200) never does.

Restates, in numpy with the SAME dtype flow and rounding points, the reference's
QP construction  /root/reference/linear_mpc/mpc.py  (line numbers cited per function):

  quat_to_zyx            utils/kinematics.py:40-49
  skew                   utils/kinematics.py:166-177
  OracleMPC.update_robot_state      mpc.py:55-79
  OracleMPC.update_mpc_if_needed    mpc.py:81-108
  OracleMPC.reference_trajectory    mpc.py:110-170
  state_space_model      mpc.py:173-192
  discretize             mpc.py:194-208   (scipy expm of the 25x25 float32 block matrix)
  qp_cost                mpc.py:211-235
  qp_constraints         mpc.py:237-260
  OracleMPC.solve_mpc    mpc.py:262-290   (drake branch: min 1/2 u'Hu + g'u, lb <= Cu <= ub)

The reference hands the QP to Drake 1.15.0 `Solve()` (-> OSQP, requirements.txt:23,51),
a third-party dependency that is absent here and not installable offline.  The QP is
strictly convex (H >= 2R > 0), so its optimum is unique; `oracle.qp_exact.solve_qp_exact`
computes it in float64 and verifies the KKT conditions.  Parity pinning: the reference
holds NO tests or golden vectors for this path, so the
construction part is pinned against outputs of the reference itself, generated here by
oracle/make_golden.py (tests/golden/*.npz); the solve is pinned by KKT verification and
independent solvers (tests/test_oracle_solver.py).
"""
from __future__ import annotations

import math

import numpy as np
from scipy.linalg import expm

from .qp_exact import solve_qp_exact

NUM_STATE = 13
NUM_INPUT = 12


def quat_to_zyx(quat) -> list:
    """(w,x,y,z) -> [roll, pitch, yaw] in Python float64 (utils/kinematics.py:40-49)."""
    w, x, y, z = (quat[0], quat[1], quat[2], quat[3])
    roll = math.atan2(2 * (w * x + y * z), 1 - 2 * (x ** 2 + y ** 2))
    pitch = math.asin(2 * (w * y - z * x))
    yaw = math.atan2(2 * (w * z + x * y), 1 - 2 * (y ** 2 + z ** 2))
    return [roll, pitch, yaw]


def quat_to_matrix(quat) -> np.ndarray:
    """(w,x,y,z) -> R_base (utils/kinematics.py:51-71)."""
    w, x, y, z = (quat[0], quat[1], quat[2], quat[3])
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (w * y + x * z)],
        [2 * (w * z + x * y), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (w * x + y * z), w * w - x * x - y * y + z * z]])


def skew(v) -> np.ndarray:
    """[v]x, float64 (utils/kinematics.py:166-177)."""
    out = np.zeros((3, 3))
    out[0, 1], out[0, 2] = -v[2], v[1]
    out[1, 0], out[1, 2] = v[2], -v[0]
    out[2, 0], out[2, 1] = -v[1], v[0]
    return out


def state_space_model(yaw: float, feet, inertia_f32: np.ndarray, mass: float):
    """Continuous single-rigid-body model, float32 storage (mpc.py:173-192)."""
    Ac = np.zeros((NUM_STATE, NUM_STATE), dtype=np.float32)
    Bc = np.zeros((NUM_STATE, NUM_INPUT), dtype=np.float32)
    c, s = np.cos(yaw), np.sin(yaw)
    Rz = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float32)
    world_I = Rz @ inertia_f32 @ Rz.T                       # float32 products
    Ac[0:3, 6:9] = Rz.T
    Ac[3:6, 9:12] = np.identity(3, dtype=np.float32)
    Ac[11, 12] = 1.0
    for leg in range(4):
        # float32 inverse times float64 skew -> float64, rounded to float32 on store
        Bc[6:9, 3 * leg:3 * leg + 3] = np.linalg.inv(world_I) @ skew(feet[leg])
        Bc[9:12, 3 * leg:3 * leg + 3] = np.identity(3, dtype=np.float32) / mass
    return Ac, Bc


def discretize(Ac: np.ndarray, Bc: np.ndarray, dt: float):
    """Zero-order hold through expm of the augmented float32 matrix (mpc.py:194-208)."""
    dim = NUM_STATE + NUM_INPUT
    M = np.zeros((dim, dim), dtype=np.float32)
    M[0:NUM_STATE, 0:NUM_STATE] = Ac * dt
    M[0:NUM_STATE, NUM_STATE:dim] = Bc * dt
    E = expm(M)
    return E[0:NUM_STATE, 0:NUM_STATE], E[0:NUM_STATE, NUM_STATE:dim]


def qp_cost(Ad, Bd, x0, x_ref, Qbar, Rbar, horizon: int, return_su: bool = False):
    """Condensed cost H = 2(Su'QSu + R), g = 2 Su'Q(Sx x0 - Xref) (mpc.py:211-235)."""
    powers = [np.identity(NUM_STATE, dtype=np.float32)]
    for i in range(horizon):
        powers.append(powers[i] @ Ad)                       # float32 products
    Sx = np.zeros((NUM_STATE * horizon, NUM_STATE), dtype=np.float32)
    Su = np.zeros((NUM_STATE * horizon, NUM_INPUT * horizon), dtype=np.float32)
    for i in range(horizon):
        Sx[NUM_STATE * i:NUM_STATE * (i + 1), :] = powers[i + 1]
        for j in range(i + 1):
            Su[NUM_STATE * i:NUM_STATE * (i + 1), NUM_INPUT * j:NUM_INPUT * (j + 1)] = powers[i - j] @ Bd
    H = 2 * (Su.T @ Qbar @ Su + Rbar)                       # float64
    g = 2 * Su.T @ Qbar @ (Sx @ x0 - x_ref)                 # (Sx x0 - Xref) in float32, rest float64
    if return_su:
        return H, g, Sx, Su
    return H, g


def qp_constraints(gait_table, mu: float, fz_max: float, horizon: int):
    """Friction pyramid + contact schedule rows (mpc.py:237-260)."""
    pyramid = np.array([[1, 0, mu], [-1, 0, mu], [0, 1, mu], [0, -1, mu], [0, 0, 1]], dtype=np.float32)
    C = np.kron(np.identity(4 * horizon, dtype=np.float32), pyramid)
    lb = np.zeros(20 * horizon, dtype=np.float32)
    ub = np.zeros(20 * horizon, dtype=np.float32)
    for k in range(4 * horizon):
        ub[5 * k:5 * k + 4] = np.inf
        ub[5 * k + 4] = gait_table[k] * fz_max
    return C, lb, ub


class RobotState:
    """The RobotData fields the path reads (utils/robot_data.py:70-76,144-149)."""

    def __init__(self, quat_base, pos_base, ang_vel_base, lin_vel_base, pos_base_feet, R_base=None):
        self.quat_base = np.asarray(quat_base)
        self.pos_base = np.asarray(pos_base)
        self.ang_vel_base = np.asarray(ang_vel_base)
        self.lin_vel_base = np.asarray(lin_vel_base)
        self.pos_base_feet = [np.asarray(p) for p in pos_base_feet]
        self.R_base = quat_to_matrix(self.quat_base) if R_base is None else np.asarray(R_base)


class OracleMPC:
    """Stateful per-robot controller with the reference's semantics (mpc.py:22-290)."""

    def __init__(self, mpc_config, robot_config):
        self.num_state, self.num_input = NUM_STATE, NUM_INPUT
        self.is_initialized = False
        self.is_first_run = True
        self.dt_control = mpc_config.dt_control
        self.iterations_between_mpc = mpc_config.iteration_between_mpc
        self.dt = 0.05                                       # hard-coded in the reference (mpc.py:38)
        self.horizon = mpc_config.horizon
        self.mu = mpc_config.friction_coef
        self.fz_max = robot_config.fz_max
        self.gravity = mpc_config.gravity
        self.base_inertia_base = robot_config.base_inertia_base
        self.mass = robot_config.mass_base
        self.com_height_des = robot_config.base_height_des
        self.Qbar = np.kron(np.identity(self.horizon), mpc_config.Q)
        self.Rbar = np.kron(np.identity(self.horizon), mpc_config.R)
        self.contact_forces = np.zeros(12)
        self.last_solution = None

    def update_robot_state(self, rd) -> None:
        if not self.is_initialized:
            self.current_state = np.zeros(13, dtype=np.float32)
            self.roll_init = 0.0
            self.pitch_init = 0.0
            self.is_initialized = True
        self._rd = rd
        rpy = quat_to_zyx(rd.quat_base)
        self.current_state[0:3] = rpy
        self.current_state[3:6] = np.array(rd.pos_base, dtype=np.float32)
        self.current_state[6:9] = np.array(rd.ang_vel_base, dtype=np.float32)
        self.current_state[9:12] = np.array(rd.lin_vel_base, dtype=np.float32)
        self.current_state[12] = -self.gravity
        self.yaw = rpy[2]
        self.pos_base_feet = rd.pos_base_feet

    def update_mpc_if_needed(self, iter_counter, base_vel_base_des, yaw_turn_rate_des, gait_table):
        vel_des = self._rd.R_base @ base_vel_base_des
        if self.is_first_run:
            self.xpos_base_desired = 0.0
            self.ypos_base_desired = 0.0
            self.yaw_desired = self.yaw
            self.is_first_run = False
        else:
            self.xpos_base_desired += self.dt_control * vel_des[0]
            self.ypos_base_desired += self.dt_control * vel_des[1]
            self.yaw_desired = self.yaw + self.dt_control * yaw_turn_rate_des
        if iter_counter % self.iterations_between_mpc == 0:
            self.ref_traj = self.reference_trajectory(vel_des, yaw_turn_rate_des)
            self.contact_forces = self.solve_mpc(self.ref_traj, gait_table)[0:12]
        return self.contact_forces[0:12]

    def reference_trajectory(self, vel_des, yaw_turn_rate) -> np.ndarray:
        x = self.current_state
        xd, yd = self.xpos_base_desired, self.ypos_base_desired
        lim = 0.1
        if xd - x[3] > lim:
            xd = x[3] + lim
        if x[3] - xd > lim:
            xd = x[3] - lim
        if yd - x[4] > lim:
            yd = x[4] + lim
        if x[4] - yd > lim:
            yd = x[4] - lim
        self.xpos_base_desired, self.ypos_base_desired = xd, yd
        if np.fabs(x[9]) > 0.2:
            self.pitch_init += self.dt * (0.0 - x[1]) / x[9]
        if np.fabs(x[10]) > 0.1:
            self.roll_init += self.dt * (0.0 - x[0]) / x[10]
        self.roll_init = np.fmin(np.fmax(self.roll_init, -0.25), 0.25)
        self.pitch_init = np.fmin(np.fmax(self.pitch_init, -0.25), 0.25)
        roll_comp = x[10] * self.roll_init
        pitch_comp = x[9] * self.pitch_init
        n, H = self.num_state, self.horizon
        X = np.zeros(n * H, dtype=np.float32)
        X[0::n] = roll_comp
        X[1::n] = pitch_comp
        X[2], X[3], X[4] = self.yaw_desired, xd, yd
        X[5::n] = self.com_height_des
        X[8::n] = yaw_turn_rate
        X[9::n] = vel_des[0]
        X[10::n] = vel_des[1]
        X[12::n] = -self.gravity
        for i in range(1, H):                                # accumulated in float32 storage
            X[2 + n * i] = X[2 + n * (i - 1)] + self.dt * yaw_turn_rate
            X[3 + n * i] = X[3 + n * (i - 1)] + self.dt * vel_des[0]
            X[4 + n * i] = X[4 + n * (i - 1)] + self.dt * vel_des[1]
        return X

    def build_qp(self, ref_traj, gait_table, with_intermediates: bool = False):
        Ac, Bc = state_space_model(self.yaw, self.pos_base_feet, self.base_inertia_base, self.mass)
        discretize(Ac, Bc, self.dt)                          # the reference calls it twice (mpc.py:267-268)
        Ad, Bd = discretize(Ac, Bc, self.dt)
        H, g = qp_cost(Ad, Bd, self.current_state, ref_traj, self.Qbar, self.Rbar, self.horizon)
        C, lb, ub = qp_constraints(gait_table, self.mu, self.fz_max, self.horizon)
        if with_intermediates:
            return H, g, C, lb, ub, dict(Ac=Ac, Bc=Bc, Ad=Ad, Bd=Bd)
        return H, g, C, lb, ub

    def solve_mpc(self, ref_traj, gait_table) -> np.ndarray:
        H, g, C, lb, ub = self.build_qp(ref_traj, gait_table)
        sol = solve_qp_exact(H, g, self.mu, ub[4::5])
        self.last_solution = sol
        return sol.u
