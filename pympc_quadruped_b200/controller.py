"""Controller interface of the reference, batched over environments.

`BatchedModelPredictiveController` keeps the semantics of the reference's
`ModelPredictiveController` (linear_mpc/mpc.py:22-290) for B robots at once:

  __init__ / _load_parameters       mpc.py:24-52    constructor takes the config CLASSES
  update_robot_state                mpc.py:55-79    quat -> ZYX angles (kinematics.py:40-49), float32 state
  update_mpc_if_needed              mpc.py:81-108   command rotation, integrators, 20-tick decimation, cached forces
  generate_reference_trajectory     mpc.py:110-170  clamp, roll/pitch compensation, X_ref fill
  _solve_mpc                        mpc.py:262-290  -> MpcqEngine.solve (sm_100a kernels via the C ABI)

The per-env bookkeeping above the solve (O(H) elementwise work) runs as torch ops on the same
device, with the reference's float32 rounding points reproduced; the hot path - QP build and
solve - is entirely inside libmpcq.so.  `ModelPredictiveController` is the B = 1 adapter with
the reference's exact signature (numpy robot_data in, numpy[12] out) so it can replace the
object constructed at scripts/isaacgym_a1.py:93 / scripts/mujoco_aliengo.py:173.
"""
from __future__ import annotations

import numpy as np
import torch

from .configs import extract_mpc_constants
from .engine import MpcqEngine


class BatchedRobotData:
    """The five RobotData fields the path reads (utils/robot_data.py:70-76,144-149), batched."""

    def __init__(self, quat_base, pos_base, ang_vel_base, lin_vel_base, pos_base_feet, R_base=None):
        self.quat_base = quat_base            # [B,4] (w,x,y,z)
        self.pos_base = pos_base              # [B,3]
        self.ang_vel_base = ang_vel_base      # [B,3] world
        self.lin_vel_base = lin_vel_base      # [B,3] world
        self.pos_base_feet = pos_base_feet    # [B,4,3] world-frame base->foot, FL FR RL RR
        self.R_base = R_base                  # [B,3,3] or None (derived from the quaternion)

    @classmethod
    def from_isaacgym(cls, root_states, foot_positions_world=None, pos_base_feet=None, env_ids=None):
        """Zero-copy-style glue for the Isaac Gym tensor API, replacing the per-robot
        `.cpu().numpy()` loop of scripts/isaacgym_a1.py:119-133 for the whole batch on the device.

        `root_states` is the wrapped actor-root-state tensor [num_actors, 13]: position 0:3, quaternion 3:7 in Isaac Gym's
        (x, y, z, w) order, linear velocity 7:10, angular velocity 10:13 (world frame).  The reference reorders the
        quaternion to (w, x, y, z) (isaacgym_a1.py:121-125); so does this.  Foot positions come either as world-frame
        foot positions [B,4,3] (e.g. from the rigid-body state tensor; base->foot = foot - base, utils/robot_data.py:144-149)
        or directly as `pos_base_feet`.  `env_ids` selects / orders the robots (LongTensor)."""
        rs = root_states if env_ids is None else root_states.index_select(0, env_ids)
        if rs.dim() != 2 or rs.shape[1] != 13:
            raise ValueError("root_states must be [num_actors, 13]")
        rs = rs.to(torch.float64)
        pos = rs[:, 0:3].contiguous()
        quat = torch.stack([rs[:, 6], rs[:, 3], rs[:, 4], rs[:, 5]], dim=1)
        if (foot_positions_world is None) == (pos_base_feet is None):
            raise ValueError("give exactly one of foot_positions_world / pos_base_feet")
        if pos_base_feet is None:
            fw = foot_positions_world.to(torch.float64).reshape(rs.shape[0], 4, 3)
            pos_base_feet = fw - pos[:, None, :]
        else:
            pos_base_feet = pos_base_feet.to(torch.float64).reshape(rs.shape[0], 4, 3)
        return cls(quat, pos, rs[:, 10:13].contiguous(), rs[:, 7:10].contiguous(), pos_base_feet.contiguous())


class BatchedModelPredictiveController:
    def __init__(self, mpc_config, robot_config, num_envs: int, device="cuda:0", dtype=torch.float32, engine=None,
                 warm_start=False, **solver_knobs):
        """`engine` is for dependency injection in tests (an object with MpcqEngine's solve / assemble signatures);
        by default the CUDA engine is created and a missing GPU / library raises.  State assembly, command integration
        and the reference trajectory always run in the device kernel (`mpcq_assemble`): there is no second backend.
        `warm_start`: start every MPC update from the previous update's active faces shifted by one horizon step
        (`mpcq_set_warm_start`); the optimum is unique, so only the number of active-set rounds changes."""
        c = extract_mpc_constants(mpc_config, robot_config)
        self.num_envs = int(num_envs)
        self.device = torch.device(device)
        self.dtype = dtype
        self.num_state, self.num_input = 13, 12
        self.dt_control = c["dt_control"]
        self.iterations_between_mpc = c["iterations_between_mpc"]
        self.dt = c["dt"]
        self.horizon = c["horizon"]
        self.mu = c["mu"]
        self.fz_max = c["fz_max"]
        self.gravity = c["gravity"]
        self.mass = c["mass"]
        self.com_height_des = c["com_height_des"]
        self.engine = engine if engine is not None else MpcqEngine(mpc_config, robot_config, dtype=dtype, device=device,
                                                                   **solver_knobs)
        B, dev = self.num_envs, self.device
        f64 = dict(dtype=torch.float64, device=dev)
        self.is_initialized = False
        self.is_first_run = True
        if not hasattr(self.engine, "assemble"):
            raise TypeError("the engine must provide assemble() (mpcq_assemble): the controller has no host-side fallback")
        self.current_state = torch.zeros((B, 13), dtype=dtype, device=dev)
        self.yaw = torch.zeros(B, dtype=dtype, device=dev)
        self._xy_des = torch.zeros((B, 2), **f64)
        self._rp_init = torch.zeros((B, 2), **f64)             # roll_init, pitch_init
        self.yaw_desired = torch.zeros(B, **f64)
        self.contact_forces = torch.zeros((B, 12), dtype=dtype, device=dev)
        self.ref_traj = torch.zeros((B, 13 * self.horizon), dtype=dtype, device=dev)
        self.last_result = None
        self.warm_start = bool(warm_start)
        self._faces = torch.zeros((B, 4 * self.horizon), dtype=torch.uint8, device=dev) if self.warm_start else None
        self._faces_in = torch.zeros_like(self._faces) if self.warm_start else None

    # reference attribute names (mpc.py:85-92,143-150) as views of the packed state
    xpos_base_desired = property(lambda self: self._xy_des[:, 0], lambda self, v: self._xy_des[:, 0].copy_(torch.as_tensor(v)))
    ypos_base_desired = property(lambda self: self._xy_des[:, 1], lambda self, v: self._xy_des[:, 1].copy_(torch.as_tensor(v)))
    roll_init = property(lambda self: self._rp_init[:, 0], lambda self, v: self._rp_init[:, 0].copy_(torch.as_tensor(v)))
    pitch_init = property(lambda self: self._rp_init[:, 1], lambda self, v: self._rp_init[:, 1].copy_(torch.as_tensor(v)))

    # ---------------------------------------------------------------------------------------
    def _t(self, a, shape):
        t = torch.as_tensor(a, dtype=torch.float64, device=self.device) if not isinstance(a, torch.Tensor) \
            else a.to(device=self.device, dtype=torch.float64)
        return t.reshape(shape)

    def update_robot_state(self, robot_data) -> None:
        """mpc.py:55-79: current_state = [rpy, pos, omega, vel, -g] rounded to float32; yaw kept in float64."""
        B = self.num_envs
        self._quat = self._t(robot_data.quat_base, (B, 4)).contiguous()
        self._pos = self._t(robot_data.pos_base, (B, 3)).contiguous()
        self._omega = self._t(robot_data.ang_vel_base, (B, 3)).contiguous()
        self._vel = self._t(robot_data.lin_vel_base, (B, 3)).contiguous()
        self.pos_base_feet = self._t(robot_data.pos_base_feet, (B, 12))
        R = getattr(robot_data, "R_base", None)
        self._R_given = None if R is None else self._t(R, (B, 3, 3)).contiguous()
        self.is_initialized = True                              # assembled on the device inside update_mpc_if_needed

    def update_mpc_if_needed(self, iter_counter: int, base_vel_base_des, yaw_turn_rate_des, gait_table,
                             solver: str = "drake", debug: bool = False, iter_debug=None):
        """mpc.py:81-108.  `solver` is accepted for signature compatibility; the drake formulation
        (lb <= Cu <= ub) is the one implemented.  Returns the cached [B,12] forces between MPC updates."""
        assert solver in ("drake", "qpsolvers")
        if solver == "qpsolvers":
            raise NotImplementedError("the reference's qpsolvers branch drops lb and solves a different QP "
                                      "(mpc.py:289); only the drake formulation is implemented")
        B = self.num_envs
        v_body = self._t(base_vel_base_des, (-1, 3)).expand(B, 3)
        yaw_rate = self._t(yaw_turn_rate_des, (-1,)).expand(B)
        do_mpc = iter_counter % self.iterations_between_mpc == 0
        self.engine.assemble(self._quat, self._pos, self._omega, self._vel, v_body.contiguous(), yaw_rate.contiguous(),
                             self._xy_des, self.yaw_desired, self._rp_init, self.is_first_run, do_mpc,
                             self.current_state, self.yaw, self.ref_traj, R_base=self._R_given)
        self.is_first_run = False
        if do_mpc:
            self.contact_forces = self._solve_mpc(self.ref_traj, gait_table)
        return self.contact_forces

    def _solve_mpc(self, ref_traj: torch.Tensor, gait_table) -> torch.Tensor:
        B, H = self.num_envs, self.horizon
        gait = gait_table if isinstance(gait_table, torch.Tensor) else torch.as_tensor(np.asarray(gait_table))
        gait = gait.to(device=self.device, dtype=torch.float32).reshape(-1, gait.shape[-1])[:, :4 * H]
        if gait.shape[0] == 1 and B > 1:
            gait = gait.expand(B, 4 * H)
        warm = {}
        if self.warm_start:
            # the previous update's faces, one horizon step later (the contact table advances one step per update)
            self._faces_in[:, :4 * (H - 1)] = self._faces[:, 4:]
            self._faces_in[:, 4 * (H - 1):] = self._faces[:, 4 * (H - 1):]
            warm = dict(faces_in=self._faces_in, faces_out=self._faces)
        res = self.engine.solve(self.current_state.to(self.dtype), self.pos_base_feet.to(self.dtype),
                                gait.contiguous(), ref_traj.to(self.dtype), yaw=self.yaw.to(self.dtype), **warm)
        self.last_result = res
        return res.forces


class ModelPredictiveController:
    """Drop-in for the reference class (same constructor and method signatures, numpy in/out)."""

    def __init__(self, mpc_config, robot_config, device="cuda:0", dtype=torch.float32, engine=None, **solver_knobs):
        self._b = BatchedModelPredictiveController(mpc_config, robot_config, 1, device=device, dtype=dtype, engine=engine,
                                                   **solver_knobs)
        self.num_state, self.num_input = 13, 12

    def __getattr__(self, name):                              # iterations_between_mpc, horizon, dt, ...
        return getattr(self._b, name)

    @property
    def ref_traj(self):
        return self._b.ref_traj[0].cpu().numpy()

    def update_robot_state(self, robot_data) -> None:
        rd = BatchedRobotData(np.asarray(robot_data.quat_base, dtype=np.float64)[None],
                              np.asarray(robot_data.pos_base, dtype=np.float64)[None],
                              np.asarray(robot_data.ang_vel_base, dtype=np.float64)[None],
                              np.asarray(robot_data.lin_vel_base, dtype=np.float64)[None],
                              np.stack([np.asarray(p, dtype=np.float64) for p in robot_data.pos_base_feet])[None],
                              np.asarray(robot_data.R_base, dtype=np.float64)[None])
        self._b.update_robot_state(rd)

    def update_mpc_if_needed(self, iter_counter, base_vel_base_des, yaw_turn_rate_des, gait_table,
                             solver="drake", debug=False, iter_debug=None) -> np.ndarray:
        f = self._b.update_mpc_if_needed(iter_counter, np.asarray(base_vel_base_des, dtype=np.float64)[None],
                                         np.asarray([yaw_turn_rate_des], dtype=np.float64),
                                         np.asarray(gait_table, dtype=np.float32)[None], solver=solver)
        return f[0].to(torch.float64).cpu().numpy()
