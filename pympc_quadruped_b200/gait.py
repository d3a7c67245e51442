"""Host-side contact schedule: produces the gait contact table the MPC consumes.

The table layout is part of the drop-in boundary: float32
``[4*horizon]``, step-major / leg-minor, legs FL, FR, RL, RR, 1 = stance,
first entry is step t+1.  Rule and named patterns follow
linear_mpc/gait.py:16-22 (patterns) and :76-100 (`set_iteration`,
`get_gait_table`); swing/stance phase follow :102-135.

Unlike the reference (an Enum whose members are process-wide singletons that read
``LinearMpcConfig.horizon`` at import time, linear_mpc/gait.py:47-50), a
`GaitSchedule` is an ordinary object with an explicit horizon, and
`gait_tables()` is the vectorised form used to feed a batch of environments.
BOUNDING10 does not exist in the reference (its bound is commented out,
linear_mpc/gait.py:20); it is synthesised with the same rule for BASELINE
config 3.
"""
from __future__ import annotations

import numpy as np

from .configs import LinearMpcConfig


class GaitSchedule:
    def __init__(self, name, num_segment, stance_offsets, stance_durations, horizon=None,
                 dt_control=LinearMpcConfig.dt_control,
                 iterations_between_mpc=LinearMpcConfig.iteration_between_mpc):
        self.name = name
        self.num_segment = int(num_segment)
        self.stance_offsets = np.asarray(stance_offsets, dtype=np.int64)
        self.stance_durations = np.asarray(stance_durations, dtype=np.int64)
        self.horizon = int(LinearMpcConfig.horizon if horizon is None else horizon)
        self._dt_control = float(dt_control)
        self._iterations_between_mpc = int(iterations_between_mpc)
        self.total_swing_time = int(self.num_segment - self.stance_durations[0])
        self.total_stance_time = int(self.stance_durations[0])
        self.stance_offsets_normalized = self.stance_offsets / self.num_segment
        self.stance_durations_normalized = self.stance_durations / self.num_segment
        self.iteration = 0.0
        self.phase = 0.0

    def with_horizon(self, horizon: int) -> "GaitSchedule":
        return GaitSchedule(self.name, self.num_segment, self.stance_offsets, self.stance_durations,
                            horizon, self._dt_control, self._iterations_between_mpc)

    # --- reference API -----------------------------------------------------
    def set_iteration(self, iterations_between_mpc: int, cur_iteration: int) -> None:
        self.iteration = np.floor(cur_iteration / iterations_between_mpc) % self.num_segment
        period = iterations_between_mpc * self.num_segment
        self.phase = (cur_iteration % period) / period

    def get_gait_table(self) -> np.ndarray:
        return gait_tables(self.stance_offsets[None], self.stance_durations[None],
                           np.array([self.num_segment]), np.array([self.iteration]), self.horizon)[0]

    def get_swing_state(self) -> np.ndarray:
        swing_offsets = self.stance_offsets_normalized + self.stance_durations_normalized
        # the reference subtracts 1 from the WHOLE vector each time one entry exceeds 1
        # (linear_mpc/gait.py:104-106); reproduced as written.
        for i in range(4):
            if swing_offsets[i] > 1:
                swing_offsets = swing_offsets - 1
        swing_durations = 1 - self.stance_durations_normalized
        state = np.full(4, self.phase, dtype=np.float32) - swing_offsets
        for i in range(4):
            if state[i] < 0:
                state[i] += 1
            state[i] = 0 if state[i] > swing_durations[i] else state[i] / swing_durations[i]
        return state

    def get_stance_state(self) -> np.ndarray:
        state = np.full(4, self.phase, dtype=np.float32) - self.stance_offsets_normalized
        for i in range(4):
            if state[i] < 0:
                state[i] += 1
            d = self.stance_durations_normalized[i]
            state[i] = 0 if state[i] > d else state[i] / d
        return state

    @property
    def swing_time(self) -> float:
        return self.get_total_swing_time(self._dt_control * self._iterations_between_mpc)

    @property
    def stance_time(self) -> float:
        return self.get_total_stance_time(self._dt_control * self._iterations_between_mpc)

    def get_total_swing_time(self, dt_mpc: float) -> float:
        return dt_mpc * self.total_swing_time

    def get_total_stance_time(self, dt_mpc: float) -> float:
        return dt_mpc * self.total_stance_time


def gait_tables(stance_offsets, stance_durations, num_segment, iteration, horizon) -> np.ndarray:
    """Contact tables for a batch: float32 [B, 4*horizon].

    table[b, 4*i+j] = 1 if ((i + 1 + iteration_b - offset_bj) mod num_segment_b) < duration_bj
    (linear_mpc/gait.py:88-98).
    """
    off = np.asarray(stance_offsets, dtype=np.int64).reshape(-1, 1, 4)
    dur = np.asarray(stance_durations, dtype=np.int64).reshape(-1, 1, 4)
    seg = np.asarray(num_segment, dtype=np.int64).reshape(-1, 1, 1)
    it = np.asarray(iteration).astype(np.int64).reshape(-1, 1, 1)
    steps = np.arange(1, horizon + 1, dtype=np.int64).reshape(1, -1, 1)
    cur = np.mod(np.mod(steps + it, seg) - off, seg)
    return (cur < dur).astype(np.float32).reshape(cur.shape[0], 4 * horizon)


class BatchedGaitSchedule:
    """Per-environment gait patterns kept on the device: the batch counterpart of the
    reference's `Gait` members (linear_mpc/gait.py:76-135).  `set_iteration` + `get_gait_table` run as ONE elementwise
    kernel (`mpcq_gait_tables`) and hand `update_mpc_if_needed` a device tensor; nothing is computed on the host."""

    def __init__(self, engine, schedules):
        import torch
        self.engine = engine
        self.horizon = engine.horizon
        dev = engine.device
        scheds = list(schedules)
        self.num_envs = len(scheds)
        seg = np.array([s.num_segment for s in scheds], dtype=np.int32)
        if np.any(seg < 1):
            raise ValueError("num_segment must be >= 1")
        self.stance_offsets = torch.as_tensor(np.stack([s.stance_offsets for s in scheds]).astype(np.int32), device=dev)
        self.stance_durations = torch.as_tensor(np.stack([s.stance_durations for s in scheds]).astype(np.int32), device=dev)
        self.num_segment = torch.as_tensor(seg, device=dev)
        self._table = torch.zeros((self.num_envs, 4 * self.horizon), dtype=torch.float32, device=dev)
        self._swing = torch.zeros((self.num_envs, 4), dtype=torch.float64, device=dev)
        self._stance = torch.zeros((self.num_envs, 4), dtype=torch.float64, device=dev)
        self._cur = torch.zeros((self.num_envs,), dtype=torch.int32, device=dev)
        # Gait.swing_time / stance_time per robot (linear_mpc/gait.py:68-74), float64 like the reference's np.float64 products
        self.swing_time = torch.as_tensor(np.array([float(s.swing_time) for s in scheds]), device=dev)
        self.stance_time = torch.as_tensor(np.array([float(s.stance_time) for s in scheds]), device=dev)

    def set_iteration(self, iterations_between_mpc: int, cur_iteration) -> None:
        """`cur_iteration`: one control tick for all robots (int) or an int tensor [B] (robots out of phase)."""
        import torch
        if torch.is_tensor(cur_iteration):
            self._cur.copy_(cur_iteration.to(torch.int32))
        else:
            self._cur.fill_(int(cur_iteration))
        self.engine.gait_tables(self.stance_offsets, self.stance_durations, self.num_segment, self._cur,
                                int(iterations_between_mpc), table=self._table, swing_state=self._swing,
                                stance_state=self._stance)

    def get_gait_table(self):
        return self._table

    def get_swing_state(self):
        return self._swing

    def get_stance_state(self):
        return self._stance


class Gait:
    """Named patterns, same names as the reference Enum (linear_mpc/gait.py:16-22)."""
    STANDING = GaitSchedule('standing', 16, [0, 0, 0, 0], [16, 16, 16, 16])
    TROTTING16 = GaitSchedule('trotting', 16, [0, 8, 8, 0], [8, 8, 8, 8])
    TROTTING10 = GaitSchedule('trotting', 10, [0, 5, 5, 0], [5, 5, 5, 5])
    JUMPING16 = GaitSchedule('jumping', 16, [0, 0, 0, 0], [4, 4, 4, 4])
    PACING16 = GaitSchedule('pacing', 16, [8, 0, 8, 0], [8, 8, 8, 8])
    PACING10 = GaitSchedule('pacing', 10, [5, 0, 5, 0], [5, 5, 5, 5])
    # not in the reference
    BOUNDING10 = GaitSchedule('bounding', 10, [5, 5, 0, 0], [5, 5, 5, 5])
