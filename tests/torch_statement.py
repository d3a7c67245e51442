"""TEST INFRASTRUCTURE (not product code): the arithmetic of the device kernel `mpcq_assemble` stated in eager PyTorch.

`assemble_statement` restates, tensor op by tensor op, what the kernel computes (mpc.py:55-79 state assembly, :81-93
command integration, :110-170 reference trajectory, with the reference's float32 storage / float64 scalar arithmetic) with
the same in-place calling convention as `MpcqEngine.assemble`.  GPU tests compare the kernel against it bit for bit; the
CPU-only tests plug it into the fake engine.  The product (pympc_quadruped_b200.controller) has the device path only.
"""
from __future__ import annotations

import torch


def quat_to_matrix(q: torch.Tensor) -> torch.Tensor:
    """(w,x,y,z) -> R_base, batched (utils/kinematics.py:51-71)."""
    w, x, y, z = q.unbind(-1)
    R = torch.stack([
        w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (w * y + x * z),
        2 * (w * z + x * y), w * w - x * x + y * y - z * z, 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (w * x + y * z), w * w - x * x - y * y + z * z], dim=-1)
    return R.reshape(q.shape[:-1] + (3, 3))


def quat_to_zyx(q: torch.Tensor) -> torch.Tensor:
    """(w,x,y,z) -> [roll, pitch, yaw] float64 (utils/kinematics.py:40-49)."""
    w, x, y, z = q.unbind(-1)
    roll = torch.atan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y))
    pitch = torch.asin(2 * (w * y - z * x))
    yaw = torch.atan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))
    return torch.stack([roll, pitch, yaw], dim=-1)


def reference_trajectory(c, x, yaw_des, xy_des, rp_init, vel_des, yaw_rate):
    """mpc.py:110-170; x = float32 current_state as float64; xy_des / rp_init updated in place.  -> float32 [B, 13 H]"""
    H, n, B = c["horizon"], 13, x.shape[0]
    dt = c["dt"]
    lim = 0.1
    xd, yd = xy_des[:, 0].clone(), xy_des[:, 1].clone()
    xd = torch.where(xd - x[:, 3] > lim, x[:, 3] + lim, xd)
    xd = torch.where(x[:, 3] - xd > lim, x[:, 3] - lim, xd)
    yd = torch.where(yd - x[:, 4] > lim, x[:, 4] + lim, yd)
    yd = torch.where(x[:, 4] - yd > lim, x[:, 4] - lim, yd)
    xy_des[:, 0], xy_des[:, 1] = xd, yd
    safe = lambda v: torch.where(v == 0, torch.ones_like(v), v)
    pitch_init = torch.where(x[:, 9].abs() > 0.2, rp_init[:, 1] + dt * (0.0 - x[:, 1]) / safe(x[:, 9]), rp_init[:, 1])
    roll_init = torch.where(x[:, 10].abs() > 0.1, rp_init[:, 0] + dt * (0.0 - x[:, 0]) / safe(x[:, 10]), rp_init[:, 0])
    rp_init[:, 0] = roll_init.clamp(-0.25, 0.25)
    rp_init[:, 1] = pitch_init.clamp(-0.25, 0.25)
    X = torch.zeros((B, H, n), dtype=torch.float32, device=x.device)
    X[:, :, 0] = (x[:, 10] * rp_init[:, 0]).to(torch.float32)[:, None]
    X[:, :, 1] = (x[:, 9] * rp_init[:, 1]).to(torch.float32)[:, None]
    X[:, :, 5] = c["com_height_des"]
    X[:, :, 8] = yaw_rate.to(torch.float32)[:, None]
    X[:, :, 9] = vel_des[:, 0].to(torch.float32)[:, None]
    X[:, :, 10] = vel_des[:, 1].to(torch.float32)[:, None]
    X[:, :, 12] = -c["gravity"]
    X[:, 0, 2] = yaw_des.to(torch.float32)
    X[:, 0, 3] = xd.to(torch.float32)
    X[:, 0, 4] = yd.to(torch.float32)
    for i in range(1, H):                                # float32 storage, float64 increments
        X[:, i, 2] = (X[:, i - 1, 2].to(torch.float64) + dt * yaw_rate).to(torch.float32)
        X[:, i, 3] = (X[:, i - 1, 3].to(torch.float64) + dt * vel_des[:, 0]).to(torch.float32)
        X[:, i, 4] = (X[:, i - 1, 4].to(torch.float64) + dt * vel_des[:, 1]).to(torch.float32)
    return X.reshape(B, n * H)


def assemble_statement(c, quat, pos, omega, vel, v_des_body, yaw_rate_des, xy_des, yaw_des, rp_init, first_run, do_mpc,
                       x0, yaw, x_ref, R_base=None):
    """Same contract as MpcqEngine.assemble (`c` = extract_mpc_constants dict): state updated in place, x0 / yaw / x_ref written."""
    B = quat.shape[0]
    rpy = quat_to_zyx(quat)
    st = torch.cat([rpy, pos, omega, vel, torch.full((B, 1), -c["gravity"], dtype=torch.float64, device=quat.device)], dim=1)
    st32 = st.to(torch.float32)
    R = quat_to_matrix(quat) if R_base is None else R_base.reshape(B, 3, 3)
    vel_des = torch.einsum("bij,bj->bi", R, v_des_body)
    if first_run:
        xy_des.zero_()
        yaw_des.copy_(rpy[:, 2])
    else:
        xy_des[:, 0] += c["dt_control"] * vel_des[:, 0]
        xy_des[:, 1] += c["dt_control"] * vel_des[:, 1]
        yaw_des.copy_(rpy[:, 2] + c["dt_control"] * yaw_rate_des)
    x0.copy_(st32.to(x0.dtype))
    yaw.copy_(rpy[:, 2].to(yaw.dtype))
    if do_mpc:
        x_ref.copy_(reference_trajectory(c, st32.to(torch.float64), yaw_des, xy_des, rp_init, vel_des, yaw_rate_des).to(x_ref.dtype))


class StatementEngine:
    """An engine whose `assemble` is the torch statement and everything else the wrapped engine's (GPU tests)."""
    def __init__(self, engine, consts):
        self._e, self._c = engine, consts
        self.horizon, self.dtype = engine.horizon, engine.dtype

    def assemble(self, *a, **k):
        assemble_statement(self._c, *a, **k)

    def __getattr__(self, name):
        return getattr(self._e, name)
