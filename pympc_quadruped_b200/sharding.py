"""Environment sharding across the GPUs of one box (one process per GPU).

Every environment is an independent QP (the reference's multi-robot loop is independent objects,
scripts/isaacgym_a1.py:92-96,119), so the batch is cut into contiguous ranges, one per rank, and
the solve loop needs NO collective.  NCCL (gloo in CPU tests) is used only for the optional
all-gather of the [B,12] ground-reaction forces and a few statistics scalars.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(num_envs: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) of rank `rank`; sizes differ by at most one, earlier ranks get the extra env."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(num_envs), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_forces(local: torch.Tensor, num_envs: int) -> torch.Tensor:
    """All-gather the per-rank [b_r,12] forces into [num_envs,12] on every rank (ragged shards allowed)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(num_envs, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * width, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * width: r * width + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)


def reduce_stats(iters_sum: float, iters_max: float, unverified: int, device) -> dict:
    """Sum / max of solver statistics over ranks."""
    s = torch.tensor([float(iters_sum), float(unverified)], dtype=torch.float64, device=device)
    m = torch.tensor([float(iters_max)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return {"iters_sum": float(s[0]), "unverified": int(s[1]), "iters_max": float(m[0])}
