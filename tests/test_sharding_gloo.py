"""The N > 1 host path on CPU: two gloo ranks each solve their contiguous shard (with the oracle-backed test
double standing in for the CUDA engine) and all-gather the forces; the result must equal the one-rank solve
bit for bit - environments are independent, no collective in the solve loop."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp

from pympc_quadruped_b200.sharding import shard_range

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_shard_ranges_partition_the_batch():
    for B in (0, 1, 7, 4096, 262144):
        for world in (1, 2, 3, 8):
            r = [shard_range(B, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
            assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, B, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fake_engine import OracleEngine
    from helpers import make_batch
    from pympc_quadruped_b200 import A1Config, Gait
    from pympc_quadruped_b200.sharding import gather_forces, reduce_stats, shard_range
    batch = make_batch(A1Config, 10, B, "mixed", (Gait.TROTTING10,), 5, solve=False)
    eng = OracleEngine(batch["cfg"], A1Config, torch.float64)
    lo, hi = shard_range(B, rank, world)
    t = lambda a: torch.as_tensor(a[lo:hi])
    res = eng.solve(t(batch["x0"]), t(batch["feet"]), t(batch["gait"]), t(batch["xref"]), yaw=t(batch["yaw"]))
    full = gather_forces(res.forces, B)
    stats = reduce_stats(float(hi - lo), float(hi - lo), 0, "cpu")
    assert stats["iters_sum"] == B
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), full.numpy())
    dist.destroy_process_group()


def test_two_rank_sharded_solve_equals_single_rank(tmp_path):
    B, world = 7, 2                                   # ragged on purpose: shards of 4 and 3
    port = 29500 + (os.getpid() % 2000)
    tmp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fake_engine import OracleEngine
    from helpers import make_batch
    from pympc_quadruped_b200 import A1Config, Gait
    batch = make_batch(A1Config, 10, B, "mixed", (Gait.TROTTING10,), 5, solve=False)
    eng = OracleEngine(batch["cfg"], A1Config, torch.float64)
    t = torch.as_tensor
    ref = eng.solve(t(batch["x0"]), t(batch["feet"]), t(batch["gait"]), t(batch["xref"]), yaw=t(batch["yaw"])).forces.numpy()
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"rank{r}.npy"))
        assert got.shape == (B, 12) and np.array_equal(got, ref)
