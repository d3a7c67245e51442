"""Build libmpcq.so (sm_100a) in-tree: pympc_quadruped_b200/csrc/_build/libmpcq.so.

nvcc cross-compiles without a GPU; the built library travels to the GPU box with the repo
snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", os.environ.get("MPCQ_OUT", "libmpcq.so"))   # MPCQ_OUT: experiment builds beside the product
SOURCES = ["mpcq_api.cu", "mpcq_probe.cu"]
DEPS = SOURCES + ["mpcq_core.cuh", "mpcq_legs.cuh", "mpcq_warp.cuh", "mpcq_host.h", os.path.join("..", "..", "include", "mpcq.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def up_to_date() -> bool:
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(os.path.join(HERE, d)) <= t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("MPCQ_EXTRA_NVCC_FLAGS", "").split()          # e.g. -DMPCQ_NW0=1 (team-size experiments)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SOURCES
    res = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
