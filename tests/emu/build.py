"""Build the TEST-ONLY host emulation of the device code (tests/emu/_build/libmpcq_emu.so).

g++ compiles pympc_quadruped_b200/csrc/mpcq_core.cuh with MPCQ_HOST_EMU: the 32 lanes of a warp
run as lock-step coroutines (tests/emu/mpcq_emu.cpp).  Used by `-m "not gpu"` tests to exercise
the kernel logic without a GPU.  Never loaded by the package.
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "..", "pympc_quadruped_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libmpcq_emu.so")
DEPS = [os.path.join(HERE, "mpcq_emu.cpp")] + [os.path.join(CSRC, f) for f in ("mpcq_core.cuh", "mpcq_legs.cuh", "mpcq_warp.cuh", "mpcq_host.h")]


def build() -> str:
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in DEPS):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", OUT, os.path.join(HERE, "mpcq_emu.cpp")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    return OUT


if __name__ == "__main__":
    print(build())
