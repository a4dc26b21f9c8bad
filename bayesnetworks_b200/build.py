"""Build libbn_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m bayesnetworks_b200.build

The shared library lands next to this file so that it travels with the repo
snapshot to the GPU box; it is git-ignored (source-only history).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbn_b200.so")
SOURCES = ["bn_api.cu", "gram.cu", "kernels.cu"]
HEADERS = ["bn_common.cuh", "chain_core.cuh", "rng_core.cuh", "score_core.cuh", "gram.h", "kernels.h",
           os.path.join("..", "..", "include", "bn_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--shared", "-Xptxas", "-v"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, f) for f in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libbn_b200.so")
    log = os.path.join(HERE, "csrc", "_ptxas.log")
    with open(log, "w") as fh:
        fh.write(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
