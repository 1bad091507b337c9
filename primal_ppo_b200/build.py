"""Builds libmapf_b200.so (the C-ABI shared library of include/mapf_b200.h) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import fcntl
import os
import subprocess
import sys

_PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_PKG, "csrc")
LIB_PATH = os.path.join(_PKG, "libmapf_b200.so")
SOURCES = ["mapf_api.cu", "step.cu", "step_wide.cu", "observe.cu", "observe_wide.cu", "step_observe.cu", "bfs.cu", "gae.cu", "glue.cu", "scenario_gen.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--use_fast_math=false",
              "-Xcompiler", "-fPIC,-fvisibility=default", "-shared", "-cudart", "shared", "--threads", "0"]


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(_PKG), "include", "mapf_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not (force or _stale()):
        return LIB_PATH
    # several ranks of one job may get here at once (torchrun): one of them builds, the others wait and re-check
    with open(os.path.join(_PKG, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not (force or _stale()):
                return LIB_PATH
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
    cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB_PATH + ".tmp"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libmapf_b200.so")
    os.replace(LIB_PATH + ".tmp", LIB_PATH)          # atomic: a concurrently loading process never sees a partial file
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
