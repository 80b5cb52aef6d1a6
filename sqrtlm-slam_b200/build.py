"""Builds the sm_100a shared library in-tree (libsqrtba.so next to this file) with nvcc.

`python sqrtlm-slam_b200/build.py` or build.build_lib(); __graft_entry__.build() calls this."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libsqrtba.so")
SOURCES = [os.path.join(HERE, "csrc", "sqrtba_solver.cu")]
DEPS = SOURCES + sorted(os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc")) if f.endswith(".cuh")) + [
    os.path.join(HERE, "host", "host_pool.h"), os.path.join(os.path.dirname(HERE), "include", "sqrtba.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build_variant(out_name: str, extra_flags: list[str]) -> str:
    """Instrumented / experimental builds next to the product library (never loaded by default)."""
    out = os.path.join(HERE, out_name)
    res = subprocess.run([nvcc_path()] + NVCC_FLAGS + extra_flags + ["-o", out] + SOURCES, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building " + out_name)
    return out


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-o", LIB] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libsqrtba.so")
    with open(os.path.join(HERE, "ptxas_info.txt"), "w") as f:  # register / spill report of the last build
        f.write(res.stderr)
    return LIB


HOST_LIB = os.path.join(HERE, "libsqrtba_host.so")
HOST_SOURCES = [os.path.join(HERE, "host", f) for f in ("sqrtbaOptimizer.cc", "harness.cc")]
HOST_DEPS = HOST_SOURCES + [os.path.join(HERE, "host", f) for f in ("Optimizer.h", "map_types.h", "host_pool.h", "map_mirror.h")]


def build_host(force: bool = False) -> str:
    """The C++ adapter (reference Optimizer API over the C ABI) + its test harness, against the header-compatible
    map types.  Plain g++; links libsqrtba.so through an $ORIGIN rpath."""
    build_lib()
    if not force and os.path.exists(HOST_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(HOST_LIB) for d in HOST_DEPS + [LIB]):
        return HOST_LIB
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", HOST_LIB] + HOST_SOURCES + [
        "-L" + HERE, "-lsqrtba", "-Wl,-rpath,$ORIGIN", "-pthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building libsqrtba_host.so")
    return HOST_LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose=True))
    print(build_host(force="--force" in sys.argv))
