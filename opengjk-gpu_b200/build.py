"""Builds libopengjk_b200.so (C ABI + sm_100a kernels) in-tree with nvcc.

    python opengjk-gpu_b200/build.py [--force]

The library is what the reference-facing C++ headers (include/) and the Python mirror
(opengjk-gpu_b200/__init__.py) bind to.  Flags: sm_100a only, -lineinfo for ncu source pages,
--fmad=false as the reference (GJK/CMakeLists.txt:32; the arithmetic additionally uses *_rn intrinsics).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libopengjk_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = [
    "-std=c++17", "-O3", "-shared", "-Xcompiler", "-fPIC",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "--fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-ffp-contract=off",
    "-cudart", "static",
]


def sources():
    return sorted(os.path.join(SRC, f) for f in os.listdir(SRC))


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + [os.path.join(ROOT, "include", "opengjk_b200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cus = [p for p in sources() if p.endswith(".cu")]
    cmd = [NVCC, *FLAGS, "-I", os.path.join(ROOT, "include"), "-I", SRC, *cus, "-o", LIB]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libopengjk_b200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
