"""Builds libmc2d.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc.

`python -m montecarlo_simulation_b200.build` or `build_library()`.  The .so is git-ignored but
travels to the GPU box with the snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libmc2d.so")
SOURCES = ["mc2d.cu"]
HEADERS = ["mc2d_kernels.cuh", "sweep_kernel.cuh", "sweep_kernel_p.cuh", "group_kernels.cuh","pair_potentials.cuh", "philox.h", os.path.join("..", "..", "include", "mc2d.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    # the image exports CC/CXX=/opt/gcc/bin/* (wrappers that link libstdc++ statically); a second
    # libstdc++ inside a Python process breaks iostreams, so pin the PATH compiler.
    "-ccbin", "/usr/bin/g++",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(tag: str, defines: list) -> str:
    """A side build for A/B measurements and fault injection: libmc2d_<tag>.so next to the product library (load it with
    MC2D_LIB=<path>).  `python -m montecarlo_simulation_b200.build --variant fi -DMC2D_FAULT_INJECTION` makes the library
    tools/check_multi_gpu.py uses to hold odd ranks back inside mc2d_upload; the product build has no such hook."""
    out = os.path.join(PKG, "libmc2d_%s.so" % tag)
    cmd = [_nvcc()] + NVCC_FLAGS + list(defines) + ["-I", os.path.join(PKG, "..", "include"), "-o", out] + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed building %s" % out)
    return out


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    extra = os.environ.get("MC2D_NVCC_EXTRA", "").split()  # e.g. -DMC2D_SWEEP_MIN_BLOCKS=4 for tuning
    cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-I", os.path.join(PKG, "..", "include"), "-o", LIB] + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    with open(os.path.join(PKG, "build.log"), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + log)
    if proc.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libmc2d.so")
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build_library(force="--force" in sys.argv, verbose=True))
