"""Builds libddm_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container;
the built .so travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libddm_b200.so")
SOURCES = ["ddm_kernels.cu", "ddm_capi.cu", "ddm_microbench.cu", "ddm_evidence.cu", "ddm_prior.cu", "ddm_reduce.cu", "ddm_exact.cu"]
HOST_SOURCES = ["ddm_wire.cpp"]  # g++ only: host threads of the compact device->host wire format
HEADERS = ["ddm_kernels.cuh", "ddm_rng.cuh", "ddm_microbench.cuh", "ddm_wire.cuh",
           os.path.join("..", "..", "include", "ddm_b200.h"),
           os.path.join("..", "..", "include", "ddm_dlpack.h")]

# -fmad=false: the fp64 validation kernels must follow the reference's unfused operation
# order (numba never contracts a*b+c); the fp32 production path asks for FMAs explicitly.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-fmad=false", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared"]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libddm_b200.so")
    return nvcc


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HOST_SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, out_path: str | None = None, defines=()) -> str:
    """Compile the library if missing or older than its sources; return its path.
    ``out_path`` / ``defines`` build an experimental variant next to it (A/B measurements)."""
    if out_path is None and not force and not is_stale():
        return LIB_PATH
    target = out_path or LIB_PATH
    objs = []
    for src in HOST_SOURCES:
        obj = os.path.join(CSRC, os.path.splitext(src)[0] + ".o")
        gcc = [os.environ.get("CXX", "g++"), "-O3", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-ffp-contract=off", "-pthread",
               "-c", src, "-o", obj]
        proc = subprocess.run(gcc, cwd=CSRC, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("g++ failed:\n" + " ".join(gcc) + "\n" + proc.stdout + proc.stderr)
        objs.append(obj)
    cmd = ([find_nvcc()] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) +
           ["-o", target] + SOURCES + objs)
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    for obj in objs:
        os.remove(obj)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return target


PHILOX7_LIB_PATH = os.path.join(HERE, "libddm_b200_philox7.so")


def build_philox7(force: bool = False) -> str:
    """The measurement-only 7-round variant (ddm_rng.cuh: DDM_PHILOX_ROUNDS), selected with DDM_B200_LIB."""
    if not force and os.path.exists(PHILOX7_LIB_PATH) and not any(
            os.path.getmtime(os.path.join(CSRC, f)) > os.path.getmtime(PHILOX7_LIB_PATH)
            for f in SOURCES + HOST_SOURCES + HEADERS if os.path.exists(os.path.join(CSRC, f))):
        return PHILOX7_LIB_PATH
    return build(out_path=PHILOX7_LIB_PATH, defines=("DDM_PHILOX_ROUNDS=7",))


CHECKED_LIB_PATH = os.path.join(HERE, "libddm_b200_checked.so")


def build_checked() -> str:
    """The bounds-checked build (ddm_kernels.cuh: DDM_CHECKED), the in-tree stand-in for compute-sanitizer's memcheck;
    select it with DDM_B200_LIB (scripts/r02_checked_run.py)."""
    return build(out_path=CHECKED_LIB_PATH, defines=("DDM_CHECKED",))


if __name__ == "__main__":
    print(build(force=True, verbose=True))
