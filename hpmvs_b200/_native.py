"""Build and load the native library (hpmvs_b200/libhpmvs_b200.so): CUDA kernels + C ABI + host scene surface.

The library is built IN-TREE with nvcc for sm_100a only; there is no fallback implementation - if it is
missing or cannot be loaded every entry point of this package raises.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HPMVS_LIB") or os.path.join(_HERE, "libhpmvs_b200.so")   # HPMVS_LIB: A/B experiments only
SOURCES = [os.path.join(_HERE, "csrc", "engine.cu"), os.path.join(_HERE, "csrc", "host_scene.cpp"),
           os.path.join(_HERE, "csrc", "host_io.cpp"), os.path.join(_HERE, "csrc", "host_pipeline.cpp")]
HEADERS = [os.path.join(_HERE, "csrc", "patch_kernels.cuh"), os.path.join(_HERE, "csrc", "patch_kernels_wf.cuh"),
           os.path.join(_HERE, "csrc", "bobyqa3.h"), os.path.join(_HERE, "csrc", "undistort_math.h"),
           os.path.join(_HERE, "..", "include", "hpmvs_b200.h")]

# -fmad=false / -ffp-contract=off: the kernels restate the reference's f32/f64 evaluation order (see
# patch_kernels.cuh); FMA contraction would change roundings and break the reproducible BOBYQA trajectory.
# -DBQ_DEFER_TRUST=1: a second trust-region step inside one optimizer round is deferred to the next round (bobyqa3.h)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false", "-DBQ_DEFER_TRUST=1",
              "-Xcompiler", "-fPIC,-ffp-contract=off", "-diag-suppress", "550,177", "-shared", "-ldl"]


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.exists(p) and os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the extension for sm_100a (cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    extra = ["-DHP_PROFILE"] if os.environ.get("HPMVS_BUILD_PROFILE") else []
    extra += os.environ.get("HPMVS_BUILD_DEFS", "").split()          # e.g. "-DBQ_DEFER_TRUST=8" for A/B builds (with HPMVS_LIB)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """The loaded C-ABI library; raises if it is missing (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
                build()
            else:
                raise RuntimeError(f"{LIB_PATH} is missing and nvcc is not available: hpmvs_b200 has no fallback path")
        _lib = C.CDLL(LIB_PATH)
    return _lib
