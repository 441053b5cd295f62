"""hpmvs_b200 - B200-native patch-optimisation engine behind HPMVS's PatchOptimizer::optimize() boundary.

Everything computational lives in libhpmvs_b200.so (hand-written sm_100a CUDA + a C ABI, see include/hpmvs_b200.h);
this package is the thin Python plumbing used by the tests and bench.py.
"""
from .engine import (Camera, Counters, Engine, HpmvsError, Options, PATCH_DTYPE, STATUS_NAMES, MAX_VIEWS, LEVELS,
                     camera_from_nvm, expand_candidates, extract_covis, seed_patches)
from . import synth  # noqa: F401

__all__ = ["Camera", "Counters", "Engine", "HpmvsError", "Options", "PATCH_DTYPE", "STATUS_NAMES", "MAX_VIEWS", "LEVELS",
           "camera_from_nvm", "expand_candidates", "extract_covis", "seed_patches", "synth"]
