"""TEST INFRASTRUCTURE - ctypes front end of oracle/_ref/libhpmvs_ref.so: the REFERENCE'S OWN sources
(/root/reference/src/hpmvs/*.cpp + vendored nlopt/CImg/stlplus3), compiled where they lie by `make -C oracle refhpmvs`
against the stand-in headers in oracle/shim/ (Eigen, glog, gflags, jpeglib are absent from this image).

Used to pin the restatement (oracle/hpmvs_oracle.cpp) against the real code path and as bench.py's reference arm.
The reference reads its scene from an NVM file + image files, so scenes are written to disk first (PPM level-0 images;
CImg's PNM reader, no JPEG codec involved).  Only tests/, smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
from typing import Optional

import numpy as np

from . import Camera, Options, PATCH_DTYPE, _p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libhpmvs_ref.so")
BIN_PATH = os.path.join(_HERE, "_ref", "hpmvs_ref")
DROPIN_BIN_PATH = os.path.join(_HERE, "_ref", "hpmvs_ref_b200")   # same CLI, integration/PatchOptimizer_b200.cpp instead of PatchOptimizer.cpp
REF_ROOT = "/root/reference"
FAIL = 100   # status of a patch for which the reference's optimize() returned false


_tried_build = False


def available() -> bool:
    """True when oracle/_ref/libhpmvs_ref.so exists; where the reference's sources are present it is built on first use."""
    global _tried_build
    if not os.path.exists(LIB_PATH) and not _tried_build and os.path.isdir(os.path.join(REF_ROOT, "src", "hpmvs")):
        _tried_build = True
        try:
            build()
        except Exception:
            pass
    return os.path.exists(LIB_PATH)


def build() -> Optional[str]:
    """Build oracle/_ref/libhpmvs_ref.so + hpmvs_ref where the reference's sources are present; else keep the prebuilt files."""
    if os.path.isdir(os.path.join(REF_ROOT, "src", "hpmvs")):
        subprocess.run(["make", "-C", _HERE, "ref", "-j8"], check=True, capture_output=True)
        subprocess.run(["make", "-C", _HERE, "refhpmvs", "-j8"], check=True, capture_output=True)
        if os.path.exists(os.path.join(_HERE, "..", "hpmvs_b200", "libhpmvs_b200.so")):
            # the reference's CLI linked against the engine instead of its own PatchOptimizer.cpp (integration/)
            subprocess.run(["make", "-C", _HERE, "dropin"], check=True, capture_output=True)
    return LIB_PATH if os.path.exists(LIB_PATH) else None


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libhpmvs_ref.so is missing (built only where /root/reference exists)")
        L = C.CDLL(LIB_PATH)
        vp, ip, fp, u8p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_uint8)
        L.refh_scene_load.restype = vp; L.refh_scene_load.argtypes = [C.c_char_p, C.POINTER(Options)]
        L.refh_scene_free.argtypes = [vp]
        L.refh_num_cameras.argtypes = [vp]; L.refh_num_points.argtypes = [vp]
        L.refh_save_nvm.argtypes = [vp, C.c_char_p]
        L.refh_get_camera.argtypes = [vp, C.c_int, C.POINTER(Camera)]
        L.refh_get_image.argtypes = [vp, C.c_int, C.c_int, u8p, C.c_int, ip, ip]
        L.refh_get_covis.argtypes = [vp, C.c_int, ip, C.c_int]
        L.refh_get_color.argtypes = [vp, C.c_int, C.c_float, C.c_float, C.c_int, fp]
        L.refh_project.argtypes = [vp, C.c_int, fp, C.c_int, fp]
        L.refh_scale_level.argtypes = [vp, C.c_int, fp, C.c_float, C.c_int, C.c_int, fp, fp, ip]
        L.refh_optimize_batch.argtypes = [vp, C.c_int, C.c_void_p, C.c_int]
        L.refh_init_patches.argtypes = [vp, C.c_void_p, C.c_int]
        L.refh_depth_reset.argtypes = [vp]
        L.refh_depth_set_batch.argtypes = [vp, C.c_int, C.c_void_p]
        L.refh_depth_unset_batch.argtypes = [vp, C.c_int, C.c_void_p]
        L.refh_get_depth.argtypes = [vp, C.c_int, C.c_int, fp, C.c_int, ip, ip]
        L.refh_accept_batch.argtypes = [vp, C.c_int, C.c_void_p, C.c_float, ip]
        _lib = L
    return _lib


class RefScene:
    """mo3d::Scene of the reference, loaded from an NVM file exactly as src/main.cpp:104-113 does."""

    def __init__(self, nvm_path: str, options: Optional[Options] = None, _tmp=None):
        self.options = options or Options.defaults()
        self._tmp = _tmp
        self._h = lib().refh_scene_load(nvm_path.encode(), C.byref(self.options))
        if not self._h:
            raise RuntimeError(f"reference failed to load {nvm_path}")

    def __del__(self):
        try:
            if self._h:
                lib().refh_scene_free(self._h)
                self._h = None
        except Exception:
            pass

    @classmethod
    def from_synth(cls, scene, options: Optional[Options] = None) -> "RefScene":
        import hpmvs_b200 as hp   # scene writer only (NVM_V3 text + PPM files)
        tmp = tempfile.TemporaryDirectory(prefix="hpmvs_ref_")
        path = os.path.join(tmp.name, "scene.nvm")
        hp.synth.write_nvm(scene, path)
        return cls(path, options, _tmp=tmp)

    def save_nvm(self, path: str) -> None:
        """NVMReader::saveNVM of the loaded model."""
        lib().refh_save_nvm(self._h, path.encode())

    @property
    def n_cameras(self) -> int:
        return lib().refh_num_cameras(self._h)

    def camera(self, i: int) -> Camera:
        c = Camera()
        lib().refh_get_camera(self._h, i, C.byref(c))
        return c

    def image(self, cam: int, level: int) -> np.ndarray:
        w, h = C.c_int32(), C.c_int32()
        need = lib().refh_get_image(self._h, cam, level, None, 0, C.byref(w), C.byref(h))
        out = np.zeros(need, np.uint8)
        lib().refh_get_image(self._h, cam, level, _p(out, C.c_uint8), need, C.byref(w), C.byref(h))
        return out.reshape(h.value, w.value, 3)

    def covis(self):
        out = []
        buf = np.zeros(4096, np.int32)
        for i in range(self.n_cameras):
            n = lib().refh_get_covis(self._h, i, _p(buf, C.c_int32), len(buf))
            out.append(buf[:n].tolist())
        return out

    def get_color(self, cam: int, x: float, y: float, level: int) -> np.ndarray:
        out = np.zeros(3, np.float32)
        lib().refh_get_color(self._h, cam, float(x), float(y), level, _p(out, C.c_float))
        return out

    def project(self, cam: int, X, level: int) -> np.ndarray:
        Xa = np.asarray(X, np.float32); out = np.zeros(3, np.float32)
        lib().refh_project(self._h, cam, _p(Xa, C.c_float), level, _p(out, C.c_float))
        return out

    def scale_level(self, cam: int, X, scale: float, level: int, max_level: int):
        Xa = np.asarray(X, np.float32)
        s, l, li = C.c_float(), C.c_float(), C.c_int32()
        lib().refh_scale_level(self._h, cam, _p(Xa, C.c_float), float(scale), level, max_level, C.byref(s), C.byref(l), C.byref(li))
        return s.value, l.value, li.value

    def optimize_batch(self, patches: np.ndarray, nthreads: int = 1) -> np.ndarray:
        """PatchOptimizer::optimize on every record; status 0 = true, FAIL = false (fields untouched, as the reference)."""
        p = np.ascontiguousarray(patches.copy())
        assert p.dtype == PATCH_DTYPE
        lib().refh_optimize_batch(self._h, len(p), p.ctypes.data, int(nthreads))
        return p

    def init_patches(self, cap: int = 1 << 20) -> np.ndarray:
        """Scene::initPatches (seeding + optimize + tree insertion); returns the patches found in the octree."""
        out = np.zeros(cap, PATCH_DTYPE)
        n = lib().refh_init_patches(self._h, out.ctypes.data, cap)
        return out[:min(n, cap)].copy()

    def depth_reset(self) -> None:
        lib().refh_depth_reset(self._h)

    def depth_unset(self, patches: np.ndarray) -> None:
        p = np.ascontiguousarray(patches)
        lib().refh_depth_unset_batch(self._h, len(p), p.ctypes.data)

    def depth_set(self, patches: np.ndarray) -> None:
        p = np.ascontiguousarray(patches)
        lib().refh_depth_set_batch(self._h, len(p), p.ctypes.data)

    def depth(self, cam: int, level: int) -> np.ndarray:
        r, c = C.c_int32(), C.c_int32()
        need = lib().refh_get_depth(self._h, cam, level, None, 0, C.byref(r), C.byref(c))
        out = np.zeros(need, np.float32)
        lib().refh_get_depth(self._h, cam, level, _p(out, C.c_float), need, C.byref(r), C.byref(c))
        return out.reshape(r.value, c.value)

    def accept(self, patches: np.ndarray, margin: float = 1.0) -> np.ndarray:
        p = np.ascontiguousarray(patches)
        out = np.zeros((len(p), 3), np.int32)
        lib().refh_accept_batch(self._h, len(p), p.ctypes.data, float(margin), _p(out, C.c_int32))
        return out


def run_cli(nvm_path: str, outdir: str, threads: int = 1, extra=(), dropin: bool = False, monotone_heap: bool = False) -> subprocess.CompletedProcess:
    """The reference's own command line (src/main.cpp): hpmvs --nvm=... --outdir=...; dropin=True runs the build whose
    PatchOptimizer is the B200 engine (needs a GPU); monotone_heap=True runs the *_det builds (oracle/ref_monotone_new.cpp)."""
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    exe = (DROPIN_BIN_PATH if dropin else BIN_PATH) + ("_det" if monotone_heap else "")
    return subprocess.run([exe, f"--nvm={nvm_path}", f"--outdir={outdir}", *extra], env=env, capture_output=True, text=True)
