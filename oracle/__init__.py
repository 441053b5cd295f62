"""TEST INFRASTRUCTURE - ctypes front end of the CPU oracle (oracle/libhpmvs_oracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  Nothing under hpmvs_b200/ does.  See oracle/hpmvs_oracle.h for what it restates.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libhpmvs_oracle.so")
MAX_VIEWS = 64
LEVELS = 6

STATUS_NAMES = ["OK", "FAIL_ADD_IMAGES", "FAIL_NCC1", "FAIL_ANGLES", "FAIL_OPT_MINIMAGES", "FAIL_OPT_ROUNDOFF",
                "FAIL_OPT_MAXEVAL", "FAIL_OPT_OTHER", "FAIL_ADD_IMAGES2", "FAIL_NCC2", "FAIL_ANGLE_FILTER",
                "FAIL_ANGLES2", "FAIL_NCC3", "FAIL_TOO_MANY_VIEWS"]


class Options(C.Structure):
    _fields_ = [("maxlevel", C.c_int32), ("minlevel", C.c_int32), ("start_level", C.c_int32),
                ("max_angle", C.c_float), ("min_angle", C.c_float), ("max_images_per_patch", C.c_int32),
                ("min_images_per_patch", C.c_int32), ("ncc_alpha_1", C.c_float), ("ncc_alpha_2", C.c_float)]

    @staticmethod
    def defaults(**kw) -> "Options":
        # HpmvsOptions.h:31-52 (float constants are formed in float there: 60.0f * M_PI / 180.0f)
        o = Options(5, 0, 4, float(np.float32(60.0 * np.pi / 180.0)), float(np.float32(10.0 * np.pi / 180.0)),
                    6, 3, 0.4, 0.5)
        for k, v in kw.items():
            setattr(o, k, v)
        return o


class Camera(C.Structure):
    _fields_ = [("P", C.c_float * 4 * 3 * LEVELS), ("center", C.c_float * 4), ("xaxis", C.c_float * 3),
                ("yaxis", C.c_float * 3), ("zaxis", C.c_float * 3), ("k00", C.c_float), ("k11", C.c_float),
                ("width", C.c_int32 * LEVELS), ("height", C.c_int32 * LEVELS)]


class Patch(C.Structure):
    _fields_ = [("center", C.c_float * 4), ("normal", C.c_float * 4), ("scale", C.c_float), ("nimages", C.c_int32),
                ("images", C.c_int32 * MAX_VIEWS), ("color", C.c_float * 3), ("ncc", C.c_float),
                ("status", C.c_int32), ("nlopt_result", C.c_int32), ("evals", C.c_int32), ("textures", C.c_int32),
                ("last_val", C.c_double)]


PATCH_DTYPE = np.dtype([("center", "<f4", 4), ("normal", "<f4", 4), ("scale", "<f4"), ("nimages", "<i4"),
                        ("images", "<i4", MAX_VIEWS), ("color", "<f4", 3), ("ncc", "<f4"), ("status", "<i4"),
                        ("nlopt_result", "<i4"), ("evals", "<i4"), ("textures", "<i4"), ("last_val", "<f8")],
                       align=True)
assert PATCH_DTYPE.itemsize == C.sizeof(Patch), (PATCH_DTYPE.itemsize, C.sizeof(Patch))


def build(force: bool = False) -> str:
    """Compile the oracle (and, when /root/reference is present, oracle/_ref from the reference's sources)."""
    ref_a = os.path.join(_HERE, "_ref", "libnlopt_ref.a")
    if os.path.isdir("/root/reference/thirdLibs/nlopt-2.4.2"):
        subprocess.run(["make", "-C", _HERE, "ref", "-j8"], check=True, capture_output=True)
    if not os.path.exists(ref_a):
        raise RuntimeError("oracle/_ref/libnlopt_ref.a missing and /root/reference not present to build it")
    subprocess.run(["make", "-C", _HERE] + (["-B", "libhpmvs_oracle.so"] if force else []), check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        vp, ip, fp, dp, u8p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_uint8)
        L.orc_scene_new.restype = vp; L.orc_scene_new.argtypes = [C.POINTER(Options)]
        L.orc_scene_free.argtypes = [vp]
        L.orc_add_camera.argtypes = [vp, C.c_double, dp, dp, C.c_int, C.c_int, u8p]
        L.orc_num_cameras.argtypes = [vp]
        L.orc_get_camera.argtypes = [vp, C.c_int, C.POINTER(Camera)]
        L.orc_get_image.restype = C.POINTER(C.c_uint8); L.orc_get_image.argtypes = [vp, C.c_int, C.c_int, ip, ip]
        L.orc_extract_covis.argtypes = [vp, C.c_int, ip, ip]
        L.orc_set_covis.argtypes = [vp, ip, ip]
        L.orc_get_covis.argtypes = [vp, C.c_int, ip, C.c_int]
        L.orc_seed_patches.argtypes = [vp, C.c_int, dp, ip, ip, C.c_void_p, u8p]
        L.orc_optimize.argtypes = [vp, C.c_void_p]
        L.orc_optimize_batch.argtypes = [vp, C.c_int, C.c_void_p, C.c_int]
        L.orc_set_inccs.argtypes = [vp, C.c_void_p, C.c_int, C.c_int, fp]
        L.orc_sample_texture.argtypes = [vp, fp, C.c_float, fp, fp, fp, C.c_int, fp]
        L.orc_objective.restype = C.c_double; L.orc_objective.argtypes = [vp, C.c_void_p, dp]
        L.orc_patch_color.argtypes = [vp, C.c_void_p, fp]
        L.orc_set_cr_asinf.argtypes = [C.c_int]
        L.orc_depth_reset.argtypes = [vp]
        L.orc_depth_set_batch.argtypes = [vp, C.c_int, C.c_void_p]
        L.orc_depth_unset_batch.argtypes = [vp, C.c_int, C.c_void_p]
        L.orc_get_depth.restype = C.POINTER(C.c_float); L.orc_get_depth.argtypes = [vp, C.c_int, C.c_int, ip, ip]
        L.orc_accept_batch.argtypes = [vp, C.c_int, C.c_void_p, C.c_float, ip]
        L.orc_expand_candidates.argtypes = [vp, C.c_int, C.c_void_p, fp, C.c_int, C.c_void_p]
        L.orc_testfunc_eval.restype = C.c_double; L.orc_testfunc_eval.argtypes = [C.c_int, dp]
        L.orc_bobyqa_testfunc.argtypes = [C.c_int, dp, dp, dp, C.c_double, C.c_int, dp, dp, dp, dp, C.c_int, ip]
        _lib = L
    return _lib


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


class OracleScene:
    """Scene as the oracle sees it: NVM cameras + level-0 u8 RGB images -> pyramids, cameras, covisibility."""

    def __init__(self, options: Optional[Options] = None):
        self.options = options or Options.defaults()
        self._h = lib().orc_scene_new(C.byref(self.options))

    def __del__(self):
        try:
            if self._h:
                lib().orc_scene_free(self._h)
                self._h = None
        except Exception:
            pass

    @classmethod
    def from_synth(cls, scene, options: Optional[Options] = None) -> "OracleScene":
        s = cls(options)
        for cam, img in zip(scene.cameras, scene.images):
            s.add_camera(cam.f, cam.q, cam.c, img)
        s.extract_covis(scene.meas_offsets, scene.meas_cam)
        return s

    def add_camera(self, f: float, q: Sequence[float], c: Sequence[float], rgb: np.ndarray) -> int:
        rgb = np.ascontiguousarray(rgb, np.uint8)
        h, w = rgb.shape[:2]
        qa = np.asarray(q, np.float64); ca = np.asarray(c, np.float64)
        return lib().orc_add_camera(self._h, float(f), _p(qa, C.c_double), _p(ca, C.c_double), w, h, _p(rgb, C.c_uint8))

    @property
    def n_cameras(self) -> int:
        return lib().orc_num_cameras(self._h)

    def camera(self, i: int) -> Camera:
        c = Camera()
        lib().orc_get_camera(self._h, i, C.byref(c))
        return c

    def image(self, cam: int, level: int) -> np.ndarray:
        w, h = C.c_int32(), C.c_int32()
        ptr = lib().orc_get_image(self._h, cam, level, C.byref(w), C.byref(h))
        return np.ctypeslib.as_array(ptr, shape=(h.value, w.value, 3)).copy()

    def extract_covis(self, meas_offsets: np.ndarray, meas_cam: np.ndarray) -> None:
        mo = np.ascontiguousarray(meas_offsets, np.int32); mc = np.ascontiguousarray(meas_cam, np.int32)
        lib().orc_extract_covis(self._h, len(mo) - 1, _p(mo, C.c_int32), _p(mc, C.c_int32))

    def set_covis(self, lists) -> None:
        offs = np.zeros(len(lists) + 1, np.int32)
        offs[1:] = np.cumsum([len(l) for l in lists])
        ids = np.asarray([v for l in lists for v in l] + [0], np.int32)
        lib().orc_set_covis(self._h, _p(offs, C.c_int32), _p(ids, C.c_int32))

    def covis(self):
        out = []
        buf = np.zeros(4096, np.int32)
        for i in range(self.n_cameras):
            n = lib().orc_get_covis(self._h, i, _p(buf, C.c_int32), len(buf))
            out.append(buf[:n].tolist())
        return out

    def seed_patches(self, xyz: np.ndarray, meas_offsets: np.ndarray, meas_cam: np.ndarray):
        xyz = np.ascontiguousarray(xyz, np.float64)
        mo = np.ascontiguousarray(meas_offsets, np.int32); mc = np.ascontiguousarray(meas_cam, np.int32)
        n = xyz.shape[0]
        out = np.zeros(n, PATCH_DTYPE)
        valid = np.zeros(n, np.uint8)
        lib().orc_seed_patches(self._h, n, _p(xyz, C.c_double), _p(mo, C.c_int32), _p(mc, C.c_int32),
                               out.ctypes.data, _p(valid, C.c_uint8))
        return out, valid.astype(bool)

    def optimize_batch(self, patches: np.ndarray, nthreads: int = 1) -> np.ndarray:
        p = np.ascontiguousarray(patches.copy())
        assert p.dtype == PATCH_DTYPE
        lib().orc_optimize_batch(self._h, len(p), p.ctypes.data, int(nthreads))
        return p

    def set_inccs(self, patch: np.ndarray, ref_idx: int = 0, robust: int = 0) -> np.ndarray:
        p = np.ascontiguousarray(patch.reshape(1).copy())
        out = np.zeros(int(p["nimages"][0]), np.float32)
        lib().orc_set_inccs(self._h, p.ctypes.data, ref_idx, robust, _p(out, C.c_float))
        return out

    def sample_texture(self, center, scale, xaxis, yaxis, zaxis, cam: int):
        c = np.asarray(center, np.float32); x = np.asarray(xaxis, np.float32)
        y = np.asarray(yaxis, np.float32); z = np.asarray(zaxis, np.float32)
        out = np.zeros(147, np.float32)
        ok = lib().orc_sample_texture(self._h, _p(c, C.c_float), float(scale), _p(x, C.c_float), _p(y, C.c_float),
                                      _p(z, C.c_float), cam, _p(out, C.c_float))
        return bool(ok), out

    def objective(self, patch: np.ndarray, x: Sequence[float]) -> float:
        p = np.ascontiguousarray(patch.reshape(1).copy())
        xa = np.asarray(x, np.float64)
        return lib().orc_objective(self._h, p.ctypes.data, _p(xa, C.c_double))

    # -- "next" rows ------------------------------------------------------------------------
    def depth_reset(self) -> None:
        lib().orc_depth_reset(self._h)

    def depth_set(self, patches: np.ndarray) -> None:
        p = np.ascontiguousarray(patches)
        lib().orc_depth_set_batch(self._h, len(p), p.ctypes.data)

    def depth_unset(self, patches: np.ndarray) -> None:
        """Scene::setDepths(patch, true) (Scene.cpp:351-381) for the records with status OK."""
        p = np.ascontiguousarray(patches)
        lib().orc_depth_unset_batch(self._h, len(p), p.ctypes.data)

    def depth(self, cam: int, level: int) -> np.ndarray:
        r, c = C.c_int32(), C.c_int32()
        ptr = lib().orc_get_depth(self._h, cam, level, C.byref(r), C.byref(c))
        return np.ctypeslib.as_array(ptr, shape=(r.value, c.value)).copy()

    def accept(self, patches: np.ndarray, margin: float = 1.0) -> np.ndarray:
        p = np.ascontiguousarray(patches)
        out = np.zeros((len(p), 3), np.int32)
        lib().orc_accept_batch(self._h, len(p), p.ctypes.data, float(margin), _p(out, C.c_int32))
        return out

    def expand_candidates(self, parents: np.ndarray, widths: np.ndarray, mode: int) -> np.ndarray:
        p = np.ascontiguousarray(parents); w = np.ascontiguousarray(widths, np.float32)
        out = np.zeros(len(p) * mode, PATCH_DTYPE)
        lib().orc_expand_candidates(self._h, len(p), p.ctypes.data, _p(w, C.c_float), mode, out.ctypes.data)
        return out

    def patch_color(self, patch: np.ndarray) -> np.ndarray:
        p = np.ascontiguousarray(patch.reshape(1).copy())
        out = np.zeros(3, np.float32)
        lib().orc_patch_color(self._h, p.ctypes.data, _p(out, C.c_float))
        return out


def set_cr_asinf(on: bool) -> None:
    """See g_cr_asinf in hpmvs_oracle.cpp: libm-independent evaluation of the one asinf on the path."""
    lib().orc_set_cr_asinf(1 if on else 0)


def testfunc(func_id: int, x: Sequence[float]) -> float:
    xa = np.asarray(x, np.float64)
    return lib().orc_testfunc_eval(func_id, _p(xa, C.c_double))


def bobyqa_testfunc(func_id: int, x0, lb, ub, xtol_rel: float = 1e-7, maxeval: int = 1000):
    """Real nlopt BOBYQA; returns (result, x, f, trace_x[n,3], trace_f[n])."""
    x0 = np.asarray(x0, np.float64); lb = np.asarray(lb, np.float64); ub = np.asarray(ub, np.float64)
    xout = np.zeros(3); fout = C.c_double(); nev = C.c_int32()
    cap = max(16, maxeval + 8)
    tx = np.zeros((cap, 3)); tf = np.zeros(cap)
    r = lib().orc_bobyqa_testfunc(func_id, _p(x0, C.c_double), _p(lb, C.c_double), _p(ub, C.c_double), xtol_rel,
                                  maxeval, _p(xout, C.c_double), C.byref(fout), _p(tx, C.c_double), _p(tf, C.c_double),
                                  cap, C.byref(nev))
    n = min(nev.value, cap)
    return r, xout, fout.value, tx[:n].copy(), tf[:n].copy()
