"""Python front end of the C ABI in include/hpmvs_b200.h (ctypes; plumbing only - all work is in the .so).

`Engine` mirrors the reference's optimizer object: constructed from options + scene data
(PatchOptimizer::PatchOptimizer(const HpmvsOptions&, const Scene*), PatchOptimizer.cpp:38-45) and then asked to
optimize patches (PatchOptimizer::optimize(Patch3d&), :78-103) - here a whole batch per call.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _native

LEVELS = 6
MAX_VIEWS = 32

STATUS_NAMES = ["OK", "FAIL_ADD_IMAGES", "FAIL_NCC1", "FAIL_ANGLES", "FAIL_OPT_MINIMAGES", "FAIL_OPT_ROUNDOFF",
                "FAIL_OPT_MAXEVAL", "FAIL_OPT_OTHER", "FAIL_ADD_IMAGES2", "FAIL_NCC2", "FAIL_ANGLE_FILTER",
                "FAIL_ANGLES2", "FAIL_NCC3", "FAIL_TOO_MANY_VIEWS"]


class Options(C.Structure):
    """hpmvs_options_t == the fields of mo3d::HpmvsOptions the path reads (HpmvsOptions.h:29-58)."""
    _fields_ = [("maxlevel", C.c_int32), ("minlevel", C.c_int32), ("start_level", C.c_int32),
                ("max_angle", C.c_float), ("min_angle", C.c_float), ("max_images_per_patch", C.c_int32),
                ("min_images_per_patch", C.c_int32), ("ncc_alpha_1", C.c_float), ("ncc_alpha_2", C.c_float)]

    @staticmethod
    def defaults(**kw) -> "Options":
        o = Options(5, 0, 4, float(np.float32(60.0 * np.pi / 180.0)), float(np.float32(10.0 * np.pi / 180.0)),
                    6, 3, 0.4, 0.5)
        for k, v in kw.items():
            setattr(o, k, v)
        return o


class Camera(C.Structure):
    _fields_ = [("P", C.c_float * 4 * 3 * LEVELS), ("center", C.c_float * 4), ("xaxis", C.c_float * 3),
                ("yaxis", C.c_float * 3), ("zaxis", C.c_float * 3), ("k00", C.c_float), ("k11", C.c_float),
                ("width", C.c_int32 * LEVELS), ("height", C.c_int32 * LEVELS)]


class Counters(C.Structure):
    _fields_ = [("patches", C.c_uint64), ("patches_ok", C.c_uint64), ("evals", C.c_uint64),
                ("textures", C.c_uint64), ("kernel_launches", C.c_uint64)]


PATCH_DTYPE = np.dtype([("center", "<f4", 4), ("normal", "<f4", 4), ("scale", "<f4"), ("nimages", "<i4"),
                        ("images", "<i4", MAX_VIEWS), ("color", "<f4", 3), ("ncc", "<f4"), ("status", "<i4"),
                        ("nlopt_result", "<i4"), ("evals", "<i4"), ("textures", "<i4"), ("score", "<f8")], align=True)
assert PATCH_DTYPE.itemsize == 208

_sigs_done = False


def _lib():
    global _sigs_done
    L = _native.lib()
    if not _sigs_done:
        vp, ip, fp, dp, u8p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_uint8)
        L.hpmvs_engine_create.argtypes = [C.POINTER(Options), C.c_int, C.POINTER(vp)]
        L.hpmvs_engine_destroy.argtypes = [vp]; L.hpmvs_engine_destroy.restype = None
        L.hpmvs_engine_set_cameras.argtypes = [vp, C.c_int, C.POINTER(Camera)]
        L.hpmvs_engine_upload_image.argtypes = [vp, C.c_int, C.c_int, u8p, C.c_int, C.c_int, C.c_size_t]
        L.hpmvs_engine_upload_image_undistort.argtypes = [vp, C.c_int, u8p, C.c_int, C.c_int, C.c_size_t, C.c_double, C.c_double]
        L.hpmvs_engine_build_pyramid.argtypes = [vp, C.c_int]
        L.hpmvs_engine_download_image.argtypes = [vp, C.c_int, C.c_int, u8p, C.c_size_t]
        L.hpmvs_engine_set_covis.argtypes = [vp, ip, ip]
        L.hpmvs_optimize_batch.argtypes = [vp, C.c_int, vp, vp, vp]
        L.hpmvs_optimize_batch_device.argtypes = [vp, C.c_int, vp, vp, vp]
        L.hpmvs_start_parameters.argtypes = [vp, C.c_int, vp, dp]
        L.hpmvs_optimize_batch_device_start.argtypes = [vp, C.c_int, vp, vp, vp, vp]
        L.hpmvs_ncc_batch.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, fp, vp]
        L.hpmvs_ncc_batch_device.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, vp, vp]
        L.hpmvs_engine_set_start_mode.argtypes = [vp, C.c_int]
        L.hpmvs_optimize_batch_submit.argtypes = [vp, C.c_int, vp, vp, vp]
        L.hpmvs_engine_counters.argtypes = [vp, C.POINTER(Counters), C.c_int]
        L.hpmvs_engine_stream.argtypes = [vp]; L.hpmvs_engine_stream.restype = vp
        L.hpmvs_engine_last_kernel_ms.argtypes = [vp]; L.hpmvs_engine_last_kernel_ms.restype = C.c_float
        L.hpmvs_error_string.argtypes = [C.c_int]; L.hpmvs_error_string.restype = C.c_char_p
        L.hpmvs_camera_from_nvm.argtypes = [C.c_double, dp, dp, C.c_int, C.c_int, C.c_int, C.POINTER(Camera)]
        L.hpmvs_extract_covis.argtypes = [C.c_int, C.c_int, ip, ip, C.c_int, ip, ip, C.c_int]
        L.hpmvs_engine_depth_reset.argtypes = [vp]
        L.hpmvs_depth_set_batch.argtypes = [vp, C.c_int, vp, vp]
        L.hpmvs_depth_unset_batch.argtypes = [vp, C.c_int, vp, vp]
        L.hpmvs_accept_batch.argtypes = [vp, C.c_int, vp, C.c_float, ip, vp]
        L.hpmvs_expand_candidates_device.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, vp]
        L.hpmvs_depth_set_batch_device.argtypes = [vp, C.c_int, vp, C.c_int, vp]
        L.hpmvs_accept_batch_device.argtypes = [vp, C.c_int, vp, C.c_float, vp, vp]
        L.hpmvs_dedup_border_device.argtypes = [vp, C.c_int, vp, vp, dp, C.c_double, vp, vp, vp]
        L.hpmvs_engine_download_depth.argtypes = [vp, C.c_int, C.c_int, fp, ip, ip]
        L.hpmvs_expand_candidates.argtypes = [C.c_int, C.POINTER(Camera), C.c_int, vp, fp, C.c_int, vp]
        L.hpmvs_seed_patches.argtypes = [C.POINTER(Options), C.c_int, C.POINTER(Camera), C.c_int, dp, ip, ip, vp, u8p]
        _sigs_done = True
    return L


class HpmvsError(RuntimeError):
    pass


def _check(rc: int) -> None:
    if rc < 0:
        raise HpmvsError(f"hpmvs_b200 error {rc}: {_lib().hpmvs_error_string(rc).decode()}")


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


# ---------------------------------------------------------------------------------------------- host scene surface
def camera_from_nvm(f: float, q: Sequence[float], c: Sequence[float], width: int, height: int, maxlevel: int = 5) -> Camera:
    """mo3d::Camera::init for one NVM camera (Camera.cpp:34-81)."""
    cam = Camera()
    qa = np.asarray(q, np.float64); ca = np.asarray(c, np.float64)
    _check(_lib().hpmvs_camera_from_nvm(float(f), _p(qa, C.c_double), _p(ca, C.c_double), int(width), int(height), maxlevel, C.byref(cam)))
    return cam


def extract_covis(ncams: int, meas_offsets: np.ndarray, meas_cam: np.ndarray, compat: bool = True) -> List[List[int]]:
    """Scene::extractCoVisiblilty (Scene.cpp:241-298); compat=True keeps the reference's index quirk."""
    mo = np.ascontiguousarray(meas_offsets, np.int32); mc = np.ascontiguousarray(meas_cam, np.int32)
    offs = np.zeros(ncams + 1, np.int32)
    ids = np.zeros(max(1, ncams * ncams), np.int32)
    n = _lib().hpmvs_extract_covis(ncams, len(mo) - 1, _p(mo, C.c_int32), _p(mc, C.c_int32), 1 if compat else 0,
                                   _p(offs, C.c_int32), _p(ids, C.c_int32), len(ids))
    _check(n)
    return [ids[offs[i]:offs[i + 1]].tolist() for i in range(ncams)]


def seed_patches(options: Options, cameras: Sequence[Camera], xyz: np.ndarray, meas_offsets: np.ndarray, meas_cam: np.ndarray):
    """Candidate construction of Scene::initPatches (Scene.cpp:116-165). Returns (patches, valid)."""
    cams = (Camera * len(cameras))(*cameras)
    xyz = np.ascontiguousarray(xyz, np.float64)
    mo = np.ascontiguousarray(meas_offsets, np.int32); mc = np.ascontiguousarray(meas_cam, np.int32)
    n = xyz.shape[0]
    out = np.zeros(n, PATCH_DTYPE)
    valid = np.zeros(n, np.uint8)
    _check(_lib().hpmvs_seed_patches(C.byref(options), len(cameras), cams, n, _p(xyz, C.c_double), _p(mo, C.c_int32),
                                     _p(mc, C.c_int32), out.ctypes.data, _p(valid, C.c_uint8)))
    return out, valid.astype(bool)


def expand_candidates(cameras: Sequence[Camera], parents: np.ndarray, widths: np.ndarray, mode: int) -> np.ndarray:
    """Candidates of CellProcessor::extend (mode 6) / ::branch (mode 4) (CellProcessor.cpp:98-119, 227-249)."""
    cams = (Camera * len(cameras))(*cameras)
    p = np.ascontiguousarray(parents); w = np.ascontiguousarray(widths, np.float32)
    assert p.dtype == PATCH_DTYPE
    out = np.zeros(len(p) * mode, PATCH_DTYPE)
    _check(_lib().hpmvs_expand_candidates(len(cameras), cams, len(p), p.ctypes.data, _p(w, C.c_float), mode, out.ctypes.data))
    return out


# ---------------------------------------------------------------------------------------------- the engine
class Engine:
    """One engine per GPU.  Holds the HBM-resident scene; `optimize` runs the fused kernel on a batch."""

    def __init__(self, options: Optional[Options] = None, device: int = 0):
        self.options = options or Options.defaults()
        self.device = device
        h = C.c_void_p()
        _check(_lib().hpmvs_engine_create(C.byref(self.options), device, C.byref(h)))
        self._h = h
        self.cameras: List[Camera] = []

    def close(self) -> None:
        if getattr(self, "_h", None):
            _lib().hpmvs_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- scene ------------------------------------------------------------------------------
    def set_cameras(self, cameras: Sequence[Camera]) -> None:
        self.cameras = list(cameras)
        arr = (Camera * len(cameras))(*cameras)
        _check(_lib().hpmvs_engine_set_cameras(self._h, len(cameras), arr))

    def upload_image(self, cam: int, level: int, rgb: np.ndarray) -> None:
        rgb = np.ascontiguousarray(rgb, np.uint8)
        h, w = rgb.shape[:2]
        _check(_lib().hpmvs_engine_upload_image(self._h, cam, level, _p(rgb, C.c_uint8), w, h, 3 * w))

    def upload_image_undistort(self, cam: int, rgb: np.ndarray, f: float, r: float) -> None:
        """Image::load's undistortion (Image.cpp:51-53, :68-149) on the GPU: distorted level-0 image in, undistorted level 0 resident."""
        rgb = np.ascontiguousarray(rgb, np.uint8)
        h, w = rgb.shape[:2]
        _check(_lib().hpmvs_engine_upload_image_undistort(self._h, cam, _p(rgb, C.c_uint8), w, h, 3 * w, float(f), float(r)))

    def build_pyramid(self, cam: int) -> None:
        _check(_lib().hpmvs_engine_build_pyramid(self._h, cam))

    def download_image(self, cam: int, level: int) -> np.ndarray:
        w, h = self.cameras[cam].width[level], self.cameras[cam].height[level]
        out = np.zeros((h, w, 3), np.uint8)
        _check(_lib().hpmvs_engine_download_image(self._h, cam, level, _p(out, C.c_uint8), 3 * w))
        return out

    def set_covis(self, lists: Sequence[Sequence[int]]) -> None:
        offs = np.zeros(len(lists) + 1, np.int32)
        offs[1:] = np.cumsum([len(l) for l in lists])
        ids = np.asarray([v for l in lists for v in l] + [0], np.int32)
        _check(_lib().hpmvs_engine_set_covis(self._h, _p(offs, C.c_int32), _p(ids, C.c_int32)))

    @classmethod
    def from_synth(cls, scene, options: Optional[Options] = None, device: int = 0, compat_covis: bool = True,
                   undistort: str = "host") -> "Engine":
        """NVM cameras + level-0 images of a hpmvs_b200.synth.SynthScene -> resident scene (pyramids built on the GPU)."""
        e = cls(options, device)
        ml = e.options.maxlevel
        cams = [camera_from_nvm(c.f, c.q, c.c, img.shape[1], img.shape[0], ml) for c, img in zip(scene.cameras, scene.images)]
        e.set_cameras(cams)
        for i, img in enumerate(scene.images):
            r = getattr(scene.cameras[i], "r", 0.0)
            if r != 0.0 and undistort == "gpu":                      # Image::load undistorts level 0 first (Image.cpp:51-53)
                e.upload_image_undistort(i, img, scene.cameras[i].f, r)
            else:
                if r != 0.0:
                    from . import io as _io
                    img = _io.undistort(img, scene.cameras[i].f, r)    # host twin, same libm as the reference: bit-exact
                e.upload_image(i, 0, img)
            e.build_pyramid(i)
        e.set_covis(extract_covis(len(cams), scene.meas_offsets, scene.meas_cam, compat_covis))
        return e

    @classmethod
    def from_stream(cls, scene_fn, options: Optional[Options] = None, device: int = 0, compat_covis: bool = True):
        """Resident scene from a generator that streams its views: scene_fn(image_sink) must call image_sink(i, nvm_cameras, image)
        once per view (in order) and return the SynthScene (without images).  Each view is uploaded and its pyramid built as it
        arrives, so the host never holds more than one level-0 image.  Returns (engine, scene)."""
        e = cls(options, device)
        ml = e.options.maxlevel

        def sink(i, nvm_cams, img):
            if i == 0:
                e.set_cameras([camera_from_nvm(c.f, c.q, c.c, c.width, c.height, ml) for c in nvm_cams])
            e.upload_image(i, 0, img)
            e.build_pyramid(i)
        scene = scene_fn(sink)
        e.set_covis(extract_covis(len(scene.cameras), scene.meas_offsets, scene.meas_cam, compat_covis))
        return e, scene

    # -- the hot path -----------------------------------------------------------------------
    def optimize(self, patches: np.ndarray, out: Optional[np.ndarray] = None, stream: int = 0) -> np.ndarray:
        """n x PatchOptimizer::optimize(Patch3d&): host records in, host records out (H2D + kernel + D2H)."""
        assert patches.dtype == PATCH_DTYPE and patches.flags.c_contiguous
        if out is None:
            out = np.empty_like(patches)
        _check(_lib().hpmvs_optimize_batch(self._h, len(patches), patches.ctypes.data, out.ctypes.data, stream or None))
        return out

    def set_start_mode(self, host_libm: bool) -> None:
        """True: optimize() evaluates parametersFromCenterNorm's asin/cos/acos on the host with this machine's libm, like
        the reference (bit-identical to a reference built here); False (default): on the device, asin correctly rounded."""
        _check(_lib().hpmvs_engine_set_start_mode(self._h, 1 if host_libm else 0))

    def optimize_ptr(self, n: int, in_ptr: int, out_ptr: int, stream: int = 0) -> None:
        """Same, on raw host pointers (e.g. pinned torch tensors)."""
        _check(_lib().hpmvs_optimize_batch(self._h, n, in_ptr, out_ptr, stream or None))

    def optimize_submit(self, n: int, in_ptr: int, out_ptr: int, stream: int) -> None:
        """Asynchronous host-buffer call (pinned pointers): enqueues H2D + kernel + D2H on `stream` and returns."""
        _check(_lib().hpmvs_optimize_batch_submit(self._h, n, in_ptr, out_ptr, stream))

    def optimize_device(self, n: int, d_in: int, d_out: int, stream: int = 0) -> None:
        """Device-resident records, asynchronous on `stream` (0 = the engine's stream)."""
        _check(_lib().hpmvs_optimize_batch_device(self._h, n, d_in, d_out, stream or None))

    def start_parameters(self, patches: np.ndarray) -> np.ndarray:
        """parametersFromCenterNorm's two start angles per record with THIS machine's libm (start mode 1) -> [n, 2] f64."""
        assert patches.dtype == PATCH_DTYPE and patches.flags.c_contiguous
        out = np.zeros((len(patches), 2), np.float64)
        _check(_lib().hpmvs_start_parameters(self._h, len(patches), patches.ctypes.data, _p(out, C.c_double)))
        return out

    def optimize_device_start(self, n: int, d_in: int, d_out: int, d_start: int, stream: int = 0) -> None:
        """Device-resident records + device-resident start angles (start mode 1), asynchronous on `stream`."""
        _check(_lib().hpmvs_optimize_batch_device_start(self._h, n, d_in, d_out, d_start or None, stream or None))

    def ncc(self, patches: np.ndarray, ref_idx: int = 0, robust: bool = False) -> np.ndarray:
        """n x PatchOptimizer::setINCCs (PatchOptimizer.cpp:448-474) -> [n, MAX_VIEWS] f32."""
        assert patches.dtype == PATCH_DTYPE and patches.flags.c_contiguous
        out = np.zeros((len(patches), MAX_VIEWS), np.float32)
        _check(_lib().hpmvs_ncc_batch(self._h, len(patches), patches.ctypes.data, ref_idx, 1 if robust else 0, _p(out, C.c_float), None))
        return out

    def ncc_device(self, n: int, d_in: int, d_inccs: int, ref_idx: int = 0, robust: bool = False, stream: int = 0) -> None:
        """Same scoring on device-resident records / results ([n, MAX_VIEWS] f32), asynchronous on `stream`."""
        _check(_lib().hpmvs_ncc_batch_device(self._h, n, d_in, ref_idx, 1 if robust else 0, d_inccs, stream or None))

    # -- "next" rows: depth maps + acceptance tests --------------------------------------------
    def depth_reset(self) -> None:
        _check(_lib().hpmvs_engine_depth_reset(self._h))

    def depth_set(self, patches: np.ndarray) -> None:
        """n x Scene::setDepths(patch, false) for the records with status OK."""
        assert patches.dtype == PATCH_DTYPE and patches.flags.c_contiguous
        _check(_lib().hpmvs_depth_set_batch(self._h, len(patches), patches.ctypes.data, None))

    def depth_unset(self, patches: np.ndarray) -> None:
        """n x Scene::setDepths(patch, true): depth cells that still hold exactly these patches' depths go back to MAX_DEPTH."""
        assert patches.dtype == PATCH_DTYPE and patches.flags.c_contiguous
        _check(_lib().hpmvs_depth_unset_batch(self._h, len(patches), patches.ctypes.data, None))

    def accept(self, patches: np.ndarray, margin: float = 1.0) -> np.ndarray:
        """[n,3] = depthTests, viewBlockTest, pixelFreeTests per patch (Scene.cpp:518-644)."""
        assert patches.dtype == PATCH_DTYPE and patches.flags.c_contiguous
        out = np.zeros((len(patches), 3), np.int32)
        _check(_lib().hpmvs_accept_batch(self._h, len(patches), patches.ctypes.data, float(margin), _p(out, C.c_int32), None))
        return out

    def expand_candidates_device(self, n: int, d_parents: int, d_widths: int, mode: int, d_out: int, stream: int = 0) -> None:
        """Candidates of CellProcessor::extend (mode 6) / ::branch (mode 4) for n device-resident parents -> mode*n device records."""
        _check(_lib().hpmvs_expand_candidates_device(self._h, int(n), d_parents, d_widths, int(mode), d_out, stream or None))

    def depth_set_device(self, n: int, d_patches: int, subtract: bool = False, stream: int = 0) -> None:
        _check(_lib().hpmvs_depth_set_batch_device(self._h, int(n), d_patches, 1 if subtract else 0, stream or None))

    def accept_device(self, n: int, d_patches: int, margin: float, d_out: int, stream: int = 0) -> None:
        _check(_lib().hpmvs_accept_batch_device(self._h, int(n), d_patches, float(margin), d_out, stream or None))

    def dedup_border_device(self, n: int, d_records: int, d_owner: int, origin, cell: float, d_keep: int, d_nkeep: int = 0, stream: int = 0) -> None:
        """hpmvs_dedup_border_device: border de-duplication of n device-resident records (owner = int32 rank per record); d_keep receives
        n uint8 flags, d_nkeep (optional device int32) the number of survivors.  Asynchronous on `stream`."""
        org = np.ascontiguousarray(origin if origin is not None else (0.0, 0.0, 0.0), np.float64)
        _check(_lib().hpmvs_dedup_border_device(self._h, int(n), d_records, d_owner, _p(org, C.c_double), float(cell), d_keep, d_nkeep or None, stream or None))

    def download_depth(self, cam: int, level: int) -> np.ndarray:
        r, c = C.c_int32(), C.c_int32()
        _check(_lib().hpmvs_engine_download_depth(self._h, cam, level, None, C.byref(r), C.byref(c)))
        out = np.zeros((r.value, c.value), np.float32)
        _check(_lib().hpmvs_engine_download_depth(self._h, cam, level, _p(out, C.c_float), C.byref(r), C.byref(c)))
        return out

    def counters(self, reset: bool = False) -> Counters:
        c = Counters()
        _check(_lib().hpmvs_engine_counters(self._h, C.byref(c), 1 if reset else 0))
        return c

    @property
    def stream(self) -> int:
        return int(_lib().hpmvs_engine_stream(self._h) or 0)

    def last_kernel_ms(self) -> float:
        return float(_lib().hpmvs_engine_last_kernel_ms(self._h))
