"""ctypes front end of the host-side file formats (NVM_V3 reader, PPM reader, ext-PLY writer) in libhpmvs_b200.so."""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Tuple

import numpy as np

from . import _native
from .engine import PATCH_DTYPE, _check
from .synth import NVMCamera, SynthScene

_done = False


def _lib():
    global _done
    L = _native.lib()
    if not _done:
        vp, ip, dp, u8p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_uint8)
        L.hpmvs_nvm_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
        L.hpmvs_nvm_close.argtypes = [vp]; L.hpmvs_nvm_close.restype = None
        for f in ("hpmvs_nvm_num_models", "hpmvs_nvm_num_cameras", "hpmvs_nvm_num_points", "hpmvs_nvm_num_measurements"):
            getattr(L, f).argtypes = [vp]
        L.hpmvs_nvm_camera.argtypes = [vp, C.c_int, C.c_char_p, C.c_int, dp, dp, dp, dp]
        L.hpmvs_nvm_points.argtypes = [vp, dp, dp, ip, ip, ip, dp]
        L.hpmvs_ppm_read.argtypes = [C.c_char_p, ip, ip, u8p]
        L.hpmvs_ply_write_ext.argtypes = [C.c_char_p, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int]
        _done = True
    return L


def read_ppm(path: str) -> np.ndarray:
    w, h = C.c_int32(), C.c_int32()
    _check(_lib().hpmvs_ppm_read(path.encode(), C.byref(w), C.byref(h), None))
    out = np.zeros((h.value, w.value, 3), np.uint8)
    _check(_lib().hpmvs_ppm_read(path.encode(), C.byref(w), C.byref(h), out.ctypes.data_as(C.POINTER(C.c_uint8))))
    return out


def read_nvm(path: str, fix_path: bool = True, load_images: bool = True) -> SynthScene:
    """NVMReader::readFile + (optionally) the level-0 images -> the same scene container the synthetic generator fills."""
    h = C.c_void_p()
    _check(_lib().hpmvs_nvm_open(path.encode(), 1 if fix_path else 0, C.byref(h)))
    try:
        L = _lib()
        nc, npnt, nm = L.hpmvs_nvm_num_cameras(h), L.hpmvs_nvm_num_points(h), L.hpmvs_nvm_num_measurements(h)
        cams: List[NVMCamera] = []
        buf = C.create_string_buffer(4096)
        for i in range(nc):
            f, r = C.c_double(), C.c_double()
            q = np.zeros(4); c = np.zeros(3)
            _check(L.hpmvs_nvm_camera(h, i, buf, 4096, C.byref(f), q.ctypes.data_as(C.POINTER(C.c_double)),
                                      c.ctypes.data_as(C.POINTER(C.c_double)), C.byref(r)))
            cams.append(NVMCamera(buf.value.decode(), f.value, q, c, r.value))
        xyz = np.zeros((npnt, 3)); offs = np.zeros(npnt + 1, np.int32); mc = np.zeros(max(nm, 1), np.int32)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        _check(L.hpmvs_nvm_points(h, xyz.ctypes.data_as(dp), None, offs.ctypes.data_as(ip), mc.ctypes.data_as(ip), None, None))
    finally:
        _lib().hpmvs_nvm_close(h)
    images = []
    if load_images:
        for cam in cams:
            img = read_ppm(cam.filename)
            cam.width, cam.height = img.shape[1], img.shape[0]
            images.append(img)
    return SynthScene(os.path.basename(path), cams, images, xyz, offs, mc[:nm])


def rewrite_nvm(src: str, dst: str, fix_path: bool = True) -> None:
    """Read an NVM_V3 file and write it back the way NVMReader::saveNVM does (NVMReader.cpp:157-183)."""
    L = _lib()
    L.hpmvs_nvm_write.argtypes = [C.c_void_p, C.c_char_p]
    h = C.c_void_p()
    _check(L.hpmvs_nvm_open(src.encode(), 1 if fix_path else 0, C.byref(h)))
    try:
        _check(L.hpmvs_nvm_write(h, dst.encode()))
    finally:
        L.hpmvs_nvm_close(h)


def undistort(rgb: np.ndarray, f: float, r: float, return_mask: bool = False):
    """Image::undistort (Image.cpp:68-149) on an [h, w, 3] u8 image; r == 0 returns a copy.  return_mask: also the [h, w] bool mask of
    the pixels that were written (the others are 0 here and uninitialised memory in the reference)."""
    img = np.ascontiguousarray(rgb, np.uint8)
    out = np.empty_like(img)
    mask = np.zeros(img.shape[:2], np.uint8)
    L = _lib()
    L.hpmvs_undistort_rgb.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    _check(L.hpmvs_undistort_rgb(img.ctypes.data, img.shape[1], img.shape[0], float(f), float(r), out.ctypes.data, mask.ctypes.data))
    return (out, mask.astype(bool)) if return_mask else out


def write_ext_ply(path: str, patches: np.ndarray, binary: bool = False, normal: bool = True, scale: bool = True,
                  visibility: bool = True) -> None:
    """DynOctTree::toExtPly for a flat list of patch records (doctree.h:525-622)."""
    p = np.ascontiguousarray(patches)
    assert p.dtype == PATCH_DTYPE
    _check(_lib().hpmvs_ply_write_ext(path.encode(), len(p), p.ctypes.data, int(binary), int(normal), int(scale), int(visibility)))
