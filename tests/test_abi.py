"""The C-ABI library: loads, exports every symbol include/hpmvs_b200.h declares, host-side surface agrees with the
oracle bit for bit, and - without a GPU - refuses to run instead of falling back (no compute calls here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import hpmvs_b200 as hp
import oracle
from hpmvs_b200 import _native
from helpers import small_plane

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "hpmvs_b200.h")).read()
    names = set(re.findall(r"\b(hpmvs_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 18
    lib = C.CDLL(_native.build())
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/hpmvs_b200.h but not exported"
    assert lib.hpmvs_abi_version() == 1


def test_record_layouts():
    assert hp.PATCH_DTYPE.itemsize == 208
    assert C.sizeof(hp.Camera) == 6 * 12 * 4 + 16 + 36 + 8 + 48
    assert C.sizeof(hp.Options) == 36


def test_no_silent_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(hp.HpmvsError, match="no CUDA device"):
        hp.Engine()


def test_camera_covis_seeds_match_oracle_bit_for_bit():
    sc, orc, _ = small_plane()
    cams = [hp.camera_from_nvm(c.f, c.q, c.c, img.shape[1], img.shape[0]) for c, img in zip(sc.cameras, sc.images)]
    for i, c in enumerate(cams):
        assert bytes(orc.camera(i)) == bytes(c)
    assert hp.extract_covis(len(cams), sc.meas_offsets, sc.meas_cam) == orc.covis()
    s1, v1 = orc.seed_patches(sc.points, sc.meas_offsets, sc.meas_cam)
    s2, v2 = hp.seed_patches(hp.Options.defaults(), cams, sc.points, sc.meas_offsets, sc.meas_cam)
    assert np.array_equal(v1, v2)
    for f in ("center", "normal", "scale", "nimages"):
        assert np.array_equal(s1[f], s2[f]), f
    assert np.array_equal(s1["images"][:, :hp.MAX_VIEWS], s2["images"])


def test_seed_edge_cases():
    sc, orc, _ = small_plane()
    cams = [hp.camera_from_nvm(c.f, c.q, c.c, img.shape[1], img.shape[0]) for c, img in zip(sc.cameras, sc.images)]
    # empty input, a point behind all cameras, a point with too few measurements
    out, valid = hp.seed_patches(hp.Options.defaults(), cams, np.zeros((0, 3)), np.zeros(1, np.int32), np.zeros(1, np.int32))
    assert len(out) == 0 and len(valid) == 0
    xyz = np.array([[0.0, 0.0, -50.0], [0.0, 0.0, 0.0]])
    offs = np.array([0, 3, 5], np.int32); mc = np.array([0, 1, 2, 0, 1], np.int32)
    o2, v2 = hp.seed_patches(hp.Options.defaults(), cams, xyz, offs, mc)
    o1, v1 = orc.seed_patches(xyz, offs, mc)
    assert not v2[0] and not v2[1] and np.array_equal(v1, v2)


def test_nvm_round_trip(tmp_path):
    sc = hp.synth.plane_scene(n_views=3, width=64, height=48, focal=60.0, n_seeds=9, seed=3, tex_size=64)
    path = str(tmp_path / "scene.nvm")
    hp.synth.write_nvm(sc, path)
    txt = open(path).read().split()
    assert txt[0] == "NVM_V3" and int(txt[1]) == 3
    assert os.path.getsize(str(tmp_path / sc.cameras[0].filename)) == len(b"P6\n64 48\n255\n") + 64 * 48 * 3


def test_host_entry_points_reject_bad_arguments():
    # no GPU needed: the host-side entry points validate before they touch a device
    lib = C.CDLL(_native.build())
    lib.hpmvs_dedup_border.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    lib.hpmvs_undistort_rgb.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    lib.hpmvs_pipeline_run.argtypes = [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 4
    assert lib.hpmvs_dedup_border(-1, None, None, None, 1.0, None) < 0
    assert lib.hpmvs_dedup_border(3, None, None, None, 1.0, None) < 0
    assert lib.hpmvs_dedup_border(0, None, None, None, 0.0, None) < 0          # cell edge must be positive
    assert lib.hpmvs_dedup_border(0, None, None, None, 1.0, None) == 0         # empty input is fine
    assert lib.hpmvs_undistort_rgb(None, 4, 4, 100.0, 0.1, None, None) < 0
    img = np.zeros((4, 4, 3), np.uint8)
    assert lib.hpmvs_undistort_rgb(img.ctypes.data, 0, 4, 100.0, 0.1, img.ctypes.data, None) < 0
    assert lib.hpmvs_pipeline_run(None, None, 0, None, None, None, None) < 0
    lib.hpmvs_free(None)                                                 # free(NULL) is a no-op


def test_undistort_identity_and_mask():
    from hpmvs_b200 import io as hio
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (40, 60, 3), dtype=np.uint8)
    same, written = hio.undistort(img, 80.0, 0.0, return_mask=True)
    assert np.array_equal(same, img) and written.all()
    und, written = hio.undistort(img, 80.0, 0.05, return_mask=True)
    assert und.shape == img.shape and 0.5 < written.mean() < 1.0 and (und[~written] == 0).all()
    # a constant image stays constant wherever it is written (bilinear weights sum to one; f32 -> u8 truncation of an exact value)
    flat = np.full((40, 60, 3), 77, np.uint8)
    und, written = hio.undistort(flat, 80.0, -0.04, return_mask=True)
    assert (und[written] == 77).all()
