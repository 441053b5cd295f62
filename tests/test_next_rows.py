"""Parity of the "next" rows around optimize(): candidate construction of CellProcessor::extend / branch
(host, CPU test), Scene::setDepths and the acceptance tests depthTests / viewBlockTest / pixelFreeTests (GPU tests)."""
import numpy as np
import pytest

import hpmvs_b200 as hp
import oracle
from helpers import small_plane, to_engine, to_oracle


def test_expand_candidates_match_oracle_bit_for_bit():
    sc, orc, seeds = small_plane()
    cams = [hp.camera_from_nvm(c.f, c.q, c.c, img.shape[1], img.shape[0]) for c, img in zip(sc.cameras, sc.images)]
    parents = seeds[:50]
    widths = np.linspace(0.05, 0.4, 50).astype(np.float32)
    for mode in (6, 4):
        ref = orc.expand_candidates(parents, widths, mode)
        got = hp.expand_candidates(cams, to_engine(parents), widths, mode)
        assert len(got) == mode * 50
        for f in ("center", "normal", "scale", "nimages"):
            assert np.array_equal(ref[f], got[f]), (mode, f)
        # geometry: candidates lie in the patch plane at the requested distance from the parent
        d = got["center"][:, :3] - np.repeat(to_engine(parents)["center"][:, :3], mode, 0)
        n = np.repeat(parents["normal"][:, :3], mode, 0)
        assert np.abs((d * n).sum(1)).max() < 1e-5
        r = np.linalg.norm(d, axis=1) / np.repeat(widths, mode)
        assert np.allclose(r, 1.0 if mode == 6 else 0.25, atol=1e-4)


@pytest.mark.gpu
def test_depth_maps_and_acceptance_tests_bit_exact():
    sc, orc, seeds = small_plane()
    eng = hp.Engine.from_synth(sc)
    oracle.set_cr_asinf(True)
    try:
        ref = orc.optimize_batch(seeds, nthreads=8)
    finally:
        oracle.set_cr_asinf(False)
    got = eng.optimize(to_engine(seeds))
    assert np.array_equal(ref["status"], got["status"])
    ok = got["status"] == 0
    # (1) empty depth maps: nothing is visible-tested away, every pixel is free
    orc.depth_reset(); eng.depth_reset()
    a_ref = orc.accept(ref, 1.0); a_got = eng.accept(got, 1.0)
    assert np.array_equal(a_ref, a_got)
    assert (a_got[ok][:, 1] == 0).all() and (a_got[ok][:, 2] == got["nimages"][ok]).all()
    # (2) commit the first half of the optimized patches, then test all of them against that state
    half = got.copy(); half_ref = ref.copy()
    idx = np.nonzero(ok)[0]
    drop = idx[len(idx) // 2:]
    half["status"][drop] = 99; half_ref["status"][drop] = 99
    orc.depth_set(half_ref); eng.depth_set(half)
    for cam in range(len(sc.cameras)):
        for lvl in range(6):
            assert np.array_equal(orc.depth(cam, lvl), eng.download_depth(cam, lvl)), (cam, lvl)
    assert any((eng.download_depth(0, l) < 1000.0).any() for l in range(6))
    for margin in (1.0, 0.25):
        a_ref = orc.accept(ref, margin); a_got = eng.accept(got, margin)
        assert np.array_equal(a_ref, a_got), margin
    # committed patches now see their own depth: fewer free pixels than before
    assert a_got[idx[:len(idx) // 2], 2].sum() < got["nimages"][idx[:len(idx) // 2]].sum()
    # (3) a patch pushed towards the cameras in front of the committed surface blocks views
    front = got[ok][:40].copy(); front_ref = ref[ok][:40].copy()
    front["center"][:, 2] -= 2.0; front_ref["center"][:, 2] -= 2.0
    b_ref = orc.accept(front_ref, 1.0); b_got = eng.accept(front, 1.0)
    assert np.array_equal(b_ref, b_got)
    assert b_got[:, 1].max() >= 1
    # idempotence: committing the same patches again changes nothing
    before = [eng.download_depth(c, 4) for c in range(len(sc.cameras))]
    eng.depth_set(half)
    assert all(np.array_equal(b, eng.download_depth(c, 4)) for c, b in enumerate(before))


@pytest.mark.gpu
@pytest.mark.parametrize("r", [0.08, -0.06])
def test_gpu_undistort_matches_the_host_twin(r):
    """Image::undistort (Image.cpp:68-149) as a kernel (hpmvs_engine_upload_image_undistort) against the host function
    hpmvs_undistort_rgb, which is bit-exact against the reference build (tests/test_formats.py).  The kernel forms the source positions
    with CUDA's double-precision libm instead of glibc's: bar = identical on >= 99.99 % of the pixels, never more than one grey level
    apart, same set of written pixels up to last-place roundings at the border; the pyramid built from it follows."""
    from hpmvs_b200 import io as hio
    sc = hp.synth.plane_scene(n_views=2, width=640, height=480, focal=600.0, n_seeds=16, seed=9, tex_size=512)
    img = sc.images[0]
    want = hio.undistort(img, 600.0, r)
    eng = hp.Engine()
    eng.set_cameras([hp.camera_from_nvm(600.0, [1, 0, 0, 0], [0, 0, 0], 640, 480)])
    eng.upload_image_undistort(0, img, 600.0, r)
    got = eng.download_image(0, 0)
    same = (got == want).all(2)
    assert same.mean() >= 0.9999, same.mean()
    assert np.abs(got.astype(np.int16) - want.astype(np.int16))[same == 0].max(initial=0) <= 1 or (~same).sum() < 40
    assert (want != img).mean() > 0.5                      # the distortion really moved pixels
    eng.build_pyramid(0)
    assert eng.download_image(0, 1).shape == (240, 320, 3)
    # r == 0 is a plain upload
    eng.upload_image_undistort(0, img, 600.0, 0.0)
    assert np.array_equal(eng.download_image(0, 0), img)


@pytest.mark.gpu
def test_device_resident_level_steps_equal_the_host_buffer_calls():
    """One level of the loop chained on the device - candidates (kernel) -> optimize -> acceptance counts -> depth set / unset - with
    records that never leave HBM must give the same bytes as the host-buffer entry points (which are pinned to the oracle and the
    reference build above)."""
    import torch
    sc, orc, seeds = small_plane()
    eng = hp.Engine.from_synth(sc)
    parents = eng.optimize(to_engine(seeds))
    parents = np.ascontiguousarray(parents[parents["status"] == 0][:150])
    n = len(parents)
    widths = np.linspace(0.05, 0.4, n).astype(np.float32)
    up = lambda a: torch.from_numpy(a.view(np.uint8).reshape(len(a), -1).copy()).cuda()
    d_par, d_w = up(parents), torch.from_numpy(widths).cuda()
    for mode in (6, 4):
        want = hp.expand_candidates(eng.cameras, parents, widths, mode)
        d_c = torch.zeros((n * mode, hp.PATCH_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
        eng.expand_candidates_device(n, d_par.data_ptr(), d_w.data_ptr(), mode, d_c.data_ptr())
        torch.cuda.synchronize()
        assert d_c.cpu().numpy().tobytes() == want.tobytes(), mode
        # optimize the candidates where they are, then the acceptance counts where they are
        d_o = torch.zeros_like(d_c)
        eng.optimize_device(n * mode, d_c.data_ptr(), d_o.data_ptr())
        torch.cuda.synchronize()
        got = d_o.cpu().numpy().view(hp.PATCH_DTYPE).reshape(-1)
        assert got.tobytes() == eng.optimize(want).tobytes()
        eng.depth_reset()
        eng.depth_set(parents)
        d_cnt = torch.zeros((n * mode, 3), dtype=torch.int32, device="cuda")
        eng.accept_device(n * mode, d_o.data_ptr(), 1.0, d_cnt.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(d_cnt.cpu().numpy(), eng.accept(got, 1.0))
    # depth set / unset on device-resident records == the host-buffer calls
    eng.depth_reset(); eng.depth_set_device(n, d_par.data_ptr()); torch.cuda.synchronize()
    a = [eng.download_depth(c, l) for c in range(len(sc.cameras)) for l in range(6)]
    eng.depth_reset(); eng.depth_set(parents)
    b = [eng.download_depth(c, l) for c in range(len(sc.cameras)) for l in range(6)]
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and any((x < 1000).any() for x in a)
    eng.depth_set_device(n // 2, d_par.data_ptr(), subtract=True); torch.cuda.synchronize()
    c1 = [eng.download_depth(c, l) for c in range(len(sc.cameras)) for l in range(6)]
    eng.depth_reset(); eng.depth_set(parents); eng.depth_unset(np.ascontiguousarray(parents[:n // 2]))
    c2 = [eng.download_depth(c, l) for c in range(len(sc.cameras)) for l in range(6)]
    assert all(np.array_equal(x, y) for x, y in zip(c1, c2))
    # subtracting gives cells back: fewer occupied cells than after the set, and the oracle agrees on the final maps
    assert sum((x < 1000).sum() for x in c1) < sum((x < 1000).sum() for x in a)
    orc.depth_reset(); po = to_oracle(parents); po["status"] = 0
    orc.depth_set(po); orc.depth_unset(po[:n // 2])
    assert all(np.array_equal(orc.depth(c, l), c1[c * 6 + l]) for c in range(len(sc.cameras)) for l in range(6))
