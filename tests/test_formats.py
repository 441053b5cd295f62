"""File formats either side of the path ("next" row f-4): NVM_V3 reader, PPM level-0 images, ext-PLY writer."""
import os
import struct

import numpy as np

import hpmvs_b200 as hp
from hpmvs_b200 import io as hio


def test_nvm_round_trip(tmp_path):
    sc = hp.synth.plane_scene(n_views=4, width=96, height=64, focal=80.0, n_seeds=25, seed=9, tex_size=64)
    path = str(tmp_path / "scene.nvm")
    hp.synth.write_nvm(sc, path)
    back = hio.read_nvm(path)
    assert len(back.cameras) == 4 and back.points.shape == sc.points.shape
    for a, b in zip(sc.cameras, back.cameras):
        assert b.filename == str(tmp_path / a.filename)                       # fixPath (NVMReader.cpp:138-144)
        assert b.f == a.f and np.array_equal(a.q, b.q) and np.array_equal(a.c, b.c) and b.r == 0.0
        assert (b.width, b.height) == (96, 64)
    assert np.array_equal(back.points, sc.points)
    assert np.array_equal(back.meas_offsets, sc.meas_offsets) and np.array_equal(back.meas_cam, sc.meas_cam)
    assert all(np.array_equal(x, y) for x, y in zip(sc.images, back.images))
    # the reloaded scene drives the same host surface bit for bit
    c0 = hp.camera_from_nvm(sc.cameras[0].f, sc.cameras[0].q, sc.cameras[0].c, 96, 64)
    c1 = hp.camera_from_nvm(back.cameras[0].f, back.cameras[0].q, back.cameras[0].c, 96, 64)
    assert bytes(c0) == bytes(c1)


def test_nvm_quirks(tmp_path):
    # quoted file names, case-insensitive tag, a second model that must be skipped, trailing empty model
    txt = ('nvm_v3\n\n2\n"imga.ppm" 500.5 1 0 0 0 0.5 -1 2 -0.01 0\nb.ppm 400 0.7071 0 0.7071 0 1 1 1 0 0\n\n'
           '2\n0.1 0.2 0.3 10 20 30 2 0 7 1.5 2.5 1 8 3.5 4.5\n1 2 3 0 0 0 0\n\n'
           '1\nc.ppm 1 1 0 0 0 0 0 0 0 0\n0\n\n0\n')
    path = tmp_path / "q.nvm"
    path.write_text(txt)
    s = hio.read_nvm(str(path), fix_path=False, load_images=False)
    assert [c.filename for c in s.cameras] == [" imga.ppm ", "b.ppm"]          # quotes become blanks (NVMReader.cpp:73)
    assert s.cameras[1].f == 400 and np.allclose(s.cameras[1].q, [0.7071, 0, 0.7071, 0])
    assert s.points.shape == (2, 3) and s.meas_offsets.tolist() == [0, 2, 2] and s.meas_cam.tolist() == [0, 1]
    bad = tmp_path / "bad.nvm"
    bad.write_text("NVM_V2\n1\n")
    try:
        hio.read_nvm(str(bad), load_images=False)
        assert False
    except hp.HpmvsError:
        pass


def _records():
    r = np.zeros(3, hp.PATCH_DTYPE)
    r["center"][:, :3] = [[1.5, -2.25, 3.0], [0.1, 0.2, 0.3], [1e-3, 2e5, -7.0]]
    r["normal"][:, :3] = [[0, 0, -1], [0.6, 0, -0.8], [1, 0, 0]]
    r["color"] = [[10.9, 200.2, 255.0], [0, 1.5, 2.5], [128, 64, 32]]
    r["scale"] = [0.125, 0.5, 3.0]
    r["nimages"] = [3, 0, 2]
    r["images"][0, :3] = [4, 1, 7]; r["images"][2, :2] = [0, 99]
    return r


def test_ext_ply_ascii(tmp_path):
    path = str(tmp_path / "a.ply")
    hio.write_ext_ply(path, _records())
    lines = open(path).read().split("\n")
    assert lines[:3] == ["ply", "format ascii 1.0", "element vertex 3"]
    assert "property float scalar_scale" in lines and "property list uint uint visible_cameras" in lines
    body = lines[lines.index("end_header") + 1:]
    assert body[0] == "1.5 -2.25 3 0 0 -1 10 200 255 0.125 "         # ostream default precision, colours truncated
    assert body[1] == "0.1 0.2 0.3 0.6 0 -0.8 0 1 2 0.5 "
    assert body[2] == "0.001 200000 -7 1 0 0 128 64 32 3 "
    assert body[3:6] == ["3 4 1 7 ", "0 ", "2 0 99 "]


def test_ext_ply_binary_and_light(tmp_path):
    path = str(tmp_path / "b.ply")
    r = _records()
    hio.write_ext_ply(path, r, binary=True)
    raw = open(path, "rb").read()
    hdr, data = raw.split(b"end_header\n", 1)
    assert b"format binary_little_endian 1.0" in hdr
    rec = struct.Struct("<3f3f3Bf")
    for i in range(3):
        x, y, z, nx, ny, nz, cr, cg, cb, sc = rec.unpack_from(data, i * rec.size)
        assert (x, y, z) == tuple(r["center"][i, :3]) and (nx, ny, nz) == tuple(r["normal"][i, :3])
        assert (cr, cg, cb) == tuple(int(v) for v in r["color"][i]) and sc == r["scale"][i]
    vis = np.frombuffer(data[3 * rec.size:], "<u4")
    assert vis.tolist() == [3, 4, 1, 7, 0, 2, 0, 99]
    # "-light" variant: xyz + rgb only (src/main.cpp:165-172)
    hio.write_ext_ply(path, r, binary=True, normal=False, scale=False, visibility=False)
    raw = open(path, "rb").read()
    hdr, data = raw.split(b"end_header\n", 1)
    assert b"nx" not in hdr and b"point_visibility" not in hdr and len(data) == 3 * 15
    # empty set
    hio.write_ext_ply(path, np.zeros(0, hp.PATCH_DTYPE))
    assert "element vertex 0" in open(path).read()


def _parse_ext_ply_ascii(path):
    """ext-PLY (ascii) as DynOctTree::toExtPly writes it (doctree.h:525-622) -> patch records."""
    lines = open(path).read().split("\n")
    n = int([l for l in lines if l.startswith("element vertex")][0].split()[2])
    h = lines.index("end_header") + 1
    rec = np.zeros(n, hp.PATCH_DTYPE)
    for i in range(n):
        f = lines[h + i].split()
        rec["center"][i, :3] = [np.float32(v) for v in f[0:3]]; rec["center"][i, 3] = 1
        rec["normal"][i, :3] = [np.float32(v) for v in f[3:6]]
        rec["color"][i] = [float(v) for v in f[6:9]]
        rec["scale"][i] = np.float32(f[9])
        g = lines[h + n + i].split()
        k = int(g[0]); rec["nimages"][i] = k; rec["images"][i, :k] = [int(v) for v in g[1:1 + k]]
    return rec


def test_ext_ply_writer_reproduces_the_reference_cli_file(tmp_path):
    # tests/golden/ref_cli_patches-40.ply was written by the REFERENCE'S OWN command line (oracle/_ref/hpmvs_ref_det, built from
    # /root/reference/src where it lies) on plane_scene(n_views=5, 320x240, f=300, 60 seeds, seed 13): the level-40 dump.
    # Our writer must produce the same bytes from the same records (format, number formatting, visibility lists).
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_cli_patches-40.ply")
    rec = _parse_ext_ply_ascii(gold)
    assert len(rec) == 268
    out = str(tmp_path / "ours.ply")
    hio.write_ext_ply(out, rec, binary=False)
    assert open(out, "rb").read() == open(gold, "rb").read()


def test_reference_cli_reads_our_nvm_and_its_ply_round_trips(tmp_path):
    from oracle import ref
    if not os.path.exists(ref.BIN_PATH):
        import pytest
        pytest.skip("oracle/_ref/hpmvs_ref not built (needs /root/reference)")
    sc = hp.synth.plane_scene(n_views=4, width=320, height=240, focal=300.0, n_seeds=80, seed=17, tex_size=256)
    nvm = str(tmp_path / "scene.nvm")
    hp.synth.write_nvm(sc, nvm)                                     # our NVM_V3 + PPM writer ...
    r = ref.run_cli(nvm, str(tmp_path / "out"), threads=2, extra=["--light_output=1"])   # ... read by the reference's NVMReader / CImg
    assert r.returncode == 0, r.stderr[-1000:]
    final = str(tmp_path / "out" / "patches-final.ply")
    rec = _parse_ext_ply_ascii(final)
    assert len(rec) > 100 and (rec["nimages"] >= 2).all()
    ours = str(tmp_path / "ours.ply")
    hio.write_ext_ply(ours, rec, binary=False)
    assert open(ours, "rb").read() == open(final, "rb").read()
    # the binary "light" variant: same header, same size
    light = open(str(tmp_path / "out" / "patches-final-light.ply"), "rb").read()
    hio.write_ext_ply(str(tmp_path / "ours_light.ply"), rec, binary=True, normal=False, scale=False, visibility=False)
    mine = open(str(tmp_path / "ours_light.ply"), "rb").read()
    assert light[:light.index(b"end_header")] == mine[:mine.index(b"end_header")] and len(light) == len(mine)


def test_nvm_writer_matches_the_reference_saveNVM(tmp_path):
    from oracle import ref
    if not ref.available():
        import pytest
        pytest.skip("oracle/_ref/libhpmvs_ref.so not built (needs /root/reference)")
    sc = hp.synth.plane_scene(n_views=4, width=96, height=64, focal=80.0, n_seeds=40, seed=9, tex_size=64)
    sc.cameras[1].r = -0.0123
    src = str(tmp_path / "scene.nvm")
    hp.synth.write_nvm(sc, src)
    theirs, ours = str(tmp_path / "theirs.nvm"), str(tmp_path / "ours.nvm")
    ref.RefScene(src).save_nvm(theirs)                      # mo3d::NVMReader::readFile + saveNVM
    hio.rewrite_nvm(src, ours)                              # hpmvs_nvm_open + hpmvs_nvm_write
    assert open(ours, "rb").read() == open(theirs, "rb").read()
    back = hio.read_nvm(ours, load_images=False)            # and it reads back
    assert len(back.cameras) == 4 and back.points.shape == (40, 3)
