"""The drop-in claim, end to end: the REFERENCE'S OWN command line (src/main.cpp, Scene::initPatches, CellProcessor, DynOctTree -
compiled from /root/reference where it lies) linked once with its own src/hpmvs/PatchOptimizer.cpp and once with
integration/PatchOptimizer_b200.cpp (= the C ABI of this repository), run on the same NVM scene with one host thread
(the reference is only deterministic single-threaded).  Both must write the same PLY files, byte for byte.

Both command lines are the *_det builds: linked with a monotone operator new (oracle/ref_monotone_new.cpp), because the
reference queues newly branched cells in POINTER order (std::set<Leaf*>, src/hpmvs/CellProcessor.cpp:289-305): with an
ordinary heap two different binaries walk the cells in different orders and only ~60 % of their final patches coincide even
when every optimize() call agrees bit for bit (measured with a CPU oracle behind the same shim)."""
import os

import numpy as np
import pytest

import hpmvs_b200 as hp
from oracle import ref

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not (os.path.exists(ref.BIN_PATH + "_det") and os.path.exists(ref.DROPIN_BIN_PATH + "_det")),
                    reason="oracle/_ref/hpmvs_ref[_b200] not built (needs /root/reference at build time)")
def test_reference_cli_on_the_engine_writes_the_same_ply(tmp_path):
    sc = hp.synth.plane_scene(n_views=5, width=320, height=240, focal=300.0, n_seeds=60, seed=13, tex_size=256)
    nvm = str(tmp_path / "scene.nvm")
    hp.synth.write_nvm(sc, nvm)
    a = ref.run_cli(nvm, str(tmp_path / "cpu"), threads=1, monotone_heap=True)
    assert a.returncode == 0, a.stderr[-2000:]
    b = ref.run_cli(nvm, str(tmp_path / "gpu"), threads=1, dropin=True, monotone_heap=True)
    assert b.returncode == 0, b.stderr[-2000:]
    pa = open(tmp_path / "cpu" / "patches-final.ply", "rb").read()
    pb = open(tmp_path / "gpu" / "patches-final.ply", "rb").read()
    n = int([l for l in pa.split(b"\n")[:20] if l.startswith(b"element vertex")][0].split()[2])
    assert n > 200, n
    if pa != pb:
        la, lb = pa.split(b"\n"), pb.split(b"\n")
        diff = sum(x != y for x, y in zip(la, lb)) + abs(len(la) - len(lb))
        raise AssertionError(f"patches-final.ply differs: {len(la)} vs {len(lb)} lines, {diff} differing")
    # every intermediate level dump as well
    for name in sorted(os.listdir(tmp_path / "cpu")):
        assert open(tmp_path / "cpu" / name, "rb").read() == open(tmp_path / "gpu" / name, "rb").read(), name
