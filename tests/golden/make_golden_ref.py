"""Mints tests/golden/ref_*.npz from the REFERENCE ITSELF: oracle/_ref/libhpmvs_ref.so = /root/reference/src/hpmvs/*.cpp
compiled where they lie (make -C oracle refhpmvs; Eigen/glog/gflags/jpeglib stood in for by oracle/shim/).
Every record is what mo3d::PatchOptimizer::optimize(Patch3d&) returned for the seed built by the reference's own
Scene::initPatches arithmetic, plus Scene::setDepths / depthTests / viewBlockTest / pixelFreeTests on the results.
Can only run where /root/reference exists.  Run from the repo root:  python tests/golden/make_golden_ref.py"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import hpmvs_b200 as hp  # noqa: E402  (scene generator + NVM/PPM writer only)
import oracle  # noqa: E402  (seed construction; itself checked against Scene::initPatches below)
from oracle import ref  # noqa: E402

SCENES = {
    "ref_plane6": ("plane_scene", dict(n_views=6, width=640, height=480, focal=600.0, n_seeds=200, seed=5, tex_size=512, arc_deg=36.0)),
    "ref_city16": ("city_scene", dict(n_views=16, width=480, height=270, focal=375.0, n_seeds=3000, seed=7, n_boxes=5)),
}


def mint(name, gen, kw):
    sc = getattr(hp.synth, gen)(**kw)
    rs = ref.RefScene.from_synth(sc)
    orc = oracle.OracleScene.from_synth(sc)
    seeds, valid = orc.seed_patches(sc.points, sc.meas_offsets, sc.meas_cam)
    seeds = np.ascontiguousarray(seeds[valid])
    out = rs.optimize_batch(seeds, nthreads=1)
    ok = out["status"] == 0
    # the seeds above are the reference's: Scene::initPatches run whole must leave exactly the accepted ones in its octree
    tree = ref.RefScene.from_synth(sc).init_patches()
    d = out["center"][:, :3] - seeds["center"][:, :3]
    moved = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2], dtype=np.float32)
    keep = ok & ~(moved > seeds["scale"] * np.float32(2))
    assert set(map(bytes, out["center"][keep])) == set(map(bytes, tree["center"])), "seed construction differs from Scene::initPatches"
    rs.depth_reset()
    rs.depth_set(out)
    accept = rs.accept(out[ok], 1.0)
    depth_sha = hashlib.sha256(b"".join(rs.depth(c, l).tobytes() for c in range(rs.n_cameras) for l in range(6))).hexdigest()
    cams = np.frombuffer(b"".join(bytes(rs.camera(i)) for i in range(rs.n_cameras)), np.uint8)
    pyr_sha = hashlib.sha256(b"".join(rs.image(c, l).tobytes() for c in range(rs.n_cameras) for l in range(6))).hexdigest()
    covis = rs.covis()
    np.savez_compressed(
        os.path.join(ROOT, "tests", "golden", name + ".npz"), generator=gen, scene_kwargs=str(kw),
        scene_sha256=hashlib.sha256(np.stack(sc.images).tobytes()).hexdigest(), cameras=cams, pyramid_sha256=pyr_sha,
        covis_offsets=np.cumsum([0] + [len(c) for c in covis]).astype(np.int32), covis_ids=np.asarray([v for c in covis for v in c], np.int32),
        seeds_center=seeds["center"], seeds_normal=seeds["normal"], seeds_scale=seeds["scale"], seeds_nimages=seeds["nimages"],
        seeds_images=seeds["images"][:, :32].astype(np.int16), ok=ok, center=out["center"], normal=out["normal"],
        nimages=out["nimages"], images=out["images"][:, :32].astype(np.int16), color=out["color"],
        tree_centers=tree["center"], accept=accept, depth_sha256=depth_sha)
    print(f"{name}: {len(seeds)} seeds, {int(ok.sum())} optimized by the reference, {len(tree)} in the octree after Scene::initPatches")


if __name__ == "__main__":
    for name, (gen, kw) in SCENES.items():
        mint(name, gen, kw)
