"""BASELINE configs[2] family, against the REAL reference: the whole expand -> optimize -> filter loop on one synthetic NVM scene,
(a) the reference's own command line (oracle/_ref/hpmvs_ref = /root/reference/src compiled where it lies) on all host cores and
(b) this repository's level-synchronous driver (hpmvs_b200/pipeline.py) on the B200 engine, configured with the reference's octree
geometry (root = bounding cube of the accepted seeds as Scene.cpp:186-193, same first and last tree level).  The two schedulers differ
(priority queues per sub-tree vs level-synchronous batches), so the patch sets are compared statistically.  Prints one JSON line.
usage: python scripts/pipeline_vs_reference.py [n_seeds]"""
import json, glob, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hpmvs_b200 as hp
from hpmvs_b200 import pipeline
from oracle import ref

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 300
cores = len(os.sched_getaffinity(0))
sc = hp.synth.plane_scene(n_views=8, width=1280, height=960, focal=1200.0, radius=8.0, arc_deg=40.0, n_seeds=n_seeds, extent=2.5,
                          seed=2, tex_size=1024)
tmp = tempfile.mkdtemp(prefix="hpmvs_pvr_")
nvm = os.path.join(tmp, "scene.nvm")
hp.synth.write_nvm(sc, nvm)


def read_ply(path):
    L = open(path).read().split("\n")
    n = int([l for l in L[:20] if l.startswith("element vertex")][0].split()[2])
    h = L.index("end_header") + 1
    return np.array([[float(x) for x in l.split()[:10]] for l in L[h:h + n]], np.float64).reshape(n, 10)


def quality(xyz, nz):
    return {"patches": int(len(xyz)), "rms_distance_to_true_plane": float(np.sqrt(np.mean(xyz[:, 2] ** 2))),
            "mean_abs_normal_z": float(np.mean(np.abs(nz)))}


# ---- (a) the reference --------------------------------------------------------------------------------------------------
t = time.perf_counter()
r = ref.run_cli(nvm, os.path.join(tmp, "ref"), threads=cores)
t_ref = time.perf_counter() - t
assert r.returncode == 0, r.stderr[-2000:]
levels = {}
for f in glob.glob(os.path.join(tmp, "ref", "patches-[0-9]*.ply")):
    v = read_ply(f)
    levels[int(os.path.basename(f)[8:-4]) // 10] = (int(len(v)), float(np.median(v[:, 9])))
fin = read_ply(os.path.join(tmp, "ref", "patches-final.ply"))
first_level, last_level = min(levels), max(levels)
ref_res = dict(seconds=t_ref, host_threads=cores, per_level={k: levels[k][0] for k in sorted(levels)}, **quality(fin[:, :3], fin[:, 5]))

# ---- (b) the engine behind the level-synchronous driver ----------------------------------------------------------------------
t0 = time.perf_counter()
eng = hp.Engine.from_synth(sc)                    # uploads the images, builds the pyramids on the device
eng.set_start_mode(True)
seeds, valid = hp.seed_patches(eng.options, eng.cameras, sc.points, sc.meas_offsets, sc.meas_cam)
seeds = np.ascontiguousarray(seeds[valid])
t_setup = time.perf_counter() - t0
# root cube as Scene::initPatches builds it (Scene.cpp:186-193) from the accepted seeds
pre = eng.optimize(seeds)
okp = pre["status"] == 0
okp &= ~(np.linalg.norm(pre["center"][:, :3] - seeds["center"][:, :3], axis=1) > pre["scale"] * 2)
c = pre["center"][okp][:, :3].astype(np.float64)
mn, mx = c.min(0), c.max(0)
width = float((mx - mn).max())
origin = (mn + mx) / 2.0 - width / 2.0
t1 = time.perf_counter()
out, stats = pipeline.run_native(eng, seeds, origin=origin, root_width=width, start_level=first_level, final_level=last_level)   # C++ driver
t_ours = time.perf_counter() - t1
t2 = time.perf_counter()
d = pipeline.WavefrontDriver(pipeline.EngineBackend(eng), origin=origin, root_width=width, start_level=first_level, final_level=last_level, cameras=eng.cameras)
out_py = d.run(seeds)
t_py = time.perf_counter() - t2
ours = dict(seconds=t_ours, seconds_numpy_driver=t_py, drivers_identical=bool(out.tobytes() == out_py.tobytes()),
            seconds_scene_upload_and_seeding=t_setup, seconds_optimize=stats.seconds_optimize, seconds_accept=stats.seconds_accept,
            optimize_calls=stats.optimized_calls, per_level={lv: int(n_ext) for lv, n_ext, _ in stats.per_level},
            **quality(out["center"][:, :3].astype(np.float64), out["normal"][:, 2]))
print(json.dumps({"workload": f"8-view 1280x960 synthetic plane, {n_seeds} NVM points, tree levels {first_level}..{last_level} (root cube {width:.4f})",
                  "reference_cli": ref_res, "b200_wavefront_driver": ours,
                  "final_patches_per_second": {"reference_cli": ref_res["patches"] / t_ref, "b200": ours["patches"] / t_ours},
                  "whole_loop_speedup": t_ref / t_ours}))
