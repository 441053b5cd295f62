"""BASELINE config 3 family: 11-view 3072x2048 synthetic scene, full expand -> optimize -> filter loop through the
level-synchronous host driver, GPU engine vs CPU oracle backend (same host logic).  Prints one JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import hpmvs_b200 as hp, oracle
from hpmvs_b200 import pipeline, synth
from test_pipeline import OracleBackend

synth.USE_GPU_RENDERER = torch.cuda.is_available()
n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
sc = synth.plane_scene(n_views=11, width=3072, height=2048, focal=2800.0, radius=10.0, arc_deg=100.0, n_seeds=n_seeds,
                       extent=3.0, seed=3, tex_size=2048, depth_noise=0.3, plane_half=9.0)
eng = hp.Engine.from_synth(sc)
seeds, valid = hp.seed_patches(eng.options, eng.cameras, sc.points, sc.meas_offsets, sc.meas_cam)
seeds = np.ascontiguousarray(seeds[valid])
width0 = float(np.median(seeds["scale"])) * 2.2
args = dict(origin=(-16.0, -16.0, -16.0), root_width=width0 * 512, start_level=9, final_level=12, final_min_level=0)   # root cube covers the scene
res = {}
for name in ("gpu", "cpu"):
    if name == "gpu":
        be = pipeline.EngineBackend(eng)
    else:
        oracle.set_cr_asinf(True)
        be = OracleBackend(oracle.OracleScene.from_synth(sc), nthreads=len(os.sched_getaffinity(0)))
    d = pipeline.WavefrontDriver(be, **args)
    t = time.perf_counter(); out = d.run(seeds); dt = time.perf_counter() - t
    res[name] = dict(seconds=dt, patches=int(len(out)), optimize_calls=d.stats.optimized_calls, optimized_ok=d.stats.optimized_ok,
                     seconds_optimize=d.stats.seconds_optimize, seconds_accept=d.stats.seconds_accept, per_level=d.stats.per_level)
    res[name + "_out"] = out
g, c = res["gpu_out"], res["cpu_out"]
same = len(g) == len(c) and np.array_equal(g["center"], c["center"])
# patches present bit-for-bit (centre + normal + view list) in both sets
key = lambda r: set(zip(r["center"].tobytes()[i * 16:(i + 1) * 16] + r["normal"].tobytes()[i * 16:(i + 1) * 16] for i in range(len(r))))
kg = {r["center"].tobytes() + r["normal"].tobytes() + r["images"][:r["nimages"]].tobytes() for r in g}
kc = {r["center"].tobytes() + r["normal"].tobytes() + r["images"][:r["nimages"]].tobytes() for r in c}
matched = len(kg & kc)
from hpmvs_b200 import io as hio
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
hio.write_ext_ply(os.path.join(ROOT, "gpurun_out", "patches-final.ply"), res["gpu_out"], binary=True)
print(json.dumps({"workload": "11-view 3072x2048 synthetic plane, %d seeds, 4 tree levels, level-synchronous driver" % len(seeds),
                  "host_threads_cpu_backend": len(os.sched_getaffinity(0)), "identical_patch_sets": bool(same),
                  "bit_identical_patches": matched, "bit_identical_fraction_of_gpu_set": matched / max(1, len(g)),
                  "gpu": {k: v for k, v in res["gpu"].items()}, "cpu": {k: v for k, v in res["cpu"].items()},
                  "optimize_stage_speedup": res["cpu"]["seconds_optimize"] / res["gpu"]["seconds_optimize"],
                  "whole_loop_speedup": res["cpu"]["seconds"] / res["gpu"]["seconds"]}))
