"""Per-seed work of the bench workload (evaluations x textures) for the multi-GPU split: python scripts/work_balance_probe.py out.npz"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, hpmvs_b200 as hp
from hpmvs_b200 import gather
sc, _ = bench.cached_scene("city100", 0)
eng = hp.Engine.from_synth(sc)
seeds, valid = hp.seed_patches(eng.options, eng.cameras, sc.points, sc.meas_offsets, sc.meas_cam)
seeds = np.ascontiguousarray(seeds[valid])
out = eng.optimize(seeds)
origin, width = gather.root_cube(seeds)
res = {}
for per_rank in (16, 64):
    for world in (2, 4, 8):
        cell, rk, ncell = gather.shard_cells(seeds, origin, width, max(100, per_rank * world), world)
        res[f"rk_{per_rank}_{world}"] = rk; res[f"cell_{per_rank}_{world}"] = cell
np.savez_compressed(sys.argv[1], nimages=seeds["nimages"], status=out["status"], evals=out["evals"], textures=out["textures"],
                    center=seeds["center"][:, :3], **res)
print("saved", len(seeds))
