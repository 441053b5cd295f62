"""BASELINE configs[3]/[4] family at small scale: the whole expand -> optimize -> filter loop sharded over N GPUs by octree cell,
NCCL gather of the final patch records + border de-duplication, compared with the same loop on one GPU.
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/pipeline_multigpu.py [n_seeds]
Rank 0 prints one JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
import numpy as np
import torch
import torch.distributed as dist
import hpmvs_b200 as hp
from hpmvs_b200 import gather, pipeline

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
sc = hp.synth.plane_scene(n_views=8, width=1280, height=960, focal=1200.0, radius=8.0, arc_deg=40.0, n_seeds=n_seeds, extent=2.5,
                          seed=2, tex_size=1024)                                   # every rank holds the whole scene (replicated)
eng = hp.Engine.from_synth(sc, device=local)
eng.set_start_mode(True)
seeds, valid = hp.seed_patches(eng.options, eng.cameras, sc.points, sc.meas_offsets, sc.meas_cam)
seeds = np.ascontiguousarray(seeds[valid])
pre = eng.optimize(seeds)
ok = (pre["status"] == 0) & ~(np.linalg.norm(pre["center"][:, :3] - seeds["center"][:, :3], axis=1) > pre["scale"] * 2)
c = pre["center"][ok][:, :3].astype(np.float64)
mn, mx = c.min(0), c.max(0)
width = float((mx - mn).max()); origin = (mn + mx) / 2.0 - width / 2.0
first, last = 5, 8
args = dict(origin=origin, root_width=width, start_level=first, final_level=last)
# the reference's sub-tree split over the seed points (getSubTrees, src/main.cpp:50-96), sub-trees dealt to the ranks
sub = pipeline.shard_subtrees(seeds, origin, width, max(100, 16 * world), world)
_, rk, nsub = gather.shard_cells(seeds, origin, width, max(100, 16 * world), world)
my_seeds = np.ascontiguousarray(seeds[rk == rank]) if world > 1 else seeds
exchanges = [0, 0.0]


def exchange(mine):                                   # per-step border hand-off: every rank gets every rank's accepted records
    t = time.perf_counter()
    allr, _ = gather.gather_patches(mine)
    exchanges[0] += 1; exchanges[1] += time.perf_counter() - t
    return allr


def quality(r):
    return {"patches": int(len(r)), "rms_distance_to_true_plane": float(np.sqrt(np.mean(r["center"][:, 2].astype(np.float64) ** 2))) if len(r) else None}


torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
mine, st = pipeline.run_native(eng, my_seeds, shard_count=world, shard_rank=rank, subtrees=sub if world > 1 else None,
                               exchange=exchange if world > 1 else None, **args)
t_shard = time.perf_counter() - t0
allr, owner = gather.gather_to_root(mine)
merged = None
if rank == 0:
    keep = gather.dedup_border(allr, owner, cell=width / (1 << last), origin=origin)
    merged = allr[keep]
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t_total = time.perf_counter() - t0
tt = torch.tensor([t_shard, t_total], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    t1 = time.perf_counter()
    single, st1 = pipeline.run_native(eng, seeds, **args)
    t_single = time.perf_counter() - t1
    # border mismatch: patches of the single-GPU result without a merged patch in the same finest cell, and vice versa
    w = width / (1 << last)
    key = lambda r: set(map(tuple, np.floor((r["center"][:, :3].astype(np.float64) - origin) / w).astype(np.int64).tolist()))
    ka, kb = key(single), key(merged)
    print(json.dumps({"workload": f"8-view 1280x960 synthetic plane, {n_seeds} NVM points, tree levels {first}..{last}, {nsub} sub-trees dealt to {world} GPUs, per-step NCCL exchange",
                      "exchange_calls": exchanges[0], "exchange_seconds_rank0": exchanges[1], "optimize_calls_rank0": int(st.optimized_calls),
                      "n_gpus": world, "sharded": dict(seconds_slowest_rank_pipeline=float(tt[0]), seconds_incl_gather_dedup=float(tt[1]),
                                                        gathered=int(len(allr)), **quality(merged)),
                      "single_gpu": dict(seconds=t_single, **quality(single)),
                      "cells_only_in_single": len(ka - kb), "cells_only_in_sharded": len(kb - ka), "cells_common": len(ka & kb),
                      "speedup_vs_single_gpu": t_single / float(tt[1])}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
