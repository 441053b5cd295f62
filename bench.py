#!/usr/bin/env python
"""bench.py - optimized patches/sec of the PatchOptimizer::optimize() hot path on synthetic N-view scenes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload city100|plane8|...]

One "step" = one pass of the hot path over one batch of seed patches.  Default workload = the configuration BASELINE.json's
north_star quotes its target on: the 100-view 1080p synthetic city block (configs[3]), all of its valid seed patches per step
(= the batch Scene::initPatches optimises, /root/reference/src/hpmvs/Scene.cpp:114-178).  Prints ONE JSON line (rank 0).

* value        : optimized (status OK) patches / s, whole job, patch records (and their start angles) already resident in HBM,
                 timed with CUDA events on the launching streams (the engine's fused kernel only).
* e2e          : same metric through the public C ABI call hpmvs_optimize_batch_submit() with PINNED HOST buffers in start mode 1
                 (the parity-certified configuration: bit-identical to the reference build): host start angles + H2D + kernel + D2H
                 inside the timed region; at N > 1 also the final gather to rank 0 + border de-duplication of every step.
* roofline     : algorithmic gather bytes (588 B per sampled 7x7x3 texture + 2*208 B record I/O per patch)
                 / kernel time, against the measured HBM peak in MEASURED_PEAKS.json.
* cpu_baseline : the reference's own PatchOptimizer (oracle/_ref/libhpmvs_ref.so, built from /root/reference's sources; falls back
                 to the oracle restatement when that prebuilt library is absent) on this box's host cores, bounded sample.
--impl reference times that CPU path alone with all host threads on the SAME batch (identical `config`).

Multi-GPU (--gpus N under torchrun): the city workloads are STRONG scaling - the fixed scene's seed batch is split by octree
sub-tree (the reference's own getSubTrees split, src/main.cpp:50-96) and the sub-trees are dealt to the ranks; the plane workloads keep
the round-1 weak-scaling form (every rank its own 10 k-seed draw) for comparison.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TEX_BYTES = 588          # 49 samples x 4 taps x 3 channels x 1 B  (PatchOptimizer.cpp:512-524, Image.h:104-113)
REC_BYTES = 208          # sizeof(hpmvs_patch_t)
STRONG = ("city100", "city500_4k", "city24")     # fixed scene, sharded by octree sub-tree at N > 1


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="city100", choices=["city100", "plane8", "plane8x100k", "city500_4k", "city24", "tiny", "fountain11"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="patches in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--inflight", type=int, default=12, help="steps in flight (each on its own stream): >1 lets the next step's CTAs start on SMs the previous step has drained")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-ncc", action="store_true", help="skip the stand-alone scoring kernel leg (profiling runs)")
    ap.add_argument("--sim-world", type=int, default=0, help="tuning aid on ONE GPU: run rank 0's shard of an N-way sub-tree split (not a bench line)")
    ap.add_argument("--subtrees-per-rank", type=int, default=16, help="the sub-tree split continues until there are max(100, this x ranks) sub-trees")
    return ap.parse_args(argv)


def workload_scene(name: str, rank: int = 0):
    """Synthetic scene + seed points for a workload.  Plane workloads: every rank gets the same images and its own seed points
    (weak scaling); city workloads: one fixed scene for all ranks (strong scaling, sharded later)."""
    from hpmvs_b200 import synth
    if name == "plane8":
        return synth.plane_scene(n_views=8, width=1280, height=960, focal=1200.0, radius=8.0, arc_deg=40.0,
                                 n_seeds=10000, extent=2.5, seed=2, point_seed=2 + 1000 * rank, tex_size=1024), \
            "8-view 1280x960 synthetic plane, 10k seed patches (BASELINE.json configs[1])"
    if name == "plane8x100k":
        return synth.plane_scene(n_views=8, width=1280, height=960, focal=1200.0, radius=8.0, arc_deg=40.0,
                                 n_seeds=100000, extent=2.5, seed=2, point_seed=2 + 1000 * rank, tex_size=1024), \
            "8-view 1280x960 synthetic plane, 100k seed patches (configs[1] scene, 10x the seeds: steady-state throughput)"
    if name == "tiny":
        return synth.plane_scene(n_views=8, width=640, height=480, focal=600.0, n_seeds=2000, seed=2, point_seed=2 + 1000 * rank,
                                 tex_size=512), "8-view 640x480 synthetic plane, 2k seed patches (smoke size)"
    try:
        import torch
        synth.USE_GPU_RENDERER = torch.cuda.is_available()    # 100 x 1080p ray casts: seconds on the GPU, minutes in numpy
    except ImportError:
        pass
    if name == "city24":
        return synth.city_scene(n_views=24, width=640, height=360, focal=500.0, n_seeds=6000, seed=4, name="city24v"), \
            "24-view 640x360 synthetic city block, 6k seed points (test size of the city family)"
    if name == "city500_4k":
        return synth.city_scene(n_views=500, width=3840, height=2160, focal=3000.0, n_seeds=400000, seed=5, name="city500v"), \
            "500-view 4K synthetic city block, 400k seed points (BASELINE.json configs[4])"
    return synth.city_scene(n_views=100, width=1920, height=1080, n_seeds=100000, seed=4, name="city100v"), \
        "100-view 1080p synthetic city block, 100k seed points -> all valid seed patches per step (BASELINE.json configs[3])"


def cached_scene(name: str, rank: int):
    """Scenes are deterministic; cache the rendered one under /tmp so repeated runs on one box skip the ray caster."""
    import pickle
    if name in STRONG:
        rank = 0                                  # one scene for every rank
    path = f"/tmp/hpmvs_b200_scene_{name}_{rank}.pkl"
    if os.path.exists(path):
        try:
            with open(path, "rb") as fh:
                return pickle.load(fh)
        except Exception:
            pass
    sc = workload_scene(name, rank)
    if name != "city500_4k":                      # 12 GB of level-0 pixels: never pickled
        try:
            tmp = f"{path}.{os.getpid()}.tmp"
            with open(tmp, "wb") as fh:
                pickle.dump(sc, fh, protocol=4)
            os.replace(tmp, path)
        except Exception:
            pass
    return sc


def bench_config(workload: str, desc: str, n_step: int, n_views: int, world: int):
    """The `config` object - identical for --impl ours and --impl reference (the driver compares them)."""
    strong = workload in STRONG
    return {"workload": desc, "views": int(n_views), "patches_per_step": int(n_step),
            "patches_per_step_scope": "whole job (fixed batch, split over the GPUs)" if strong else "per GPU (every GPU its own seed draw)",
            "start_angles": "host libm (start mode 1): results bit-identical to the reference build on this machine",
            "l2": "GPU arm: 256 MiB device-to-device copy on the step's stream before every step (> 126 MB L2); CPU arm: n/a"}


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons with nvidia-smi while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.samples.append(line.strip())
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        time.sleep(0.15)
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_arm(scene):
    """The CPU implementation that is timed beside the engine: the reference's own PatchOptimizer (oracle/_ref/libhpmvs_ref.so,
    compiled from /root/reference/src/hpmvs/*.cpp where they lie) when that prebuilt library is present -> kind "reference";
    otherwise the oracle restatement linked to the reference's BOBYQA -> kind "port".  Returns (scene object, kind, description)."""
    import oracle
    from oracle import ref
    if ref.available() and not os.environ.get("HPMVS_BENCH_FORCE_PORT"):
        return ref.RefScene.from_synth(scene), "reference", \
            "the reference's own mo3d::PatchOptimizer::optimize (its sources compiled with -O3 + OpenMP as CMakeLists.txt:4-7; " \
            "Eigen/glog stood in for by oracle/shim), one optimizer per thread as src/main.cpp:123-125, OpenMP over patches as Scene.cpp:114"
    return oracle.OracleScene.from_synth(scene), "port", \
        "oracle restatement of PatchOptimizer + the reference's real nlopt BOBYQA, OpenMP over patches as Scene.cpp:114"


def cpu_reference_rate(scene, seeds_or, sample: int, threads: int, repeats: int = 1):
    """CPU arm on host cores: optimized patches/s on the first `sample` seeds (bounded CPU work)."""
    orc, kind, how = cpu_arm(scene)
    batch = seeds_or[:sample]
    best = None
    ok = 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        out = orc.optimize_batch(batch, nthreads=threads)
        dt = time.perf_counter() - t0
        ok = int((out["status"] == 0).sum())
        best = dt if best is None else min(best, dt)
    return ok / best, best, ok, len(batch), kind, how


def to_oracle(p_en):
    import oracle
    out = np.zeros(len(p_en), oracle.PATCH_DTYPE)
    for f in ("center", "normal", "scale", "nimages"):
        out[f] = p_en[f]
    out["images"][:, :p_en["images"].shape[1]] = p_en["images"]
    return out


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path with all host threads (see cpu_arm), on the same batch and
    with the same `config` as the GPU arm.  Under torchrun rank 0 alone works."""
    import oracle
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    scene, desc = cached_scene(args.workload, 0)
    seeds, valid = oracle.OracleScene.from_synth(scene).seed_patches(scene.points, scene.meas_offsets, scene.meas_cam)
    seeds = seeds[valid]
    orc, kind, how = cpu_arm(scene)
    threads = host_threads()
    n = len(seeds)
    sample = int(args.cpu_sample) if args.cpu_sample else n       # the whole batch of the step, like the GPU arm
    batch = seeds[:sample]
    for _ in range(min(args.warmup, 3)):
        orc.optimize_batch(batch[: max(64, sample // 8)], nthreads=threads)
    t_tot, ok_tot = 0.0, 0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        out = orc.optimize_batch(batch, nthreads=threads)
        t_tot += time.perf_counter() - t0
        ok_tot += int((out["status"] == 0).sum())
    val = ok_tot / t_tot
    line = {"impl": "reference", "metric": "optimized patches/sec", "value": val, "unit": "patches/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.workload in STRONG else "weak", "vs_baseline": None, "dtype": "f32 samples / f64 optimizer",
            "data": "synthetic", "config": bench_config(args.workload, desc, n, len(scene.cameras), world),
            "cpu_baseline": {"value": val, "unit": "patches/s", "cores": threads, "kind": kind,
                             "sample": f"{'all' if sample == n else 'first ' + str(sample) + ' of'} {n} seed patches per step; {how}"},
            "e2e": {"value": val, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _OUT.write(json.dumps(line) + "\n"); _OUT.flush()


# ---- configs[2]: fountain-P11-style 11-view scene, the FULL expand -> optimize -> filter loop, 1 B200 vs the reference's own CLI --------
def fountain_scene():
    from hpmvs_b200 import synth
    try:
        import torch
        synth.USE_GPU_RENDERER = torch.cuda.is_available()
    except ImportError:
        pass
    n_pts = int(os.environ.get("HPMVS_FOUNTAIN_POINTS", "300"))
    sc = synth.plane_scene(n_views=11, width=3072, height=2048, focal=2800.0, radius=10.0, arc_deg=100.0, n_seeds=n_pts,
                           extent=3.0, seed=3, tex_size=2048, depth_noise=0.3, plane_half=9.0)
    return sc, f"fountain-P11-style 11-view 3072x2048 synthetic NVM, {n_pts} NVM points, full expand->optimize->filter loop (BASELINE.json configs[2])"


def fountain_quality(xyz, nz):
    return {"patches": int(len(xyz)), "rms_distance_to_true_plane": float(np.sqrt(np.mean(xyz[:, 2] ** 2))) if len(xyz) else None,
            "mean_abs_normal_z": float(np.mean(np.abs(nz))) if len(xyz) else None}


def run_fountain(args):
    """One step = one whole run of the loop on the scene.  GPU arm: hpmvs_pipeline_run (C++ level-synchronous driver behind the C ABI:
    every batch crosses the ABI with host buffers, so value == e2e); reference arm: the reference's own command line
    (oracle/_ref/hpmvs_ref = /root/reference/src compiled where it lies) on all host cores.  The schedulers differ (batches per level vs
    priority queues per sub-tree), so the clouds are compared by count and accuracy, and the rate is patches of the FINAL cloud per second."""
    import glob
    import tempfile
    sc, desc = fountain_scene()
    config = {"workload": desc, "views": 11, "patches_per_step": "the final cloud of one run of the loop",
              "rate": "patches in the final cloud / wall time of the loop (scene upload and seeding outside, as the reference logs it: main.cpp:142-185)"}
    steps = max(1, min(args.steps, 3))
    if args.impl == "reference":
        from oracle import ref
        tmp = tempfile.mkdtemp(prefix="hpmvs_f11_")
        nvm = os.path.join(tmp, "scene.nvm")
        import hpmvs_b200 as hp
        hp.synth.write_nvm(sc, nvm)
        threads = host_threads()
        secs, fin = [], None
        for k in range(steps):
            t = time.perf_counter()
            r = ref.run_cli(nvm, os.path.join(tmp, f"ref{k}"), threads=threads)
            secs.append(time.perf_counter() - t)
            assert r.returncode == 0, r.stderr[-2000:]
            # the CLI reports its own loop time ("Done within X seconds", main.cpp:183-185); fall back to the wall time of the process
            L = open(os.path.join(tmp, f"ref{k}", "patches-final.ply")).read().split("\n")
            n = int([l for l in L[:20] if l.startswith("element vertex")][0].split()[2])
            h = L.index("end_header") + 1
            fin = np.array([[float(x) for x in l.split()[:10]] for l in L[h:h + n]], np.float64).reshape(n, 10)
            import re
            m = re.search(r"Done within ([0-9.eE+-]+) seconds", r.stderr + r.stdout)
            if m:
                secs[-1] = float(m.group(1))
        levels = sorted(int(os.path.basename(f)[8:-4]) // 10 for f in glob.glob(os.path.join(tmp, f"ref{steps - 1}", "patches-[0-9]*.ply")))
        val = len(fin) / (sum(secs) / len(secs))
        line = {"impl": "reference", "metric": "optimized patches/sec", "value": val, "unit": "patches/s", "n_gpus": args.gpus, "steps": steps,
                "warmup": 0, "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 samples / f64 optimizer", "data": "synthetic", "config": config,
                "run": dict(tree_levels=levels, **fountain_quality(fin[:, :3], fin[:, 5])),
                "cpu_baseline": {"value": val, "unit": "patches/s", "cores": threads, "kind": "reference",
                                 "sample": "the whole scene through the reference's own CLI (src/main.cpp), its loop time as it logs it"},
                "e2e": {"value": val, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        _OUT.write(json.dumps(line) + "\n"); _OUT.flush()
        return
    import torch
    import hpmvs_b200 as hp
    from hpmvs_b200 import gather, pipeline
    torch.cuda.set_device(0)
    eng = hp.Engine.from_synth(sc)
    eng.set_start_mode(True)
    seeds, valid = hp.seed_patches(eng.options, eng.cameras, sc.points, sc.meas_offsets, sc.meas_cam)
    seeds = np.ascontiguousarray(seeds[valid])
    # root cube and start level as Scene::initPatches forms them (Scene.cpp:167-199, DynOctTree::add(e, width), doctree.h:379-394)
    pre = eng.optimize(seeds)
    okp = (pre["status"] == 0) & ~(np.linalg.norm(pre["center"][:, :3] - seeds["center"][:, :3], axis=1) > pre["scale"] * 2)
    origin, width = gather.root_cube(np.ascontiguousarray(pre[okp]))
    sc3 = np.maximum(pre["scale"][okp], np.float32(width / 1024.0))
    first_level = int(np.ceil(np.log2(width / (2.0 * sc3.astype(np.float64)))).min())
    secs, out, stats = [], None, None
    eng.counters(reset=True)
    for k in range(steps + 1):                       # the first run warms the engine up (graph + slot creation)
        t = time.perf_counter()
        out, stats = pipeline.run_native(eng, seeds, origin=origin, root_width=width, start_level=first_level, final_level=20)
        if k:
            secs.append(time.perf_counter() - t)
        else:
            eng.counters(reset=True)
    cnt = eng.counters(reset=True)
    t_mean = sum(secs) / len(secs)
    val = len(out) / t_mean
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6650.0
    alg = (TEX_BYTES * cnt.textures + 2 * REC_BYTES * cnt.patches) / len(secs)
    line = {"metric": "optimized patches/sec", "value": val, "unit": "patches/s", "n_gpus": 1, "steps": steps, "warmup": 1,
            "ms_per_step": 1e3 * t_mean, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 samples / f64 optimizer", "data": "synthetic", "config": config,
            "run": dict(tree_levels=[lv for lv, _, _ in stats.per_level], per_level=stats.per_level, optimize_calls=int(stats.optimized_calls),
                        optimized_ok=int(stats.optimized_ok), seconds_optimize=stats.seconds_optimize, seconds_accept=stats.seconds_accept,
                        optimize_calls_per_second=stats.optimized_calls / t_mean,
                        **fountain_quality(out["center"][:, :3].astype(np.float64), out["normal"][:, 2])),
            "e2e": {"value": val, "unit": "patches/s", "h2d_bytes_per_step": int(cnt.patches * REC_BYTES / len(secs)),
                    "d2h_bytes_per_step": int(cnt.patches * REC_BYTES / len(secs)), "ms_per_step": 1e3 * t_mean,
                    "note": "every batch of the loop crosses the C ABI with host buffers: value == e2e"},
            "gpu_launches": int(cnt.kernel_launches),
            "roofline": {"bound": "hbm", "achieved": (alg / stats.seconds_optimize / 1e9) if stats.seconds_optimize else None,
                         "peak": peak, "unit": "GB/s", "frac": (alg / stats.seconds_optimize / 1e9 / peak) if stats.seconds_optimize else None,
                         "traffic": None, "kernel": "the fused-path kernels of all batches of one run (see the default workload)",
                         "note": "algorithmic bytes of one run / time spent inside hpmvs_optimize_batch during that run"},
            "cpu_baseline": None}
    _OUT.write(json.dumps(line) + "\n"); _OUT.flush()


_OUT = sys.stdout


def _claim_stdout():
    """stdout carries exactly ONE JSON line: keep a private handle on the real stdout and point fd 1 at stderr, so that whatever
    libraries print there (NCCL's version banner, OpenMP notices ...) cannot end up next to it."""
    real = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    return real


def shard_for_rank(workload, seeds_all, rank, world, per_rank=16):
    """Strong-scaling split of a fixed seed batch: the reference's sub-tree split of the octree over the seed points
    (hpmvs_shard_cells = getSubTrees, src/main.cpp:50-96), sub-trees dealt to the ranks.  Returns (my seeds, info)."""
    from hpmvs_b200 import gather
    if world == 1 or workload not in STRONG:
        return seeds_all, None
    origin, width = gather.root_cube(seeds_all)
    cell, rk, ncell = gather.shard_cells(seeds_all, origin, width, max(100, per_rank * world), world)
    counts = np.bincount(rk[rk >= 0], minlength=world)
    info = {"subtrees": int(ncell), "seeds_per_rank": counts.tolist(), "origin": origin.tolist(), "root_width": width}
    return np.ascontiguousarray(seeds_all[rk == rank]), info


def main():
    args = parse_args()
    global _OUT
    _OUT = _claim_stdout()
    if args.workload == "fountain11":
        if int(os.environ.get("RANK", "0")) == 0:
            run_fountain(args)
        return
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import hpmvs_b200 as hp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: hpmvs_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own banner / debug output (NCCL_DEBUG) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    strong = args.workload in STRONG

    # ---- scene replicated on every GPU; seed batch: sharded by octree sub-tree (strong) or one draw per rank (weak) ----------------
    t_setup0 = time.perf_counter()
    opts = hp.Options.defaults()
    if args.workload == "city500_4k":
        # 500 x 4K level-0 images are 12 GB: rendered, uploaded and dropped one view at a time (never pickled, never all on the host)
        from hpmvs_b200 import synth
        synth.USE_GPU_RENDERER = True
        nv = int(os.environ.get("HPMVS_CITY500_VIEWS", "500")); ns = int(os.environ.get("HPMVS_CITY500_SEEDS", "400000"))
        desc = f"{nv}-view 4K synthetic city block, {ns // 1000}k seed points (BASELINE.json configs[4])"
        eng, scene = hp.Engine.from_stream(lambda sink: synth.city_scene(n_views=nv, width=3840, height=2160, focal=3000.0, n_seeds=ns,
                                                                         seed=5, name="city500v", image_sink=sink), opts, device=local_rank)
    else:
        scene, desc = cached_scene(args.workload, rank)
        eng = hp.Engine.from_synth(scene, opts, device=local_rank)
    t_upload = time.perf_counter() - t_setup0
    seeds_all, valid = hp.seed_patches(opts, eng.cameras, scene.points, scene.meas_offsets, scene.meas_cam)
    seeds_all = np.ascontiguousarray(seeds_all[valid])
    seeds, shard_info = shard_for_rank(args.workload, seeds_all, rank, args.sim_world or world, args.subtrees_per_rank)
    n = len(seeds)
    n_step_cfg = len(seeds_all)
    hbm_used = torch.cuda.mem_get_info()
    hbm_used_gb = (hbm_used[1] - hbm_used[0]) / 1e9

    F = max(1, int(args.inflight))
    streams = [torch.cuda.Stream() for _ in range(F)]     # real (non-legacy) streams: their handles are what the C ABI launches on
    stream = streams[0]
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    assert sptr != 0
    rec = torch.from_numpy(seeds.view(np.uint8).reshape(n, REC_BYTES))
    h_ins, h_outs = [], []
    for _ in range(F):
        hi = torch.empty((n, REC_BYTES), dtype=torch.uint8).pin_memory(); hi.copy_(rec)
        h_ins.append(hi); h_outs.append(torch.empty((n, REC_BYTES), dtype=torch.uint8).pin_memory())
    h_in, h_out = h_ins[0], h_outs[0]
    d_in = h_in.to("cuda", non_blocking=False)
    d_outs = [torch.empty_like(d_in) for _ in range(F)]
    # start mode 1 on device-resident records: the two start angles per patch from THIS machine's libm (as the reference forms them)
    d_start = torch.from_numpy(eng.start_parameters(seeds)).cuda()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    flush_src = torch.zeros_like(flush)

    def flush_l2():
        # a 256 MiB device-to-device copy (copy engine): evicts L2 without needing an SM, so it cannot be held up behind the
        # persistent CTAs of a step that is still draining
        flush.copy_(flush_src, non_blocking=True)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def launch(k):
        eng.optimize_device_start(n, d_in.data_ptr(), d_outs[k % F].data_ptr(), d_start.data_ptr(), streams[k % F].cuda_stream)

    # ---- warm-up ---------------------------------------------------------------------------------------------
    n_warm = max(3, args.warmup, F)       # at least one launch per in-flight slot: the engine creates a slot's buffers + graph on first use
    for k in range(n_warm):
        launch(k)
    torch.cuda.synchronize()
    eng.counters(reset=True)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)

    # ---- timed: K steps, kernel only, records resident in HBM, L2 flushed before every step.  With --inflight F > 1 step k runs
    # on stream k % F, so the CTAs of step k+1 start on the SMs that step k has already drained; the region is timed with one
    # event pair: start after all streams are idle, end after every stream has finished (a join stream waits for all of them) ------
    barrier()
    join = torch.cuda.Stream()
    ev_start = torch.cuda.Event(enable_timing=True); ev_end = torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev_start.record(join)
    for st in streams:
        st.wait_event(ev_start)
    for k in range(args.steps):
        st = streams[k % F]
        with torch.cuda.stream(st):
            flush_l2()
        launch(k)
    for st in streams:
        join.wait_stream(st)
    ev_end.record(join)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    t_dev = ev_start.elapsed_time(ev_end) / 1e3
    cnt = eng.counters(reset=True)
    ok_per_step = cnt.patches_ok / args.steps
    tex_per_step = cnt.textures / args.steps
    evals_per_step = cnt.evals / args.steps
    launches_timed = int(cnt.kernel_launches)

    # ---- timed: e2e through the C ABI with pinned host buffers in start mode 1 (host start angles + H2D + kernel + D2H per step), same
    # streams; at N > 1 every step ends with the gather of the ranks' accepted records onto rank 0 and the border de-duplication there,
    # overlapped with the next step's kernel (step k+1 is submitted before step k is collected) --------------------------------------
    from hpmvs_b200 import gather
    eng.set_start_mode(True)
    merged = [0, 0]
    gather_s = [0.0]

    dedup_cell = float(np.median(seeds_all["scale"])) * 2.0        # one patch per cell of the tree level the seeds are inserted at
    d_nkeep = torch.zeros(1, dtype=torch.int32, device="cuda")
    gstream = torch.cuda.Stream()                                  # the exchange runs beside the next step's kernels

    def collect(k):
        streams[k % F].synchronize()
        out_k = h_outs[k % F].numpy().view(hp.PATCH_DTYPE).reshape(n)
        okc = int((out_k["status"] == 0).sum())
        if dist is not None and strong:
            tg = time.perf_counter()
            with torch.cuda.stream(gstream):
                # the step's records go back to the device once (pinned, 4.6 MB per 22 k patches), the accepted ones are compacted
                # there, gathered onto rank 0 over NCCL send/recv and de-duplicated there by the engine's kernels
                d_res = h_outs[k % F].to("cuda", non_blocking=True)
                mine = d_res[d_res.view(torch.int32)[:, gather.STATUS_WORD] == 0]
                allr, owner = gather.gather_to_root_device(mine)
                if rank == 0:
                    keep = torch.empty(len(allr), dtype=torch.uint8, device="cuda")
                    eng.dedup_border_device(len(allr), allr.data_ptr(), owner.data_ptr(), shard_info["origin"], dedup_cell,
                                            keep.data_ptr(), d_nkeep.data_ptr(), gstream.cuda_stream)
                    merged[0], merged[1] = int(len(allr)), int(d_nkeep.item())       # the D2H read of the step's result
            gstream.synchronize()
            gather_s[0] += time.perf_counter() - tg
        return okc

    for k in range(F):
        eng.optimize_submit(n, h_ins[k % F].data_ptr(), h_outs[k % F].data_ptr(), streams[k % F].cuda_stream)
    for k in range(F):
        collect(k)            # warm-up of the exchange as well: NCCL sets up its send/recv connections on first use (hundreds of ms)
    gather_s[0] = 0.0
    barrier()
    t0 = time.perf_counter()
    ok_e2e = 0
    for k in range(args.steps):
        st = streams[k % F]
        with torch.cuda.stream(st):
            flush_l2()
        eng.optimize_submit(n, h_ins[k % F].data_ptr(), h_outs[k % F].data_ptr(), st.cuda_stream)
        if k >= F - 1:
            ok_e2e = collect(k - (F - 1))
    for k in range(max(0, args.steps - (F - 1)), args.steps):
        ok_e2e = collect(k)
    barrier()
    t_e2e = time.perf_counter() - t0
    gather_s[0] = 0.0 if dist is None else gather_s[0]
    eng.set_start_mode(False)
    # subtract nothing: the flush is a few hundred microseconds per step and stays inside (conservative)
    clocks = sampler.finish() if sampler else None
    out_np = h_out.numpy().view(hp.PATCH_DTYPE).reshape(n)
    status_hist = np.bincount(out_np["status"], minlength=14).tolist()
    for ho in h_outs[1:min(F, args.steps)]:
        assert np.array_equal(ho.numpy(), h_out.numpy()), "overlapping batches must return identical records"
    torch.cuda.set_stream(stream)

    # ---- the stand-alone scoring kernel (K1 = PatchOptimizer::setINCCs for a batch: the gather + NCC part of the path without the
    # optimizer around it), device-resident records, timed with CUDA events; reported as `roofline_ncc` ------------------------------
    ncc = None
    if not args.no_ncc and n > 0:
        reps = max(1, 200000 // max(n, 1))
        d_big = d_in.repeat(reps, 1).contiguous()
        nb = n * reps
        d_inc = torch.empty((nb, hp.MAX_VIEWS), dtype=torch.float32, device="cuda")
        for _ in range(2):
            eng.ncc_device(nb, d_big.data_ptr(), d_inc.data_ptr(), 0, False, sptr)
        torch.cuda.synchronize()
        eng.counters(reset=True)
        ncc_evs = []
        for _ in range(max(3, args.steps // 2)):
            flush_l2()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            eng.ncc_device(nb, d_big.data_ptr(), d_inc.data_ptr(), 0, False, sptr)
            e1.record(stream)
            ncc_evs.append((e0, e1))
        torch.cuda.synchronize()
        ncc_ms = [a.elapsed_time(b) for a, b in ncc_evs]
        ncc_cnt = eng.counters(reset=True)
        ncc = {"tex_per_launch": ncc_cnt.textures / len(ncc_ms), "launch_s": sum(ncc_ms) / len(ncc_ms) / 1e3, "nb": nb,
               "tma_staged_windows": bool(int(os.environ.get("HPMVS_NCC_TMA", "0")))}
        del d_big, d_inc

    # ---- reduce over ranks -------------------------------------------------------------------------------------
    stats = torch.tensor([t_dev, t_e2e, ok_per_step, tex_per_step, float(n), evals_per_step, float(ok_e2e), gather_s[0]],
                         dtype=torch.float64, device="cuda")
    per_rank_ms = None
    if dist is not None:
        allt = torch.zeros(world, dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allt, stats[:1].clone())
        per_rank_ms = [round(1e3 * float(v) / args.steps, 3) for v in allt.cpu().tolist()]
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        t_dev_max, t_e2e_max, gather_max = float(mx[0]), float(mx[1]), float(mx[7])
        ok_all, tex_all, n_all, evals_all, ok_e2e_all = float(sm[2]), float(sm[3]), float(sm[4]), float(sm[5]), float(sm[6])
    else:
        t_dev_max, t_e2e_max, gather_max = t_dev, t_e2e, 0.0
        ok_all, tex_all, n_all, evals_all, ok_e2e_all = ok_per_step, tex_per_step, float(n), evals_per_step, float(ok_e2e)

    if rank == 0:
        value = ok_all * args.steps / t_dev_max
        e2e_val = ok_e2e_all * args.steps / t_e2e_max
        # roofline of the dominant (only) kernel on THIS rank: algorithmic bytes per launch / mean launch time
        alg_bytes = TEX_BYTES * tex_per_step + 2 * REC_BYTES * n
        mean_launch_s = t_dev / args.steps
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak = 6650.0; peak_src = "fallback (B200_PROFILING.md)"
        achieved = alg_bytes / mean_launch_s / 1e9

        def traffic_of(key):
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(key)
                return int(tj["dram_bytes_read"]) + int(tj["dram_bytes_write"]) if tj else None
            except Exception:
                return None
        # CPU baseline on a bounded sample of the same batch (rank 0)
        cpu = None
        if not args.no_cpu and len(scene.images) > 0:              # (the streamed 500-view scene keeps no host images: no CPU leg)
            threads = host_threads()
            sample = args.cpu_sample or int(min(len(seeds_all), max(512, 3000 * threads)))
            rate, dt, okc, ns, kind, how = cpu_reference_rate(scene, to_oracle(seeds_all), sample, threads)
            cpu = {"value": rate, "unit": "patches/s", "cores": threads, "kind": kind,
                   "sample": f"first {ns} of {len(seeds_all)} seed patches of the same batch, {okc} optimized, {dt:.2f} s wall; {how}"}
        line = {"metric": "optimized patches/sec", "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * t_dev_max / args.steps, "higher_is_better": True,
                "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32 samples / f64 optimizer", "data": "synthetic",
                "config": bench_config(args.workload, desc, n_step_cfg, len(scene.cameras), world),
                "run": {"patches_per_step_this_rank": int(n), "patches_per_step_all_ranks": n_all, "optimized_per_step": ok_all,
                        "evals_per_step": evals_all, "textures_per_step": tex_all, "steps_in_flight": F,
                        "warmup_launches": n_warm,      # max(W, 3, steps in flight): every in-flight slot is created before the timed region
                        "status_histogram_rank0": status_hist, "too_many_views_rank0": status_hist[13],
                        "parallelism": (f"octree sub-trees dealt to {world} ranks (getSubTrees split), scene replicated, no data-path collective; "
                                        "final gather to rank 0 + border de-dup inside e2e") if strong else
                                       f"patch shards x{world}, scene replicated, no data-path collective",
                        "shards": shard_info, "device_ms_per_step_by_rank": per_rank_ms, "wall_s_timed_region": t_wall, "e2e_gather_dedup_ms_per_step": 1e3 * gather_max / args.steps,
                        "patches_gathered_kept": merged if dist is not None and strong else None,
                        "scene_upload_s": t_upload, "scene_upload_note": "render + upload + pyramid, streamed view by view" if args.workload == "city500_4k" else "upload + pyramid",
                        "hbm_used_gb": hbm_used_gb},
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": "patches/s", "h2d_bytes_per_step": int(n * (REC_BYTES + 16)), "d2h_bytes_per_step": int(n * REC_BYTES),
                        "ms_per_step": 1e3 * t_e2e_max / args.steps, "start_mode": 1},
                "gpu_launches": launches_timed,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic_of(args.workload), "peak_source": peak_src,
                             "kernel": "one step of the fused path = hp::wf_post_kernel<fill> + rounds x (hp::wf_advance_kernel, hp::wf_eval_kernel, hp::wf_post_kernel) in a CUDA-graph WHILE loop (batches >= 4000 patches), else hp::optimize_kernel",
                             "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": 1e3 * mean_launch_s,
                             "note": "algorithmic gather bytes (588 B/texture, no reuse credit) of one step / device time of a step (timed region / steps; steps overlap when steps_in_flight > 1); traffic = DRAM bytes summed over all kernels of one step (profiles/traffic.json); the footprints are L1/L2 resident, the scoring kernel is instruction-issue bound and the optimizer kernel latency bound, see DESIGN.md section 5"},
                "cpu_baseline": cpu}
        if ncc:
            nb = ncc["nb"]
            a = (TEX_BYTES * ncc["tex_per_launch"] + (REC_BYTES + 4 * hp.MAX_VIEWS) * nb) / ncc["launch_s"] / 1e9
            line["roofline_ncc"] = {"bound": "hbm", "kernel": "hp::ncc_kernel (setINCCs for a batch: projection, 7x7 bilinear RGB gather, normalise, NCC)",
                                    "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak,
                                    "patches_per_launch": int(nb), "textures_per_launch": ncc["tex_per_launch"], "launch_ms": 1e3 * ncc["launch_s"],
                                    "patch_scores_per_s": nb / ncc["launch_s"], "traffic": traffic_of(args.workload + "_ncc"),
                                    "tma_staged_windows": ncc["tma_staged_windows"],
                                    "note": "secondary figure: the scoring part of the path alone; not counted in value/e2e"}
        _OUT.write(json.dumps(line) + "\n"); _OUT.flush()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
