"""Synchronous hpmvs_optimize_batch latency by batch size, persistent (HPMVS_WF=0) vs automatic choice: python scripts/sync_call_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, hpmvs_b200 as hp
sc, _ = bench.cached_scene("plane8x100k", 0)
res = {}
for mode in ("0", "-1", "1"):
    os.environ["HPMVS_WF"] = mode
    eng = hp.Engine.from_synth(sc)
    seeds, valid = hp.seed_patches(eng.options, eng.cameras, sc.points, sc.meas_offsets, sc.meas_cam)
    seeds = np.ascontiguousarray(seeds[valid])
    for n in (300, 1000, 3000, 8000, 16000, 24000, 40000, 64000, 96000):
        b = np.ascontiguousarray(seeds[:n])
        eng.optimize(b); eng.optimize(b)
        t = time.perf_counter(); eng.optimize(b); eng.optimize(b); dt = (time.perf_counter() - t) / 2
        res[(mode, n)] = dt
        print(f"HPMVS_WF={mode:>2s} n={n:6d}: {1e3*dt:8.2f} ms per synchronous call ({n/dt/1e3:7.1f} k patches/s)", flush=True)
    del eng
