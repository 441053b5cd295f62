"""End-to-end drop-in run: the REFERENCE'S OWN command line (oracle/_ref/hpmvs_ref, all host threads) against the same command line
linked with integration/PatchOptimizer_b200.cpp (oracle/_ref/hpmvs_ref_b200: scheduler, octree, depth tests, PLY writer are the
reference's; optimize() is the B200 engine, concurrent calls coalesced into batches) on one synthetic NVM scene.
Prints one JSON line.  usage: python scripts/dropin_bench.py [n_seeds] [gpu_host_threads] [subtrees]"""
import json, os, re, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hpmvs_b200 as hp
from oracle import ref

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 300
gpu_threads = int(sys.argv[2]) if len(sys.argv) > 2 else 512
subtrees = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
cores = len(os.sched_getaffinity(0))
sc = hp.synth.plane_scene(n_views=8, width=1280, height=960, focal=1200.0, radius=8.0, arc_deg=40.0, n_seeds=n_seeds, extent=2.5,
                          seed=2, tex_size=1024)
tmp = tempfile.mkdtemp(prefix="hpmvs_dropin_")
nvm = os.path.join(tmp, "scene.nvm")
hp.synth.write_nvm(sc, nvm)


def read_ply(path):
    L = open(path).read().split("\n")
    n = int([l for l in L[:20] if l.startswith("element vertex")][0].split()[2])
    h = L.index("end_header") + 1
    return np.array([[float(x) for x in l.split()[:6]] for l in L[h:h + n]], np.float64).reshape(n, 6)


def run(exe, out, threads, extra=()):
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), HPMVS_DROPIN_STATS="1", OMP_STACKSIZE="2M")
    t = time.perf_counter()
    r = subprocess.run([exe, f"--nvm={nvm}", f"--outdir={out}", *extra], env=env, capture_output=True, text=True)
    dt = time.perf_counter() - t
    if r.returncode != 0:
        raise SystemExit(f"{exe} failed ({r.returncode}): {r.stderr[-2000:]}")
    v = read_ply(os.path.join(out, "patches-final.ply"))
    m = re.search(r"(\d+) optimize\(\) calls in (\d+) batches \(largest (\d+)\)", r.stderr)
    return {"seconds": dt, "host_threads": threads, "final_patches": int(len(v)),
            "rms_distance_to_true_plane": float(np.sqrt(np.mean(v[:, 2] ** 2))) if len(v) else None,
            "mean_abs_normal_z": float(np.mean(np.abs(v[:, 5]))) if len(v) else None,
            "optimize_calls": int(m.group(1)) if m else None, "batches": int(m.group(2)) if m else None,
            "largest_batch": int(m.group(3)) if m else None}


res = {"workload": f"8-view 1280x960 synthetic plane, {n_seeds} NVM points, full reference pipeline (src/main.cpp) to its final level",
       "host_cores": cores}
res["reference_cli"] = run(ref.BIN_PATH, os.path.join(tmp, "cpu"), cores)
res["reference_cli_on_engine"] = run(ref.DROPIN_BIN_PATH, os.path.join(tmp, "gpu"), gpu_threads, [f"--subtrees={subtrees}"])
res["speedup_whole_run"] = res["reference_cli"]["seconds"] / res["reference_cli_on_engine"]["seconds"]
print(json.dumps(res))
