"""Per-round timeline of one wavefront launch (HPMVS_WF_LOG=1): python scripts/wf_roundlog.py <workload> <out.csv>"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("HPMVS_WF", "1"); os.environ["HPMVS_WF_LOG"] = "1"
import numpy as np
import bench, hpmvs_b200 as hp
from hpmvs_b200 import engine as E
wl, out = sys.argv[1], sys.argv[2]
sc, _ = bench.cached_scene(wl, 0)
eng = hp.Engine.from_synth(sc)
seeds, valid = hp.seed_patches(eng.options, eng.cameras, sc.points, sc.meas_offsets, sc.meas_cam)
seeds = np.ascontiguousarray(seeds[valid])
for _ in range(3):
    got = eng.optimize(seeds)
print(wl, len(seeds), "kernel ms", eng.last_kernel_ms(), "evals mean/max", got["evals"].mean(), got["evals"].max())
L = E._lib(); L.hpmvs_engine_dump_round_log.argtypes = [C.c_void_p, C.c_char_p]
n = L.hpmvs_engine_dump_round_log(eng._h, out.encode())
rows = np.loadtxt(out, delimiter=",", skiprows=1)
print("rounds", n, "total us", rows[-1, 1])
for a in range(0, n, max(1, n // 40)):
    b = min(n - 1, a + max(1, n // 40))
    print(f"round {a:5d}: {rows[a,1]:9.1f} us  evals {int(rows[a,2]):6d} posts {int(rows[a,3]):5d} dead {int(rows[a,4]):6d}   {(rows[b,1]-rows[a,1])/max(1,b-a):7.1f} us/round")
