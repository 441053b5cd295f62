"""Do the per-round kernel chains of two independent wavefront launches overlap on the GPU?  Two engines, two host threads, two streams,
each optimising half of the city100 batch; compared with one engine optimising the whole batch.  HPMVS_WF=2 (host-driven rounds,
plain launches) vs HPMVS_WF=1 (CUDA-graph WHILE loops)."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("HPMVS_WF_PARTS", "1"); os.environ.setdefault("HPMVS_WF_SPLIT", "0")
import numpy as np, torch
import bench, hpmvs_b200 as hp
wl = sys.argv[1] if len(sys.argv) > 1 else "city100"
nthreads = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sc, _ = bench.cached_scene(wl, 0)
engs = [hp.Engine.from_synth(sc) for _ in range(nthreads)]
seeds, valid = hp.seed_patches(engs[0].options, engs[0].cameras, sc.points, sc.meas_offsets, sc.meas_cam)
seeds = np.ascontiguousarray(seeds[valid]); n = len(seeds)
d_in = torch.from_numpy(seeds.view(np.uint8).reshape(n, -1).copy()).cuda(); d_out = torch.zeros_like(d_in)
streams = [torch.cuda.Stream() for _ in range(nthreads)]
def whole():
    engs[0].optimize_device(n, d_in.data_ptr(), d_out.data_ptr(), streams[0].cuda_stream); streams[0].synchronize()
def part(k):
    a, b = n * k // nthreads, n * (k + 1) // nthreads
    engs[k].optimize_device(b - a, d_in[a:].data_ptr(), d_out[a:].data_ptr(), streams[k].cuda_stream); streams[k].synchronize()
for _ in range(2): whole()
t = time.perf_counter(); whole(); t_whole = time.perf_counter() - t
ref = d_out.cpu().numpy().copy()
for _ in range(2):
    th = [threading.Thread(target=part, args=(k,)) for k in range(nthreads)]; [x.start() for x in th]; [x.join() for x in th]
t = time.perf_counter()
th = [threading.Thread(target=part, args=(k,)) for k in range(nthreads)]; [x.start() for x in th]; [x.join() for x in th]
t_parts = time.perf_counter() - t
t = time.perf_counter(); part(0); t_one_part = time.perf_counter() - t
assert np.array_equal(ref, d_out.cpu().numpy())
print(f"mode HPMVS_WF={os.environ.get('HPMVS_WF')} {wl}: whole batch {1e3*t_whole:.1f} ms; {nthreads} parts concurrently {1e3*t_parts:.1f} ms; one part alone {1e3*t_one_part:.1f} ms")
