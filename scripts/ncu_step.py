"""One hot-path step under ncu: a warm-up call and ONE profiled call of the engine on a bench workload, with host-driven rounds
(HPMVS_WF=2) so that every kernel of the step is an ordinary launch.  python scripts/ncu_step.py <workload> [ncc]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, hpmvs_b200 as hp
wl = sys.argv[1]; ncc = len(sys.argv) > 2 and sys.argv[2] == "ncc"
sc, _ = bench.cached_scene(wl, 0)
eng = hp.Engine.from_synth(sc)
seeds, valid = hp.seed_patches(eng.options, eng.cameras, sc.points, sc.meas_offsets, sc.meas_cam)
seeds = np.ascontiguousarray(seeds[valid]); n = len(seeds)
d_in = torch.from_numpy(seeds.view(np.uint8).reshape(n, -1).copy()).cuda()
if ncc:
    d_inc = torch.empty((n, hp.MAX_VIEWS), dtype=torch.float32, device="cuda")
    for _ in range(2):
        eng.ncc_device(n, d_in.data_ptr(), d_inc.data_ptr(), 0, False)
    torch.cuda.synchronize()
    print("ncc", n, eng.last_kernel_ms(), "ms", eng.counters().textures)
else:
    d_out = torch.zeros_like(d_in)
    eng.optimize_device(n, d_in.data_ptr(), d_out.data_ptr())
    torch.cuda.synchronize()
    print("step", n, eng.last_kernel_ms(), "ms", eng.counters().textures)
