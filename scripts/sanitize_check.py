"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): 120 patches through every kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hpmvs_b200 as hp
sc = hp.synth.plane_scene(n_views=6, width=320, height=240, focal=300.0, n_seeds=120, seed=3, tex_size=256)
eng = hp.Engine.from_synth(sc)
seeds, valid = hp.seed_patches(eng.options, eng.cameras, sc.points, sc.meas_offsets, sc.meas_cam)
seeds = np.ascontiguousarray(seeds[valid])
out = eng.optimize(seeds)
# the parked variant (pool state staged with cp.async.bulk), start mode 1 and two overlapping asynchronous submits
os.environ["HPMVS_PARKED"] = "1"
eng2 = hp.Engine.from_synth(sc)
eng2.set_start_mode(True)
out2 = eng2.optimize(seeds)
assert np.array_equal(out2["status"], out["status"])
import torch
raw = torch.from_numpy(seeds.view(np.uint8).reshape(len(seeds), -1).copy())
h_in = [raw.clone().pin_memory() for _ in range(2)]; h_out = [torch.zeros_like(raw).pin_memory() for _ in range(2)]
st = [torch.cuda.Stream(), torch.cuda.Stream()]
for k in range(2):
    eng2.optimize_submit(len(seeds), h_in[k].data_ptr(), h_out[k].data_ptr(), st[k].cuda_stream)
torch.cuda.synchronize()
assert np.array_equal(h_out[0].numpy(), h_out[1].numpy())
inc = eng.ncc(seeds, 0, True)
eng.depth_reset(); eng.depth_set(out); acc = eng.accept(out, 1.0)
print("ok", int((out["status"] == 0).sum()), "of", len(out), "accept sum", acc.sum(0).tolist())
