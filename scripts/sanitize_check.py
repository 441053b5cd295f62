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
inc = eng.ncc(seeds, 0, True)
eng.depth_reset(); eng.depth_set(out); acc = eng.accept(out, 1.0)
print("ok", int((out["status"] == 0).sum()), "of", len(out), "accept sum", acc.sum(0).tolist())
