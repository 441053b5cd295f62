"""Mints tests/golden/plane4_small.npz from the CPU oracle (correctly-rounded asinf mode, so that the GPU engine
must reproduce it bit for bit).  Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import hashlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import hpmvs_b200 as hp  # noqa: E402
import oracle  # noqa: E402

KW = dict(n_views=6, width=640, height=480, focal=600.0, n_seeds=200, seed=5, tex_size=512, arc_deg=36.0)
sc = hp.synth.plane_scene(**KW)
orc = oracle.OracleScene.from_synth(sc)
seeds, valid = orc.seed_patches(sc.points, sc.meas_offsets, sc.meas_cam)
oracle.set_cr_asinf(True)
out = orc.optimize_batch(seeds[valid], nthreads=1)
inc = np.stack([np.pad(orc.set_inccs(seeds[valid][i:i + 1], 0, 0), (0, 8))[:8] for i in range(int(valid.sum()))])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "plane4_small.npz"), scene_kwargs=str(KW), scene_sha256=hashlib.sha256(np.stack(sc.images).tobytes()).hexdigest(),
                    seeds_center=seeds[valid]["center"], seeds_normal=seeds[valid]["normal"], seeds_scale=seeds[valid]["scale"],
                    seeds_nimages=seeds[valid]["nimages"], seeds_images=seeds[valid]["images"], inccs=inc,
                    **{f: out[f] for f in ("status", "center", "normal", "nimages", "images", "color", "evals", "textures", "last_val")})
print("golden:", int(valid.sum()), "seeds,", int((out["status"] == 0).sum()), "optimized")
