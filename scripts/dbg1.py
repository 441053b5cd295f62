import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, hashlib
import hpmvs_b200 as hp, oracle
from helpers import compare_outputs, to_engine
g = np.load("tests/golden/plane4_small.npz")
kw = eval(str(g["scene_kwargs"]))
sc = hp.synth.plane_scene(**kw)
print("hash equal", hashlib.sha256(np.stack(sc.images).tobytes()).hexdigest() == str(g["scene_sha256"]))
eng = hp.Engine.from_synth(sc)
seeds = np.zeros(len(g["seeds_scale"]), hp.PATCH_DTYPE)
seeds["center"] = g["seeds_center"]; seeds["normal"] = g["seeds_normal"]; seeds["scale"] = g["seeds_scale"]
seeds["nimages"] = g["seeds_nimages"]; seeds["images"] = g["seeds_images"][:, :hp.MAX_VIEWS]
inc = eng.ncc(seeds, 0, False)
print("inccs equal", np.array_equal(inc[:, :8], g["inccs"]), np.abs(inc[:, :8]-g["inccs"]).max())
bad = np.nonzero((inc[:, :8] != g["inccs"]).any(1))[0]
print("bad rows", bad[:5], inc[bad[:2], :8], g["inccs"][bad[:2]], g["seeds_nimages"][bad[:2]])
got = eng.optimize(seeds)
print("status equal", np.array_equal(got["status"], g["status"]))
ok = g["status"] == 0
for f in ("center", "normal", "nimages", "color", "evals"):
    print(f, np.array_equal(got[f][ok], g[f][ok]))
# smoke scene
sc = hp.synth.plane_scene(n_views=8, width=640, height=480, focal=600.0, n_seeds=200, seed=11, tex_size=512)
eng = hp.Engine.from_synth(sc, device=0)
seeds, valid = hp.seed_patches(eng.options, eng.cameras, sc.points, sc.meas_offsets, sc.meas_cam)
seeds = np.ascontiguousarray(seeds[valid])
got = eng.optimize(seeds)
orc = oracle.OracleScene.from_synth(sc)
oracle.set_cr_asinf(True)
so = np.zeros(len(seeds), oracle.PATCH_DTYPE)
for f in ("center", "normal", "scale", "nimages"): so[f] = seeds[f]
so["images"][:, :hp.MAX_VIEWS] = seeds["images"]
ref = orc.optimize_batch(so, nthreads=4)
print(compare_outputs(ref, got))
import collections
print(collections.Counter(ref["status"].tolist()), collections.Counter(got["status"].tolist()))
d = np.nonzero(ref["status"] != got["status"])[0]
print("status diff idx", d[:10], ref["status"][d[:10]], got["status"][d[:10]], ref["evals"][d[:10]], got["evals"][d[:10]])
