"""Scratch GPU parity check with verbose statistics (the pytest -m gpu tests are the real gate)."""
import sys, os, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import hpmvs_b200 as hp, oracle
from helpers import small_plane, to_engine, compare_outputs

sc, orc, seeds = small_plane()
eng = hp.Engine.from_synth(sc)
# pyramid parity
for cam in (0, 5):
    for lvl in range(6):
        a = orc.image(cam, lvl); b = eng.download_image(cam, lvl)
        print("pyramid cam", cam, "lvl", lvl, a.shape, "equal" if np.array_equal(a, b) else "DIFF %d" % (a != b).sum())
pe = to_engine(seeds)
# K1 parity
g = eng.ncc(pe, 0, False)
bad = 0
for i in range(len(seeds)):
    r = orc.set_inccs(seeds[i:i+1], 0, 0)
    if not np.array_equal(r, g[i, :len(r)]):
        bad += 1
        if bad < 4: print("ncc mismatch", i, r, g[i, :len(r)])
print("K1 setINCCs mismatches:", bad, "of", len(seeds))
oracle.set_cr_asinf(os.environ.get('ORC_CR','1')=='1')
t = time.time(); ref = orc.optimize_batch(seeds, nthreads=8); t_cpu = time.time() - t
t = time.time(); got = eng.optimize(pe); t_gpu = time.time() - t
print("cpu %.3fs gpu %.3fs kernel %.3f ms" % (t_cpu, t_gpu, eng.last_kernel_ms()))
print("oracle status", collections.Counter(ref["status"].tolist()))
print("engine status", collections.Counter(got["status"].tolist()))
print(compare_outputs(ref, got))
c = eng.counters()
print("counters", c.patches, c.patches_ok, c.evals, c.textures, c.kernel_launches)
ok = (ref["status"] == 0) & (got["status"] == 0)
idx = np.nonzero(ok)[0]
nb = 0
for i in idx:
    if not (np.array_equal(ref["center"][i], got["center"][i]) and ref["evals"][i] == got["evals"][i]):
        nb += 1
        if nb <= 5:
            print("diff patch", i, "evals", ref["evals"][i], got["evals"][i], "center", ref["center"][i], got["center"][i], "score", ref["last_val"][i], got["score"][i])
