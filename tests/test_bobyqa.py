"""The product's ask/tell BOBYQA (hpmvs_b200/csrc/bobyqa3.h, host build) against the REAL vendored nlopt
(oracle/_ref, built from /root/reference/thirdLibs/nlopt-2.4.2): every evaluated point must be bit-identical.

Known-answer anchor: nlopt's own 3-D bounded test objective Box-Betts
(thirdLibs/nlopt-2.4.2/test/testfuncs.c:65-89, xmin (1,10,1), minf 0, bounds :87-89)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INF = float("inf")
HP_LB = [-INF, -23.99999, -23.99999]     # PatchOptimizer.cpp:326-337
HP_UB = [INF, 23.99999, 23.99999]


@pytest.fixture(scope="module")
def bq3():
    so = os.path.join(ROOT, "tests", "cpp", "libbq3_host.so")
    src = os.path.join(ROOT, "tests", "cpp", "bobyqa_parity.cpp")
    hdr = os.path.join(ROOT, "hpmvs_b200", "csrc", "bobyqa3.h")
    oracle.lib()
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-o", so, src,
                        os.path.join(ROOT, "oracle", "libhpmvs_oracle.so"), "-Wl,-rpath," + os.path.join(ROOT, "oracle")], check=True)
    L = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    L.bq3_run_testfunc.argtypes = [C.c_int, dp, dp, dp, C.c_double, C.c_int, dp, dp, dp, dp, C.c_int, C.POINTER(C.c_int)]
    L.bq3_last_rescues.restype = C.c_int

    def run(fid, x0, lb, ub, xtol=1e-7, maxeval=1000):
        x0 = np.asarray(x0, float); lb = np.asarray(lb, float); ub = np.asarray(ub, float)
        xo = np.zeros(3); fo = C.c_double(); ne = C.c_int(); cap = maxeval + 8
        tx = np.zeros((cap, 3)); tf = np.zeros(cap)
        p = lambda a: a.ctypes.data_as(dp)
        r = L.bq3_run_testfunc(fid, p(x0), p(lb), p(ub), xtol, maxeval, p(xo), C.byref(fo), p(tx), p(tf), cap, C.byref(ne))
        n = min(ne.value, cap)
        return r, xo, fo.value, tx[:n].copy(), tf[:n].copy(), L.bq3_last_rescues()

    L.bq3_run_testfunc_tile.argtypes = L.bq3_run_testfunc.argtypes + [C.c_int, C.c_int, C.POINTER(C.c_int)]

    def run_tile(fid, x0, lb, ub, xtol=1e-7, maxeval=1000, lane=0, split=True):
        """Same optimiser on lane `lane` of a 32-state tile (bq3::StateTile), advanced phase by phase like the wavefront kernels."""
        x0 = np.asarray(x0, float); lb = np.asarray(lb, float); ub = np.asarray(ub, float)
        xo = np.zeros(3); fo = C.c_double(); ne = C.c_int(); ny = C.c_int(); cap = maxeval + 8
        tx = np.zeros((cap, 3)); tf = np.zeros(cap)
        p = lambda a: a.ctypes.data_as(dp)
        r = L.bq3_run_testfunc_tile(fid, p(x0), p(lb), p(ub), xtol, maxeval, p(xo), C.byref(fo), p(tx), p(tf), cap, C.byref(ne),
                                    lane, 1 if split else 0, C.byref(ny))
        n = min(ne.value, cap)
        return r, xo, fo.value, tx[:n].copy(), tf[:n].copy(), L.bq3_last_rescues(), ny.value
    run.tile = run_tile
    run.tile_bytes = L.bq3_tile_bytes()
    return run


def _same(a, b):
    r1, x1, f1, tx1, tf1 = a
    r2, x2, f2, tx2, tf2 = b[:5]
    return r1 == r2 and len(tx1) == len(tx2) and np.array_equal(tx1, tx2) and np.array_equal(x1, x2) and f1 == f2


def test_boxbetts_known_answer(bq3):
    # the reference library itself reaches nlopt's documented minimiser ...
    r, x, f, tx, tf = oracle.bobyqa_testfunc(0, [1.0, 10.5, 1.1], [0.9, 9, 0.9], [1.2, 11.2, 1.2], 1e-7, 1000)
    assert r in (1, 4)
    assert np.allclose(x, [1.0, 10.0, 1.0], atol=1e-5) and f < 1e-12
    # ... and ours visits exactly the same points
    assert _same((r, x, f, tx, tf), bq3(0, [1.0, 10.5, 1.1], [0.9, 9, 0.9], [1.2, 11.2, 1.2]))


@pytest.mark.parametrize("fid,x0,lb,ub,maxeval", [
    (1, [-1.2, 1.0, 0.5], [-5, -5, -5], [5, 5, 5], 1000),
    (1, [0, 0, 0], [-INF] * 3, [INF] * 3, 1000),
    (2, [0, 0, 0], HP_LB, HP_UB, 1000), (3, [0, 0, 0], HP_LB, HP_UB, 1000), (4, [0, 0, 0], HP_LB, HP_UB, 1000),
    (5, [0, 0, 0], HP_LB, HP_UB, 1000), (6, [0, 0, 0], HP_LB, HP_UB, 1000), (7, [0, 0, 0], HP_LB, HP_UB, 1000),
    (2, [0, 20.0, -23.0], HP_LB, HP_UB, 1000),              # start within 0.75*gap of a bound (options.c:703-710)
    (3, [0, -23.99999, 23.99999], HP_LB, HP_UB, 1000),      # start ON the bounds
    (5, [0, 10, 10], HP_LB, HP_UB, 30),                     # MAXEVAL_REACHED inside the main loop
    (2, [0, 0, 0], HP_LB, HP_UB, 5),                        # MAXEVAL_REACHED inside PRELIM
    (8, [0, 0, 0], HP_LB, HP_UB, 1000),
])
def test_trace_identical(bq3, fid, x0, lb, ub, maxeval):
    ref = oracle.bobyqa_testfunc(fid, x0, lb, ub, 1e-7, maxeval)
    assert _same(ref, bq3(fid, x0, lb, ub, 1e-7, maxeval))


def test_fuzz_including_rescue_and_roundoff(bq3):
    """Seeded families of noisy / quantised / badly scaled objectives: exercises ROUNDOFF_LIMITED, MAXEVAL and RESCUE."""
    rng = np.random.default_rng(0)
    results, rescues = {}, 0
    for fid in list(range(100, 400)) + list(range(5000, 5400)):
        x0 = [rng.normal(0, 0.3), rng.uniform(-23.9, 23.9), rng.uniform(-23.9, 23.9)] if fid % 2 else [0, 0, 0]
        ref = oracle.bobyqa_testfunc(fid, x0, HP_LB, HP_UB, 1e-7, 1000)
        got = bq3(fid, x0, HP_LB, HP_UB, 1e-7, 1000)
        assert _same(ref, got), fid
        results[ref[0]] = results.get(ref[0], 0) + 1
        rescues += got[5] > 0
    assert results.get(-4, 0) > 0 and results.get(5, 0) > 0 and results.get(1, 0) > 0 and results.get(4, 0) > 0
    assert rescues > 10      # the RESCUE branch really ran, and still matched


def test_tiled_state_and_phase_split_visit_the_same_points(bq3):
    """The wavefront kernels keep 32 optimiser states interleaved in one tile (bq3::StateTile: every member a 256-byte cell, lane l's
    value at byte 8*l) and advance them phase by phase (A: absorb the objective value, T: trust-region step, B: shift / geometry step /
    Lagrange values), yielding in front of every heavy block of another phase.  Same points, same result, on every lane, and no
    byte outside the lane's own column is touched."""
    assert bq3.tile_bytes % 256 == 0 and bq3.tile_bytes // 256 >= 200
    rng = np.random.default_rng(1)
    yields = 0
    for k, fid in enumerate([0, 1, 2, 3, 5, 8] + list(range(100, 160)) + list(range(5000, 5060))):
        if fid == 0:
            x0, lb, ub = [1.0, 10.5, 1.1], [0.9, 9, 0.9], [1.2, 11.2, 1.2]
        else:
            x0 = [rng.normal(0, 0.3), rng.uniform(-23.9, 23.9), rng.uniform(-23.9, 23.9)] if fid % 2 else [0, 0, 0]
            lb, ub = HP_LB, HP_UB
        want = bq3(fid, x0, lb, ub, 1e-7, 1000)
        for split in (False, True):
            got = bq3.tile(fid, x0, lb, ub, 1e-7, 1000, lane=(7 * k) % 32, split=split)
            assert got[0] != -1000, "a neighbouring lane's bytes were written"
            assert _same(want[:5], got), (fid, split)
            assert got[5] == want[5]
            yields += got[6]
    assert yields > 1000
