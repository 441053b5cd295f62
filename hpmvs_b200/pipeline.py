"""Level-synchronous host driver for the expand -> optimize -> filter loop around the engine.

The reference's scheduler (CellProcessor + DynOctTree, /root/reference/src/hpmvs/CellProcessor.cpp:369-420) walks one
priority queue per octree sub-tree and commits every accepted patch before it creates the next candidate.  That
serial order cannot feed tens of thousands of patches to a GPU, and rebuilding the octree scheduler is out of scope
(SURVEY section 8); this driver is the batching stand-in the survey asks for (section 7, "wavefront"): per tree level it
collects the candidates of ALL cells, optimises them in one batch, runs the acceptance tests against a snapshot of
the depth maps and then commits the survivors in a deterministic order.  It keeps the reference's per-cell rules:

  * seeds:   reject if the optimised centre moved more than 2*scale (Scene.cpp:171), one patch per cell, the
             best-supported patch wins (CellProcessor::filter, CellProcessor.cpp:43-82)
  * extend:  6 candidates at one cell width, scale = width*0.9/2 (CellProcessor.cpp:98-119); accepted when
             width/2 < 2*scale < width, drift < 1.5*width, depthTests >= MIN_IMAGES, viewBlockTest < MIN_IMAGES,
             pixelFreeTests >= MIN_IMAGES-1 and > 75 % of the views (CellProcessor.cpp:129-142)
  * branch:  4 candidates at width/4, scale = width*0.45/2, kept when they stay inside the parent cell
             (CellProcessor.cpp:227-264); children live in cells of half the width

It is written against a small backend protocol so that the SAME host logic runs on the GPU engine and on the CPU
oracle - tests/test_pipeline.py requires identical patch sets from both.
"""
from __future__ import annotations

import dataclasses
import time
from typing import Dict, List, Tuple

import numpy as np

MIN_IMAGES = 3
DEPTH_TEST_FACTOR = 1.0


class EngineBackend:
    """The B200 engine (hpmvs_b200.Engine) behind the driver protocol."""

    def __init__(self, engine):
        from . import engine as E
        self.e = engine
        self.dtype = E.PATCH_DTYPE
        self._expand = E.expand_candidates
        self.e.depth_reset()

    def optimize(self, rec): return self.e.optimize(np.ascontiguousarray(rec))
    def accept(self, rec, margin): return self.e.accept(np.ascontiguousarray(rec), margin)
    def depth_set(self, rec): self.e.depth_set(np.ascontiguousarray(rec))
    def expand(self, parents, widths, mode): return self._expand(self.e.cameras, np.ascontiguousarray(parents), widths, mode)


@dataclasses.dataclass
class PipelineStats:
    optimized_calls: int = 0
    optimized_ok: int = 0
    seconds_optimize: float = 0.0
    seconds_accept: float = 0.0
    per_level: List[Tuple[int, int, int]] = dataclasses.field(default_factory=list)   # (level, extended, branched)


def _cell_keys(centers: np.ndarray, origin: np.ndarray, width: float) -> np.ndarray:
    return np.floor((centers[:, :3].astype(np.float64) - origin) / width).astype(np.int64)


def _pack(keys: np.ndarray) -> np.ndarray:
    k = keys + (1 << 20)
    return (k[:, 0] << 42) | (k[:, 1] << 21) | k[:, 2]


class WavefrontDriver:
    def __init__(self, backend, origin, root_width: float, start_level: int, final_level: int, max_rounds: int = 64,
                 final_min_level: int = 9, cameras=None, shard_count: int = 1, shard_rank: int = 0, shard_level: int = 0):
        # final_min_level = HpmvsOptions::PATCH_FINAL_MINLEVEL as the CLI sets it (src/main.cpp:44,234): a cell below that tree
        # level whose patch yields no child when it branches is split anyway and loses its patch (CellProcessor.cpp:266-283)
        self.final_min_level = final_min_level
        # multi-GPU: the cells of tree level shard_level are dealt to shard_count ranks, this driver keeps only shard_rank's (host_pipeline.cpp)
        self.shard_count, self.shard_rank, self.shard_level = shard_count, shard_rank, shard_level
        # cameras (hpmvs_camera_t list, optional): enables the per-round image-space de-duplication below
        self.P0 = None
        if cameras is not None:
            self.P0 = np.stack([np.ctypeslib.as_array(c.P)[0].astype(np.float64) for c in cameras])     # [ncams, 3, 4]
            self.k00 = np.array([c.k00 for c in cameras], np.float64)
        self.b = backend
        self.origin = np.asarray(origin, np.float64)
        self.root_width = float(root_width)
        self.start_level, self.final_level, self.max_rounds = start_level, final_level, max_rounds
        self.stats = PipelineStats()

    def width(self, level: int) -> float:
        return self.root_width / (1 << level)

    # -- helpers ------------------------------------------------------------------------------------------
    def _optimize(self, rec):
        t = time.perf_counter()
        out = self.b.optimize(rec)
        self.stats.seconds_optimize += time.perf_counter() - t
        self.stats.optimized_calls += len(rec)
        self.stats.optimized_ok += int((out["status"] == 0).sum())
        return out

    def _insert(self, level_cells: Dict[int, np.ndarray], rec: np.ndarray, width: float) -> np.ndarray:
        """One patch per cell; on a collision the patch with more views wins (filter), then the earlier one.
        Returns a mask of the records that now live in the grid."""
        keys = _pack(_cell_keys(rec["center"], self.origin, width))
        live = np.zeros(len(rec), bool)
        for i, k in enumerate(keys.tolist()):
            old = level_cells.get(k)
            if old is None or rec["nimages"][i] > old["nimages"]:
                level_cells[k] = rec[i].copy()
                live[i] = True
        return live

    def _mine(self, centers: np.ndarray) -> np.ndarray:
        if self.shard_count <= 1:
            return np.ones(len(centers), bool)
        k = _cell_keys(centers, self.origin, self.width(self.shard_level))
        cell = (k[:, 0] * 73856093) ^ (k[:, 1] * 19349663) ^ (k[:, 2] * 83492791)
        return np.mod(cell, self.shard_count) == self.shard_rank

    def _first_per_ref_pixel(self, rec: np.ndarray, width: float) -> np.ndarray:
        """The reference commits one patch at a time, so a patch that has just been accepted occupies its depth-map pixels and blocks
        the next candidate that lands on them (pixelFreeTests, Scene.cpp:587-611).  A batched round tests all candidates against the
        SAME snapshot; to keep the one-patch-per-image-cell behaviour the accepted candidates of a round are thinned to the first one
        per cell of their reference view, a cell being the image footprint of one tree cell (width * f / depth pixels)."""
        if self.P0 is None or len(rec) == 0:
            return np.ones(len(rec), bool)
        ref = rec["images"][:, 0].astype(np.int64)
        X = rec["center"].astype(np.float64)
        P = self.P0[ref]
        r = ((P[:, :, 0] * X[:, 0:1] + P[:, :, 1] * X[:, 1:2]) + P[:, :, 2] * X[:, 2:3]) + P[:, :, 3] * X[:, 3:4]   # fixed order (host_pipeline.cpp)
        z = np.maximum(r[:, 2], 1e-9)
        cell_px = np.maximum(width * self.k00[ref] / z, 1e-6)
        ku = np.floor(r[:, 0] / z / cell_px).astype(np.int64); kv = np.floor(r[:, 1] / z / cell_px).astype(np.int64)
        key = (ref << 44) | ((ku + (1 << 20)) << 22) | (kv + (1 << 20))
        _, firsts = np.unique(key, return_index=True)
        keep = np.zeros(len(rec), bool); keep[firsts] = True
        return keep

    # -- the loop -----------------------------------------------------------------------------------------
    def run(self, seeds: np.ndarray) -> np.ndarray:
        out = self._optimize(seeds)
        ok = out["status"] == 0
        moved = np.linalg.norm(out["center"][:, :3] - seeds["center"][:, :3], axis=1)
        ok &= ~(moved > out["scale"] * 2)                                    # Scene.cpp:171
        ok &= self._mine(out["center"])
        cells: Dict[int, np.ndarray] = {}
        level = self.start_level
        first = out[ok]
        self._insert(cells, first, self.width(level))
        self.b.depth_set(np.array(list(cells.values()), dtype=self.b.dtype) if cells else first[:0])
        final: List[np.ndarray] = []
        while True:
            w = self.width(level)
            frontier = np.array(list(cells.values()), dtype=self.b.dtype) if cells else first[:0]
            n_ext = 0
            for _ in range(self.max_rounds):                                  # extend until the wavefront dies out
                if len(frontier) == 0:
                    break
                cand = self.b.expand(frontier, np.full(len(frontier), w, np.float32), 6)
                parent = np.repeat(np.arange(len(frontier)), 6)
                keys = _pack(_cell_keys(cand["center"], self.origin, w))
                free = np.array([k not in cells for k in keys.tolist()], bool)
                # one candidate per free cell and round (first parent wins), like a cell being filled once
                _, firsts = np.unique(keys, return_index=True)
                uniq = np.zeros(len(keys), bool); uniq[firsts] = True
                sel = free & uniq
                if not sel.any():
                    break
                cand, parent = cand[sel], parent[sel]
                res = self._optimize(cand)
                good = res["status"] == 0
                good &= (res["scale"] * 2.0 < w) & (res["scale"] * 2.0 > w / 2.0)
                drift = np.linalg.norm(res["center"][:, :3] - frontier["center"][parent][:, :3], axis=1)
                good &= drift < w * 1.5
                # a patch that leaves the root cube is handed to the neighbouring sub-tree, and dropped when there is none
                # (CellProcessor.cpp:147-153, distributeBorderCell :533-540): the cloud never grows beyond the octree's root
                rel = (res["center"][:, :3].astype(np.float64) - self.origin) / self.root_width
                good &= ((rel >= 0.0) & (rel < 1.0)).all(axis=1)
                good &= self._mine(res["center"])
                t = time.perf_counter()
                counts = self.b.accept(res, DEPTH_TEST_FACTOR)
                self.stats.seconds_accept += time.perf_counter() - t
                nimg = np.maximum(res["nimages"], 1)
                good &= (counts[:, 0] >= MIN_IMAGES) & (counts[:, 1] < MIN_IMAGES)
                good &= (counts[:, 2] >= MIN_IMAGES - 1) & (counts[:, 2] * 1.0 / nimg > 0.75)
                acc = res[good]
                acc = acc[self._first_per_ref_pixel(acc, w)]
                if len(acc) == 0:
                    break
                live = self._insert(cells, acc, w)
                new = acc[live]
                self.b.depth_set(new)
                n_ext += len(new)
                frontier = new
            patches = np.array(list(cells.values()), dtype=self.b.dtype) if cells else first[:0]
            if level >= self.final_level or len(patches) == 0:
                final.append(patches)
                self.stats.per_level.append((level, n_ext, 0))
                break
            # branch into the next level: 4 candidates per patch, kept when they stay inside the parent's cell
            cand = self.b.expand(patches, np.full(len(patches), w, np.float32), 4)
            parent = np.repeat(np.arange(len(patches)), 4)
            pkeys = _pack(_cell_keys(patches["center"], self.origin, w))[parent]
            inside = _pack(_cell_keys(cand["center"], self.origin, w)) == pkeys
            cand, parent, pkeys = cand[inside], parent[inside], pkeys[inside]
            res = self._optimize(cand) if len(cand) else cand
            keep = (res["status"] == 0) & (_pack(_cell_keys(res["center"], self.origin, w)) == pkeys) if len(cand) else np.zeros(0, bool)
            children = res[keep]
            branched_parents = np.unique(parent[keep]) if len(cand) else np.zeros(0, np.int64)
            # cells that did not branch keep their patch as a final result - from PATCH_FINAL_MINLEVEL on (CellProcessor.cpp:266-269)
            stay = np.ones(len(patches), bool); stay[branched_parents] = False
            if level >= self.final_min_level:
                final.append(patches[stay])
            self.stats.per_level.append((level, n_ext, int(len(children))))
            cells = {}
            level += 1
            self._insert(cells, children, self.width(level))
            self.b.depth_set(np.array(list(cells.values()), dtype=self.b.dtype) if cells else children[:0])
        return np.concatenate(final) if final else seeds[:0]


# ----------------------------------------------------------------------------------------------------------------------
# The same driver in C++ behind the C ABI (hpmvs_pipeline_run, hpmvs_b200/csrc/host_pipeline.cpp): the product's host path.
# ----------------------------------------------------------------------------------------------------------------------
def run_native(engine, seeds: np.ndarray, origin, root_width: float, start_level: int, final_level: int, final_min_level: int = 9,
               max_rounds: int = 64, dedup_ref_pixel: bool = True, shard_count: int = 1, shard_rank: int = 0, shard_level: int = 0):
    """Runs hpmvs_pipeline_run on `engine`; returns (final patch records, PipelineStats).  shard_count > 1: only the cells of tree
    level shard_level that are dealt to shard_rank are grown (one call per GPU; merge with gather.gather_patches + dedup_border)."""
    import ctypes as C
    from . import engine as E
    L = E._lib()

    class Params(C.Structure):
        _fields_ = [("origin", C.c_double * 3), ("root_width", C.c_double), ("start_level", C.c_int32), ("final_level", C.c_int32),
                    ("final_min_level", C.c_int32), ("max_rounds", C.c_int32), ("dedup_ref_pixel", C.c_int32), ("ncams", C.c_int32),
                    ("cams", C.POINTER(E.Camera)), ("shard_count", C.c_int32), ("shard_rank", C.c_int32), ("shard_level", C.c_int32)]

    class Stats(C.Structure):
        _fields_ = [("optimize_calls", C.c_int64), ("optimized_ok", C.c_int64), ("seconds_optimize", C.c_double), ("seconds_accept", C.c_double),
                    ("nlevels", C.c_int32), ("level", C.c_int32 * 24), ("extended", C.c_int64 * 24), ("branched", C.c_int64 * 24)]

    cams = (E.Camera * len(engine.cameras))(*engine.cameras)
    prm = Params((C.c_double * 3)(*[float(v) for v in origin]), float(root_width), start_level, final_level, final_min_level, max_rounds,
                 1 if dedup_ref_pixel else 0, len(engine.cameras), cams, shard_count, shard_rank, shard_level)
    st = Stats()
    s = np.ascontiguousarray(seeds)
    assert s.dtype == E.PATCH_DTYPE
    out_p = C.c_void_p(); n = C.c_int32()
    L.hpmvs_pipeline_run.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(Stats)]
    L.hpmvs_free.argtypes = [C.c_void_p]
    E._check(L.hpmvs_pipeline_run(engine._h, C.byref(prm), len(s), s.ctypes.data, C.byref(out_p), C.byref(n), C.byref(st)))
    try:
        buf = (C.c_char * (n.value * E.PATCH_DTYPE.itemsize)).from_address(out_p.value) if n.value else b""
        out = np.frombuffer(buf, dtype=E.PATCH_DTYPE, count=n.value).copy() if n.value else np.zeros(0, E.PATCH_DTYPE)
    finally:
        L.hpmvs_free(out_p)
    ps = PipelineStats(int(st.optimize_calls), int(st.optimized_ok), float(st.seconds_optimize), float(st.seconds_accept),
                       [(int(st.level[i]), int(st.extended[i]), int(st.branched[i])) for i in range(st.nlevels)])
    return out, ps
