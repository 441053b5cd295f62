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
    def depth_unset(self, rec): self.e.depth_unset(np.ascontiguousarray(rec))
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
                 final_min_level: int = 9, cameras=None, shard_count: int = 1, shard_rank: int = 0, shard_level: int = 0,
                 minlevel: int = 0, level_cameras=None):
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
        # level support (Scene::getLevelSupport, Scene.cpp:335-344) needs the camera centres and k00 + k11
        lc = level_cameras if level_cameras is not None else cameras
        self.minlevel = minlevel
        self.cam_center = None
        if lc is not None:
            self.cam_center = np.stack([np.ctypeslib.as_array(c.center).astype(np.float32) for c in lc])
            self.cam_ksum = np.array([np.float32(c.k00) + np.float32(c.k11) for c in lc], np.float32)
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

    @staticmethod
    def _filter_pick(members) -> int:
        """CellProcessor::filter (CellProcessor.cpp:43-82): keep the patch with the smallest mean signed distance of the OTHERS' centres
        along its own normal; first minimum in arrival order.  f32, same operation order as host_pipeline.cpp."""
        f = np.float32
        best, bestd = 0, f(3.402823466e+38)
        for a, A in enumerate(members):
            n = A["normal"][:3].astype(np.float32).copy()
            z = f(f(n[0] * n[0] + n[1] * n[1]) + n[2] * n[2])
            if z > 0:
                n = n / np.sqrt(z, dtype=np.float32)
            dist = f(0)
            for b, B in enumerate(members):
                if a == b:
                    continue
                d = (B["center"][:3] - A["center"][:3]).astype(np.float32)
                dist = f(dist + f(f(n[0] * d[0] + n[1] * d[1]) + n[2] * d[2]))
            dist = f(dist / f(len(members) - 1))
            if dist < bestd:
                best, bestd = a, dist
        return best

    def _insert(self, level_cells: Dict[int, np.ndarray], rec: np.ndarray, width: float):
        """One patch per cell; patches that share a cell go through CellProcessor::filter's rule.  Returns (mask of the records that
        now live in the grid, list of resident patches that lost their cell - their depths must be subtracted)."""
        keys = _pack(_cell_keys(rec["center"], self.origin, width)).tolist()
        live = np.zeros(len(rec), bool)
        removed = []
        groups: Dict[int, List[int]] = {}
        for i, k in enumerate(keys):
            groups.setdefault(k, []).append(i)
        for k, g in groups.items():                                   # dicts keep first-appearance order, like the C++ `order` list
            old = level_cells.get(k)
            members = ([old] if old is not None else []) + [rec[i] for i in g]
            b = self._filter_pick(members) if len(members) > 1 else 0
            if old is not None:
                if b == 0:
                    continue
                removed.append(old)
                level_cells[k] = rec[g[b - 1]].copy()
                live[g[b - 1]] = True
            else:
                level_cells[k] = rec[g[b]].copy()
                live[g[b]] = True
        return live, removed

    def _level_support(self, rec: np.ndarray) -> np.ndarray:
        """Scene::getLevelSupport(patch, MINLEVEL): views whose rounded pyramid level for this patch is above MINLEVEL (f32 / f64 mix
        of Camera::getLevel, Camera.cpp:92-95, as host_pipeline.cpp::host_level)."""
        out = np.zeros(len(rec), np.int64)
        if self.cam_center is None:
            return out + 1
        f = np.float32
        for i in range(len(rec)):
            c = rec["center"][i].astype(np.float32)
            for k in range(int(rec["nimages"][i])):
                cam = int(rec["images"][i, k])
                d = (c - self.cam_center[cam]).astype(np.float32)
                p = (d * d).astype(np.float32)
                fz = np.sqrt(f(f(p[0] + p[2]) + f(p[1] + p[3])), dtype=np.float32)
                lvl = f(np.log2(np.float64(f(rec["scale"][i] * self.cam_ksum[cam])) / (2.0 * np.float64(fz))))
                r = int(np.floor(abs(float(lvl)) + 0.5)) * (1 if lvl >= 0 else -1)         # std::round: half away from zero
                out[i] += r > self.minlevel
        return out

    def _commit(self, cells, rec, width):
        """Insert the accepted records, subtract the depths of the residents they displace, set the depths of the new residents."""
        live, removed = self._insert(cells, rec, width)
        if removed:
            self.b.depth_unset(np.array(removed, dtype=self.b.dtype))
        fresh = rec[live]
        if len(fresh):
            self.b.depth_set(fresh)
        return fresh

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
        self._commit(cells, first, self.width(level))
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
                new = self._commit(cells, acc, w)
                n_ext += len(new)
                frontier = new
            patches = np.array(list(cells.values()), dtype=self.b.dtype) if cells else first[:0]
            if level >= self.final_level or len(patches) == 0:
                final.append(patches)
                self.stats.per_level.append((level, n_ext, 0))
                break
            # branch into the next level (CellProcessor::branch).  A patch without level support is exhausted (CellProcessor.cpp:222-225):
            # its cell keeps it and it is a final result at whatever level
            exhausted = self._level_support(patches) < 1
            if exhausted.any():
                final.append(patches[exhausted])
            pidx = np.nonzero(~exhausted)[0]
            parents = patches[pidx]
            # 4 candidates per patch, kept when they stay inside the parent's cell
            cand = self.b.expand(parents, np.full(len(parents), w, np.float32), 4) if len(parents) else patches[:0]
            parent = pidx[np.repeat(np.arange(len(parents)), 4)]
            pkeys = _pack(_cell_keys(patches["center"], self.origin, w))[parent] if len(parents) else np.zeros(0, np.int64)
            inside = _pack(_cell_keys(cand["center"], self.origin, w)) == pkeys
            cand, parent, pkeys = cand[inside], parent[inside], pkeys[inside]
            res = self._optimize(cand) if len(cand) else cand
            keep = (res["status"] == 0) & (_pack(_cell_keys(res["center"], self.origin, w)) == pkeys) if len(cand) else np.zeros(0, bool)
            children = res[keep]
            branched_parents = np.unique(parent[keep]) if len(cand) else np.zeros(0, np.int64)
            # cells that did not branch keep their patch as a final result - from PATCH_FINAL_MINLEVEL on (CellProcessor.cpp:266-269)
            stay = ~exhausted; stay[branched_parents] = False
            gone = ~exhausted
            if level >= self.final_min_level:
                final.append(patches[stay])
                gone &= ~stay
            # every other cell is split and its patch leaves the tree: Scene::setDepths(old, true) (CellProcessor.cpp:271-279)
            if gone.any():
                self.b.depth_unset(patches[gone])
            self.stats.per_level.append((level, n_ext, int(len(children))))
            cells = {}
            level += 1
            self._commit(cells, children, self.width(level))
        return np.concatenate(final) if final else seeds[:0]


# ----------------------------------------------------------------------------------------------------------------------
# The same driver in C++ behind the C ABI (hpmvs_pipeline_run, hpmvs_b200/csrc/host_pipeline.cpp): the product's host path.
# ----------------------------------------------------------------------------------------------------------------------
EXCHANGE_FN = None


def shard_subtrees(patches: np.ndarray, origin, root_width: float, min_subtrees: int, nranks: int):
    """hpmvs_shard_subtrees: the reference's sub-tree split (getSubTrees, src/main.cpp:50-96) of a patch set as a table
    (level[], key[n,3], rank[]) for run_native(subtrees=...)."""
    import ctypes as C
    from . import engine as E
    L = E._lib()
    p = np.ascontiguousarray(patches); org = np.ascontiguousarray(origin, np.float64)
    cap = 4096
    lvl = np.zeros(cap, np.int32); key = np.zeros((cap, 3), np.int64); rk = np.zeros(cap, np.int32)
    L.hpmvs_shard_subtrees.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    m = L.hpmvs_shard_subtrees(len(p), p.ctypes.data, org.ctypes.data, float(root_width), int(min_subtrees), int(nranks), cap,
                               lvl.ctypes.data, key.ctypes.data, rk.ctypes.data)
    E._check(m)
    return lvl[:m].copy(), key[:m].copy(), rk[:m].copy()


def run_native(engine, seeds: np.ndarray, origin, root_width: float, start_level: int, final_level: int, final_min_level: int = 9,
               max_rounds: int = 64, dedup_ref_pixel: bool = True, shard_count: int = 1, shard_rank: int = 0, shard_level: int = 0,
               minlevel: int = 0, subtrees=None, exchange=None):
    """Runs hpmvs_pipeline_run on `engine`; returns (final patch records, PipelineStats).  shard_count > 1: only the cells of tree
    level shard_level that are dealt to shard_rank are grown (one call per GPU; merge with gather.gather_patches + dedup_border)."""
    import ctypes as C
    from . import engine as E
    L = E._lib()

    class Params(C.Structure):
        _fields_ = [("origin", C.c_double * 3), ("root_width", C.c_double), ("start_level", C.c_int32), ("final_level", C.c_int32),
                    ("final_min_level", C.c_int32), ("max_rounds", C.c_int32), ("dedup_ref_pixel", C.c_int32), ("ncams", C.c_int32),
                    ("cams", C.POINTER(E.Camera)), ("shard_count", C.c_int32), ("shard_rank", C.c_int32), ("shard_level", C.c_int32),
                    ("minlevel", C.c_int32), ("nsub", C.c_int32), ("sub_level", C.c_void_p), ("sub_key", C.c_void_p), ("sub_rank", C.c_void_p),
                    ("exchange", C.c_void_p), ("exchange_user", C.c_void_p)]

    class Stats(C.Structure):
        _fields_ = [("optimize_calls", C.c_int64), ("optimized_ok", C.c_int64), ("seconds_optimize", C.c_double), ("seconds_accept", C.c_double),
                    ("nlevels", C.c_int32), ("level", C.c_int32 * 24), ("extended", C.c_int64 * 24), ("branched", C.c_int64 * 24),
                    ("exchanged", C.c_int64)]

    cams = (E.Camera * len(engine.cameras))(*engine.cameras)
    keep_alive = []
    nsub, p_lvl, p_key, p_rk = 0, None, None, None
    if subtrees is not None:
        lvl, key, rk = (np.ascontiguousarray(subtrees[0], np.int32), np.ascontiguousarray(subtrees[1], np.int64), np.ascontiguousarray(subtrees[2], np.int32))
        keep_alive += [lvl, key, rk]
        nsub, p_lvl, p_key, p_rk = len(lvl), lvl.ctypes.data, key.ctypes.data, rk.ctypes.data
    cb = None
    recv_bufs = []
    if exchange is not None:
        # exchange(records) -> all ranks' records concatenated in rank order (every rank calls it the same number of times)
        CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int))

        def _cb(user, n_send, send, recv, n_recv):
            try:
                mine = np.frombuffer((C.c_char * (n_send * E.PATCH_DTYPE.itemsize)).from_address(send), dtype=E.PATCH_DTYPE, count=n_send).copy() \
                    if n_send else np.zeros(0, E.PATCH_DTYPE)
                allr = np.ascontiguousarray(exchange(mine))
                recv_bufs.append(allr)                             # valid until the next call
                del recv_bufs[:-2]
                recv[0] = allr.ctypes.data
                n_recv[0] = len(allr)
                return 0
            except Exception:                                      # never let an exception cross the C boundary
                import traceback
                traceback.print_exc()
                return -1
        cb = CB(_cb)
        keep_alive.append(cb)
    prm = Params((C.c_double * 3)(*[float(v) for v in origin]), float(root_width), start_level, final_level, final_min_level, max_rounds,
                 1 if dedup_ref_pixel else 0, len(engine.cameras), cams, shard_count, shard_rank, shard_level, minlevel,
                 nsub, p_lvl, p_key, p_rk, C.cast(cb, C.c_void_p) if cb is not None else None, None)
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
