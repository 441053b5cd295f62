"""World-size-2 gloo tests (CPU) of the N>1 host logic: shard assignment, variable-length gather, border de-dup."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import hpmvs_b200 as hp
from hpmvs_b200 import gather

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _make(n, rank, rng):
    r = np.zeros(n, hp.PATCH_DTYPE)
    r["center"][:, :3] = rng.uniform(-1, 1, (n, 3)); r["center"][:, 3] = 1
    r["nimages"] = rng.integers(3, 9, n); r["score"] = rng.uniform(0, 0.1, n); r["scale"] = 0.01 * (rank + 1)
    return r


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    mine = _make(5 + 7 * rank, rank, rng)            # ragged: 5 and 12 records
    if rank == 1:
        mine["center"][0] = [0.5, 0.5, 0.5, 1]; mine["nimages"][0] = 8
        mine["status"][1] = 2                          # a rejected record travels but never survives de-dup
    else:
        mine["center"][0] = [0.5004, 0.5003, 0.5001, 1]; mine["nimages"][0] = 4    # same cell as rank 1's patch 0
    allr, owner = gather.gather_patches(mine)
    keep = gather.dedup_border(allr, owner, cell=0.01)
    empty, own0 = gather.gather_patches(np.zeros(0, hp.PATCH_DTYPE))   # empty shards on every rank
    q.put((rank, allr.tobytes(), owner.tolist(), keep.tolist(), len(empty)))
    dist.barrier()
    dist.destroy_process_group()


def _worker_root(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)                       # the same batch on every rank, sharded by octree sub-tree
    allp = _make(900, 0, rng)
    origin, width = gather.root_cube(allp)
    cell, rk, ncell = gather.shard_cells(allp, origin, width, 16, world)
    mine = np.ascontiguousarray(allp[rk == rank])
    got, owner = gather.gather_to_root(mine)
    e, _ = gather.gather_to_root(np.zeros(0, hp.PATCH_DTYPE))
    if rank == 0:
        q.put((rank, got.tobytes(), owner.tolist(), len(e), rk.tolist()))
    else:
        q.put((rank, b"" if got is None else b"x", [] if owner is None else [1], -1 if e is None else len(e), rk.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_unpadded_gather_to_root_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker_root, args=(r, 2, port, q)) for r in range(2)]
    for p in ps: p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps: p.join(60)
    assert all(p.exitcode == 0 for p in ps)
    (r0, b0, o0, e0, rk0), (r1, b1, o1, e1, rk1) = res
    assert rk0 == rk1                                          # every rank computes the same partition
    assert b1 == b"" and o1 == [] and e1 == -1 and e0 == 0     # only the root receives
    allp = _make(900, 0, np.random.default_rng(7))
    rk = np.asarray(rk0)
    want = np.concatenate([allp[rk == 0], allp[rk == 1]])
    assert b0 == want.tobytes() and o0 == [0] * int((rk == 0).sum()) + [1] * int((rk == 1).sum())


def test_shard_cells_is_the_reference_subtree_split():
    """hpmvs_shard_cells against a plain restatement of getSubTrees (src/main.cpp:50-96): split the root into its non-empty children,
    keep splitting the fullest sub-tree until there are enough; every patch in exactly one sub-tree; ranks balanced greedily."""
    rng = np.random.default_rng(3)
    p = _make(5000, 0, rng)
    p["center"][:2500, :3] *= 0.2                          # a dense clump: forces deep splits in one corner
    origin, width = gather.root_cube(p)
    c32 = p["center"][:, :3]
    assert np.isclose(width, float((c32.max(0) - c32.min(0)).max())) and np.allclose(origin + width / 2, (c32.max(0) + c32.min(0)) / 2, atol=1e-6)
    for want_trees, world in ((2, 1), (8, 2), (100, 4), (100, 8), (1, 3)):
        cell, rk, ncell = gather.shard_cells(p, origin, width, want_trees, world)
        assert (cell >= 0).all() and (rk >= 0).all() and rk.max() < world and cell.max() == ncell - 1
        # restatement: cells as (level, ix, iy, iz) with point lists
        rel = (p["center"][:, :3].astype(np.float64) - origin) / width
        q = np.minimum(np.floor(rel * (1 << 20)), (1 << 20) - 1).astype(np.int64)
        def split(level, idx):
            sh = 20 - (level + 1)
            code = ((q[idx, 0] >> sh) & 1) | (((q[idx, 1] >> sh) & 1) << 1) | (((q[idx, 2] >> sh) & 1) << 2)
            return [(level + 1, idx[code == c]) for c in range(8) if (code == c).any()]
        subs = [(0, np.arange(len(p)))] if want_trees < 2 else split(0, np.arange(len(p)))
        while want_trees >= 2 and len(subs) < want_trees:
            big = max(range(len(subs)), key=lambda i: (len(subs[i][1]), -i))
            if len(subs[big][1]) < 100:
                break
            subs = split(*subs[big]) + [s for i, s in enumerate(subs) if i != big]
        assert ncell == len(subs)
        for i, (_, idx) in enumerate(subs):
            assert (cell[idx] == i).all()
            assert len(set(rk[idx].tolist())) == 1           # a sub-tree is never split over ranks
        # sub-trees are dealt costliest-first by the sum of their patches' view counts: the greedy bound holds for that weight
        load = np.bincount(rk, weights=p["nimages"].astype(np.float64), minlength=world)
        assert load.max() - load.min() <= max(p["nimages"][s[1]].sum() for s in subs)
    # a patch outside the cube belongs to nobody
    far = p[:3].copy(); far["center"][0, 0] = 1e3
    cell, rk, _ = gather.shard_cells(far, origin, width, 8, 2)
    assert cell[0] == -1 and rk[0] == -1 and (cell[1:] >= 0).all()


def test_gather_and_dedup_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps: p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps: p.join(60)
    assert all(p.exitcode == 0 for p in ps)
    (r0, b0, o0, k0, e0), (r1, b1, o1, k1, e1) = res
    assert b0 == b1 and o0 == o1 and k0 == k1 and e0 == e1 == 0      # every rank ends with the same merged set
    allr = np.frombuffer(b0, hp.PATCH_DTYPE)
    assert len(allr) == 17 and o0 == [0] * 5 + [1] * 12
    assert 5 in k0 and 0 not in k0                                     # rank 1's 8-view patch beats rank 0's 4-view one
    assert 6 not in k0                                                 # rejected record dropped
    assert len(k0) == 15


def test_dedup_native_equals_numpy_twin():
    rng = np.random.default_rng(5)
    for trial in range(5):
        parts = [_make(int(rng.integers(0, 400)), r, rng) for r in range(4)]
        allr = np.concatenate(parts)
        owner = np.concatenate([np.full(len(p), r, np.int32) for r, p in enumerate(parts)])
        allr["status"][rng.random(len(allr)) < 0.1] = 2
        allr["nimages"][rng.random(len(allr)) < 0.5] = 5           # ties on the view count -> score, then rank decide
        cell = [0.5, 0.2, 0.05, 1.5, 0.01][trial]
        assert gather.dedup_border(allr, owner, cell).tolist() == gather.dedup_border_numpy(allr, owner, cell).tolist()


def test_dedup_single_rank_is_identity():
    rng = np.random.default_rng(1)
    r = _make(50, 0, rng)
    r["center"][10] = r["center"][11]
    keep = gather.dedup_border(r, np.zeros(50, np.int32), cell=0.5)
    assert keep.tolist() == list(range(50))


def test_bench_shards_are_disjoint_and_deterministic():
    sys.path.insert(0, ROOT)
    import bench
    a, _ = bench.workload_scene("tiny", 0)
    b, _ = bench.workload_scene("tiny", 1)
    a2, _ = bench.workload_scene("tiny", 0)
    # both arms print the same `config` object (the driver compares them)
    assert bench.bench_config("city100", "d", 44000, 100, 4) == bench.bench_config("city100", "d", 44000, 100, 4)
    assert np.array_equal(a.points, a2.points) and not np.array_equal(a.points, b.points)
    assert all(np.array_equal(x, y) for x, y in zip(a.images, b.images))   # the scene is replicated, seeds are sharded


@pytest.mark.gpu
def test_device_dedup_equals_host_dedup():
    """hpmvs_dedup_border_device (hash insert + atomic reductions on the GPU) against the host function on the same gathered records:
    same survivors, incl. ties on the view count (-> score, then rank decide), rejected records and a non-zero grid origin."""
    rng = np.random.default_rng(11)
    eng = hp.Engine()
    for trial in range(5):
        parts = [_make(int(rng.integers(1, 3000)), r, rng) for r in range(4)]
        allr = np.concatenate(parts)
        owner = np.concatenate([np.full(len(p), r, np.int32) for r, p in enumerate(parts)])
        allr["status"][rng.random(len(allr)) < 0.1] = 2
        allr["nimages"][rng.random(len(allr)) < 0.5] = 5
        allr["score"][rng.random(len(allr)) < 0.3] = 0.05              # ties on the score as well -> rank, then index
        cell = [0.5, 0.2, 0.05, 1.5, 0.01][trial]
        origin = np.array([-1.03, -0.97, -1.01]) if trial % 2 else None
        want = gather.dedup_border(allr, owner, cell, origin=origin)
        d_rec = torch.from_numpy(allr.view(np.uint8).reshape(len(allr), -1).copy()).cuda()
        d_own = torch.from_numpy(owner).cuda()
        keep = torch.zeros(len(allr), dtype=torch.uint8, device="cuda"); nk = torch.zeros(1, dtype=torch.int32, device="cuda")
        eng.dedup_border_device(len(allr), d_rec.data_ptr(), d_own.data_ptr(), origin, cell, keep.data_ptr(), nk.data_ptr())
        torch.cuda.synchronize()
        got = np.nonzero(keep.cpu().numpy())[0]
        assert got.tolist() == want.tolist(), trial
        assert int(nk.item()) == len(want)


def test_subtree_table_agrees_with_the_per_patch_assignment():
    """hpmvs_shard_subtrees (the split as a table of (level, cell key, rank), what hpmvs_pipeline_run looks centres up in) and
    hpmvs_shard_cells (the split as a per-patch assignment) are two views of one partition."""
    from hpmvs_b200 import pipeline
    rng = np.random.default_rng(4)
    p = _make(4000, 0, rng)
    p["center"][:1500, :3] *= 0.15
    origin, width = gather.root_cube(p)
    for want_trees, world in ((8, 2), (64, 4), (200, 8)):
        cell, rk, ncell = gather.shard_cells(p, origin, width, want_trees, world)
        lvl, key, srk = pipeline.shard_subtrees(p, origin, width, want_trees, world)
        assert len(lvl) == ncell and srk.max() < world
        assert (cell >= 0).all()                               # the cube is the bounding box of these centres: nobody is outside
        rel = np.clip((p["center"][:, :3].astype(np.float64) - origin) / width, 0.0, 1.0)
        q = np.minimum(np.floor(rel * (1 << 20)), (1 << 20) - 1).astype(np.int64)
        for s in range(ncell):
            k = q >> (20 - int(lvl[s]))
            inside = (k == key[s][None, :]).all(1)
            assert np.array_equal(np.nonzero(inside)[0], np.nonzero(cell == s)[0]), s
            assert (rk[inside] == srk[s]).all()
