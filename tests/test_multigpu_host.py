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
    assert np.array_equal(a.points, a2.points) and not np.array_equal(a.points, b.points)
    assert all(np.array_equal(x, y) for x, y in zip(a.images, b.images))   # the scene is replicated, seeds are sharded
