"""Final multi-GPU exchange of the patch-optimisation path: variable-length gather of patch records over
torch.distributed (NCCL on GPUs, gloo in the CPU tests) and border de-duplication.

The reference is one process with shared memory; its analogue of cross-shard traffic is the border hand-off
between sub-trees (/root/reference/src/hpmvs/CellProcessor.cpp:487-540).  Here every rank optimises its own shard of
sub-trees and only the final patch sets are exchanged: all_gather of the counts, then one padded all_gather of the
208-byte records.  Patches from different ranks that fall into the same finest-level cell are then reduced to the
best-supported one, the rule of CellProcessor::filter (src/hpmvs/CellProcessor.cpp:43-82: most views wins).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .engine import PATCH_DTYPE

REC = PATCH_DTYPE.itemsize


def gather_patches(records: np.ndarray, device: Optional[torch.device] = None) -> Tuple[np.ndarray, np.ndarray]:
    """All ranks receive the concatenation of every rank's records (rank order) and the owning rank of each."""
    assert records.dtype == PATCH_DTYPE
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return records.copy(), np.zeros(len(records), np.int32)
    world = dist.get_world_size()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    n = torch.tensor([len(records)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    buf = torch.zeros((cap, REC), dtype=torch.uint8, device=device)
    if len(records):
        buf[:len(records)] = torch.from_numpy(np.ascontiguousarray(records).view(np.uint8).reshape(len(records), REC)).to(device)
    out = torch.empty((world, cap, REC), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(out.view(world * cap, REC), buf)
    host = out.cpu().numpy()
    parts = [host[r, :counts[r]].reshape(-1).view(PATCH_DTYPE) for r in range(world)]
    owner = np.concatenate([np.full(counts[r], r, np.int32) for r in range(world)]) if sum(counts) else np.zeros(0, np.int32)
    return np.concatenate(parts) if sum(counts) else np.zeros(0, PATCH_DTYPE), owner


def gather_to_root(records: np.ndarray, device: Optional[torch.device] = None, root: int = 0):
    """Unpadded variable-length gather of the ranks' records onto rank `root` only: an all_gather of the counts (8 bytes per rank),
    then one grouped point-to-point exchange - every other rank sends exactly its own records, the root posts one receive per rank
    straight into its slice of the output (NCCL send/recv over NVLink; gloo in the CPU tests).  Cost on the root = the bytes it
    receives; the other ranks pay for their own records only, so the exchange does not grow with the world size on them.
    Returns (records, owner) on the root and (None, None) elsewhere; `records` on the root is a host array."""
    assert records.dtype == PATCH_DTYPE
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return records.copy(), np.zeros(len(records), np.int32)
    world, rank = dist.get_world_size(), dist.get_rank()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    n = torch.tensor([len(records)], dtype=torch.int64, device=device)
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(counts, n)
    counts = counts.cpu().tolist()
    mine = torch.from_numpy(np.ascontiguousarray(records).view(np.uint8).reshape(len(records), REC)).to(device)
    if rank != root:
        if len(records):
            for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, mine, root)]):
                w.wait()
        return None, None
    total = int(sum(counts))
    out = torch.empty((total, REC), dtype=torch.uint8, device=device)
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    ops = []
    for r in range(world):
        if counts[r] == 0:
            continue
        if r == root:
            out[offs[r]:offs[r + 1]] = mine
        else:
            ops.append(dist.P2POp(dist.irecv, out[offs[r]:offs[r + 1]], r))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    host = out.cpu().numpy().reshape(-1).view(PATCH_DTYPE) if total else np.zeros(0, PATCH_DTYPE)
    owner = np.concatenate([np.full(counts[r], r, np.int32) for r in range(world)]) if total else np.zeros(0, np.int32)
    return host, owner


STATUS_WORD = PATCH_DTYPE.fields["status"][1] // 4      # int32 index of hpmvs_patch_t.status inside a record


def gather_to_root_device(d_records: torch.Tensor, root: int = 0):
    """Device-resident form of gather_to_root: `d_records` is a uint8 [n, 208] tensor on this rank's GPU (NCCL) - the accepted records
    of this rank; the root receives every rank's slice straight into one device tensor (no host staging, no padding).
    Returns (records uint8 [total, 208], owner int32 [total]) on the root, (None, None) elsewhere."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return d_records, torch.zeros(len(d_records), dtype=torch.int32, device=d_records.device)
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = d_records.device
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, torch.tensor([len(d_records)], dtype=torch.int64, device=dev))
    counts = counts.cpu().tolist()
    if rank != root:
        if len(d_records):
            for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, d_records.contiguous(), root)]):
                w.wait()
        return None, None
    total = int(sum(counts))
    out = torch.empty((total, REC), dtype=torch.uint8, device=dev)
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    ops = []
    for r in range(world):
        if counts[r] == 0:
            continue
        if r == root:
            out[offs[r]:offs[r + 1]] = d_records
        else:
            ops.append(dist.P2POp(dist.irecv, out[offs[r]:offs[r + 1]], r))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    owner = torch.repeat_interleave(torch.arange(world, dtype=torch.int32, device=dev), torch.tensor(counts, device=dev))
    return out, owner


def root_cube(patches: np.ndarray):
    """hpmvs_root_cube: the octree's root cube as Scene::initPatches forms it (Scene.cpp:186-193) -> (origin[3], width)."""
    import ctypes as C
    from . import engine as E
    L = E._lib()
    p = np.ascontiguousarray(patches)
    origin = np.zeros(3, np.float64); width = C.c_double()
    L.hpmvs_root_cube.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
    E._check(L.hpmvs_root_cube(len(p), p.ctypes.data, origin.ctypes.data, C.byref(width)))
    return origin, float(width.value)


def shard_cells(patches: np.ndarray, origin, width: float, min_subtrees: int, nranks: int):
    """hpmvs_shard_cells: the reference's sub-tree split (getSubTrees, src/main.cpp:50-96) of a patch set, sub-trees dealt to `nranks`
    ranks biggest-first to the least loaded rank -> (cell_of[n], rank_of[n], number of sub-trees)."""
    import ctypes as C
    from . import engine as E
    L = E._lib()
    p = np.ascontiguousarray(patches)
    org = np.ascontiguousarray(origin, np.float64)
    cell = np.zeros(max(len(p), 1), np.int32); rk = np.zeros(max(len(p), 1), np.int32)
    L.hpmvs_shard_cells.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    m = L.hpmvs_shard_cells(len(p), p.ctypes.data, org.ctypes.data, float(width), int(min_subtrees), int(nranks), cell.ctypes.data, rk.ctypes.data)
    E._check(m)
    return cell[:len(p)], rk[:len(p)], int(m)


def dedup_border(records: np.ndarray, owner: np.ndarray, cell: float, origin=None) -> np.ndarray:
    """hpmvs_dedup_border (host C++ behind the C ABI, hpmvs_b200/csrc/host_pipeline.cpp); dedup_border_numpy below is its twin.
    Cells are counted from `origin` (the octree's low corner; None = world origin)."""
    import ctypes as C
    from . import engine as E
    L = E._lib()
    rec = np.ascontiguousarray(records); own = np.ascontiguousarray(owner, np.int32)
    keep = np.zeros(max(len(rec), 1), np.int32)
    org = None if origin is None else np.ascontiguousarray(origin, np.float64)
    L.hpmvs_dedup_border.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    m = L.hpmvs_dedup_border(len(rec), rec.ctypes.data, own.ctypes.data, None if org is None else org.ctypes.data, float(cell), keep.ctypes.data)
    E._check(m)
    return keep[:m].astype(np.int64)


def dedup_border_numpy(records: np.ndarray, owner: np.ndarray, cell: float, origin=None) -> np.ndarray:
    """Keep one patch per cubic cell of edge `cell` when ranks disagree: most views first (CellProcessor::filter),
    then the lower final score, then the lower rank.  Patches of a single rank are never merged (that is the
    host scheduler's job inside a sub-tree).  Returns the indices kept, ascending."""
    ok = records["status"] == 0
    idx = np.nonzero(ok)[0]
    if len(idx) == 0:
        return idx
    org = np.zeros(3) if origin is None else np.asarray(origin, np.float64)
    key = np.floor((records["center"][idx, :3].astype(np.float64) - org) / cell).astype(np.int64)
    order = np.lexsort((owner[idx], records["score"][idx], -records["nimages"][idx], key[:, 2], key[:, 1], key[:, 0]))
    k = key[order]
    first = np.ones(len(order), bool)
    first[1:] = (k[1:] != k[:-1]).any(1)
    cell_id = np.cumsum(first) - 1
    winner_owner = owner[idx][order][first][cell_id]
    keep = first | (owner[idx][order] == winner_owner)       # same-rank neighbours of the winner stay
    return np.sort(idx[order][keep])
