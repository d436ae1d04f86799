"""Column sharding of the spatial pooler across ranks.

Each rank owns a contiguous range of columns (permanence rows, mask rows, duty
cycles).  Per timestep there is ONE exchange: every rank contributes its best
``min(k, C/world)`` candidates as (float64 key, int32 global column) pairs in
ascending column order; after an all-gather in rank order the concatenation is in
ascending column order too, so the canonical rule (larger key first, ties -> lower
column) can be applied to it on every rank identically.  The temporal memory is
replicated: every rank computes it from the same active-column list and the same
MT19937 stream, so no second exchange is needed and results are bit-identical to the
single-GPU run by construction.

Segment sharding of the temporal memory (``segment_shard``) distributes the synapse
rows -- the HBM-heavy part -- by segment id and adds ONE more exchange per timestep:
``gather_records`` all-gathers each rank's fixed-size record of matching segments and
lowest recyclable segment ids (see ``include/bithtm_b200.h``).
"""

from __future__ import annotations


def gather_candidates(keys, cols, group=None):
    """All-gather (keys[k_loc] float64, cols[k_loc] int32) in rank order.  Works on
    CUDA tensors (NCCL) and CPU tensors (gloo)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    out_keys = torch.empty(world * keys.numel(), dtype=keys.dtype, device=keys.device)
    out_cols = torch.empty(world * cols.numel(), dtype=cols.dtype, device=cols.device)
    if keys.is_cuda:
        dist.all_gather_into_tensor(out_keys, keys, group=group)
        dist.all_gather_into_tensor(out_cols, cols, group=group)
    else:  # gloo has no all_gather_into_tensor on every build: use list form
        ks = [torch.empty_like(keys) for _ in range(world)]
        cs = [torch.empty_like(cols) for _ in range(world)]
        dist.all_gather(ks, keys, group=group)
        dist.all_gather(cs, cols, group=group)
        out_keys, out_cols = torch.cat(ks), torch.cat(cs)
    return out_keys, out_cols


def gather_columns(local, group=None):
    """All-gather a per-column array (this rank's slice) into the full [C] array."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if local.is_cuda:
        out = torch.empty(world * local.numel(), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local.contiguous(), group=group)
    return torch.cat(parts)


def gather_records(send, recv, group=None):
    """All-gather the per-rank exchange records (int32, equal sizes) in rank order into
    ``recv``.  CUDA tensors (NCCL) or CPU tensors (gloo)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    assert recv.numel() == world * send.numel()
    if send.is_cuda:
        dist.all_gather_into_tensor(recv, send, group=group)
    else:
        parts = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(parts, send, group=group)
        recv.copy_(torch.cat(parts))
    return recv


def merge_segment_parts(parts):
    """Per-rank (ids, count, cells, perm) of the held segments -> arrays ordered by segment id."""
    import numpy as np

    ids = np.concatenate([p[0] for p in parts])
    order = np.argsort(ids, kind="stable")
    return tuple(np.concatenate([p[i] for p in parts])[order] for i in range(1, 4))


def map_exchange_regions(engine, group=None):
    """Allocate this rank's exchange region of the fused sharded step in memory the peer GPUs
    can store into over NVLink, exchange the mappings, and hand all bases to the engine.
    Symmetric memory (torch.distributed._symmetric_memory) when it works, else CUDA IPC
    handles passed through the process group."""
    import torch
    import torch.distributed as dist

    n = engine.exchange_region_ints()
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    assert world == engine.seg_world and rank == engine.seg_rank
    try:
        import torch.distributed._symmetric_memory as symm_mem

        t = symm_mem.empty(n, dtype=torch.int32, device=engine.device)
        t.zero_()
        torch.cuda.synchronize(engine.device)
        handle = symm_mem.rendezvous(t, group=group if group is not None else dist.group.WORLD)
        ptrs = [int(p) for p in handle.buffer_ptrs]
        keep = (t, handle)
        how = "symmetric memory"
    except Exception:
        t = torch.zeros(n, dtype=torch.int32, device=engine.device)
        torch.cuda.synchronize(engine.device)
        mine = t.untyped_storage()._share_cuda_()
        handles = [None] * world
        dist.all_gather_object(handles, mine, group=group)
        opened, ptrs = [], []
        for r in range(world):
            if r == rank:
                ptrs.append(t.data_ptr())
            else:
                st = torch.UntypedStorage._new_shared_cuda(*handles[r])
                opened.append(st)
                ptrs.append(st.data_ptr())
        keep = (t, opened)
        how = "CUDA IPC"
    dist.barrier(group=group)
    engine.set_exchange_regions(ptrs, keepalive=keep)
    return how


def held_segment_ids(n_segments, rank, world):
    """Ids < n_segments of the segments whose synapse rows rank `rank` of `world` stores: ids are
    dealt in blocks of 64, round-robin (csrc/common.cuh: seg_held / seg_row / seg_gid), ascending --
    which is also the order of the rank's local rows."""
    import numpy as np

    ids = np.arange(int(n_segments), dtype=np.int64)
    if world > 1:
        ids = ids[(ids >> 6) % world == rank]
    return ids


def local_row(segment_id, world):
    """Local synapse row of a held segment (csrc/common.cuh: seg_row)."""
    return segment_id if world <= 1 else ((((segment_id >> 6) // world) << 6) | (segment_id & 63))
