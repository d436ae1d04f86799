#!/usr/bin/env python
"""BASELINE configs[4]: ONE network of 1 048 576 columns x 16 384 inputs, k = 20 972 active columns, 32 cells,
column-sharded over the 8 ranks of a torchrun job (131 072 columns = 16 GiB of float64 permanence per GPU).  No
oracle run is possible at this size (128 GiB of permanence); checked here: every rank computes the same
replicated state (active columns, segments per cell, winner lists), the invariants of a timestep hold, no capacity
status is raised.  (tools/cfg5_lockstep.py checks the same k against the oracle on one GPU with a short input;
tests/test_multi.py checks sharded == unsharded.)

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/cfg5_run.py [steps] [C] [I]
"""
import json
import os
import sys
import time
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist

    import bithtm_b200 as bithtm
    from bithtm_b200.projections import DenseProjection

    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    C = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
    I = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    c, k = 32, round(C * 0.02)
    patterns = 3
    g = np.random.default_rng(0)
    base = g.random((patterns, I)) < 0.2
    xs = base[np.arange(steps) % patterns] ^ (g.random((steps, I)) < 0.05)
    rows = C // world
    perm = torch.empty(rows, I, dtype=torch.float64, device="cuda")
    # drawn in 64 row blocks with per-block seeds: the same matrix for every world size (a short input lets ONE GPU
    # hold the whole network, so the 8-GPU run can be compared with the single-GPU one line by line)
    blk = C // 64
    for b in range(64):
        lo = b * blk - rank * rows
        if lo < 0 or lo >= rows:
            continue
        gen = torch.Generator(device="cuda")
        gen.manual_seed(7700 + b)
        perm[lo:lo + blk] = torch.randn(blk, I, dtype=torch.float64, device="cuda", generator=gen) * 0.1
    np.random.seed(0)
    t0 = time.time()
    sp = bithtm.SpatialPooler(I, C, k, proximal_projection=DenseProjection(I, C, permanence=perm))
    htm = bithtm.HierarchicalTemporalMemory(I, C, c, k, spatial_pooler=sp, rng_sync="lazy",
                                            column_shard=True if world > 1 else None, max_segments=1 << 22,
                                            max_synapses_per_segment=128, fused="shard" if world > 1 else "grid",
                                            ring_len=steps)
    del perm
    sp.proximal_projection._host_permanence = None
    torch.cuda.empty_cache()
    eng = htm.engine
    if rank == 0:
        print(f"engine: {eng.arena_bytes / 2**30:.1f} GiB arena per rank, skip table {eng.ctx.skip_polys} x {eng.ctx.skip_gran}, "
              f"xch_ll {eng.ctx.xch_ll}, ring {eng.ctx.rng_ring_words} words, built in {time.time() - t0:.1f} s", flush=True)
    htm.temporal_memory._rng.before(eng)
    eng.load_ring(xs)
    g1 = eng.graph(1, learning=True)
    log = []
    for t in range(steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.launch_graph(g1, 1)
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        sc = eng.scalars()
        cur = (int(sc[0]) - 1) & 1
        cols = eng.buf["active_cols"][cur * k:(cur + 1) * k].cpu().numpy()
        nseg = eng.buf["cell_nseg"]
        sig = torch.tensor([zlib.crc32(cols.tobytes()), int(sc[2]), int(sc[4]), int(sc[5 + cur]), int(nseg.sum().item())],
                           dtype=torch.int64, device="cuda")
        sigs = [torch.empty_like(sig) for _ in range(world)]
        if world > 1:
            dist.all_gather(sigs, sig)
        else:
            sigs = [sig]
        agree = all(bool(torch.equal(s, sigs[0])) for s in sigs)
        ok_cols = bool(np.all(np.diff(cols) > 0) and cols[0] >= 0 and cols[-1] < C)
        entry = dict(step=t, ms=round(float(ms), 3), status=int(sc[12]), active_columns_crc32=int(sig[0]), segments=int(sc[2]), matching=int(sc[4]),
                     winners=int(sc[5 + cur]), learning_rows=int(sc[8]), growing_rows_local=int(sc[23]),
                     segments_per_cell_sum=int(sig[4]), ranks_agree=agree, active_columns_sorted_distinct=ok_cols)
        log.append(entry)
        if rank == 0:
            print(entry, flush=True)
        eng.check_status()
        assert agree and ok_cols and int(sig[4]) == int(sc[2])
    if rank == 0:
        print(json.dumps({"workload": f"cfg5: {C} columns x {I} inputs, k={k}, {c} cells, ONE network over {world} GPU(s)",
                          "per_rank_permanence_GiB": rows * I * 8 / 2**30, "steps": log}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
