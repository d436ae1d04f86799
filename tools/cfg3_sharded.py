#!/usr/bin/env python
"""cfg3 (65536 columns x 16384 inputs, 32 cells, k=1311) as ONE network sharded over the
ranks of a torchrun job: spatial pooler by column, temporal memory by segment id, two
exchanges per timestep.  Permanence rows are drawn on each device (performance run; parity
of the sharded path is tests/test_multi.py).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/cfg3_sharded.py [steps] [C] [I] [mode]

mode = "fused" (default: one cooperative kernel per shard and per 50 steps, exchanges in-kernel
over NVLink peer memory) or "nccl" (per-stage kernels, NCCL all-gathers issued by the host).
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def measure(steps=300, C=65536, I=16384, mode="fused", verbose=False):
    """Every rank of the (already initialised, when world > 1) process group calls this; returns
    the result dict on every rank (times are the max over ranks)."""
    import torch
    import torch.distributed as dist

    import bithtm_b200 as bithtm
    from bithtm_b200.projections import DenseProjection

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    c, k = 32, round(C * 0.02)
    patterns, chunk = 50, 50
    g = np.random.default_rng(0)
    base = g.random((patterns, I)) < 0.2
    xs = base[np.arange(2 * patterns) % patterns] ^ (g.random((2 * patterns, I)) < 0.05)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234 + rank)
    perm = torch.randn(C // world, I, dtype=torch.float64, device="cuda", generator=gen) * 0.1
    np.random.seed(0)
    sp = bithtm.SpatialPooler(I, C, k, proximal_projection=DenseProjection(I, C, permanence=perm))
    htm = bithtm.HierarchicalTemporalMemory(I, C, c, k, spatial_pooler=sp, rng_sync="lazy",
                                            column_shard=True if world > 1 else None,
                                            max_segments=1 << 21, max_synapses_per_segment=64,
                                            fused="shard" if mode == "fused" else "off",
                                            ring_len=2 * patterns if mode == "fused" else 0)
    del perm
    sp.proximal_projection._host_permanence = None
    torch.cuda.empty_cache()
    eng = htm.engine
    htm.temporal_memory._rng.before(eng)
    words = [eng.pack_input(x) for x in xs]
    graph = None
    if mode == "fused":
        eng.load_ring(xs)
        graph = eng.graph(chunk, learning=True)
    torch.cuda.synchronize()
    times = []
    for t0 in range(0, steps, chunk):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if graph is not None:
            eng.launch_graph(graph, chunk)
        else:
            for t in range(t0, min(t0 + chunk, steps)):
                htm.process(words[t % len(words)], return_state=False)
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b) / chunk], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        times.append(float(ms))
        if verbose and rank == 0:
            sc = eng.scalars()
            print(f"step {t0 + chunk}: {times[-1]:.3f} ms/step  S={sc[2]} M={sc[4]} L={sc[8]} status={sc[12]}",
                  flush=True)
    eng.check_status()
    phases = None
    if mode == "fused":
        names = ["overlap", "local top-k", "candidate record", "exchange 1", "unpack + global top-k", "SP learn + winner bits",
                 "lists + learning flags", "learning lists + draw 2", "stream chunks", "learn + post", "segment scan",
                 "record", "exchange 2", "merge", "draw 3 + jitter + predictions"]
        st = eng.buf["blk"][7 * 1024:7 * 1024 + 2 * (len(names) + 1)].cpu().numpy().view(np.uint64).astype(np.float64)
        phases = {n: round(float(v) / 1e3, 2) for n, v in zip(names, np.diff(st))}
        if verbose and rank == 0:
            for n, v in phases.items():
                print(f"  {n:32s} {v:8.2f} us")
            if os.environ.get("BH_TOPK_STAMPS"):  # library built with -DBH_TOPK_STAMPS
                tk = eng.buf["blk"][7 * 1024 + 80:7 * 1024 + 80 + 32].cpu().numpy().view(np.uint64)
                print("  topk_grid stages of the LAST call (ns):", np.diff(tk[:7].astype(np.int64)).tolist(),
                      "candidates", int(tk[15] & 0xffffffff), "range published", int(tk[15] >> 32))
    out = {"workload": f"cfg3 as ONE network: {C} columns x {I} inputs, k={k}, sharded over {world} GPU(s)",
           "n_gpus": world, "ms_per_step": times[-1], "steps_per_s": 1e3 / times[-1], "mode": mode,
           "exchanges_per_step": 2 if world > 1 else 0,
           "transport": getattr(htm, "exchange_transport", "local") if mode == "fused" else "NCCL all-gather",
           "steps": steps, "phase_us_rank0_last_step": phases}
    del htm, eng, sp
    torch.cuda.empty_cache()
    return out


def main():
    import torch
    import torch.distributed as dist

    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    C = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    I = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
    mode = sys.argv[4] if len(sys.argv) > 4 else "fused"
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = measure(steps, C, I, mode, verbose=True)
    if int(os.environ.get("RANK", 0)) == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
