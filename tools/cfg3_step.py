#!/usr/bin/env python
"""Whole SP+TM step at cfg3 size (65536 columns x 16384 inputs, 32 cells, k=1311) on
one GPU: ms/step and the per-phase split of the fused cooperative kernel.
Permanence is drawn on the device (performance run, not a parity run).

    python tools/cfg3_step.py [steps] [C] [I]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bithtm_b200 as bithtm
from bithtm_b200.projections import DenseProjection

PHASES = ["P0 overlap+draw1", "P1 topk", "P2 sp_learn+duty+select_a", "P3 select_b+learn_select_a",
          "P4 learn_select_b+draw2", "P4b rng chunks", "P5 learn_apply", "P6 post", "P7 activate_a", "P8 draw3", "P9 activate_b"]


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    C = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    I = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
    c, k = 32, round(C * 0.02)
    patterns = 50
    g = np.random.default_rng(0)
    base = g.random((patterns, I)) < 0.2
    xs = base[np.arange(steps) % patterns] ^ (g.random((steps, I)) < 0.05)
    torch.manual_seed(0)
    perm = torch.randn(C, I, dtype=torch.float64, device="cuda") * 0.1
    np.random.seed(0)
    sp = bithtm.SpatialPooler(I, C, k, proximal_projection=DenseProjection(I, C, permanence=perm))
    htm = bithtm.HierarchicalTemporalMemory(I, C, c, k, spatial_pooler=sp, rng_sync="lazy", ring_len=steps,
                                            max_segments=1 << 21, max_synapses_per_segment=64,
                                            fused=os.environ.get("BH_FUSED", "grid"),
                                            **({"tail_chunks": int(os.environ["TAIL_CHUNKS"])} if "TAIL_CHUNKS" in os.environ else {}))
    del perm
    sp.proximal_projection._host_permanence = None
    torch.cuda.empty_cache()
    eng = htm.engine
    htm.temporal_memory._rng.before(eng)
    eng.load_ring(xs)
    g1 = eng.graph(1, learning=True)
    acc = np.zeros(len(PHASES))
    times = []
    for t in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.launch_graph(g1, 1)
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
        if t >= steps - 50:
            st = eng.buf["blk"][7 * 1024:7 * 1024 + 2 * (len(PHASES) + 1)].cpu().numpy().view(np.uint64).astype(np.float64)
            acc += np.diff(st) / 50
        if t % 50 == 49:
            sc = eng.scalars()
            print(f"step {t + 1}: {np.mean(times[-50:]):.3f} ms/step  S={sc[2]} M={sc[4]} L={sc[8]} P={sc[9]} "
                  f"W={sc[5 + (t & 1)]} status={sc[12]}", flush=True)
    print(f"last 50 steps: {np.mean(times[-50:]):.3f} ms/step -> {1e3 / np.mean(times[-50:]):.1f} steps/s")
    for name, v in zip(PHASES, acc):
        print(f"  {name:32s} {v / 1e3:9.2f} us")
    if os.environ.get("BH_TOPK_STAMPS"):  # library built with -DBH_TOPK_STAMPS: stages of topk_grid (last step)
        st = eng.buf["blk"][7 * 1024 + 80:7 * 1024 + 80 + 32].cpu().numpy().view(np.uint64)
        print("  topk_grid stages (ns):", np.diff(st[:7].astype(np.int64)).tolist(), "candidates", int(st[15] & 0xffffffff),
              "range published", int(st[15] >> 32))
        rs = eng.buf["blk"][7 * 1024 + 118:7 * 1024 + 124].cpu().numpy().view(np.uint64).astype(np.int64)
        print("  rng_chunk CTA 0 (ns): start", int(rs[1] - rs[0]), "generate", int(rs[2] - rs[1]))


if __name__ == "__main__":
    main()
