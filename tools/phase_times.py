#!/usr/bin/env python
"""Per-phase times inside the fused step kernel (globaltimer stamps of CTA 0),
averaged over steps at steady state.  Usage: python tools/phase_times.py [fused] [ctas] [warm_steps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bithtm_b200 as bithtm
from bench import CFG2, make_inputs

PHASES = ["P0 overlap+draw1", "P1 topk", "P2 sp_learn+duty+select_a", "P3 select_b+learn_select_a",
          "P4 learn_select_b+draw2", "P4b rng chunks", "P5 learn_apply", "P6 post", "P7 activate_a", "P8 draw3", "P9 activate_b"]

def main():
    fused = sys.argv[1] if len(sys.argv) > 1 else "cluster"
    ctas = int(sys.argv[2]) if len(sys.argv) > 2 else None
    warm = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
    cfg = CFG2
    n = warm + 300
    xs = make_inputs(cfg, n, 0)
    np.random.seed(0)
    htm = bithtm.HierarchicalTemporalMemory(cfg["input_dim"], cfg["column_dim"], cfg["cell_dim"],
                                            cfg["active_columns"], rng_sync="lazy", ring_len=n,
                                            max_segments=1 << 17, fused=fused, fused_ctas=ctas)
    eng = htm.engine
    htm.temporal_memory._rng.before(eng)
    eng.load_ring(xs)
    g = eng.graph(100, learning=True)
    for _ in range(warm // 100):
        eng.launch_graph(g, 100)
    torch.cuda.synchronize()
    g1 = eng.graph(1, learning=True)
    acc = np.zeros(len(PHASES))
    reps = 200
    for _ in range(reps):
        eng.launch_graph(g1, 1)
        torch.cuda.synchronize()
        st = eng.buf["blk"][7 * 1024:7 * 1024 + 2 * (len(PHASES) + 1)].cpu().numpy().view(np.uint64).astype(np.float64)
        acc += np.diff(st)
    acc /= reps
    sc = eng.scalars()
    print(f"fused={fused} ctas={eng.ctx.fused_ctas}  S={sc[2]} M={sc[4]} L={sc[8]} P={sc[9]}  total {acc.sum() / 1e3:.1f} us")
    for name, v in zip(PHASES, acc):
        print(f"  {name:32s} {v / 1e3:7.2f} us")


if __name__ == "__main__":
    main()
