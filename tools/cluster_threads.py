#!/usr/bin/env python
"""cfg2, ONE network: L2-resident step time of the cluster kernel for several CTA sizes
(fused_threads) and cluster sizes.   python tools/cluster_threads.py [CTASxTHREADS ...]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bithtm_b200 as bithtm
from bench import CFG2, make_inputs


def measure(ctas, threads, steps=1500, warm=1000):
    cfg = CFG2
    total = warm + steps
    np.random.seed(0)
    htm = bithtm.HierarchicalTemporalMemory(cfg["input_dim"], cfg["column_dim"], cfg["cell_dim"], cfg["active_columns"],
                                            rng_sync="lazy", ring_len=total, max_segments=1 << 17, fused="cluster",
                                            fused_ctas=ctas, fused_threads=threads)
    eng = htm.engine
    htm.temporal_memory._rng.before(eng)
    eng.load_ring(make_inputs(cfg, total, 0))
    g = eng.graph(50, learning=True)
    for _ in range(warm // 50):
        eng.launch_graph(g, 50)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps // 50):
        eng.launch_graph(g, 50)
    b.record()
    torch.cuda.synchronize()
    eng.check_status()
    return a.elapsed_time(b) / steps * 1e3


if __name__ == "__main__":
    out = {}
    pairs = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(16, 1024), (16, 768), (16, 512), (12, 1024), (8, 512)]
    for ctas, threads in pairs:
        out[f"{ctas}x{threads or 1024}"] = round(measure(ctas, threads), 2)
    print(json.dumps({"workload": "cfg2, one network, us/step L2-resident", "us_per_step": out}))
