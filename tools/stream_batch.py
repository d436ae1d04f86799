#!/usr/bin/env python
"""Aggregate throughput of B independent cfg2 streams advanced side by side on one GPU
(BASELINE configs[3]: independent 2048-column SP+TM streams).

    python tools/stream_batch.py [B] [fused_ctas] [steps] [warm] [seed0] [fused_threads]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def measure(B=64, ctas=16, steps=400, warm=200, seed0=0, threads=None):
    import torch

    import bithtm_b200 as bithtm
    from bench import CFG2, make_inputs

    cfg = CFG2
    total = warm + steps
    nets, inputs = [], []
    for i in range(B):
        np.random.seed(seed0 + i)
        nets.append(bithtm.HierarchicalTemporalMemory(cfg["input_dim"], cfg["column_dim"], cfg["cell_dim"],
                                                      cfg["active_columns"], rng_sync="lazy", ring_len=total,
                                                      max_segments=1 << 15, fused="cluster", fused_ctas=ctas,
                                                      fused_threads=threads))
        inputs.append(make_inputs(cfg, total, seed0 + i))
    batch = bithtm.StreamBatch(nets)
    batch.load_inputs(inputs)
    per = 50
    for _ in range(warm // per):
        batch.run(per)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps // per):
        batch.run(per)
    b.record()
    torch.cuda.synchronize()
    batch.check_status()
    ms = a.elapsed_time(b)
    n = (steps // per) * per
    return {"streams": B, "fused_ctas": ctas, "fused_threads": threads or 1024, "steps_per_stream": n, "ms": ms,
            "aggregate_steps_per_s": B * n / (ms * 1e-3), "per_stream_steps_per_s": n / (ms * 1e-3)}


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    print(json.dumps(measure(*(a + [64, 16, 400, 200, 0, None][len(a):]))))
