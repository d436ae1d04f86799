#!/usr/bin/env python
"""BASELINE configs[1] (2048 columns x 1024-bit input, k = 41, 32 cells): the latency-bound network.  One
step = one launch of the cluster kernel; times with L2 flushed before every step, L2-resident (graphs of
50 steps) and end to end from host arrays."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def measure(cfg, steps=400, warm=100):
    import torch

    import bithtm_b200 as bithtm
    from bench import make_inputs

    total = warm + steps
    xs = make_inputs(cfg, total, cfg["seed"])

    def build(ring_len, rng_sync):
        np.random.seed(cfg["seed"])
        return bithtm.HierarchicalTemporalMemory(cfg["input_dim"], cfg["column_dim"], cfg["cell_dim"],
                                                 cfg["active_columns"], rng_sync=rng_sync, ring_len=ring_len,
                                                 max_segments=1 << 17)

    htm = build(total, "lazy")
    eng = htm.engine
    htm.temporal_memory._rng.before(eng)
    eng.load_ring(xs)
    g1 = eng.graph(1, learning=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        eng.launch_graph(g1, 1)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush.fill_(1)
        a.record()
        eng.launch_graph(g1, 1)
        b.record()
    torch.cuda.synchronize()
    cold_us = float(np.mean([a.elapsed_time(b) for a, b in ev])) * 1e3
    per = 50
    g50 = eng.graph(per, learning=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(max(1, steps // per)):
        eng.launch_graph(g50, per)
    e1.record()
    torch.cuda.synchronize()
    warm_us = e0.elapsed_time(e1) * 1e3 / (max(1, steps // per) * per)
    eng.check_status()
    del htm, eng
    htm2 = build(0, "step")
    for t in range(warm):
        htm2.process(xs[t])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(warm, total):
        htm2.process(xs[t])
    torch.cuda.synchronize()
    e2e_us = (time.perf_counter() - t0) / steps * 1e6
    return {"workload": f"cfg2: {cfg['column_dim']} columns x {cfg['input_dim']}-bit input, k={cfg['active_columns']}, "
                        f"{cfg['cell_dim']} cells/column, learning on, one 16-CTA cluster kernel per step",
            "state": f"steps {warm}..{total} of a fresh network",
            "us_per_step_l2_flushed": round(cold_us, 2), "steps_per_s_l2_flushed": round(1e6 / cold_us, 1),
            "us_per_step_l2_resident": round(warm_us, 2), "steps_per_s_l2_resident": round(1e6 / warm_us, 1),
            "us_per_step_e2e_host": round(e2e_us, 2), "steps_per_s_e2e_host": round(1e6 / e2e_us, 1)}


if __name__ == "__main__":
    import json

    from bench import CFG2

    print(json.dumps(measure(CFG2)))
