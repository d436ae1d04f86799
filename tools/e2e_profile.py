#!/usr/bin/env python
"""Where the end-to-end step (HierarchicalTemporalMemory.process(host array) + a read of the result) spends its
time on the host: wall clock per step, then a cProfile of the same loop.

    python tools/e2e_profile.py [cfg2|cfg3] [steps]
"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bithtm_b200 as bithtm
from bench import CFG2, CFG3, build_network, make_inputs


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else (2000 if which == "cfg2" else 400)
    warm = 500 if which == "cfg2" else 300
    if which == "cfg3":
        cfg = CFG3
        xs = make_inputs(cfg, 100, 0)
        htm = build_network(cfg, 0, 1, 100, "step")
    else:
        cfg = CFG2
        xs = make_inputs(cfg, 3000, 0)
        np.random.seed(0)
        htm = bithtm.HierarchicalTemporalMemory(cfg["input_dim"], cfg["column_dim"], cfg["cell_dim"],
                                                cfg["active_columns"], max_segments=1 << 17)
    for t in range(warm):
        htm.process(xs[t % len(xs)])

    def run():
        b = 0
        for t in range(n):
            sp, tm = htm.process(xs[(warm + t) % len(xs)])
            b += int(tm.active_column_bursting.sum())
        return b

    t0 = time.perf_counter()
    run()
    dt = time.perf_counter() - t0
    print(f"{which}: e2e {dt / n * 1e6:.1f} us/step")
    pr = cProfile.Profile()
    pr.enable()
    run()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(18)


if __name__ == "__main__":
    main()
