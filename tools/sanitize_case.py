"""Smallest complete case for compute-sanitizer: a 256-column network, 30 steps in each
execution mode (cluster kernel, cooperative-grid kernel, one kernel per stage).
    compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bithtm_b200 as bithtm

g = np.random.default_rng(0)
base = g.random((5, 64)) < 0.25
for mode in (sys.argv[1:] or ["cluster", "grid", "off"]):
    np.random.seed(3)
    htm = bithtm.HierarchicalTemporalMemory(64, 256, 8, 20, fused=mode, max_segments=4096)
    for t in range(30):
        sp, tm = htm.process(base[t % 5] ^ (g.random(64) < 0.05))
    print(mode, "ok", tm.n_segments, int(htm.engine.scalars()[12]))
