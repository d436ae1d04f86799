import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bithtm_b200 as bithtm
from bench import CFG2, make_inputs
cfg = CFG2
xs = make_inputs(cfg, 3000, 0)
np.random.seed(0)
htm = bithtm.HierarchicalTemporalMemory(cfg["input_dim"], cfg["column_dim"], cfg["cell_dim"], cfg["active_columns"], max_segments=1 << 17)
for t in range(500): htm.process(xs[t])
def run():
    b = 0
    for t in range(500, 2500):
        sp, tm = htm.process(xs[t]); b += int(tm.active_column_bursting.sum())
    return b
t0 = time.perf_counter(); run(); dt = time.perf_counter() - t0
print("e2e us/step", dt / 2000 * 1e6)
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
