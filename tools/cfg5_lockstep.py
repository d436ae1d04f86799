#!/usr/bin/env python
"""BASELINE configs[4]'s temporal-memory scale on ONE GPU against the oracle: 1 048 576 columns, k = 20 972
active columns, 32 cells -- with a short input (1024 bits) so that the float64 permanence (8 GiB) fits one GPU
and the host.  rand(L, W+1) is 4.4e8 doubles per timestep here (projections.py:120, quadratic in k): the device
only steps over it (lazy draws); the oracle really draws it (~3.5 GB per step), which bounds this run to a few
steps.  Every State field of every step and the learned state are compared.

    python tools/cfg5_lockstep.py [steps] [columns] [input_bits]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch

    import bithtm_b200 as bithtm
    from bithtm_b200.projections import DenseProjection
    from helpers import diff_records, gpu_record, oracle_record
    from oracle.digest import canonical_from_rows, state_digest
    from oracle.htm_oracle import HTMOracle, OracleConfig

    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    C = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
    I = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    c, k, seed = 32, round(C * 0.02), 3
    g = np.random.default_rng(8)
    base = g.random((2, I)) < 0.2
    xs = base[np.arange(steps) % 2] ^ (g.random((steps, I)) < 0.02)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    perm = torch.randn(C, I, dtype=torch.float64, device="cuda", generator=gen) * 0.1
    host_perm = perm.cpu().numpy()
    np.random.seed(seed)
    sp = bithtm.SpatialPooler(I, C, k, proximal_projection=DenseProjection(I, C, permanence=perm))
    t0 = time.time()
    htm = bithtm.HierarchicalTemporalMemory(I, C, c, k, spatial_pooler=sp, max_segments=1 << 21,
                                            max_synapses_per_segment=128)
    del perm
    eng = htm.engine
    print(f"engine: fused_mode {eng.ctx.fused_mode}, skip table {eng.ctx.skip_polys} x {eng.ctx.skip_gran} words, "
          f"lazy policy {eng.ctx.lazy_policy}, ring {eng.ctx.rng_ring_words} words, step words {eng.ctx.rng_step_words}, "
          f"arena {eng.arena_bytes / 2**30:.1f} GiB, built in {time.time() - t0:.1f} s", flush=True)
    orc = HTMOracle(OracleConfig(I, C, c, k), rng=np.random.RandomState(seed), overlap="packed", permanence=host_perm)
    stats = []
    for t in range(steps):
        t1 = time.time()
        sp_state, tm_state = htm.process(xs[t])
        torch.cuda.synchronize()
        t_gpu = time.time() - t1
        t1 = time.time()
        rec = orc.step(xs[t])
        t_cpu = time.time() - t1
        problems = diff_records(gpu_record(htm, sp_state, tm_state), oracle_record(rec))
        if problems:
            print(f"step {t}: MISMATCH " + "; ".join(problems) + f"; sc={eng.scalars()[:24]}")
            sys.exit(1)
        stats.append(dict(step=t, draws=int(rec.draws), learning=len(rec.learning_segment), matching=len(rec.matching_segment),
                          segments=rec.n_segments, gpu_ms=round(t_gpu * 1e3, 2), oracle_s=round(t_cpu, 1)))
        print(stats[-1], flush=True)
    eng.check_status()
    tmp = htm.temporal_memory.distal_projection
    owner, count, cells, perm_rows = tmp.export_segments()
    same = state_digest(np.zeros(1), htm.spatial_pooler.boosting.duty_cycle, tmp.bundle_segments,
                        canonical_from_rows(owner, cells, perm_rows)) == \
        state_digest(np.zeros(1), orc.duty, orc.cell_nseg, orc.canonical_synapses())
    rng_same = bool(np.array_equal(np.random.random_sample(64), orc.rng.random_sample(64)))
    print(json.dumps({"workload": f"{C} columns x {I} inputs, k={k}, {c} cells, {steps} steps lock-step vs the oracle",
                      "bit_exact_every_step": True, "learned_state_equal": bool(same), "np_random_position_equal": rng_same,
                      "steps": stats}))
    sys.exit(0 if same and rng_same else 1)


if __name__ == "__main__":
    main()
