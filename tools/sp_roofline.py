#!/usr/bin/env python
"""Time the spatial-pooler kernels one by one at an HBM-bound size (default cfg3:
65536 columns x 16384 inputs, k = 1311) with CUDA events and report achieved GB/s
against the measured HBM peak.  Permanence is drawn on the device (performance only).

    python tools/sp_roofline.py [C] [I] [iters]
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bithtm_b200 as bithtm
from bithtm_b200 import _native as nat
from bithtm_b200.projections import DenseProjection


def measure(Ccol=65536, I=16384, iters=20):
    k = round(Ccol * 0.02)
    torch.manual_seed(0)
    perm = torch.randn(Ccol, I, dtype=torch.float64, device="cuda") * 0.1
    proj = DenseProjection(I, Ccol, permanence=perm)
    sp = bithtm.SpatialPooler(I, Ccol, k, proximal_projection=proj)
    eng = sp._ensure_engine()
    del perm, proj._host_permanence
    proj._host_permanence = None
    torch.cuda.empty_cache()
    g = np.random.default_rng(0)
    xs = [eng.pack_input(g.random(I) < 0.2) for _ in range(4)]
    for x in xs:  # warm-up: a few complete SP steps so duty cycles / active sets are realistic
        sp.process(x, learning=True)
    torch.cuda.synchronize()
    st = eng.stream
    calls = {
        "sp_overlap_boost": lambda x: nat.lib.bh_sp_step,  # placeholder, replaced below
    }
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    mask_bytes = Ccol * eng.ctx.mask_stride * 4
    rows = [
        ("sp_overlap", lambda x: nat.lib.bh_sp_overlap(eng.ref, x.data_ptr(), st), mask_bytes + I / 8 + 4 * Ccol),
        ("boost", lambda x: nat.lib.bh_boost(eng.ref, st), 16 * Ccol),
        ("topk", lambda x: nat.lib.bh_inhibit(eng.ref, st), 8 * Ccol + 4 * k),
        ("sp_learn", lambda x: nat.lib.bh_sp_learn(eng.ref, x.data_ptr(), st), 16 * k * I + k * eng.ctx.mask_stride * 4 + I / 8),
        ("duty_update", lambda x: nat.lib.bh_duty_update(eng.ref, st), 9 * Ccol),
    ]
    out = []
    for name, fn, nbytes in rows:
        for i in range(3):
            nat.check(fn(xs[i % 4]))
        evs = []
        for i in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            nat.check(fn(xs[i % 4]))
            b.record()
            evs.append((a, b))
            if name == "sp_learn":
                # rotate the active set so consecutive launches touch different rows (no L2 reuse)
                nat.check(nat.lib.bh_sp_overlap(eng.ref, xs[(i + 1) % 4].data_ptr(), st))
                nat.check(nat.lib.bh_boost(eng.ref, st))
                nat.check(nat.lib.bh_inhibit(eng.ref, st))
                nat.check(nat.lib.bh_duty_update(eng.ref, st))
                nat.check(nat.lib.bh_advance_step(eng.ref, st))
        torch.cuda.synchronize()
        ms = float(np.median([a.elapsed_time(b) for a, b in evs]))
        gbs = nbytes / (ms * 1e-3) / 1e9
        out.append({"kernel": name, "us": round(ms * 1e3, 2), "algorithmic_bytes": int(nbytes),
                    "achieved_gbs": round(gbs, 1), "peak_gbs": peak, "frac": round(gbs / peak, 4)})
    return {"workload": f"SP {Ccol} columns x {I}-bit input, k={k}", "kernels": out}


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    r = measure(*(a + [65536, 16384, 20][len(a):]))
    print(json.dumps(r))
    for kk in r["kernels"]:
        print(f"  {kk['kernel']:14s} {kk['us']:9.2f} us  {kk['achieved_gbs']:8.1f} GB/s  {100 * kk['frac']:5.1f} % of {kk['peak_gbs']}")
