#!/usr/bin/env python
"""Per-kernel times of the whole SP+TM step at an HBM-bound size (default cfg3: 65536
columns x 16384 inputs, 32 cells, k = 1311) on one GPU, one kernel per stage, CUDA events
after every launch (bh_profile_step), against the measured HBM peak.  Permanence is drawn
on the device (performance run, not a parity run).

    python tools/cfg3_kernels.py [C] [I] [warm_steps] [profiled_steps]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def algorithmic_bytes(cfg, name, st):
    """Bytes one launch of kernel `name` must move.  cfg: column_dim,
    input_dim, active_columns, cell_dim; st: S (segments), synapses, M, L, P, W, n_row (mean
    synapses of a learning row), rng_words (stream words produced by the launch)."""
    C, I, k, c = cfg["column_dim"], cfg["input_dim"], cfg["active_columns"], cfg["cell_dim"]
    S, syn, M, L, W = (st[n] for n in ("S", "synapses", "M", "L", "W"))
    P, n_row = st.get("P", 0), st.get("n_row", 40)
    table = {
        "sp_overlap_boost": C * I / 8 + I / 8 + 16 * C,          # mask + input + duty read, overlaps/boosted write
        "sp_overlap": C * I / 8 + I / 8 + 4 * C,
        "topk": 8 * C + 4 * k,                                    # keys once + the k winners
        "topk_multi": 8 * C + 4 * k,
        "sp_learn": 16 * k * I + k * I / 8 + I / 8,               # fp64 RMW of k rows + their mask rows
        "duty_update": 9 * C,
        "tm_activate_a": 8 * syn + 12 * S,                        # every live synapse (cell + perm) once
        "tm_learn_apply": 16 * (L + P) * n_row + 8 * L * (W + 1),  # learning rows RMW + their priority rows
        "tm_draw2": 8 * L * (W + 1),                              # stream words written (4 B each, 2 per double)
        "rng_chunks": 8 * L * (W + 1),
    }
    # the fused kernel moves the whole step (SP + TM algorithmic bytes)
    sp = C * I / 8 + I / 8 + 16 * k * I + k * I / 8 + 28 * C
    tm = 8 * syn + 16 * (L + P) * n_row + 4 * (C * c) / 8 + 8 * (k * c + L * (W + 1) + M)
    table["step_fused_cluster"] = table["step_fused_grid"] = sp + tm
    return table.get(name)


def network_stats(eng):
    sc = eng.scalars()
    S = int(sc[2])
    counts = eng.buf["seg_count"][:S].cpu().numpy()
    syn = int(counts.sum())
    return dict(S=S, synapses=syn, M=int(sc[4]), L=int(sc[8]), P=int(sc[9]), W=int(sc[5 + ((int(sc[0]) - 1) & 1)]),
                n_row=float(syn) / max(S, 1))


def measure(Ccol=65536, I=16384, warm=250, profiled=20):
    import torch

    import bithtm_b200 as bithtm
    from bithtm_b200.projections import DenseProjection

    c, k = 32, round(Ccol * 0.02)
    patterns = 50
    g = np.random.default_rng(0)
    base = g.random((patterns, I)) < 0.2
    xs = base[np.arange(2 * patterns) % patterns] ^ (g.random((2 * patterns, I)) < 0.05)
    torch.manual_seed(0)
    perm = torch.randn(Ccol, I, dtype=torch.float64, device="cuda") * 0.1
    np.random.seed(0)
    sp = bithtm.SpatialPooler(I, Ccol, k, proximal_projection=DenseProjection(I, Ccol, permanence=perm))
    htm = bithtm.HierarchicalTemporalMemory(I, Ccol, c, k, spatial_pooler=sp, rng_sync="lazy", ring_len=len(xs),
                                            max_segments=1 << 21, max_synapses_per_segment=64, fused="grid")
    del perm
    sp.proximal_projection._host_permanence = None
    torch.cuda.empty_cache()
    eng = htm.engine
    htm.temporal_memory._rng.before(eng)
    eng.load_ring(xs)
    words = [eng.pack_input(x) for x in xs]
    # the whole step as one cooperative kernel (what a user runs): graphs of 50 steps from the device ring
    per = 50
    graph = eng.graph(per, learning=True)
    for _ in range(max(1, warm // per)):
        eng.launch_graph(graph, per)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(2):
        eng.launch_graph(graph, per)
    b.record()
    torch.cuda.synchronize()
    fused_us = a.elapsed_time(b) * 1e3 / (2 * per)
    # the same step one kernel per stage (the library is stateless: the mode is a field of the context)
    eng.ctx.fused_mode = 0
    eng.ctx.pipe_ctas = 0  # (a field of the fused kernels)
    prof, order = {}, []
    for t in range(profiled):
        for name, ms in eng.profile_step(words[t % len(words)], learning=True):
            if name not in prof:
                order.append(name)
            prof[name] = prof.get(name, 0.0) + ms / profiled
    eng.check_status()
    st = network_stats(eng)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    cfg = dict(column_dim=Ccol, input_dim=I, active_columns=k, cell_dim=c)
    rows = []
    for name in order:
        ab = algorithmic_bytes(cfg, name, st)
        us = prof[name] * 1e3
        row = {"kernel": name, "us": round(us, 2)}
        if ab:
            gbs = ab / (us * 1e-6) / 1e9
            row.update(algorithmic_bytes=int(ab), achieved_gbs=round(gbs, 1), frac=round(gbs / peak, 4))
        rows.append(row)
    total = sum(prof.values()) * 1e3
    sp_bytes = algorithmic_bytes(cfg, "step_fused_grid", st)
    return {"workload": f"SP {Ccol} columns x {I}-bit input, k={k}, TM {c} cells/column",
            "fused_step_us": round(fused_us, 1), "fused_steps_per_s": round(1e6 / fused_us, 1),
            "fused_algorithmic_bytes": int(sp_bytes), "fused_achieved_gbs": round(sp_bytes / (fused_us * 1e-6) / 1e9, 1),
            "fused_frac": round(sp_bytes / (fused_us * 1e-6) / 1e9 / peak, 4),
            "per_stage_step_us": round(total, 1), "peak_gbs": peak, "state": st, "kernels": rows,
            "note": "fused_* = whole step as one cooperative kernel (graphs of 50 steps, device input ring); "
                    "kernels = the same step one kernel per stage, CUDA events after every launch"}


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    r = measure(*(a + [65536, 16384, 250, 20][len(a):]))
    print(json.dumps(r))
    for kk in r["kernels"]:
        extra = f"{kk['achieved_gbs']:8.1f} GB/s  {100 * kk['frac']:5.1f} % of {r['peak_gbs']}" if "frac" in kk else ""
        print(f"  {kk['kernel']:20s} {kk['us']:9.2f} us  {extra}")
    print(f"  per-stage sum {r['per_stage_step_us']} us; fused step {r['fused_step_us']} us -> {r['fused_steps_per_s']} "
          f"steps/s = {r['fused_achieved_gbs']} GB/s algorithmic = {100 * r['fused_frac']:.1f} % of peak")
