#!/usr/bin/env python
"""Shared-mask batched overlap (DenseProjection.process_batch; a loop of projections.py:18-21):
the int8 tensor-core contraction (bh_sp_overlap_batched_tc) next to the AND + popcount kernel
(bh_sp_overlap_batched), CUDA-event timed, results compared bit for bit.

    python tools/batched_overlap.py [B] [C] [I] [iters]

Default: BASELINE configs[3]'s shape, 1024 inputs x 2048 columns x 1024 input bits.
Prints one JSON line.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bithtm_b200 as bithtm
from bithtm_b200 import _native as nat
from bithtm_b200.projections import DenseProjection


REP = 20
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tensor_peak_tops():
    """Dense int8 tensor peak of this B200: 2 x the measured dense bf16 throughput (MEASURED_PEAKS.json,
    burst figure; cuBLAS bf16 8192^3), else 2 x the profiling recipe's fallback."""
    try:
        return 2.0 * float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]), "2 x MEASURED_PEAKS.json bf16_tflops"
    except Exception:
        return 2.0 * 1665.0, "2 x fallback bf16 1665 TF/s (B200_PROFILING.md)"


def measure(B=1024, C=2048, I=1024, iters=200):
    torch.manual_seed(0)
    perm = torch.randn(C, I, dtype=torch.float64, device="cuda") * 0.1
    np.random.seed(0)
    proj = DenseProjection(I, C, permanence=perm)
    sp = bithtm.SpatialPooler(I, C, max(1, round(0.02 * C)), proximal_projection=proj)
    sp._ensure_engine()
    eng = proj._need_engine()
    words = eng.ctx.input_words
    g = torch.Generator(device="cuda").manual_seed(1)
    packed = torch.randint(-2 ** 31, 2 ** 31 - 1, (B, words), dtype=torch.int64, device="cuda", generator=g)
    packed &= torch.randint(-2 ** 31, 2 ** 31 - 1, (B, words), dtype=torch.int64, device="cuda", generator=g)  # ~25 % ones
    if I % 32:
        packed[:, -1] &= (1 << (I % 32)) - 1
    packed = packed.to(torch.int32).contiguous()
    outs, res = {}, {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    names = ["bh_sp_overlap_batched_tc", "bh_sp_overlap_batched_mma", "bh_sp_overlap_batched"]
    if words % 4 == 0:
        names.insert(1, "bh_sp_overlap_batched_tc5")
    for name in names:
        fn = getattr(nat.lib, name)
        out = torch.full((B, C), -1, dtype=torch.int32, device="cuda")
        for _ in range(5):
            nat.check(fn(eng.ref, packed.data_ptr(), B, out.data_ptr(), eng.stream), name)
        torch.cuda.synchronize()
        warm, cold = [], []
        for it in range(iters):
            # cold: one launch after an L2 flush (includes the host's launch latency: the GPU idles
            # between the event and the kernel); warm: REP launches back to back, per launch
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            rep = 1 if it % 2 else REP
            if it % 2:
                flush.fill_(it & 0xFF)
            a.record()
            for _ in range(rep):
                nat.check(fn(eng.ref, packed.data_ptr(), B, out.data_ptr(), eng.stream), name)
            b.record()
            torch.cuda.synchronize()
            (cold if it % 2 else warm).append(a.elapsed_time(b) * 1e3 / rep)
        outs[name] = out
        us = float(np.median(warm))
        res[name] = {"us_l2_flushed_single_launch": round(float(np.median(cold)), 2), "us_back_to_back": round(us, 2),
                     "out_gbs": round(4.0 * B * C / us / 1e3, 1),
                     "int8_tops": round(2.0 * B * C * words * 32 / us / 1e6, 1)}
    same = all(bool(torch.equal(outs[n], outs["bh_sp_overlap_batched"])) for n in names)
    if int(eng.buf["sc"][nat.SC_T5_ERR].item()):
        same = False
    # spot check against the definition on a few rows
    mask = (perm >= 0.0)
    bits = ((packed[:4].to(torch.int64)[:, :, None] >> torch.arange(32, device="cuda")) & 1).reshape(4, -1)[:, :I].bool()
    want = (mask[None, :, :] & bits[:, None, :]).sum(dim=2).to(torch.int32)
    ok = bool(torch.equal(outs["bh_sp_overlap_batched_tc"][:4], want))
    tc = res["bh_sp_overlap_batched_tc"]
    peak, src = tensor_peak_tops()
    best = "bh_sp_overlap_batched_tc5" if "bh_sp_overlap_batched_tc5" in res else "bh_sp_overlap_batched_mma"
    return {"workload": f"{B} inputs x {C} columns x {I} bits, one shared mask", "iters": iters,
            "tensor_core_auto": tc, "tcgen05": res.get("bh_sp_overlap_batched_tc5"), "mma_sync": res["bh_sp_overlap_batched_mma"],
            "popcount": res["bh_sp_overlap_batched"],
            "roofline": {"bound": "tensor", "kernel": "k_sp_overlap_batched_t5" if best.endswith("tc5") else "k_sp_overlap_batched_tc",
                         "achieved": res[best]["int8_tops"], "peak": round(peak, 1), "unit": "TOP/s (dense int8)",
                         "frac": round(res[best]["int8_tops"] / peak, 4), "peak_source": src},
            "bit_identical": same, "matches_definition": ok,
            "algorithmic_bytes": int(4 * B * C + (B + C) * words * 4)}


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    C = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    I = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    iters = int(sys.argv[4]) if len(sys.argv) > 4 else 200
    r = measure(B, C, I, iters)
    print(json.dumps(r))
    if not (r["bit_identical"] and r["matches_definition"]):
        sys.exit(1)


if __name__ == "__main__":
    main()
