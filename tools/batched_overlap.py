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


# int8 operations per second the mma.sync path can sustain on B200 at 1965 MHz, as ncu reports it
# (sm__ops_path_tensor_op_imma_src_int8_sparsity_off.sum.peak_sustained = 606 208 ops/cycle)
IMMA_PEAK_TOPS = 606208 * 1.965e9 / 1e12


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
    for name in ("bh_sp_overlap_batched_tc", "bh_sp_overlap_batched"):
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
    same = bool(torch.equal(outs["bh_sp_overlap_batched_tc"], outs["bh_sp_overlap_batched"]))
    # spot check against the definition on a few rows
    mask = (perm >= 0.0)
    bits = ((packed[:4].to(torch.int64)[:, :, None] >> torch.arange(32, device="cuda")) & 1).reshape(4, -1)[:, :I].bool()
    want = (mask[None, :, :] & bits[:, None, :]).sum(dim=2).to(torch.int32)
    ok = bool(torch.equal(outs["bh_sp_overlap_batched_tc"][:4], want))
    tc = res["bh_sp_overlap_batched_tc"]
    return {"workload": f"{B} inputs x {C} columns x {I} bits, one shared mask", "iters": iters,
            "tensor_core": tc, "popcount": res["bh_sp_overlap_batched"],
            "roofline": {"bound": "tensor", "kernel": "k_sp_overlap_batched_tc", "achieved": tc["int8_tops"],
                         "peak": round(IMMA_PEAK_TOPS, 1), "unit": "TOP/s (int8, mma.sync path)",
                         "frac": round(tc["int8_tops"] / IMMA_PEAK_TOPS, 4),
                         "peak_source": "ncu sm__ops_path_tensor_op_imma_src_int8 peak_sustained x 1965 MHz"},
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
