"""Debug: one network as a single cooperative kernel vs the same network as `world` shard kernels running
concurrently on one GPU; prints the bookkeeping scalars of every step until they diverge."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bithtm_b200 as bithtm
from bithtm_b200.projections import DenseProjection

world, k, ctas = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 80
I, C, c = int(os.environ.get("I", 4096)), int(os.environ.get("C", 32768)), 32
PAT = int(os.environ.get("PATTERNS", 20))
g = np.random.default_rng(3)
base = g.random((PAT, I)) < 0.2
xs = base[np.arange(steps) % PAT] ^ (g.random((steps, I)) < 0.05)
gen = torch.Generator(device="cuda")
gen.manual_seed(99)
perm = torch.randn(C, I, dtype=torch.float64, device="cuda", generator=gen) * 0.1
kw = dict(rng_sync="lazy", max_segments=int(os.environ.get("MAXSEG", 1 << 18)),
          max_synapses_per_segment=int(os.environ.get("SYN", 128)))
CELLS = {"0": False, "1": True}.get(os.environ.get("CELLS", ""), "auto")  # exchange protocol of the shards
if os.environ.get("LAZY") == "0":
    kw["lazy_rng"] = False


def build(**extra):
    np.random.seed(4)
    rows = perm
    if "column_shard" in extra:
        r, w = extra["column_shard"]
        rows = perm[r * C // w:(r + 1) * C // w]
    sp = bithtm.SpatialPooler(I, C, k, proximal_projection=DenseProjection(I, C, permanence=rows))
    return bithtm.HierarchicalTemporalMemory(I, C, c, k, spatial_pooler=sp, **kw, **extra)


whole = build(fused=os.environ.get("WHOLE", "grid"), **({"lazy_rng": False} if os.environ.get("WLAZY") == "0" else {}))
shards = [build(column_shard=(r, world), fused="shard", fused_ctas=ctas, exchange_cells=CELLS) for r in range(world)]
print(f"whole: fused_mode {whole.engine.ctx.fused_mode} lazy_policy {whole.engine.ctx.lazy_policy} | shards: xch_ll "
      f"{shards[0].engine.ctx.xch_ll} lazy_policy {shards[0].engine.ctx.lazy_policy} tail_chunks {shards[0].engine.ctx.tail_chunks}",
      flush=True)
regions = [torch.zeros(h.engine.exchange_region_ints(), dtype=torch.int32, device="cuda") for h in shards]
for h in shards + [whole]:
    h.temporal_memory._rng.before(h.engine)
for h in shards:
    h.engine.set_exchange_regions([r.data_ptr() for r in regions], keepalive=regions)
streams = [torch.cuda.Stream() for _ in shards]
names = ["step", "prev", "S", "Snext", "M", "W0", "W1", "L0", "L", "P", "NU", "NR", "status", "mtpos", "XM", "XRA", "XRT"]
for t in range(steps):
    words = whole.engine.pack_input(xs[t])
    whole.process(words, return_state=False)
    torch.cuda.synchronize()
    for h, st in zip(shards, streams):
        with torch.cuda.stream(st):
            h.process(words, return_state=False)
    torch.cuda.synchronize()
    a = whole.engine.scalars()[:17]
    bad = False
    for r, h in enumerate(shards):
        b = h.engine.scalars()[:17]
        idx = [i for i in (2, 4, 5, 6, 7, 8, 9, 10, 11, 12) if a[i] != b[i]]
        if idx:
            bad = True
            print(f"step {t} shard {r}: " + ", ".join(f"{names[i]} {b[i]} vs {a[i]}" for i in idx), "| XM/XRA/XRT", b[14:17], "status", b[12])
    if bad:
        wk = whole.engine.k
        cur = (int(a[0]) - 1) & 1
        wa = whole.engine.buf["active_cols"][cur * wk:(cur + 1) * wk].cpu().numpy()
        sa = shards[0].engine.buf["active_cols"][cur * wk:(cur + 1) * wk].cpu().numpy()
        print("active columns equal:", np.array_equal(wa, sa))
        break
    if t % 10 == 0 or os.environ.get("VERBOSE"):
        print(f"step {t}: S={a[2]} M={a[4]} L={a[8]} NU={a[10]} NR={a[11]} status={a[12]} | shard0 XM/XRA/XRT "
              f"{shards[0].engine.scalars()[14:17]} status {[int(h.engine.scalars()[12]) for h in shards]}", flush=True)
print("done")
