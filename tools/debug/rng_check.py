"""Debug: run one grid-kernel network and compare the device stream ring with the MT19937 stream numpy
would produce, region by region, after every step."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bithtm_b200 as bithtm
from bithtm_b200 import _mtjump
from bithtm_b200.projections import DenseProjection

k, steps = int(sys.argv[1]), int(sys.argv[2])
I, C, c = 4096, 32768, 32
g = np.random.default_rng(3)
base = g.random((20, I)) < 0.2
xs = base[np.arange(steps) % 20] ^ (g.random((steps, I)) < 0.05)
gen = torch.Generator(device="cuda")
gen.manual_seed(99)
perm = torch.randn(C, I, dtype=torch.float64, device="cuda", generator=gen) * 0.1
np.random.seed(4)
sp = bithtm.SpatialPooler(I, C, k, proximal_projection=DenseProjection(I, C, permanence=perm))
htm = bithtm.HierarchicalTemporalMemory(I, C, c, k, spatial_pooler=sp, rng_sync="lazy", max_segments=1 << 18,
                                        max_synapses_per_segment=128, fused="grid",
                                        lazy_rng=os.environ.get("LAZYMODE", "auto"),
                                        **({"fused_ctas": int(os.environ["CTAS"])} if os.environ.get("CTAS") else {}))
eng = htm.engine
link = htm.temporal_memory._rng
link.before(eng)
key = link._key.copy()
total = steps * (2 * k * (k + 1) + 4 * k * c) + 1000000
print("expected stream words", total, flush=True)
want = _mtjump.raw_stream(key, total)
ringw = eng.ctx.rng_ring_words
ring = eng.buf["rng_ring"]


def check(lo, hi, what, t):
    if hi <= lo:
        return True
    idx = (torch.arange(lo, hi, device="cuda") & (ringw - 1))
    got = ring[idx].cpu().numpy().view(np.uint32)
    bad = np.flatnonzero(got != want[lo:hi])
    if bad.size:
        print(f"step {t}: {what} [{lo}, {hi}): {bad.size} wrong words, first at +{bad[0]} (abs {lo + bad[0]}), last at +{bad[-1]}; "
              f"got {got[bad[0]]:#x} want {want[lo + bad[0]]:#x}")
        return False
    return True


for t in range(steps):
    htm.process(eng.pack_input(xs[t]), return_state=False)
    torch.cuda.synchronize()
    r = eng.buf["rng64"].cpu().numpy()
    sc = eng.scalars()
    produced, cursor, off1, off2, off3, n2, n3 = (int(v) for v in r[:7])
    region_lo, dense_lo, lazy, jbase, tail, tchunks, twords = (int(v) for v in r[12:19])
    kc2 = 2 * k * c
    ok = check(off1, off1 + kc2, "draw1", t)
    ok &= check(off3, off3 + 2 * n3, "draw3", t)
    ok &= check(max(region_lo, cursor), produced, "region after cursor", t)
    print(f"step {t}: lazy={lazy} G={sc[23]} L={sc[8]} W={sc[5 + ((t) & 1)]} M={sc[4]} S={sc[2]} off2={off2} n2={n2} tail={tail} chunks={tchunks}x{twords} "
          f"produced={produced} cursor={cursor} region_lo={region_lo} {'OK' if ok else 'BAD'}", flush=True)
    if not ok:
        jb = eng.buf["rng_jump"].cpu().numpy().reshape(-1, 640)
        nz = [(i, int(np.count_nonzero(jb[i]))) for i in range(jb.shape[0]) if np.count_nonzero(jb[i])]
        print("nonzero jump slots:", nz[:40])
        gl = eng.buf["grow_list"][:3 * int(sc[23])].cpu().numpy().reshape(-1, 3)
        print("grow list rows:", gl[:, 0].tolist())
        rw = 2 * (int(sc[5 + ((t + 1) & 1)]) + 1)
        for row in gl[:6, 0]:
            check(off2 + int(row) * rw, off2 + (int(row) + 1) * rw, f"row {row}", t)
        break
