"""Names of the globaltimer phase stamps the fused step kernels leave in ctx.blk row 7 (CTA 0, last step)."""
import numpy as np

FUSED_GRID = ["overlap + draw 1", "top-k", "SP learn + duty + winner bits", "winner lists + learning flags",
              "learning lists + draw 2", "stream production", "learn", "post", "segment scan", "draw 3",
              "matching list + jitter + predictions"]
SHARD = ["overlap", "local top-k", "candidate record", "exchange 1", "unpack + global top-k", "SP learn + winner bits",
         "lists + learning flags", "learning lists + draw 2", "stream production", "learn + post", "segment scan",
         "record", "exchange 2", "merge", "draw 3 + jitter + predictions"]


def read(eng):
    names = SHARD if eng.ctx.fused_mode == 3 else FUSED_GRID
    st = eng.buf["blk"][7 * 1024:7 * 1024 + 2 * (len(names) + 1)].cpu().numpy().view(np.uint64).astype(np.float64)
    return {n: round(float(v) / 1e3, 2) for n, v in zip(names, np.diff(st))}


LL_SELECT = ["send histogram", "gather histograms", "threshold bin", "per-rank counts", "collect + send columns",
             "gather columns", "rank members + write list"]
LL_SEGS = ["sort + send record", "gather headers", "merge"]


def read_ll(eng):
    """Sub-phases of the two one-CTA exchange phases of the fused sharded step (csrc/shard_ll.cuh), last step."""
    raw = eng.buf["blk"][7 * 1024:7 * 1024 + 2 * 64].cpu().numpy().view(np.uint64).astype(np.float64)
    sel = {n: round(float(v) / 1e3, 2) for n, v in zip(LL_SELECT, np.diff(raw[40:47]))}
    seg = {n: round(float(v) / 1e3, 2) for n, v in zip(LL_SEGS, np.diff(raw[52:56]))}
    return {"selection": sel, "segments": seg}
