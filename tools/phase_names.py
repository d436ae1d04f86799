"""Names of the globaltimer phase stamps the fused step kernels leave in ctx.blk row 7 (CTA 0, last step)."""
import numpy as np

FUSED_GRID = ["overlap + draw 1", "top-k", "SP learn + duty + winner bits", "winner lists + learning flags",
              "learning lists + draw 2", "stream production", "learn", "post", "segment scan", "draw 3",
              "matching list + jitter + predictions"]
SHARD = ["overlap", "local top-k", "candidate record", "exchange 1", "unpack + global top-k", "SP learn + winner bits",
         "lists + learning flags", "learning lists + draw 2", "stream production", "learn + post", "segment scan",
         "record", "exchange 2", "merge", "draw 3 + jitter + predictions"]


PIPE_SP = ["SP learn + duty (step s)", "overlap + histogram (step s+1)", "selection (step s+1)", "waits for the TM team"]
PIPE_TM = ["draw 1", "winner bits", "winner lists + learning flags", "learning lists + draw 2", "stream production",
           "learn", "segment scan", "draw 3", "matching list + jitter + predictions", "waits for the SP team"]


PIPE_SHARD_SP = ["SP learn + duty (step s)", "overlap + histogram (step s+1)", "selection exchange (step s+1)",
                 "waits for the TM team"]
PIPE_SHARD_TM = ["draw 1", "bookkeeping (sub-team)", "stream production", "learn", "segment scan", "segment exchange + merge",
                 "draw 3 + jitter + predictions", "waits for the SP team"]


def read(eng):
    if eng.ctx.fused_mode == 3 and eng.ctx.pipe_ctas > 0:
        raw = eng.buf["blk"][7 * 1024:7 * 1024 + 2 * 80].cpu().numpy().view(np.uint64).astype(np.float64)
        sp = {n: round(float(v) / 1e3, 2) for n, v in zip(PIPE_SHARD_SP, np.diff(raw[2:7]))}
        # (stamp 1 -- after draw 1 -- is taken inside the bookkeeping sub-team, which the stamping CTA is not part
        # of when the TM team is larger than the sub-team: draw 1 and the bookkeeping are reported together)
        tmr = np.concatenate([raw[64:65], raw[66:73]])
        tm = {n: round(float(v) / 1e3, 2) for n, v in zip(["draw 1 + bookkeeping (sub-team)"] + PIPE_SHARD_TM[2:], np.diff(tmr))}
        return {"SP team": sp, "TM team": tm}
    if eng.ctx.fused_mode == 2 and eng.ctx.pipe_ctas > 0:
        # two-pipeline kernel: the last pipelined iteration of the last launch; both teams start together
        raw = eng.buf["blk"][7 * 1024:7 * 1024 + 2 * 80].cpu().numpy().view(np.uint64).astype(np.float64)
        sp = {n: round(float(v) / 1e3, 2) for n, v in zip(PIPE_SP, np.diff(raw[2:7]))}
        tm = {n: round(float(v) / 1e3, 2) for n, v in zip(PIPE_TM, np.diff(raw[64:75]))}
        return {"SP team": sp, "TM team": tm}
    names = SHARD if eng.ctx.fused_mode == 3 else FUSED_GRID
    st = eng.buf["blk"][7 * 1024:7 * 1024 + 2 * (len(names) + 1)].cpu().numpy().view(np.uint64).astype(np.float64)
    return {n: round(float(v) / 1e3, 2) for n, v in zip(names, np.diff(st))}


LL_SELECT = ["send histogram", "gather histograms", "threshold bin + per-rank counts", "collect + send columns",
             "gather columns", "rank members + write list"]
LL_SEGS = ["sort + send record", "gather headers", "merge"]


def read_ll(eng):
    """Sub-phases of the two one-CTA exchange phases of the fused sharded step (csrc/shard_ll.cuh), last step."""
    raw = eng.buf["blk"][7 * 1024:7 * 1024 + 2 * 64].cpu().numpy().view(np.uint64).astype(np.float64)
    sel = {n: round(float(v) / 1e3, 2) for n, v in zip(LL_SELECT, np.diff(raw[40:47]))}
    seg = {n: round(float(v) / 1e3, 2) for n, v in zip(LL_SEGS, np.diff(raw[52:56]))}
    return {"selection": sel, "segments": seg}
