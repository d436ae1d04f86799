"""Device-side state of one SP+TM network: the ``bh_ctx`` handed to the C ABI and
the single HBM arena it points into.  PyTorch is used for device memory, pinned
host memory and the current stream only -- all arithmetic is in the CUDA library.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _native as nat

_TORCH_DTYPES = None


def _torch():
    import torch

    global _TORCH_DTYPES
    if _TORCH_DTYPES is None:
        _TORCH_DTYPES = {"float64": torch.float64, "float32": torch.float32, "int32": torch.int32,
                         "uint8": torch.uint8, "int64": torch.int64}
    return torch


def require_cuda(device=None):
    torch = _torch()
    if not torch.cuda.is_available():
        raise nat.NativeError("bithtm_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def _round_up(x, m):
    return (x + m - 1) // m * m


class Engine:
    """Owns the arena and the context struct for one network (or one column
    shard of it)."""

    def __init__(self, input_dim, column_dim, cell_dim, active_columns, *, device=None,
                 max_segments=None, max_synapses_per_segment=128, match_capacity=None,
                 learn_capacity=None, rand_capacity=None, ring_len=0, tm_blocks=None,
                 fused="auto", fused_ctas=None, fused_threads=None, column_shard=None, parallel_rng="auto", segment_shard=None,
                 exchange_match_capacity=None, exchange_recycle_capacity=None, lazy_rng="auto", skip_gran=None,
                 skip_min=None, skip_polys=None, tail_chunks=None, exchange_cells="auto", pipeline=None):
        """``column_shard=(rank, world)``: this engine owns columns
        [rank*C/world, (rank+1)*C/world) of the spatial pooler (permanence, mask, duty
        cycles).  ``segment_shard=(rank, world)``: it holds the synapse rows of the
        segments whose 64-id block is dealt to ``rank`` (the per-cell state and the
        temporal-memory bookkeeping are replicated; one exchange per step, see
        ``include/bithtm_b200.h``).  Without ``segment_shard`` the temporal memory is
        replicated whole."""
        torch = _torch()
        self.device = require_cuda(device)
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None) or (
            lambda i: torch.cuda.current_stream(i).cuda_stream)
        if not (1 <= cell_dim <= 32):
            raise NotImplementedError("bithtm_b200 supports 1..32 cells per column (one bit-word per column)")
        I, Ccol, c, k = int(input_dim), int(column_dim), int(cell_dim), int(active_columns)
        N = Ccol * c
        sm, major, minor = C.c_int(0), C.c_int(0), C.c_int(0)
        nat.check(nat.lib.bh_device_info(self._dev_index, C.byref(sm), C.byref(major), C.byref(minor)),
                  "bh_device_info")
        self.sm_count = sm.value
        if max_segments is None:
            max_segments = min(8 * N, 1 << 22)
        max_segments = max(int(max_segments), 64)
        if match_capacity is None:
            match_capacity = max_segments
        if learn_capacity is None:
            learn_capacity = max_segments + k * c
        if rand_capacity is None:  # float64 draws one step may take: rand(k, c) + rand(L, W+1) + rand(M)
            rand_capacity = max(1 << 18, 2 * k * (k + 1) + k * c + min(match_capacity, 4 * k * c))
        if tm_blocks is None:
            tm_blocks = self.sm_count if max_segments <= (1 << 20) else min(1024, self.sm_count * 6)
        ctx = nat.BhCtx()
        ctx.device = self._dev_index  # the library makes it current around every call
        ctx.input_dim, ctx.input_words = I, (I + 31) // 32
        ctx.mask_stride = _round_up(ctx.input_words, 4)
        ctx.column_dim, ctx.cell_dim, ctx.active_columns = Ccol, c, k
        if column_shard is None:
            self.shard_rank, self.shard_world = 0, 1
        else:
            self.shard_rank, self.shard_world = int(column_shard[0]), int(column_shard[1])
            if Ccol % self.shard_world:
                raise ValueError("column_dim must be divisible by the number of column shards")
            if fused != "shard":
                fused = "off"  # the all-gather sits between kernels
        if segment_shard is None:
            self.seg_rank, self.seg_world = 0, 1
        else:
            self.seg_rank, self.seg_world = int(segment_shard[0]), int(segment_shard[1])
            if fused != "shard":
                fused = "off"  # the exchange sits between kernels
        ctx.seg_rank, ctx.seg_world = self.seg_rank, self.seg_world
        if self.seg_world > 1 or fused == "shard":
            ctx.xm_cap = int(exchange_match_capacity or min(int(match_capacity), 8 * k + 1024))
            ctx.xr_cap = int(exchange_recycle_capacity or (2 * k + 64))
        if fused == "shard":
            # exchanges as {word, sequence} cells (csrc/shard_ll.cuh) while the one-CTA selection / merge fit
            # in shared memory; very wide networks use the copy + flag protocol
            # (and while the matching segments of all ranks -- about one per active column in steady state --
            # fit the one-CTA merge: 16384 in all, 4096 per rank).  ``exchange_cells=False`` forces the latter.
            fits = 4 * (6400 + k + Ccol // self.shard_world // 16) <= 150 * 1024 and k <= 4096
            ctx.xch_ll = 1 if (fits if exchange_cells == "auto" else bool(exchange_cells)) else 0
            if ctx.xch_ll and not exchange_match_capacity:
                ctx.xm_cap = min(ctx.xm_cap, 4096)  # what one CTA sorts per rank
        ctx.col_local = Ccol // self.shard_world
        ctx.col_lo = self.shard_rank * ctx.col_local
        self.C_local, self.col_lo = ctx.col_local, ctx.col_lo
        self.k_local = min(k, ctx.col_local)
        ctx.seg_capacity, ctx.syn_capacity = max_segments, _round_up(int(max_synapses_per_segment), 32)
        ctx.match_capacity, ctx.learn_capacity = int(match_capacity), int(learn_capacity)
        ctx.tm_blocks, ctx.sm_count = int(tm_blocks), self.sm_count
        # the stream ring holds raw MT19937 words (2 per float64 draw)
        from . import _mtjump

        step_words = max(4 * int(rand_capacity), 1 << 18)  # most words one step may draw (2x margin)
        typical = 2 * (k * (k + 1) + 2 * k * c)
        huge = step_words > (1 << 29)
        if huge:
            # rand(L, W+1) is quadratic in k (BASELINE configs[4]: 4.4e8 doubles per step at k = 20972): such a
            # network can only step over it (lazy draws, csrc/mt19937.cuh).  The ring must still span one
            # step's draws (absolute stream index -> slot), with a small margin instead of 2x.
            step_words = 2 * int(rand_capacity) + int(rand_capacity) // 8
            if lazy_rng == "auto":
                lazy_rng = "always"
            if parallel_rng == "auto":
                parallel_rng = False  # its jump table would have > 10^5 polynomials
        if parallel_rng == "auto":  # many-CTA stream production once a step draws enough words
            parallel_rng = typical >= 8 * _mtjump.CHUNK_WORDS
        need = step_words + (1 << 24) if huge else 2 * step_words
        if parallel_rng:
            # history for the sparse chunk start (csrc/mt19937.cuh): 19937 << s words, where 623 << s
            # covers one step's production
            depth = _mtjump.DEGREE
            while depth // _mtjump.DEGREE * _mtjump.MAX_SHIFT <= typical + 2 * _mtjump.CHUNK_WORDS:
                depth *= 2
            need = max(need, depth + 2 * step_words)
        ring_words = 1 << 20
        while ring_words < need:
            ring_words <<= 1
        if ring_words > (1 << 31):
            raise NotImplementedError(
                f"{k} active columns draw about {typical // 2:,} random numbers per timestep (rand(L, W+1), "
                "projections.py:120, is quadratic in the number of active columns); the device stream ring is limited "
                "to 2^31 words. Pass rand_capacity= to bound the draws per step explicitly (see DESIGN.md, cfg5).")
        if huge and fused not in ("grid", "shard", "auto"):
            raise NotImplementedError("a network that draws this much per step needs the lazy draws of the cooperative "
                                      'step kernels: fused="grid" or "shard"')
        # whole step as one kernel: on one thread-block cluster while the step is
        # latency-bound (mask <= 8 MiB), else on a cooperative grid with one CTA per SM.  The cluster
        # kernel has no many-CTA stream production phase: a network that draws enough words per step to
        # want it runs on the grid, and an explicit fused="cluster" produces the stream on one CTA.
        if fused == "auto":
            fused = "cluster" if Ccol * ctx.mask_stride * 4 <= (8 << 20) and not parallel_rng else "grid"
        if fused == "cluster":
            parallel_rng = False
        ctx.rng_ring_words, ctx.rng_step_words, ctx.ring_len = ring_words, step_words, int(ring_len)
        ctx.jump_polys = (step_words + step_words // 2) // _mtjump.CHUNK_WORDS + 3 if parallel_rng else 0
        ctx.rng_lookahead = min(2 * (k * c + 4 * k) + 2 * nat.MT_N, step_words // 2) if parallel_rng else 0
        # lazy draws (csrc/mt19937.cuh): rand(L, W+1) is only stepped over and the rows that are read are
        # produced by table jumps -- the cooperative-grid kernels of networks that draw a lot per step
        if lazy_rng == "auto":
            lazy_rng = bool(parallel_rng) and fused in ("grid", "shard")
        if lazy_rng and fused in ("grid", "shard"):
            gran = int(skip_gran) if skip_gran else 4096
            while skip_gran is None and (step_words + step_words // 4) // gran > 4096:
                gran *= 2  # table of at most ~4096 polynomials (2496 B each)
            ctx.skip_gran = gran
            ctx.skip_polys = int(skip_polys) if skip_polys else (step_words + step_words // 4) // gran + 8
            ctx.skip_min = int(skip_min) if skip_min else max(2 * gran, 1 << 16)
            ctx.job_cap = 64 + step_words // _mtjump.WINDOW_WORDS + 2
            ctx.lazy_policy = 1 if lazy_rng == "always" else 0
            # a shard's phases are short: more, shorter tail chunks (one jump each) keep their generation under the scan
            ctx.tail_chunks = int(tail_chunks or (8 if fused == "shard" and self.shard_world > 2 else 0))
        ctx.fused_mode = {"off": 0, "cluster": 1, "grid": 2, "shard": 3}[fused]
        if fused_ctas is None:
            fused_ctas = 16 if fused == "cluster" else self.sm_count
        ctx.fused_ctas = int(fused_ctas)
        # two-pipeline grid kernel (csrc/fused.cuh, k_step_pipe): the spatial pooler of step s+1 beside the
        # temporal memory of step s; ``pipeline`` = CTAs of the temporal-memory team (None: BH_PIPE or off)
        # (a launch of a single step -- the end-to-end host call -- always runs the one-pipeline kernel)
        if pipeline is None:
            pipeline = os.environ.get("BH_PIPE", "auto")
            pipeline = pipeline if pipeline == "auto" else int(pipeline)
        able = (fused == "grid" and Ccol >= 16384) or (fused == "shard" and ctx.xch_ll)
        if pipeline == "auto":
            # steady state: the SP passes scale with the CTAs they get, the TM chain is mostly dependent round trips;
            # the more shards, the less SP work per rank and the larger the TM team should be (measured at cfg3)
            share = 0.31 if fused == "grid" else {1: 0.31, 2: 0.35, 4: 0.62}.get(self.shard_world, 0.75)
            pipeline = int(round(ctx.fused_ctas * share)) if ctx.fused_ctas >= 64 else 0
        ctx.pipe_ctas = int(pipeline) if (pipeline and able and 2 <= int(pipeline) < ctx.fused_ctas) else 0
        # threads per CTA of the cluster kernel: 1024 (one CTA per SM) is fastest for ONE network; several
        # independent networks side by side (StreamBatch) run faster with smaller CTAs sharing the SMs
        if fused_threads is not None:
            if fused != "cluster":
                raise ValueError('fused_threads applies to fused="cluster" only')
            if not (256 <= int(fused_threads) <= 1024 and int(fused_threads) % 32 == 0):
                raise ValueError("fused_threads must be a multiple of 32 in [256, 1024]")
            ctx.fused_threads = int(fused_threads)
        self.ctx = ctx
        self.I, self.C, self.c, self.k, self.N = I, Ccol, c, k, N

        nbytes = nat.lib.bh_layout(C.byref(ctx), None)
        self.arena = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        base = self.arena.data_ptr()
        got = nat.lib.bh_layout(C.byref(ctx), C.c_void_p(base))
        assert got == nbytes
        self.arena_bytes = nbytes
        # typed views into the arena
        counts = self._counts()
        self.buf = {}
        for name, dt in nat.DEVICE_BUFFERS.items():
            ptr = getattr(ctx, name) or 0
            n = counts[name]
            if n == 0:
                self.buf[name] = torch.empty(0, dtype=_TORCH_DTYPES[dt], device=self.device)
                continue
            off = ptr - base
            item = torch.empty(0, dtype=_TORCH_DTYPES[dt]).element_size()
            assert 0 <= off and off + n * item <= nbytes, name
            self.buf[name] = self.arena[off:off + n * item].view(_TORCH_DTYPES[dt])
        # pinned host staging for bh_step_host
        self.input_pinned = torch.zeros(ctx.mask_stride, dtype=torch.int32).pin_memory()
        self.summary_pinned = torch.zeros(nat.summary_ints(k) + 4, dtype=torch.int32).pin_memory()
        ctx.input_pinned = self.input_pinned.data_ptr()
        ctx.summary_pinned = self.summary_pinned.data_ptr()
        self._summary_np = self.summary_pinned.numpy()
        self._summary_out = np.zeros(nat.summary_ints(k), dtype=np.int32)
        self._summary_out_ptr = self._summary_out.ctypes.data  # (ndarray.ctypes builds an object per access)
        if ctx.jump_polys:
            tab = _mtjump.jump_table(ctx.jump_polys, cache_dir=os.path.dirname(nat.LIB_PATH))
            self.buf["mt_jump"].copy_(torch.from_numpy(tab.view(np.int32).reshape(-1)).to(self.device))
        if ctx.skip_polys:
            tab = _mtjump.skip_table(ctx.skip_polys, ctx.skip_gran, cache_dir=os.path.dirname(nat.LIB_PATH))
            self.buf["mt_skip"].copy_(torch.from_numpy(tab.view(np.int32).reshape(-1)).to(self.device))
        nat.check(nat.lib.bh_init(C.byref(ctx), self.stream), "bh_init")
        if self.seg_world > 1:  # this rank's exchange record / the gathered records of all ranks
            n = int(nat.lib.bh_tm_shard_xch_ints(C.byref(ctx)))
            self.xch_send = torch.zeros(n, dtype=torch.int32, device=self.device)
            self.xch_recv = torch.zeros(n * self.seg_world, dtype=torch.int32, device=self.device)
        self.tm_deferred = False  # the last step left a deferred jitter draw / no winner cells (networks.py)
        # stand-alone PredictiveProjection calls take arbitrary cell lists, so the sparse clean-up the fused
        # phases do (by active column) cannot be relied on afterwards: countdown of full clears, see
        # begin_regular_step
        self.standalone_dirty = 0
        self._dirty_seen = -1
        self.epoch = 0  # bumped by every completed step; lazily fetched State fields check it
        self._graphs = {}
        self.host_graph = True  # bh_step_host as one CUDA graph launch (constants are frozen at capture)

    # ------------------------------------------------------------------ plumbing
    def _counts(self):
        x = self.ctx
        C_, I, c, k = x.column_dim, x.input_dim, x.cell_dim, x.active_columns
        CL = x.col_local
        N, S, E, M = C_ * 32, x.seg_capacity, x.syn_capacity, x.match_capacity  # device cell id = col*32+cell
        W = max(1, x.seg_world)
        rows = S if W == 1 else ((S + 63) // 64 + W - 1) // W * 64  # locally held synapse rows
        self.seg_rows = rows
        return {
            "sp_perm": CL * I, "sp_mask": CL * x.mask_stride, "duty": CL, "overlaps": CL, "boosted": CL,
            "active_cols": 3 * k, "col_active": C_, "col_pred": C_, "col_act": 2 * C_, "col_win": C_,
            "cell_nseg": N, "cell_maxjit": N, "cell_npred": N, "cell_widx": N,
            "seg_owner": S, "seg_count": S, "seg_pot": S, "seg_conn": S, "syn_cell": rows * E, "syn_perm": rows * E,
            "row_pred": k, "row_act": k, "row_win": k, "row_unacc": k, "winners": 2 * k * c, "unacc": k * c,
            "m_seg": M, "m_conn": M, "m_jit": M, "m_flag": M, "learn_list": x.learn_capacity, "punish_list": M,
            "recyc_list": W * x.xr_cap if W > 1 else 0,
            "x_send": max((3 * min(k, CL) + 3) // 4 * 4, (4 + 3 * x.xm_cap + x.xr_cap + 3) // 4 * 4)
            + ((16 + 3 * x.xm_cap + x.xr_cap) if x.xch_ll else 0) if x.fused_mode == 3 else 0,
            "xk_keys": W * min(k, CL) if x.fused_mode == 3 else 0,
            "xk_cols": W * min(k, CL) if x.fused_mode == 3 else 0, "blk": 8 * 1024, "topk_ws": 81920, "mt_key": nat.MT_N, "rng_ring": x.rng_ring_words, "mt_jump": x.jump_polys * nat.MT_N,
            "rng64": nat.R_COUNT, "sc": nat.SC_COUNT,
            "mt_skip": x.skip_polys * nat.MT_N, "rng_jump": x.job_cap * 640 if x.skip_polys else 0,
            "grow_list": x.learn_capacity * 3 if x.skip_polys else 0,
            "input_ring": x.ring_len * x.input_words, "input_dev": x.mask_stride,
            "summary_dev": nat.summary_ints(k),
        }

    @property
    def stream(self):
        """torch's current stream on this device, as a raw cudaStream_t."""
        return C.c_void_p(self._raw_stream(self._dev_index))

    @property
    def ref(self):
        return C.byref(self.ctx)

    # device cell ids are column * 32 + cell; the reference's flat ids are column * c + cell
    def cells_to_flat(self, dev_ids: np.ndarray) -> np.ndarray:
        d = np.asarray(dev_ids).astype(np.int64)
        return (d >> 5) * self.c + (d & 31)

    def per_cell(self, name: str) -> np.ndarray:
        """A per-cell device array as the reference's flat [C * c] array."""
        a = self.buf[name].cpu().numpy().reshape(self.C, 32)[:, :self.c]
        return np.ascontiguousarray(a).reshape(-1)

    # ------------------------------------------------------------------ segment shards
    def held_segment_ids(self, S: int) -> np.ndarray:
        """Ids < S of the segments whose synapse rows this rank stores, ascending (= local row order)."""
        from ._shard import held_segment_ids

        return held_segment_ids(S, self.seg_rank, self.seg_world)

    # fused sharded step (fused="shard"): every rank's exchange region must be visible to its peers
    def exchange_region_ints(self) -> int:
        return int(nat.lib.bh_xch_region_ints(self.ref))

    def set_exchange_regions(self, pointers, keepalive=None):
        """pointers[r] = device address of rank r's (zero-filled) exchange region as mapped into THIS
        process; pointers[own rank] is the local one."""
        world = max(1, self.seg_world)
        assert len(pointers) == world <= nat.MAX_RANKS
        for r, p in enumerate(pointers):
            self.ctx.xpeer[r] = int(p)
        self._xch_keepalive = keepalive

    def tm_shard_pre(self, flags=1):
        nat.check(nat.lib.bh_tm_shard_pre(self.ref, int(flags), self.xch_send.data_ptr(), self.stream),
                  "bh_tm_shard_pre")
        return self.xch_send

    def tm_shard_post(self, gathered, want_jitter=True):
        assert gathered.numel() == self.xch_recv.numel() and gathered.dtype == self.xch_recv.dtype
        nat.check(nat.lib.bh_tm_shard_post_ex(self.ref, gathered.data_ptr(), int(bool(want_jitter)), self.stream),
                  "bh_tm_shard_post_ex")
        self.epoch += 1

    def scalars(self) -> np.ndarray:
        return self.buf["sc"].cpu().numpy()

    def check_status(self, sc=None):
        st = int(self.scalars()[nat.SC_STATUS] if sc is None else sc)
        fatal = st & nat.ST_FATAL
        if fatal:
            msgs = [m for bit, m in nat.ST_NAMES.items() if fatal & bit]
            raise nat.NativeError("bithtm_b200 capacity overflow: " + "; ".join(msgs))
        return st

    # ------------------------------------------------------------------ inputs
    def pack_input(self, x) -> "torch.Tensor":
        """bool[I] (numpy, any array-like, or a CUDA tensor) -> packed int32 words on the device."""
        torch = _torch()
        if isinstance(x, torch.Tensor) and x.is_cuda:
            if x.dtype == torch.int32 and x.numel() == self.ctx.input_words:
                return x
            src = x.to(torch.uint8).contiguous()
            out = torch.empty(self.ctx.input_words, dtype=torch.int32, device=self.device)
            nat.check(nat.lib.bh_pack_input(self.ref, src.data_ptr(), out.data_ptr(), self.stream), "bh_pack_input")
            return out
        return torch.from_numpy(self.pack_host(x)).to(self.device)

    def pack_host(self, x) -> np.ndarray:
        xb = np.asarray(x).astype(bool).reshape(-1)
        if xb.size != self.I:
            raise ValueError(f"input has {xb.size} bits, expected {self.I}")
        pad = self.ctx.input_words * 32 - self.I
        if pad:
            xb = np.concatenate([xb, np.zeros(pad, dtype=bool)])
        return np.packbits(xb, bitorder="little").view(np.int32)

    def load_ring(self, inputs):
        """Upload [T, I] bool inputs into the device input ring (T == ring_len)."""
        torch = _torch()
        rows = np.stack([self.pack_host(r) for r in inputs])
        assert rows.shape[0] == self.ctx.ring_len
        self.buf["input_ring"].copy_(torch.from_numpy(rows.reshape(-1)).to(self.device))
        self.buf["sc"][nat.SC_INPUT_POS] = 0

    # ------------------------------------------------------------------ RNG (legacy MT19937 of np.random)
    def set_rng_state(self, key: np.ndarray, pos: int):
        torch = _torch()
        k32 = np.ascontiguousarray(key, dtype=np.uint32).view(np.int32)
        self.buf["mt_key"].copy_(torch.from_numpy(k32).to(self.device))
        self.buf["sc"][nat.SC_MT_POS] = int(pos)
        nat.check(nat.lib.bh_rng_import(self.ref, self.stream), "bh_rng_import")

    def get_rng_state(self):
        nat.check(nat.lib.bh_rng_export(self.ref, self.stream), "bh_rng_export")
        key = self.buf["mt_key"].cpu().numpy().view(np.uint32).copy()
        pos = int(self.buf["sc"][nat.SC_MT_POS].item())
        return key, pos

    def rng_fill(self, count: int) -> np.ndarray:
        torch = _torch()
        out = torch.empty(max(count, 1), dtype=torch.float64, device=self.device)
        nat.check(nat.lib.bh_rng_fill(self.ref, out.data_ptr(), int(count), self.stream), "bh_rng_fill")
        return out[:count].cpu().numpy()

    def begin_regular_step(self):
        """Called at the start of every timestep driven through the network classes.  After stand-alone
        projection calls: step 1 clears the per-column winner words and active flags, step 2 the activation
        buffer that was filled by the stand-alone call (it is this step's, and must start empty)."""
        if not self.standalone_dirty or self._dirty_seen == self.epoch:
            return
        self._dirty_seen = self.epoch
        if self.standalone_dirty == 2:
            self.buf["col_win"].zero_()
            self.buf["col_active"].zero_()
        else:
            half = self.epoch & 1
            self.buf["col_act"][half * self.C:(half + 1) * self.C].zero_()
        self.standalone_dirty -= 1

    # ------------------------------------------------------------------ checkpoint / resume
    _SNAPSHOT = ("sp_perm", "sp_mask", "duty", "overlaps", "boosted", "active_cols", "col_active", "col_pred",
                 "col_act", "col_win", "cell_nseg", "cell_maxjit", "cell_npred", "cell_widx", "seg_owner",
                 "seg_count", "seg_pot", "seg_conn", "syn_cell", "syn_perm", "row_pred", "row_act", "row_win",
                 "row_unacc", "winners", "unacc", "m_seg", "m_conn", "m_jit", "m_flag", "learn_list",
                 "punish_list", "recyc_list")

    def snapshot(self):
        """Everything a resumed run needs, as host tensors: the learned state, the previous timestep's
        context, the scalars and the MT19937 state at the stream cursor."""
        torch = _torch()
        torch.cuda.synchronize(self.device)
        S = int(self.scalars()[nat.SC_NSEG])
        E = self.ctx.syn_capacity
        rows = len(self.held_segment_ids(S))
        out = {"shape": (self.I, self.C, self.c, self.k, self.shard_rank, self.shard_world, self.seg_rank,
                         self.seg_world), "n_segments": S}
        for name in self._SNAPSHOT:
            t = self.buf[name]
            if name in ("seg_owner", "seg_count", "seg_pot", "seg_conn"):
                t = t[:S]
            elif name in ("syn_cell", "syn_perm"):
                t = t[:rows * E].view(rows, E)
            out[name] = t.cpu().clone()
        sc = self.buf["sc"].cpu().clone()
        sc[nat.SC_BAR_COUNT] = 0
        out["sc"] = sc
        out["rng"] = self.get_rng_state()
        out["epoch"], out["tm_deferred"] = self.epoch, self.tm_deferred
        return out

    def restore(self, snap):
        """Inverse of :meth:`snapshot` on an engine of the same shape (capacities may differ)."""
        torch = _torch()
        shape = (self.I, self.C, self.c, self.k, self.shard_rank, self.shard_world, self.seg_rank, self.seg_world)
        if tuple(snap["shape"]) != shape:
            raise ValueError(f"snapshot of a {tuple(snap['shape'])} network cannot be loaded into {shape}")
        S, E = int(snap["n_segments"]), self.ctx.syn_capacity
        if S > self.ctx.seg_capacity or (S > 0 and int(snap["seg_count"].max()) > E):
            raise nat.NativeError("snapshot holds more segments / synapses per segment than this network's capacity")
        for name in self._SNAPSHOT:
            src, dst = snap[name], self.buf[name]
            if name in ("syn_cell", "syn_perm"):
                rows, w = src.shape[0], min(src.shape[1], E)
                dst.zero_()
                dst[:rows * E].view(rows, E)[:, :w].copy_(src[:, :w].to(self.device))
            elif name in ("m_seg", "m_conn", "m_jit", "m_flag", "learn_list", "punish_list"):
                n = min(src.numel(), dst.numel())  # list capacities follow max_segments
                dst[:n].copy_(src[:n].to(self.device))
            else:
                dst.zero_()
                dst[:src.numel()].copy_(src.reshape(-1).to(self.device))
        keep = self.buf["sc"].cpu()
        sc = snap["sc"].clone()
        for i in (nat.SC_BAR_COUNT, nat.SC_BAR_GEN, nat.SC_INPUT_POS):
            sc[i] = keep[i]
        self.buf["sc"].copy_(sc.to(self.device))
        key, pos = snap["rng"]
        self.set_rng_state(key, pos)
        self.epoch, self.tm_deferred = int(snap["epoch"]), bool(snap["tm_deferred"])
        self._graphs_epoch = None

    # ------------------------------------------------------------------ steps
    def step_device(self, words, learning=True):
        """``learning``: bool, or the flag word of ``bh_step`` (1 = learning, 2 = no winner cells)."""
        nat.check(nat.lib.bh_step(self.ref, words.data_ptr(), int(learning), self.stream), "bh_step")
        self.epoch += 1

    def step_host(self, x_bool: np.ndarray, learning=True) -> np.ndarray:
        xb = x_bool
        if not (isinstance(xb, np.ndarray) and xb.dtype == np.bool_ and xb.flags.c_contiguous):
            xb = np.ascontiguousarray(x_bool, dtype=np.bool_)  # one byte per bit, 0/1: what bh_step_host reads
        if xb.size != self.I:
            raise ValueError(f"input has {xb.size} bits, expected {self.I}")
        learning = int(learning)  # bool, or the flag word of bh_step
        key = ("host", learning, bytes(self.ctx))  # kernel arguments are frozen in the graph
        handle = self._graphs.get(key)
        if handle is None and self.host_graph:
            # one CUDA graph = H2D copy node, the step, D2H copy node (captured on a side stream)
            torch = _torch()
            handle = C.c_void_p()
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                nat.check(nat.lib.bh_host_graph_create(self.ref, int(learning), self.stream, C.byref(handle)),
                          "bh_host_graph_create")
            torch.cuda.current_stream(self.device).wait_stream(side)
            self._graphs[key] = handle
        if handle is not None:
            nat.check(nat.lib.bh_step_host_graph(self.ref, handle, xb.__array_interface__["data"][0],
                                                 self._summary_out_ptr, self.stream), "bh_step_host_graph")
        else:
            nat.check(nat.lib.bh_step_host(self.ref, xb.__array_interface__["data"][0], int(learning),
                                           self._summary_out_ptr, self.stream), "bh_step_host")
        self.epoch += 1
        return self._summary_out

    def profile_step(self, words, learning=True):
        """One step with a CUDA event after every launch -> [(kernel name, milliseconds)]."""
        ms = (C.c_float * 48)()
        names = (C.c_char_p * 48)()
        n = nat.lib.bh_profile_step(self.ref, words.data_ptr(), int(bool(learning)), self.stream, ms, names, 48)
        if n < 0:
            nat.check(n, "bh_profile_step")
        self.epoch += 1
        return [(names[i].decode(), float(ms[i])) for i in range(n)]

    def summary(self) -> np.ndarray:
        nat.check(nat.lib.bh_summary(self.ref, self._summary_out.ctypes.data, self.stream), "bh_summary")
        return self._summary_out

    def graph(self, steps_per_graph: int, learning=True):
        key = (steps_per_graph, int(learning))
        if key not in self._graphs:
            torch = _torch()
            handle = C.c_void_p()
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                nat.check(nat.lib.bh_graph_create(self.ref, int(steps_per_graph), int(learning),
                                                  self.stream, C.byref(handle)), "bh_graph_create")
            torch.cuda.current_stream(self.device).wait_stream(side)
            self._graphs[key] = handle
        return self._graphs[key]

    def launch_graph(self, handle, steps_per_graph: int):
        nat.check(nat.lib.bh_graph_launch(handle, self.stream), "bh_graph_launch")
        self.epoch += steps_per_graph

    def __del__(self):
        try:
            for h in getattr(self, "_graphs", {}).values():
                nat.lib.bh_graph_destroy(h)
        except Exception:
            pass
