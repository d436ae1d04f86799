"""Projection plugins -- device-backed mirrors of ``bithtm/projections.py``.

``DenseProjection`` (SP proximal synapses, projections.py:6-24) and
``PredictiveProjection`` (TM distal segments, projections.py:194-293, built in the
reference on ``SparseProjection`` :27-192).  Same class names, constructor
arguments and defaults.  The arithmetic runs in ``libbithtm_b200.so``; these
classes hold hyper-parameters, evaluate the constants with the reference's own
Python expressions, and materialise results as NumPy arrays in the reference's
dtypes when they are read.
"""

from __future__ import annotations

import numpy as np

from . import _native as nat
from .regularizations import eng_set_active


class DenseProjection:
    """projections.py:6-24.  The float64 permanence matrix and its bit-packed
    connected mask live in HBM; ``permanence`` downloads a copy."""

    def __init__(self, input_dim, output_dim, permanence_mean=0.0, permanence_std=0.1,
                 permanence_threshold=0.0, permanence_increment=0.03, permanence_decrement=0.015,
                 *, permanence=None):
        """``permanence`` (extension, keyword-only): a ready [output_dim, input_dim] float64
        matrix (NumPy array or CUDA tensor) to use instead of drawing one -- e.g. a matrix
        drawn on the device for sizes where the host draw (8 GiB at 65536 x 16384) is the
        bottleneck.  When given, the global np.random stream is NOT consumed."""
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.permanence_threshold = permanence_threshold
        self.permanence_increment = permanence_increment
        self.permanence_decrement = permanence_decrement
        # projections.py:16 -- the identical call, so the global np.random stream is
        # consumed exactly as the reference consumes it (float64, first consumer).
        if permanence is None:
            self._host_permanence = np.random.randn(output_dim, input_dim) * permanence_std + permanence_mean
        else:
            if permanence.shape[1] != input_dim or permanence.shape[0] > output_dim:
                raise ValueError("permanence must have shape (output_dim, input_dim) "
                                 "(or this rank's rows of it when column-sharded)")
            self._host_permanence = permanence
        self._engine = None

    def _constants(self):
        both = self.permanence_increment + self.permanence_decrement
        return dict(
            sp_threshold=float(self.permanence_threshold),  # :19
            sp_delta_on=float(1.0 * both - self.permanence_decrement),  # :24 with input bit True
            sp_delta_off=float(0.0 * both - self.permanence_decrement),  # :24 with input bit False
        )

    def _bind(self, engine):
        import torch

        if self._engine is engine:
            return
        if self._engine is not None:
            raise NotImplementedError("this DenseProjection is already attached to another network")
        if (engine.I, engine.C) != (self.input_dim, self.output_dim):
            raise ValueError("DenseProjection shape does not match the network")
        self._engine = engine
        for k, v in self._constants().items():
            setattr(engine.ctx, k, v)
        src = self._host_permanence
        if engine.shard_world > 1 and src.shape[0] == self.output_dim:
            src = src[engine.col_lo:engine.col_lo + engine.C_local]  # this rank's rows
        if tuple(src.shape) != (engine.C_local, self.input_dim):
            raise ValueError("permanence rows do not match this rank's column shard")
        if isinstance(src, torch.Tensor):
            engine.buf["sp_perm"].copy_(src.reshape(-1).to(device=engine.device, dtype=torch.float64))
        else:
            engine.buf["sp_perm"].copy_(torch.from_numpy(np.ascontiguousarray(src, dtype=np.float64).reshape(-1))
                                        .to(engine.device))
        self._host_permanence = None
        nat.check(nat.lib.bh_sp_build_mask(engine.ref, engine.stream), "bh_sp_build_mask")

    def _need_engine(self):
        if self._engine is None:
            raise RuntimeError("DenseProjection is not attached to a network yet; construct a "
                               "bithtm_b200 SpatialPooler with it (proximal_projection=...) first")
        return self._engine

    @property
    def permanence(self):
        if self._engine is None:
            return self._host_permanence
        eng = self._engine  # column-sharded: this rank's rows
        return eng.buf["sp_perm"].cpu().numpy().reshape(eng.C_local, self.input_dim)

    def process(self, input_activation):
        """projections.py:18-21 -> int64 overlaps."""
        eng = self._need_engine()
        words = eng.pack_input(input_activation)
        nat.check(nat.lib.bh_sp_overlap(eng.ref, words.data_ptr(), eng.stream), "bh_sp_overlap")
        return eng.buf["overlaps"].cpu().numpy().astype(np.int64)

    def process_batch(self, input_activations, tensor_core=True):
        """Extension: ``process`` (projections.py:18-21) for a batch of inputs [B, input_dim] against
        this one projection -> int64 overlaps [B, output_dim] (a CUDA tensor in, a CUDA tensor out
        when given one).  Column-sharded: this rank's columns.  ``tensor_core``: ``True`` = the int8
        tensor-core contraction, tcgen05 + TMEM + TMA for all but small or oddly sized problems
        (``bh_sp_overlap_batched_tc``); ``"tcgen05"`` / ``"mma"`` force one of its two kernels; ``False`` =
        AND + popcount on the integer pipe (``bh_sp_overlap_batched``).  All are exact and agree bit for bit."""
        import torch

        eng = self._need_engine()
        is_cuda = isinstance(input_activations, torch.Tensor) and input_activations.is_cuda
        x = input_activations if is_cuda else torch.from_numpy(np.ascontiguousarray(input_activations)).to(eng.device)
        x = x.reshape(-1, self.input_dim).to(torch.uint8).contiguous()  # one byte per bit
        B, words = x.shape[0], eng.ctx.input_words
        packed = torch.empty(B, words, dtype=torch.int32, device=eng.device)
        nat.check(nat.lib.bh_pack_inputs(eng.ref, x.data_ptr(), B, words, packed.data_ptr(), eng.stream), "bh_pack_inputs")
        out = torch.empty(B, eng.C_local, dtype=torch.int32, device=eng.device)
        if tensor_core and words * 32 >= 1 << 24:
            raise ValueError("process_batch(tensor_core=True) supports fewer than 2**24 input bits; "
                             "pass tensor_core=False")
        name = {True: "bh_sp_overlap_batched_tc", "tcgen05": "bh_sp_overlap_batched_tc5",
                "mma": "bh_sp_overlap_batched_mma", False: "bh_sp_overlap_batched"}[tensor_core]
        nat.check(getattr(nat.lib, name)(eng.ref, packed.data_ptr(), B, out.data_ptr(), eng.stream), name)
        if tensor_core in (True, "tcgen05") and int(eng.buf["sc"][nat.SC_T5_ERR].item()):
            raise nat.NativeError("bithtm_b200: the tcgen05 batched overlap timed out on a pipeline barrier")
        out = out.to(torch.int64)
        return out if is_cuda else out.cpu().numpy()

    def update(self, input_activation, learning_output):
        """projections.py:23-24."""
        eng = self._need_engine()
        for k, v in self._constants().items():
            setattr(eng.ctx, k, v)
        words = eng.pack_input(input_activation)
        eng_set_active(eng, learning_output)
        nat.check(nat.lib.bh_sp_learn(eng.ref, words.data_ptr(), eng.stream), "bh_sp_learn")


class _Lazy:
    """Fetch-on-first-read attribute holder tied to an engine epoch: device
    buffers are overwritten by the next timestep, so a stale read raises instead
    of returning another step's data."""

    def __init__(self, engine):
        self._engine = engine
        self._epoch = engine.epoch
        self._cache = {}

    def _get(self, name, fn):
        if name not in self._cache:
            if self._engine.epoch != self._epoch:
                raise RuntimeError(
                    f"State.{name} was not read before the next timestep ran; its device buffer has been "
                    "overwritten. Read it (or call .materialize()) right after process().")
            self._cache[name] = fn()
        return self._cache[name]


class PredictiveProjection:
    """projections.py:194-293.  Segment store in HBM: compact rows of
    (presynaptic cell, float32 permanence), per-segment owner cell and count."""

    class State(_Lazy):
        """projections.py:195-203; every field is fetched from the device on first read."""

        def __init__(self, engine, projection, have_jitter=True):
            super().__init__(engine)
            self._p = projection
            self._have_jitter = have_jitter  # False: return_jittered_potential_info=False (projections.py:253)

        def _scalars(self):
            return self._get("_sc", self._engine.scalars)

        @property
        def _S(self):
            return int(self._scalars()[nat.SC_NSEG])

        @property
        def _M(self):
            return int(self._scalars()[nat.SC_M])

        @property
        def prediction(self):  # :251 float64 count of active segments per cell
            return self._get("prediction", lambda: self._engine.per_cell("cell_npred").astype(np.float64))

        @property
        def segment_potential(self):  # :246 int64 [S]
            def fetch():
                eng = self._engine
                pot = eng.buf["seg_pot"][:self._S].cpu().numpy().astype(np.int64)
                if eng.seg_world > 1:  # COLLECTIVE: each rank scanned only the segments it stores
                    import torch
                    import torch.distributed as dist

                    held = np.zeros(self._S, dtype=np.int64)
                    ids = eng.held_segment_ids(self._S)
                    held[ids] = pot[ids]
                    t = torch.from_numpy(held).to(eng.device)
                    dist.all_reduce(t, group=self._p._group)
                    pot = t.cpu().numpy()
                return pot

            return self._get("segment_potential", fetch)

        @property
        def matching_segment(self):  # :247 int64 [M], ascending
            return self._get("matching_segment",
                             lambda: self._engine.buf["m_seg"][:self._M].cpu().numpy().astype(np.int64))

        @property
        def matching_segment_activation(self):  # :249
            return self._get("matching_segment_activation",
                             lambda: self._engine.buf["m_conn"][:self._M].cpu().numpy().astype(np.int64))

        @property
        def matching_segment_active(self):  # :250
            return self.matching_segment_activation >= self._p.segment_activation_threshold

        @property
        def max_jittered_potential(self):  # :236-238 float32 [N]; None until drawn (projections.py:196-203)
            if not self._have_jitter:
                return None
            return self._get("max_jittered_potential", lambda: self._engine.per_cell("cell_maxjit"))

        @property
        def matching_segment_jittered_potential(self):  # :234-235 float32 [M]
            if not self._have_jitter:
                return None
            return self._get("matching_segment_jittered_potential",
                             lambda: self._engine.buf["m_jit"][:self._M].cpu().numpy())

        def materialize(self):
            for n in ("prediction", "segment_potential", "matching_segment", "matching_segment_activation",
                      "max_jittered_potential", "matching_segment_jittered_potential"):
                getattr(self, n)
            return self

    def __init__(self, output_dim, permanence_initial=0.21, permanence_threshold=0.5, permanence_increment=0.1,
                 permanence_decrement=0.1, permanence_punishment=0.01, segment_activation_threshold=15,
                 segment_matching_threshold=15, segment_sampling_synapses=32,
                 segment_bundle_growth_exponential=True, *, cell_dim=None, active_columns=None, **engine_kwargs):
        """``cell_dim`` / ``active_columns`` (extensions, keyword-only): let a projection that is NOT owned by
        a bithtm_b200 TemporalMemory allocate its own device state on first use -- e.g. when it is plugged
        into the reference's TemporalMemory (networks.py:48-55), which calls ``process`` / ``update`` itself.
        ``active_columns`` bounds the winner / active cell lists (active_columns * cell_dim cells)."""
        assert segment_activation_threshold >= segment_matching_threshold  # projections.py:211
        self.output_dim = output_dim
        self._cell_dim, self._active_columns, self._engine_kwargs = cell_dim, active_columns, engine_kwargs
        self._rng_link = None
        self._winners_epoch = -1  # engine epoch at which update() supplied the winner cells
        self.permanence_initial = permanence_initial
        self.permanence_threshold = permanence_threshold
        self.permanence_increment = permanence_increment
        self.permanence_decrement = permanence_decrement
        self.permanence_punishment = permanence_punishment
        self.segment_activation_threshold = segment_activation_threshold
        self.segment_matching_threshold = segment_matching_threshold
        self.segment_sampling_synapses = segment_sampling_synapses
        self._engine = None
        self._group = None  # torch.distributed group of the segment shards

    @staticmethod
    def _deltas(active_change, inactive_change):
        # projections.py:102: bool array * float + float -> float64, evaluated per bit value
        on = (np.array([True]) * (active_change - inactive_change) + inactive_change)[0]
        off = (np.array([False]) * (active_change - inactive_change) + inactive_change)[0]
        return float(on), float(off), int(min(active_change, inactive_change) < 0)  # :105

    def _constants(self, epsilon=1e-8):
        l_on, l_off, l_del = self._deltas(self.permanence_increment, -self.permanence_decrement)  # :287
        p_on, p_off, p_del = self._deltas(-self.permanence_punishment, 0.0)  # :292
        return dict(
            tm_learn_on=l_on, tm_learn_off=l_off, tm_learn_can_delete=l_del,
            tm_punish_on=p_on, tm_punish_off=p_off, tm_punish_can_delete=p_del,
            tm_perm_initial=float(np.float32(self.permanence_initial)),  # :149
            tm_perm_threshold=float(np.float32(self.permanence_threshold)),  # :171
            epsilon=float(np.float32(epsilon)),
            seg_activation_threshold=int(self.segment_activation_threshold),
            seg_matching_threshold=int(self.segment_matching_threshold),
            seg_sampling_synapses=int(self.segment_sampling_synapses),
        )

    def _bind(self, engine, epsilon=1e-8):
        if self._engine is not None and self._engine is not engine:
            raise NotImplementedError("this PredictiveProjection is already attached to another network")
        if engine.N != self.output_dim:
            raise ValueError("PredictiveProjection output_dim does not match column_dim * cell_dim")
        self._engine = engine
        for k, v in self._constants(epsilon).items():
            setattr(engine.ctx, k, v)

    # ---- read-only views of the learned state, in the reference's vocabulary -------
    @property
    def bundle_segments(self):  # projections.py:227 int32 [N]
        if self._engine is None:
            return np.zeros(self.output_dim, dtype=np.int32)
        return self._engine.per_cell("cell_nseg")

    @property
    def n_segments(self):
        return 0 if self._engine is None else int(self._engine.scalars()[nat.SC_NSEG])

    @property
    def segment_bundle(self):  # projections.py:226 int32 [S, 1]
        S = self.n_segments
        if S == 0:
            return np.zeros((0, 1), dtype=np.int32)
        eng = self._engine
        return eng.cells_to_flat(eng.buf["seg_owner"][:S].cpu().numpy()).astype(np.int32).reshape(S, 1)

    def export_local_segments(self):
        """(ids, count, cells, perm) of the segments whose synapse rows THIS rank stores
        (all of them when not segment-sharded), ids ascending."""
        S = self.n_segments
        eng = self._engine
        E = eng.ctx.syn_capacity
        ids = eng.held_segment_ids(S)
        n = len(ids)
        count = eng.buf["seg_count"][:S].cpu().numpy()[ids]
        cells = eng.cells_to_flat(eng.buf["syn_cell"][:n * E].cpu().numpy()).reshape(n, E)
        perm = eng.buf["syn_perm"][:n * E].cpu().numpy().reshape(n, E).copy()
        free = np.arange(E)[None, :] >= count[:, None]
        cells[free] = -1
        perm[free] = -1.0
        return ids, count, cells, perm

    def export_segments(self, parts=None):
        """(owner[S], count[S], cells[S, E], perm[S, E]) with free slots = -1 / -1.0 --
        the row form ``oracle.digest.canonical_from_rows`` and an export shim consume.
        Segment-sharded: a COLLECTIVE (every rank contributes its rows), unless the
        per-rank ``export_local_segments()`` results are passed as ``parts``."""
        S = self.n_segments
        eng = self._engine
        owner = eng.cells_to_flat(eng.buf["seg_owner"][:S].cpu().numpy())
        if eng.seg_world == 1:
            _, count, cells, perm = self.export_local_segments()
            return owner, count, cells, perm
        from ._shard import merge_segment_parts

        if parts is None:
            import torch.distributed as dist

            parts = [None] * eng.seg_world
            dist.all_gather_object(parts, self.export_local_segments(), group=self._group)
        count, cells, perm = merge_segment_parts(parts)
        return owner, count, cells, perm

    @property
    def segment_projection(self):
        """Read-only snapshot of the synapse store in the reference's storage vocabulary
        (``SparseProjection``, projections.py:27-68): what
        ``reference_implementations.TemporalMemory.copy_custom`` reads
        (reference_implementations.py:51-70)."""
        return SegmentProjectionView(self)

    # ---- the plugin methods themselves (projections.py:229-293), with explicit arguments ----
    # TemporalMemory.process here runs the whole timestep in fused kernels; these are the same device
    # phases behind the reference's own method signatures, for callers that orchestrate a timestep
    # themselves the way networks.py:91-128 does.
    def _need_engine(self):
        if self._engine is None:
            if self._cell_dim is None or self._active_columns is None:
                raise RuntimeError("PredictiveProjection is not attached to a network: construct a bithtm_b200 "
                                   "TemporalMemory with it, or pass cell_dim= and active_columns= to use it "
                                   "stand-alone")
            from ._engine import Engine

            self._bind(Engine(1, self.output_dim // self._cell_dim, self._cell_dim, self._active_columns,
                              fused="off", **self._engine_kwargs))
        if self._engine.seg_world > 1:
            raise NotImplementedError("stand-alone process / update are not available on segment shards")
        if self._rng_link is None:
            from ._rnglink import _RngLink

            self._rng_link = _RngLink("step")
        return self._engine

    def _device_cells(self, flat):
        """Reference flat cell ids (column * cell_dim + cell) -> device int32 tensor of column * 32 + cell."""
        import torch

        eng = self._engine
        f = np.asarray(flat, dtype=np.int64).reshape(-1)
        if f.size and (f.min() < 0 or f.max() >= eng.N):
            raise ValueError("cell index out of range")
        dev = ((f // eng.c) * 32 + f % eng.c).astype(np.int32)
        return torch.from_numpy(dev).to(eng.device), int(f.size)

    def fill_jittered_potential_info(self, state, matching_segment_bundle=None):
        """projections.py:229-239: draws rand(M) for an activation that was asked not to."""
        eng = self._need_engine()
        if state._have_jitter:
            return
        if state._epoch != eng.epoch:
            raise RuntimeError("only the latest activation state can still draw its jitter")
        self._rng_link.before(eng)
        nat.check(nat.lib.bh_tm_fill_jitter(eng.ref, eng.stream), "bh_tm_fill_jitter")
        self._rng_link.after(eng)
        state._have_jitter = True

    def get_jittered_potential_info(self, state, matching_segment_bundle=None):
        """projections.py:241-243."""
        self.fill_jittered_potential_info(state, matching_segment_bundle)
        return state.max_jittered_potential, state.matching_segment_jittered_potential

    def process(self, active_input, return_jittered_potential_info=True):
        """projections.py:245-255 for an explicit list of active cells (flat ids).  Completes a timestep
        on the device: an ``update`` call that belongs to the same timestep must come first, as in
        networks.py:106-121."""
        eng = self._need_engine()
        cells, n = self._device_cells(active_input)
        jit = bool(return_jittered_potential_info)
        self._rng_link.before(eng)
        nat.check(nat.lib.bh_tm_activate_cells(eng.ref, cells.data_ptr(), n, int(jit),
                                               int(self._winners_epoch == eng.epoch), eng.stream),
                  "bh_tm_activate_cells")
        eng.epoch += 1
        eng.tm_deferred = True  # row lists of the fused summary were not formed: next step takes the staged path
        eng.standalone_dirty = 2
        if jit:
            self._rng_link.after(eng)
        eng.check_status()
        self._last_state = self.State(eng, self, jit)
        return self._last_state

    def update(self, prev_state, input_activation, learning_output, output_punishment, winner_input=None,
               output_learning=None, epsilon=1e-8):
        """projections.py:257-293.  ``prev_state`` must be the State of the latest activation (the device
        holds exactly one); ``output_punishment`` must be uniform within a column (it is
        ``np.repeat(column_punishment, cell_dim)`` at the reference's only call site, networks.py:111)."""
        if prev_state is None:  # :258-259
            return
        eng = self._need_engine()
        import torch

        if getattr(prev_state, "_engine", None) is not eng or prev_state._epoch != eng.epoch:
            raise NotImplementedError("prev_state must be the State this projection's latest activation returned "
                                      "(the previous distal state lives on the device)")
        if output_learning is not None:
            raise NotImplementedError("output_learning= is derived from learning_output on the device")
        C_, c = eng.C, eng.c
        act = np.asarray(input_activation, dtype=bool).reshape(C_, c)
        pun = np.asarray(output_punishment, dtype=bool).reshape(C_, c)
        if not (pun == pun[:, :1]).all():
            raise NotImplementedError("output_punishment must be the same for all cells of a column")
        padded = np.zeros((C_, 32), dtype=bool)
        padded[:, :c] = act
        words = np.packbits(padded, axis=1, bitorder="little").view(np.uint32).reshape(-1).view(np.int32)
        col_active = (~pun[:, 0]).astype(np.uint8)
        win, n_win = self._device_cells(learning_output)
        if winner_input is None:
            prev_win, n_prev = win, -1
        else:
            prev_win, n_prev = self._device_cells(winner_input)
        if max(n_win, n_prev) > eng.k * c:
            raise ValueError("more winner cells than active_columns * cell_dim")
        self._bind(eng, epsilon)
        w_dev = torch.from_numpy(words).to(eng.device)
        a_dev = torch.from_numpy(col_active).to(eng.device)
        self._rng_link.before(eng)
        nat.check(nat.lib.bh_tm_learn_args(eng.ref, win.data_ptr(), n_win, prev_win.data_ptr(), n_prev,
                                           w_dev.data_ptr(), a_dev.data_ptr(), eng.stream), "bh_tm_learn_args")
        self._rng_link.after(eng)
        eng.check_status()
        self._winners_epoch = eng.epoch
        if not prev_state._have_jitter:  # the update drew it (get_jittered_potential_info, :263)
            prev_state._have_jitter = True

    # ---- state import: the inverse of export_segments (checkpoint / resume of the learned state) ----
    def import_segments(self, owner, count, cells, perm):
        """Load a segment store in the row form ``export_segments`` returns (or what
        ``reference_implementations.TemporalMemory.copy_custom`` reads, reference_implementations.py:51-70):
        owner[S] flat cell of each segment, count[S] valid synapses, cells[S, E'] presynaptic flat cells with
        negative = free slot, perm[S, E'] float32.  Rows are re-compacted; the previous timestep's context is
        forgotten (``bh_tm_reset``), the learned state is replaced.  Segment shards keep the rows they hold."""
        import torch

        eng = self._engine
        if eng is None:
            raise RuntimeError("attach the projection to a network first")
        owner = np.asarray(owner, dtype=np.int64).reshape(-1)
        S = len(owner)
        cells = np.asarray(cells, dtype=np.int64).reshape(S, -1)
        perm = np.asarray(perm, dtype=np.float32).reshape(S, -1)
        E = eng.ctx.syn_capacity
        valid = cells >= 0
        n_valid = valid.sum(axis=1)
        if count is not None and not np.array_equal(n_valid, np.asarray(count).reshape(-1)):
            raise ValueError("count does not match the number of non-negative cells per row")
        if S > eng.ctx.seg_capacity or (S and n_valid.max() > E):
            raise nat.NativeError("import_segments: more segments / synapses per segment than this network's capacity")
        nat.check(nat.lib.bh_tm_reset(eng.ref, eng.stream), "bh_tm_reset")
        order = np.argsort(~valid, axis=1, kind="stable")  # valid slots first, original order kept
        cells_c = np.take_along_axis(cells, order, axis=1)
        perm_c = np.take_along_axis(perm, order, axis=1)
        width = cells.shape[1]
        rows_cell = np.zeros((S, E), dtype=np.int32)
        rows_perm = np.full((S, E), -1.0, dtype=np.float32)
        w = min(width, E)
        dev_cells = (cells_c // eng.c) * 32 + cells_c % eng.c
        rows_cell[:, :w] = np.where(cells_c[:, :w] >= 0, dev_cells[:, :w], 0)
        rows_perm[:, :w] = np.where(cells_c[:, :w] >= 0, perm_c[:, :w], -1.0)
        ids = eng.held_segment_ids(S)
        n = len(ids)
        owner_dev = ((owner // eng.c) * 32 + owner % eng.c).astype(np.int32)

        def put(name, arr, length):
            eng.buf[name][:length].copy_(torch.from_numpy(np.ascontiguousarray(arr).reshape(-1)).to(eng.device))

        for name in ("seg_owner", "seg_count", "seg_pot", "seg_conn"):
            eng.buf[name].zero_()
        put("seg_owner", owner_dev, S)
        put("seg_count", n_valid.astype(np.int32), S)
        put("syn_cell", rows_cell[ids], n * E)
        put("syn_perm", rows_perm[ids], n * E)
        nseg = np.bincount(owner_dev, minlength=eng.C * 32).astype(np.int32)
        put("cell_nseg", nseg, eng.C * 32)
        eng.buf["sc"][nat.SC_NSEG] = S
        eng.buf["sc"][nat.SC_NSEG_NEXT] = S
        eng.tm_deferred = False


class SegmentProjectionView:
    """Export shim with the attribute names of the reference's ``SparseProjection``
    (projections.py:32-68).  ``output_edge`` holds the presynaptic flat cell directly
    (slot index 0), so ``get_output_edge_target`` is the reference's ``% (input_dim+1)``
    and free slots carry ``invalid_output_edge == input_dim`` (:36)."""

    def __init__(self, projection):
        self.input_dim = projection.output_dim  # presynaptic cells = all cells (networks.py:55)
        self.invalid_input_edge = 0
        self.invalid_output_edge = self.input_dim
        if projection._engine is None:
            self.output_dim = 0
            self.output_edges = np.zeros((0, 1), dtype=np.int32)
            self.output_edge = np.zeros((0, 0), dtype=np.int32)
            self.output_permanence = np.zeros((0, 0), dtype=np.float32)
            return
        owner, count, cells, perm = projection.export_segments()
        width = int(count.max()) if len(count) else 0
        self.output_dim = len(owner)
        self.output_edges = count.astype(np.int32).reshape(-1, 1)  # :42
        edge = cells[:, :width].astype(np.int32)
        edge[edge < 0] = self.invalid_output_edge
        self.output_edge = edge  # :43
        self.output_permanence = perm[:, :width].astype(np.float32)  # :44 (-1.0 = free)

    def get_output_edge_target(self, output_edge):  # projections.py:60-61
        return output_edge % (self.input_dim + 1)
