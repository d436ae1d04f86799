"""Network layer -- device-backed mirrors of ``bithtm/networks.py``:
``SpatialPooler`` (:7-35), ``TemporalMemory`` (:38-128) and
``HierarchicalTemporalMemory`` (:131-149) with the same constructor signatures,
``process`` methods and State attribute names.  Extra keyword-only arguments
size the device buffers (the reference grows its arrays without bound).

Randomness: the reference draws from the global legacy ``np.random`` stream.
Here the MT19937 state is advanced on the device; with ``rng_sync="step"``
(default) the global ``np.random`` state is uploaded before and written back
after every timestep, so a caller that also uses ``np.random`` (example.py:52)
sees exactly the reference's stream.  ``rng_sync="lazy"`` uploads once and writes
back only on :meth:`sync_rng` (no per-step host round trip).
"""

from __future__ import annotations

import numpy as np

from . import _native as nat
from ._engine import Engine
from ._rnglink import _RngLink
from .projections import DenseProjection, PredictiveProjection, _Lazy
from .regularizations import ExponentialBoosting, GlobalInhibition, eng_set_active


def _bits(words: np.ndarray, c: int) -> np.ndarray:
    """uint32 [n] -> bool [n, c] (bit b = cell b)."""
    w = np.ascontiguousarray(words, dtype=np.uint32).reshape(-1, 1)
    return ((w >> np.arange(c, dtype=np.uint32)) & np.uint32(1)).astype(bool)


class SpatialPooler:
    """networks.py:7-35."""

    class State(_Lazy):
        """networks.py:8-12.  ``overlaps`` (int64) and ``boosted_overlaps`` (float64)
        are fetched from the device on first read."""

        def __init__(self, engine, active_column=None):
            super().__init__(engine)
            if active_column is not None:
                self._cache["active_column"] = active_column
            # engine.epoch == completed steps == device step counter: parity of the
            # ping-pong buffer this step's active columns were written to
            self._parity = engine.epoch & 1
            self._bh_engine_epoch = (engine, engine.epoch)
            self._group = None

        @property
        def active_column(self):
            def fetch():
                eng, cur = self._engine, self._parity
                return eng.buf["active_cols"][cur * eng.k:(cur + 1) * eng.k].cpu().numpy().astype(np.int64)

            return self._get("active_column", fetch)

        def _columns(self, name):
            eng = self._engine
            if eng.shard_world > 1:  # COLLECTIVE: every rank must read this field
                from ._shard import gather_columns

                return gather_columns(eng.buf[name], self._group).cpu().numpy()
            return eng.buf[name].cpu().numpy()

        @property
        def overlaps(self):
            return self._get("overlaps", lambda: self._columns("overlaps").astype(np.int64))

        @property
        def boosted_overlaps(self):
            return self._get("boosted_overlaps", lambda: self._columns("boosted"))

    def __init__(self, input_dim, column_dim, active_columns, proximal_projection=None, boosting=None,
                 inhibition=None, **engine_kwargs):
        self.input_dim = input_dim
        self.column_dim = column_dim
        self.active_columns = active_columns
        self.proximal_projection = proximal_projection or DenseProjection(input_dim, column_dim)  # :22
        self.boosting = boosting or ExponentialBoosting(column_dim, active_columns)  # :23
        self.inhibition = inhibition or GlobalInhibition(active_columns)  # :24
        if not isinstance(self.proximal_projection, DenseProjection) or not isinstance(self.boosting, ExponentialBoosting):
            raise TypeError("bithtm_b200.SpatialPooler needs bithtm_b200 DenseProjection / ExponentialBoosting "
                            "objects (device-backed); only `inhibition` may be an arbitrary host object")
        self._engine_kwargs = engine_kwargs
        self._engine = None
        self._standalone_tm = False
        self._group = None  # torch.distributed group of the column shards

    @property
    def _native_inhibition(self):
        return getattr(self.inhibition, "_bh_native", False)

    def _attach(self, engine):
        self._engine = engine
        self.proximal_projection._bind(engine)
        self.boosting._bind(engine)
        if self._native_inhibition:
            self.inhibition._bind(engine)

    def _ensure_engine(self):
        if self._engine is None:
            kw = dict(max_segments=64, max_synapses_per_segment=32)
            kw.update(self._engine_kwargs)
            self._attach(Engine(self.input_dim, self.column_dim, 1, self.active_columns, **kw))
            self._standalone_tm = True
        return self._engine

    def _process_sharded(self, eng, words, learning):
        """Column-sharded SP step: local overlap/boost + local candidates, one all-gather,
        identical global selection on every rank, local learning (see _shard.py)."""
        import torch

        from ._shard import gather_candidates

        if not self._native_inhibition:
            raise NotImplementedError("column sharding needs the built-in GlobalInhibition (canonical rule)")
        k_loc = eng.k_local
        keys = torch.empty(k_loc, dtype=torch.float64, device=eng.device)
        cols = torch.empty(k_loc, dtype=torch.int32, device=eng.device)
        nat.check(nat.lib.bh_sp_shard_local(eng.ref, words.data_ptr(), keys.data_ptr(), cols.data_ptr(), eng.stream),
                  "bh_sp_shard_local")
        all_keys, all_cols = gather_candidates(keys, cols, self._group)
        nat.check(nat.lib.bh_sp_shard_finish(eng.ref, words.data_ptr(), all_keys.data_ptr(), all_cols.data_ptr(),
                                             int(all_keys.numel()), int(bool(learning)), eng.stream),
                  "bh_sp_shard_finish")
        return self.State(eng)

    def process(self, input, learning=True):
        """networks.py:26-35.  Leaves the active columns on the device for the TM."""
        eng = self._ensure_engine()
        eng.begin_regular_step()
        self.boosting._bind(eng)
        words = eng.pack_input(input)
        if eng.shard_world > 1:
            state = self._process_sharded(eng, words, learning)
            if self._standalone_tm:
                self._complete_step(state)
            return state
        if self._native_inhibition:
            nat.check(nat.lib.bh_sp_step(eng.ref, words.data_ptr(), int(bool(learning)), eng.stream), "bh_sp_step")
            state = self.State(eng)
            if self._standalone_tm:
                self._complete_step(state)
            return state
        # host inhibition (any object with .process(boosted) -> ordered active columns)
        nat.check(nat.lib.bh_sp_overlap(eng.ref, words.data_ptr(), eng.stream), "bh_sp_overlap")
        nat.check(nat.lib.bh_boost(eng.ref, eng.stream), "bh_boost")
        boosted = eng.buf["boosted"].cpu().numpy()
        active_column = np.asarray(self.inhibition.process(boosted))
        eng_set_active(eng, active_column)
        if learning:
            nat.check(nat.lib.bh_sp_learn(eng.ref, words.data_ptr(), eng.stream), "bh_sp_learn")
        nat.check(nat.lib.bh_duty_update(eng.ref, eng.stream), "bh_duty_update")
        state = self.State(eng, active_column=active_column)
        state._cache["boosted_overlaps"] = boosted
        if self._standalone_tm:
            self._complete_step(state)
        return state

    def _complete_step(self, state):
        """No temporal memory follows: rotate the ping-pong buffers ourselves."""
        eng = self._engine
        nat.check(nat.lib.bh_advance_step(eng.ref, eng.stream), "bh_advance_step")
        eng.epoch += 1
        state._epoch = eng.epoch


class TemporalMemory:
    """networks.py:38-128."""

    class State(_Lazy):
        """networks.py:39-46.  Built from the per-step summary (active columns and
        three bit-words per active column); ``cell_prediction`` and ``distal_state``
        fields are fetched from the device on first read."""

        def __init__(self, engine, projection, summary, have_winner=True, have_jitter=True):
            super().__init__(engine)
            k, c = engine.k, engine.c
            self._c, self._C, self._k = c, engine.C, k
            self._head = summary[:4 + 4 * k].copy()  # the summary buffer is reused by the next step
            self._tail = summary[4 + 4 * k + nat.MT_N + 1:4 + 4 * k + nat.MT_N + 5].copy()
            self._have_winner = have_winner
            self.n_segments = int(self._head[2])
            self.distal_state = PredictiveProjection.State(engine, projection, have_jitter)

        @property
        def _active_column(self):  # host data (the copied summary): readable at any time
            a = self._cache.get("_active_column")
            if a is None:
                a = self._cache["_active_column"] = self._head[4:4 + self._k].astype(np.int64)
            return a

        def _words(self, i):  # row_pred / row_act / row_win of the summary
            return self._head[4 + i * self._k:4 + (i + 1) * self._k].view(np.uint32)

        @property
        def _row_pred(self):
            return self._words(1)

        @property
        def _row_act(self):
            return self._words(2)

        @property
        def _row_win(self):
            return self._words(3)

        def _cells(self, words):
            rows, cells = np.nonzero(_bits(words, self._c))
            return (self._active_column[rows], cells)

        @property
        def active_cell(self):  # networks.py:116-117
            return self._get("active_cell", lambda: self._cells(self._row_act))

        @property
        def winner_cell(self):  # networks.py:103-104
            if not self._have_winner:
                return None
            return self._get("winner_cell", lambda: self._cells(self._row_win))

        @winner_cell.setter
        def winner_cell(self, value):
            self._cache["winner_cell"] = value

        @property
        def active_column_bursting(self):  # networks.py:97, bool [k, 1]
            return (self._row_pred == 0).reshape(-1, 1)

        @property
        def column_metrics(self):
            """Extension: the three counts example.py:55-57 prints, from the step summary alone (no read-back
            of ``cell_prediction``): bursting active columns, active columns that were predicted, predicted
            columns that did not become active; plus the columns predicted for the next step."""
            bursting = int((self._row_pred == 0).sum())
            correct = self._k - bursting
            return {"bursting": bursting, "correct": correct, "incorrect": int(self._tail[0]) - correct,
                    "predicted_columns": int(self._tail[1])}

        @property
        def cell_activation(self):  # networks.py:118-119, bool [C, c]
            def build():
                out = np.zeros((self._C, self._c), dtype=bool)
                out[self._active_column] = _bits(self._row_act, self._c)
                return out

            return self._get("cell_activation", build)

        @property
        def cell_prediction(self):  # networks.py:122, bool [C, c]
            return self._get("cell_prediction", lambda: _bits(
                self._engine.buf["col_pred"].cpu().numpy().view(np.uint32), self._c))

        def materialize(self):
            self.cell_prediction
            self.distal_state.materialize()
            return self

    class _EmptyState:
        """networks.py:59-65."""

        def __init__(self, column_dim, cell_dim):
            self.active_cell = (np.empty(0, dtype=np.int32), np.empty(0, dtype=np.int32))
            self.winner_cell = None
            self.cell_activation = np.zeros((column_dim, cell_dim), dtype=np.bool_)
            self.cell_prediction = np.zeros((column_dim, cell_dim), dtype=np.bool_)
            self.active_column_bursting = np.empty(0, dtype=np.bool_)
            self.distal_state = None

    def __init__(self, column_dim, cell_dim, distal_projection=None, rng_sync="step", **engine_kwargs):
        self.column_dim = column_dim
        self.cell_dim = cell_dim
        self.distal_projection = distal_projection or PredictiveProjection(column_dim * cell_dim)  # :55
        if not isinstance(self.distal_projection, PredictiveProjection):
            raise TypeError("bithtm_b200.TemporalMemory needs a bithtm_b200 PredictiveProjection (device-backed)")
        self._engine_kwargs = engine_kwargs
        self._engine = None
        self._rng = _RngLink(rng_sync)
        self.last_state = self.get_empty_state()  # :57

    def get_empty_state(self):
        return self._EmptyState(self.column_dim, self.cell_dim)

    def flatten_cell(self, cell):  # networks.py:67-71
        if cell is None:
            return None
        assert len(cell) == 2 and len(cell[0].shape) == 1
        return cell[0] * self.cell_dim + cell[1]

    def _attach(self, engine, epsilon=1e-8):
        self._engine = engine
        self.distal_projection._bind(engine, epsilon)
        self.distal_projection._rng_link = self._rng  # one link per np.random stream

    def sync_rng(self):
        """Write the device MT19937 state back into the global np.random."""
        if self._engine is not None:
            self._rng.sync(self._engine)

    def _finish(self, summary, have_winner=True, have_jitter=True):
        eng = self._engine
        eng.check_status(summary[1])
        state = self.State(eng, self.distal_projection, summary, have_winner, have_jitter)
        self.last_state = state
        return state

    def process(self, sp_state, prev_state=None, learning=True, return_winner_cell=True, epsilon=1e-8,
                return_state=True):
        """networks.py:91-128.  ``return_state=False`` (extension) enqueues the step without
        reading anything back (no host synchronisation; needs ``rng_sync="lazy"``)."""
        forget = False
        if prev_state is not None and prev_state is not self.last_state:
            # the previous state lives on the device; the one other state that can be named without it is
            # the empty state (networks.py:59-65): start a new sequence, keeping everything learned
            if (getattr(prev_state, "distal_state", 0) is None and getattr(prev_state, "winner_cell", 0) is None
                    and not np.any(prev_state.cell_prediction)):
                forget = True
            else:
                raise NotImplementedError("bithtm_b200.TemporalMemory keeps the previous state on the device; "
                                          "prev_state may be last_state or an empty state (get_empty_state())")
        want = bool(learning or return_winner_cell)  # networks.py:99
        active_column = None
        tag = getattr(sp_state, "_bh_engine_epoch", None)
        if self._engine is None:
            if tag is not None:
                raise RuntimeError("attach SpatialPooler and TemporalMemory through HierarchicalTemporalMemory")
            active_column = np.asarray(sp_state.active_column)
            self._attach(Engine(1, self.column_dim, self.cell_dim, len(active_column), **self._engine_kwargs), epsilon)
        eng = self._engine
        eng.begin_regular_step()
        self.distal_projection._bind(eng, epsilon)
        if forget:
            nat.check(nat.lib.bh_tm_reset(eng.ref, eng.stream), "bh_tm_reset")
            self.last_state = self.get_empty_state()
        on_device = tag is not None and tag[0] is eng and tag[1] == eng.epoch
        if not on_device:
            eng_set_active(eng, sp_state.active_column if active_column is None else active_column)
        self._rng.before(eng)
        if eng.seg_world > 1:  # segment shards: local scan, ONE all-gather, merge (see _shard.py)
            from ._shard import gather_records

            send = eng.tm_shard_pre(int(bool(learning)) | (0 if return_winner_cell else 2))
            eng.tm_shard_post(gather_records(send, eng.xch_recv, self.distal_projection._group),
                              want_jitter=bool(return_winner_cell))
        else:
            nat.check(nat.lib.bh_tm_step_ex(eng.ref, int(bool(learning)), int(bool(return_winner_cell)),
                                            int(bool(return_winner_cell)), eng.stream), "bh_tm_step_ex")
            eng.epoch += 1
        # a step without winner cells / without the jitter draw leaves state only the per-stage kernels
        # interpret (deferred rand(M), "winner_cell is None"): the next step must take this path too
        eng.tm_deferred = not bool(return_winner_cell)
        if not return_state:
            if self._rng.mode != "lazy":
                raise ValueError('return_state=False needs rng_sync="lazy" (no per-step read-back)')
            self.last_state = None
            return None
        summary = eng.summary()
        self._rng.after(eng, summary)
        return self._finish(summary, have_winner=want, have_jitter=bool(return_winner_cell))


class HierarchicalTemporalMemory:
    """networks.py:131-149."""

    def __init__(self, input_dim, column_dim, cell_dim, active_columns=None, spatial_pooler=None,
                 temporal_memory=None, rng_sync="step", device=None, column_shard=None, process_group=None,
                 segment_shard=None, **engine_kwargs):
        """Extensions (keyword-only in spirit): ``column_shard=True`` shards the network over
        the ranks of ``process_group`` (default: the world group of an initialised
        torch.distributed; every rank must construct the network after the same
        ``np.random.seed`` and feed the same inputs): the spatial pooler by column and --
        unless ``segment_shard=False`` -- the temporal memory's synapse rows by segment id.
        ``(rank, world)`` tuples select a shard explicitly (single-process tests).  Other
        keyword arguments size the device buffers (see ``Engine``)."""
        if active_columns is None:
            active_columns = round(column_dim * 0.02)  # :136-137
        self.input_dim = input_dim
        self.column_dim = column_dim
        self.cell_dim = cell_dim
        self.active_columns = active_columns
        self.spatial_pooler = spatial_pooler or SpatialPooler(input_dim, column_dim, active_columns)  # :143
        self.temporal_memory = temporal_memory or TemporalMemory(column_dim, cell_dim, rng_sync=rng_sync)  # :144
        sp, tm = self.spatial_pooler, self.temporal_memory
        if not isinstance(sp, SpatialPooler) or not isinstance(tm, TemporalMemory):
            raise TypeError("bithtm_b200.HierarchicalTemporalMemory composes bithtm_b200 SpatialPooler / "
                            "TemporalMemory objects")
        if temporal_memory is None:
            tm._rng = _RngLink(rng_sync)
        if sp._engine is not None or tm._engine is not None:
            raise NotImplementedError("spatial_pooler / temporal_memory were already used stand-alone")
        shard = None
        if column_shard:
            import torch.distributed as dist

            if column_shard is True:
                shard = (dist.get_rank(process_group), dist.get_world_size(process_group))
            else:
                shard = tuple(column_shard)
            sp._group = process_group
        if segment_shard is None:
            segment_shard = shard is not None
        if segment_shard is True:
            segment_shard = shard
        seg = tuple(segment_shard) if segment_shard else None
        if seg is not None and seg[1] <= 1:
            seg = None
        tm.distal_projection._group = process_group
        # buffer-sizing keyword arguments given to a user-built SpatialPooler / TemporalMemory are honoured
        # (explicit arguments of this constructor win)
        merged = dict(sp._engine_kwargs)
        merged.update(tm._engine_kwargs)
        merged.update(engine_kwargs)
        self._engine = Engine(input_dim, column_dim, cell_dim, sp.active_columns, device=device,
                              column_shard=shard, segment_shard=seg, **merged)
        sp._attach(self._engine)
        tm._attach(self._engine)
        if self._engine.ctx.fused_mode == 3:  # one kernel per shard, exchanges over peer memory
            import torch

            eng = self._engine
            if eng.seg_world > 1 and column_shard is True:
                from ._shard import map_exchange_regions

                self.exchange_transport = map_exchange_regions(eng, process_group)
            elif eng.seg_world <= 1:  # a single shard exchanges with itself
                region = torch.zeros(eng.exchange_region_ints(), dtype=torch.int32, device=eng.device)
                eng.set_exchange_regions([region.data_ptr()], keepalive=region)
            # explicit (rank, world) tuples: the caller wires the regions (single-process tests)

    @property
    def engine(self):
        return self._engine

    def sync_rng(self):
        self.temporal_memory.sync_rng()

    def reset_sequence(self):
        """Extension: forget the previous timestep (as ``TemporalMemory.process(prev_state=
        get_empty_state())`` does, networks.py:59-65, 91-93); everything learned is kept."""
        eng = self._engine
        nat.check(nat.lib.bh_tm_reset(eng.ref, eng.stream), "bh_tm_reset")
        self.temporal_memory.last_state = self.temporal_memory.get_empty_state()
        eng.tm_deferred = False

    def state_dict(self):
        """Extension (checkpoint): the learned state, the previous timestep's context and the MT19937
        stream position of this network (of this rank's shard), as host tensors."""
        return self._engine.snapshot()

    def load_state_dict(self, state):
        """Extension (resume): inverse of :meth:`state_dict`; the run continues bit-identically.  With
        ``rng_sync="step"`` the global ``np.random`` state is set to the checkpoint's as well."""
        eng, tm = self._engine, self.temporal_memory
        eng.restore(state)
        key, pos = state["rng"]
        if tm._rng.mode == "step":
            np.random.set_state(("MT19937", np.asarray(key, dtype=np.uint32), int(pos), 0, 0.0))
        tm._rng.adopt(eng, ("MT19937", key, pos))
        tm.last_state = None  # lives on the device; the next process() returns the first readable state

    def process(self, input, learning=True, return_state=True, return_winner_cell=True):
        """networks.py:146-149 (``return_winner_cell`` -- an extension here -- is passed on to
        ``TemporalMemory.process``; with ``learning=False`` it gives the inference-only step).  Host inputs go through ``bh_step_host`` (one H2D of
        the packed input, the whole step on the device, one D2H of the step summary).
        ``return_state=False`` (extension; device inputs, ``rng_sync="lazy"``) only enqueues
        the step: nothing is read back and the host does not wait."""
        sp, tm, eng = self.spatial_pooler, self.temporal_memory, self._engine
        eng.begin_regular_step()
        is_host = not (hasattr(input, "is_cuda") and input.is_cuda)
        want = bool(learning or return_winner_cell)  # networks.py:99
        flags = int(bool(learning)) | (0 if return_winner_cell else 2)  # BH_STEP_LEARNING | BH_STEP_NO_WINNER_CELLS
        mode = eng.ctx.fused_mode
        if mode == 3:  # the shard's whole step is one kernel (exchanges inside)
            if not sp._native_inhibition:
                raise NotImplementedError('fused="shard" needs the built-in GlobalInhibition')
        whole_kernel = sp._native_inhibition and (mode == 3 or (eng.shard_world == 1 and eng.seg_world == 1))
        if not whole_kernel:
            # an arbitrary host object in the inhibition slot, or NCCL exchanges between the stages of a shard
            sp_state = sp.process(input, learning=learning)
            sp_state._group = sp._group
            tm_state = tm.process(sp_state, learning=learning, return_state=return_state,
                                  return_winner_cell=return_winner_cell)
            if not return_state:
                return None
            sp_state._epoch = eng.epoch  # its buffers stay valid until the next step
            return sp_state, tm_state
        sp.boosting._bind(eng)
        tm._rng.before(eng)
        if mode == 3 or not is_host or not return_state:
            # device input (or nothing to read back): enqueue the step; the summary is fetched only when asked for
            words = eng.pack_input(input)
            eng.step_device(words, learning=flags)
            eng.tm_deferred = not return_winner_cell
            if not return_state:
                if tm._rng.mode != "lazy":
                    raise ValueError('return_state=False needs rng_sync="lazy" (no per-step read-back)')
                tm.last_state = None
                return None
            summary = eng.summary()
        else:
            # host input: ONE call -- H2D of the packed input, the whole step, D2H of the step summary
            summary = eng.step_host(np.asarray(input).reshape(-1), learning=flags)
            eng.tm_deferred = not return_winner_cell
        tm._rng.after(eng, summary)
        tm_state = tm._finish(summary, have_winner=want, have_jitter=bool(return_winner_cell))
        sp_state = sp.State(eng, active_column=tm_state._active_column)
        sp_state._parity ^= 1  # created after the step completed
        sp_state._group = sp._group
        return sp_state, tm_state
