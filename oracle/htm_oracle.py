"""CPU oracle for the bitHTM spatial-pooler + temporal-memory timestep.

TEST INFRASTRUCTURE ONLY.  Nothing under ``bithtm_b200/`` may import this
module; it is the checker for the CUDA path (``tests/``, ``__graft_entry__.smoke``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``).

This is a *restatement* in NumPy of the algorithm of cokwa/bitHTM, not a copy:
the reference keeps TM synapses in growable 2-D arrays plus a cell->segment
forward index (``bithtm/projections.py:27-192``, ``bithtm/utils.py:79-135``);
here a segment is a row of (presynaptic cell, permanence) slots and every
consumer is a count/sum over the row, which is all the reference's results
depend on ("slot layout does not affect any result").

Parity status: the reference ships no tests or golden vectors, so the pin is
the reference itself: ``tests/golden/make_golden.py`` runs the unmodified
reference from ``/root/reference`` lock-step against this oracle and writes the
digests committed under ``tests/golden/``.  Parity claim is scoped to NumPy 2.x
on an AVX2-or-better host (``np.exp`` on float32 is SIMD-path dependent,
SURVEY.md Appendix B).

Reference citations are ``file:line`` into cokwa/bitHTM.
"""

from __future__ import annotations

import dataclasses

import numpy as np


F32 = np.float32
F64 = np.float64


@dataclasses.dataclass
class OracleConfig:
    """Hyper-parameters; defaults are the reference's constructor defaults."""

    input_dim: int
    column_dim: int
    cell_dim: int
    active_columns: int | None = None  # networks.py:136-137 -> round(0.02*C)
    # DenseProjection (projections.py:7-10)
    sp_permanence_mean: float = 0.0
    sp_permanence_std: float = 0.1
    sp_permanence_threshold: float = 0.0
    sp_permanence_increment: float = 0.03
    sp_permanence_decrement: float = 0.015
    # ExponentialBoosting (regularizations.py:5-7)
    boost_intensity: float = 0.3
    boost_momentum: float = 0.99
    # PredictiveProjection (projections.py:205-209)
    tm_permanence_initial: float = 0.21
    tm_permanence_threshold: float = 0.5
    tm_permanence_increment: float = 0.1
    tm_permanence_decrement: float = 0.1
    tm_permanence_punishment: float = 0.01
    segment_activation_threshold: int = 15
    segment_matching_threshold: int = 15
    segment_sampling_synapses: int = 32
    epsilon: float = 1e-8  # networks.py:91

    def __post_init__(self):
        if self.active_columns is None:
            self.active_columns = round(self.column_dim * 0.02)


def canonical_topk(keys: np.ndarray, k: int) -> np.ndarray:
    """Deterministic global inhibition: larger key first, ties -> lower column
    index, result in ascending column index.  Replaces
    ``np.argpartition(x, -k)[-k:]`` (regularizations.py:28-29), whose tie-break
    and output order are NumPy-dispatch dependent."""
    order = np.lexsort((np.arange(len(keys)), -keys))  # primary: -keys, secondary: index
    return np.sort(order[:k]).astype(np.int64)


class CanonicalGlobalInhibition:
    """Plugin for the *reference's* ``SpatialPooler(..., inhibition=...)`` slot
    (networks.py:16,24) that applies :func:`canonical_topk`."""

    def __init__(self, active_outputs):
        self.active_outputs = active_outputs

    def process(self, input_activation):
        return canonical_topk(input_activation, self.active_outputs)


@dataclasses.dataclass
class StepRecord:
    """Everything one timestep produces, in the reference's dtypes."""

    overlaps: np.ndarray  # int64 [C]            projections.py:18-21
    boosted: np.ndarray  # float64 [C]           regularizations.py:15-17
    active_column: np.ndarray  # int64 [k]       regularizations.py:28-29
    bursting: np.ndarray  # bool [k]             networks.py:97
    winner_cell: np.ndarray  # int64 [W] flat    networks.py:102-104
    active_cell: np.ndarray  # int64 [A] flat    networks.py:115-117
    learning_segment: np.ndarray  # int64 [L]    projections.py:264-281
    punished_segment: np.ndarray  # int64 [P]    projections.py:269
    n_segments: int
    matching_segment: np.ndarray  # int64 [M]    projections.py:247
    matching_activation: np.ndarray  # int64 [M] projections.py:249
    matching_jit: np.ndarray  # float32 [M]      projections.py:234-235
    draws: int  # float64 uniforms consumed this step
    undefined_tie: bool  # np.argsort tie at the n_add boundary (SURVEY 8c residual)


class HTMOracle:
    """SP + TM step.  ``rng`` is anything with ``randn``/``random_sample``
    (the ``np.random`` module or a ``RandomState``); draws are consumed in the
    reference's order: construction ``randn(C, I)`` (projections.py:16), then per
    step ``rand(k, c)`` (networks.py:87), ``rand(L, W+1)`` (projections.py:120),
    ``rand(M)`` (projections.py:235)."""

    def __init__(self, cfg: OracleConfig, rng=None, inhibition="canonical", overlap="dense",
                 permanence=None):
        self.cfg = cfg
        self.rng = np.random if rng is None else rng
        self.inhibition = inhibition
        self.overlap_mode = overlap
        C, I, c = cfg.column_dim, cfg.input_dim, cfg.cell_dim
        self.C, self.I, self.c, self.k = C, I, c, cfg.active_columns
        self.N = C * c

        # --- spatial pooler state -------------------------------------------------
        if permanence is None:
            # projections.py:16 -- float64
            permanence = self.rng.randn(C, I) * cfg.sp_permanence_std + cfg.sp_permanence_mean
        self.permanence = permanence
        self.duty = np.zeros(C, dtype=F32)  # regularizations.py:13
        density = self.k / C  # regularizations.py:9
        # regularizations.py:16: python float is a weak scalar -> float32 multiply
        self.boost_coef = F32(-(cfg.boost_intensity / density))
        # projections.py:24: bool * float - float -> float64, folded per bit value
        both = cfg.sp_permanence_increment + cfg.sp_permanence_decrement
        self.sp_delta_on = 1.0 * both - cfg.sp_permanence_decrement
        self.sp_delta_off = 0.0 * both - cfg.sp_permanence_decrement
        if overlap == "packed":
            self._repack(np.arange(C))

        # --- temporal memory state --------------------------------------------------
        self.n_seg = 0
        self.seg_owner = np.zeros(0, dtype=np.int64)  # projections.py:226 segment_bundle
        self.seg_count = np.zeros(0, dtype=np.int64)  # projections.py:42 output_edges
        self.syn_cell = np.full((0, 32), -1, dtype=np.int64)  # -1 = free slot
        self.syn_perm = np.full((0, 32), -1.0, dtype=F32)
        self.cell_nseg = np.zeros(self.N, dtype=np.int32)  # projections.py:227 bundle_segments
        self.have_prev = False  # prev distal state is None (networks.py:74, projections.py:258)
        self.cell_prediction = np.zeros((C, c), dtype=bool)  # networks.py:63
        self.cell_activation = np.zeros((C, c), dtype=bool)  # networks.py:62
        self.prev_winners = None  # networks.py:61 / :112
        self.m_seg = np.zeros(0, dtype=np.int64)
        self.m_act = np.zeros(0, dtype=np.int64)
        self.m_jit = np.zeros(0, dtype=F32)
        self.seg_potential = np.zeros(0, dtype=np.int64)
        self.max_jit = np.zeros(self.N, dtype=F32)
        self.npred = np.zeros(self.N, dtype=np.int64)
        # TM permanence deltas, evaluated as the reference does (Appendix A):
        # act*(da-di)+di with a bool array -> float64
        self.d_learn = self._deltas(cfg.tm_permanence_increment, -cfg.tm_permanence_decrement)
        self.d_punish = self._deltas(-cfg.tm_permanence_punishment, 0.0)

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _deltas(active_change, inactive_change):
        # projections.py:102
        on = np.array([True]) * (active_change - inactive_change) + inactive_change
        off = np.array([False]) * (active_change - inactive_change) + inactive_change
        return float(on[0]), float(off[0]), min(active_change, inactive_change) < 0

    def _repack(self, rows):
        conn = self.permanence[rows] >= self.cfg.sp_permanence_threshold
        pad = (-self.I) % 64
        if pad:
            conn = np.concatenate([conn, np.zeros((len(rows), pad), dtype=bool)], axis=1)
        packed = np.packbits(conn, axis=1, bitorder="little").view(np.uint64)
        if not hasattr(self, "mask"):
            self.mask = np.zeros((self.C, packed.shape[1]), dtype=np.uint64)
        self.mask[rows] = packed

    def _grow_rows(self, n_new):
        E = self.syn_cell.shape[1]
        need = self.n_seg + n_new
        cap = self.syn_cell.shape[0]
        if need > cap:
            new_cap = max(need, 2 * cap, 64)
            for name, fill in (("syn_cell", -1), ("syn_perm", F32(-1.0))):
                old = getattr(self, name)
                new = np.full((new_cap, E), fill, dtype=old.dtype)
                new[:cap] = old
                setattr(self, name, new)
            for name in ("seg_owner", "seg_count"):
                old = getattr(self, name)
                new = np.zeros(new_cap, dtype=old.dtype)
                new[:cap] = old
                setattr(self, name, new)

    def _grow_cols(self, width):
        E = self.syn_cell.shape[1]
        if width <= E:
            return
        new_E = max(width, 2 * E)
        cap = self.syn_cell.shape[0]
        nc = np.full((cap, new_E), -1, dtype=np.int64)
        nc[:, :E] = self.syn_cell
        npm = np.full((cap, new_E), -1.0, dtype=F32)
        npm[:, :E] = self.syn_perm
        self.syn_cell, self.syn_perm = nc, npm

    # ------------------------------------------------------------------ SP
    def sp_overlap(self, x):
        """projections.py:18-21"""
        if self.overlap_mode == "packed":
            pad = (-self.I) % 64
            xb = np.concatenate([x, np.zeros(pad, dtype=bool)]) if pad else x
            xw = np.packbits(xb, bitorder="little").view(np.uint64)
            return np.bitwise_count(self.mask & xw).sum(axis=1).astype(np.int64)
        weight = self.permanence >= self.cfg.sp_permanence_threshold
        return (weight & x).sum(axis=1)

    def sp_boost(self, overlaps):
        """regularizations.py:15-17 (float32 exp, float64 product)"""
        factor = np.exp(self.boost_coef * self.duty)
        return factor * overlaps

    def sp_inhibit(self, boosted):
        """regularizations.py:28-29 / canonical rule"""
        if self.inhibition == "canonical":
            return canonical_topk(boosted, self.k)
        if self.inhibition == "argpartition":
            return np.argpartition(boosted, -self.k)[-self.k:]
        return np.asarray(self.inhibition(boosted))

    def sp_learn(self, x, active_column):
        """projections.py:23-24 (float64, no clip)"""
        self.permanence[active_column] += np.where(x, self.sp_delta_on, self.sp_delta_off)
        if self.overlap_mode == "packed":
            self._repack(active_column)

    def sp_duty_update(self, active_column):
        """regularizations.py:19-21 (two separately rounded float32 ops)"""
        self.duty *= F32(self.cfg.boost_momentum)
        self.duty[active_column] += F32(1.0 - self.cfg.boost_momentum)

    # ------------------------------------------------------------------ TM pieces
    def _update_rows(self, segs, prev_active_flat, deltas):
        """projections.py:97-109 on the given segment rows.  Returns nothing;
        deletes synapses whose float64 sum went negative."""
        if len(segs) == 0:
            return
        d_on, d_off, can_delete = deltas
        cells = self.syn_cell[segs]
        valid = cells >= 0
        act = prev_active_flat[np.where(valid, cells, 0)] & valid
        summed = self.syn_perm[segs].astype(F64) + np.where(valid, np.where(act, d_on, d_off), 0.0)
        self.syn_perm[segs] = np.where(valid, summed.astype(F32), self.syn_perm[segs])
        if can_delete:  # projections.py:105
            dead = valid & (summed < 0.0)
            self.seg_count[segs] -= dead.sum(axis=1)
            self.syn_cell[segs] = np.where(dead, -1, cells)

    def _grow(self, segs, prev_active_flat, winners_prev):
        """projections.py:111-161.  Consumes rand(L, W+1).  Returns whether an
        argsort tie straddled a selection boundary (reference-undefined)."""
        cfg = self.cfg
        L, W = len(segs), len(winners_prev)
        if L == 0 or (W + 1) == 0:
            return False, 0
        pri = self.rng.random_sample((L, W + 1)).astype(F32)  # projections.py:120
        tie = False
        sample = cfg.segment_sampling_synapses
        widx = np.full(self.N, -1, dtype=np.int64)
        widx[winners_prev] = np.arange(W)
        for r in range(L):
            s = segs[r]
            cells = self.syn_cell[s]
            valid = cells >= 0
            vcells = cells[valid]
            n_active = int(prev_active_flat[vcells].sum())  # projections.py:114
            n_add = int(np.clip(sample - n_active, 0, min(sample, W)))  # :115
            if n_add == 0:
                continue
            p = pri[r, :W].copy()
            have = widx[vcells]
            p[have[have >= 0]] = np.inf  # :121
            absent = p < F32(1.0)  # :123
            order = np.argsort(p, kind="stable")  # canonical: stable by candidate index
            chosen = order[:n_add]
            if n_add < W and p[order[n_add - 1]] == p[order[n_add]] and np.isfinite(p[order[n_add]]):
                tie = True
            chosen = np.sort(chosen[absent[chosen]])
            if len(chosen) == 0:
                continue
            free = np.flatnonzero(~valid)
            if len(free) < len(chosen):
                self._grow_cols(int(valid.sum()) + len(chosen))
                free = np.flatnonzero(self.syn_cell[s] < 0)
            slots = free[: len(chosen)]
            self.syn_cell[s, slots] = winners_prev[chosen]
            self.syn_perm[s, slots] = F32(cfg.tm_permanence_initial)  # :149 scalar -> float32
            self.seg_count[s] += len(chosen)  # :161
        return tie, L * (W + 1)

    def _fill_jitter(self):
        """projections.py:229-239, run lazily (networks.py:76 -> projections.py:241-243) when the
        previous activation was asked not to draw it (return_jittered_potential_info=False)."""
        if not getattr(self, "jit_pending", False):
            return 0
        u = self.rng.random_sample(len(self.m_seg))  # :235
        jit = (self.seg_potential[self.m_seg].astype(F64) + u).astype(F32)
        self.max_jit = np.zeros(self.N, dtype=F32)
        np.maximum.at(self.max_jit, self.seg_owner[self.m_seg], jit)  # :236-237
        self.m_jit = jit
        self.jit_pending = False
        return len(self.m_seg)

    def tm_select(self, active_column, want=True):
        """networks.py:95-104: bursting columns and winner cells (`want` = learning or
        return_winner_cell, networks.py:99)."""
        cfg, c = self.cfg, self.c
        eps = F32(cfg.epsilon)
        acp = self.cell_prediction[active_column]
        burst = ~acp.any(axis=1)
        k = len(active_column)
        if not want:
            return acp, burst, None, 0
        filled = self._fill_jitter() if self.have_prev else 0
        if self.have_prev:  # networks.py:73-82
            mj = self.max_jit.reshape(self.C, c)[active_column]
            colmax = mj.max(axis=1, keepdims=True)
            col_matching = colmax >= F32(cfg.segment_matching_threshold)
            best = np.abs(mj - colmax) < eps
        else:
            col_matching = np.zeros((k, 1), dtype=bool)
            best = np.zeros((k, c), dtype=bool)
        # networks.py:84-89: float32 array += float64 draws -> f32(f64(count) + u)
        u = self.rng.random_sample((k, c))
        jittered = (self.cell_nseg.reshape(self.C, c)[active_column].astype(F64) + u).astype(F32)
        least = np.abs(jittered - jittered.min(axis=1, keepdims=True)) < eps
        win = acp | (burst[:, None] & np.where(col_matching, best, least))
        rows, cells = np.nonzero(win)
        winners = active_column[rows] * c + cells
        return acp, burst, winners.astype(np.int64), filled + k * c

    def tm_learn(self, active_column, winners):
        """projections.py:257-293."""
        cfg, c = self.cfg, self.c
        eps = F32(cfg.epsilon)
        empty = np.zeros(0, dtype=np.int64)
        if not self.have_prev:  # projections.py:258-259
            return empty, empty, False, 0
        m = self.m_seg
        owner = self.seg_owner[m]
        is_winner = np.zeros(self.N, dtype=bool)
        is_winner[winners] = True
        unpredicted = self.npred[owner] == 0  # :266 prediction < eps
        best = np.abs(self.m_jit - self.max_jit[owner]) < eps  # :267
        m_active = self.m_act >= cfg.segment_activation_threshold
        learn = m[is_winner[owner] & (m_active | (unpredicted & best))]  # :268
        col_off = np.ones(self.C, dtype=bool)
        col_off[active_column] = False
        punish = m[col_off[owner // c]]  # :269
        unacc = winners[self.max_jit[winners] < eps]  # :271-273
        if len(unacc):
            n = len(unacc)
            recycled = np.flatnonzero(self.seg_count[: self.n_seg] < cfg.segment_matching_threshold)[:n]  # :80-81
            n_new = n - len(recycled)
            # :82-85 / :275-278
            np.subtract.at(self.cell_nseg, self.seg_owner[recycled], 1)
            self.syn_cell[recycled] = -1
            self.syn_perm[recycled] = F32(-1.0)
            self.seg_count[recycled] = 0
            self.cell_nseg[unacc] += 1  # :277
            self.seg_owner[recycled] = unacc[: len(recycled)]
            fresh = np.arange(self.n_seg, self.n_seg + n_new, dtype=np.int64)
            if n_new:
                self._grow_rows(n_new)
                self.seg_owner[fresh] = unacc[len(recycled):]
                self.seg_count[fresh] = 0
                self.n_seg += n_new
            learn = np.concatenate([learn, recycled, fresh])  # :281
        prev_active_flat = self.cell_activation.reshape(-1)
        self._update_rows(learn, prev_active_flat, self.d_learn)  # :284-289
        tie, draws = False, 0
        if self.prev_winners is not None:  # :191
            tie, draws = self._grow(learn, prev_active_flat, self.prev_winners)
        self._update_rows(punish, prev_active_flat, self.d_punish)  # :290-293
        return learn, punish, tie, draws

    def tm_activate(self, active_flat_mask, want_jitter=True):
        """projections.py:245-255 + :229-239 (the jitter only when asked, networks.py:121)."""
        cfg = self.cfg
        S = self.n_seg
        cells = self.syn_cell[:S]
        valid = cells >= 0
        hit = active_flat_mask[np.where(valid, cells, 0)] & valid
        potential = hit.sum(axis=1).astype(np.int64)  # :175-178
        matching = np.flatnonzero(potential >= cfg.segment_matching_threshold)  # :247
        conn = (hit[matching] & (self.syn_perm[:S][matching] >= F32(cfg.tm_permanence_threshold))).sum(axis=1)  # :167-173
        owner = self.seg_owner[matching]
        active = conn >= cfg.segment_activation_threshold  # :250
        self.npred = np.bincount(owner, weights=active, minlength=self.N).astype(np.int64)  # :251
        if not want_jitter:  # the draw is deferred until (and unless) somebody needs it
            self.m_seg, self.m_act, self.m_jit = matching.astype(np.int64), conn.astype(np.int64), None
            self.max_jit = np.zeros(self.N, dtype=F32)
            self.seg_potential = potential
            self.cell_prediction = (self.npred > 0).reshape(self.C, self.c)
            self.have_prev = True
            self.jit_pending = True
            return 0
        self.jit_pending = False
        u = self.rng.random_sample(len(matching))  # :235
        jit = (potential[matching].astype(F64) + u).astype(F32)
        self.max_jit = np.zeros(self.N, dtype=F32)
        np.maximum.at(self.max_jit, owner, jit)  # :236-237
        self.m_seg, self.m_act, self.m_jit = matching.astype(np.int64), conn.astype(np.int64), jit
        self.seg_potential = potential
        self.cell_prediction = (self.npred > 0).reshape(self.C, self.c)  # networks.py:122
        self.have_prev = True
        return len(matching)

    # ------------------------------------------------------------------ full step
    def step(self, x, learning=True, active_column=None, return_winner_cell=True) -> StepRecord:
        """One ``HierarchicalTemporalMemory.process`` (networks.py:146-149).
        ``active_column`` overrides inhibition (host-inhibition parity mode);
        ``return_winner_cell`` is ``TemporalMemory.process``'s flag (networks.py:91)."""
        x = np.asarray(x, dtype=bool)
        overlaps = self.sp_overlap(x)
        boosted = self.sp_boost(overlaps)
        if active_column is None:
            active_column = self.sp_inhibit(boosted)
        active_column = np.asarray(active_column, dtype=np.int64)
        if learning:
            self.sp_learn(x, active_column)
        self.sp_duty_update(active_column)  # networks.py:33: always

        acp, burst, winners, d1 = self.tm_select(active_column, want=learning or return_winner_cell)
        learn = punish = np.zeros(0, dtype=np.int64)
        tie, d2 = False, 0
        if learning:
            learn, punish, tie, d2 = self.tm_learn(active_column, winners)
        act_rows = acp | burst[:, None]  # networks.py:115
        rows, cells = np.nonzero(act_rows)
        active_cell = (active_column[rows] * self.c + cells).astype(np.int64)
        activation = np.zeros((self.C, self.c), dtype=bool)
        activation[active_column] = act_rows  # networks.py:118-119
        d3 = self.tm_activate(activation.reshape(-1), want_jitter=return_winner_cell)
        self.cell_activation = activation
        self.prev_winners = winners
        return StepRecord(
            overlaps=overlaps, boosted=boosted, active_column=active_column, bursting=burst,
            winner_cell=winners if winners is not None else np.zeros(0, dtype=np.int64), active_cell=active_cell,
            learning_segment=learn, punished_segment=punish, n_segments=self.n_seg, matching_segment=self.m_seg,
            matching_activation=self.m_act,
            matching_jit=self.m_jit if self.m_jit is not None else np.zeros(0, dtype=F32), draws=d1 + d2 + d3,
            undefined_tie=tie,
        )

    # ------------------------------------------------------------------ canonical state
    def canonical_synapses(self):
        """Per segment id: owner cell and the sorted (presynaptic cell,
        permanence bits) list -- the layout-independent form parity is checked in."""
        from .digest import canonical_from_rows

        S = self.n_seg
        return canonical_from_rows(self.seg_owner[:S], self.syn_cell[:S], self.syn_perm[:S])
