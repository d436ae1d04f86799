"""GPU parity tests (`-m gpu`): the CUDA path, called through the reference-facing
classes / C ABI, against the NumPy oracle on the same seeded inputs and against
the golden traces recorded from the unmodified reference.  Bit-exact everywhere.
"""

import os

import numpy as np
import pytest

from helpers import (diff_records, golden_inputs, gpu_record, gpu_state_digest, load_golden, make_inputs,
                     oracle_record, oracle_state_digest, step_digest)
from oracle.digest import canonical_from_rows, state_digest
from oracle.htm_oracle import HTMOracle, OracleConfig, canonical_topk

pytestmark = pytest.mark.gpu


def _engine(I=64, C=128, c=8, k=10, **kw):
    from bithtm_b200._engine import Engine

    return Engine(I, C, c, k, max_segments=256, **kw)


# ------------------------------------------------------------------ building blocks
def test_device_mt19937_matches_numpy_stream():
    """bh_rng_fill == np.random.random_sample for any count / position, including
    odd word positions and block boundaries (networks.py:87, projections.py:120,235)."""
    eng = _engine()
    rs = np.random.RandomState(12345)
    rs.randn(7)  # leave the state mid-block
    rs.randint(0, 10, size=3)  # odd number of 32-bit words consumed
    st = rs.get_state()
    eng.set_rng_state(st[1], st[2])
    for count in [0, 1, 5, 311, 312, 313, 623, 624, 625, 1, 1000, 4096, 100_003]:
        got = eng.rng_fill(count)
        want = rs.random_sample(count)
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), f"count {count}"
    key, pos = eng.get_rng_state()
    st = rs.get_state()
    # same continuation: (key, pos) may differ in representation only at pos == 624
    rs2 = np.random.RandomState()
    rs2.set_state(("MT19937", key, pos, 0, 0.0))
    assert np.array_equal(rs2.random_sample(1000), rs.random_sample(1000))


def test_device_mt19937_parallel_production_matches_numpy_stream():
    """Many-CTA production (jump-ahead polynomials, csrc/mt19937.cuh rng_chunk) yields the
    same stream as the serial generator and as NumPy, from a fresh state (window not yet
    generated), mid-stream, with odd word offsets, and exports the same state."""
    eng = _engine(parallel_rng=True, rand_capacity=1 << 21)
    assert eng.ctx.jump_polys > 0
    rs = np.random.RandomState(777)
    rs.randint(0, 10, size=1)  # odd word position
    st = rs.get_state()
    eng.set_rng_state(st[1], st[2])
    for count in [600_000, 3, 40_000, 1_000_001, 0, 150_000, 2_000_000, 12]:
        got = eng.rng_fill(count)
        want = rs.random_sample(count)
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), f"count {count}"
    key, pos = eng.get_rng_state()
    rs2 = np.random.RandomState()
    rs2.set_state(("MT19937", key, pos, 0, 0.0))
    assert np.array_equal(rs2.random_sample(1000), rs.random_sample(1000))
    # re-import mid-way: the ring restarts from the 624-word key
    st = rs.get_state()
    eng.set_rng_state(st[1], st[2])
    got = eng.rng_fill(300_000)
    assert np.array_equal(got.view(np.uint64), rs.random_sample(300_000).view(np.uint64))


def test_device_np_expf_matches_numpy():
    """The boost kernel's exp == np.exp(float32) bit for bit (regularizations.py:16)."""
    import ctypes

    import torch

    from bithtm_b200 import _native as nat

    g = np.random.default_rng(7)
    x = np.concatenate([
        -g.random(4_000_000, dtype=np.float32) * np.float32(15.5),
        np.array([0.0, -0.0, -1e-30, -15.5, -14.985366], dtype=np.float32),
    ]).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    yd = torch.empty_like(xd)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    nat.check(nat.lib.bh_test_np_expf(xd.data_ptr(), yd.data_ptr(), x.size, st))
    y = yd.cpu().numpy()
    assert np.array_equal(np.exp(x).view(np.uint32), y.view(np.uint32))


@pytest.mark.parametrize("I,C,k", [(64, 128, 10), (200, 300, 25), (1000, 2048, 41), (1024, 2048, 41), (4100, 512, 20),
                                   (96, 20480, 410)])  # the last one takes the grid-wide top-k
def test_spatial_pooler_operators(I, C, k):
    """DenseProjection.process/update, ExponentialBoosting.process/update and
    GlobalInhibition.process one by one against the oracle (projections.py:18-24,
    regularizations.py:15-29)."""
    import bithtm_b200 as bithtm

    seed = 11
    np.random.seed(seed)
    sp = bithtm.SpatialPooler(I, C, k)
    orc = HTMOracle(OracleConfig(I, C, 4, k), rng=np.random.RandomState(seed))
    assert np.array_equal(sp.proximal_projection.permanence, orc.permanence)
    g = np.random.default_rng(seed)
    for t in range(25):
        x = g.random(I) < 0.2
        st = sp.process(x, learning=True)
        ov = orc.sp_overlap(x)
        bo = orc.sp_boost(ov)
        ac = orc.sp_inhibit(bo)
        orc.sp_learn(x, ac)
        orc.sp_duty_update(ac)
        assert np.array_equal(st.overlaps, ov), f"overlaps step {t}"
        assert st.overlaps.dtype == np.int64 and st.boosted_overlaps.dtype == np.float64
        assert np.array_equal(st.boosted_overlaps.view(np.uint64), bo.view(np.uint64)), f"boosted step {t}"
        assert np.array_equal(st.active_column, ac), f"active columns step {t}"
        assert np.array_equal(sp.boosting.duty_cycle.view(np.uint32), orc.duty.view(np.uint32)), f"duty step {t}"
    assert np.array_equal(sp.proximal_projection.permanence.view(np.uint64), orc.permanence.view(np.uint64))
    # plugin-granularity calls
    x = g.random(I) < 0.2
    assert np.array_equal(sp.proximal_projection.process(x), orc.sp_overlap(x))


def test_topk_ties_take_lowest_index():
    """All-equal keys (zero duty, equal overlaps happen in the first steps): the
    canonical rule picks the lowest indices, ascending."""
    import bithtm_b200 as bithtm

    np.random.seed(0)
    sp = bithtm.SpatialPooler(64, 512, 7)
    sp._ensure_engine()
    keys = np.zeros(512)
    keys[[5, 100, 300]] = 2.0
    keys[[7, 9, 200, 400, 500]] = 1.0
    got = sp.inhibition.process(keys)
    assert got.tolist() == [5, 7, 9, 100, 200, 300, 400]


def test_topk_grid_wide_ties():
    """Grid-wide top-k (>= 16384 columns): more than 1024 identical keys at the cut
    (tie mode) and a mixed case, lowest column index first."""
    import bithtm_b200 as bithtm

    np.random.seed(0)
    C, k = 20480, 2000
    sp = bithtm.SpatialPooler(64, C, k)
    sp._ensure_engine()
    from oracle.htm_oracle import canonical_topk

    g = np.random.default_rng(1)
    cases = [np.zeros(C), np.where(g.random(C) < 0.05, 3.0, 1.0), g.integers(0, 50, C).astype(np.float64),
             g.random(C) * 100]
    for keys in cases:
        got = sp.inhibition.process(keys)
        assert np.array_equal(got, canonical_topk(keys, k))


# ------------------------------------------------------------------ lock-step SP+TM
def _lockstep(name, steps=None, check_every=1, **engine_kw):
    import bithtm_b200 as bithtm

    info = load_golden(name)
    g = info["g"]
    steps = info["steps"] if steps is None else steps
    xs = golden_inputs(info, steps)
    np.random.seed(info["seed"])
    htm = bithtm.HierarchicalTemporalMemory(info["I"], info["C"], info["c"], info["k"], **engine_kw)
    orc = HTMOracle(OracleConfig(info["I"], info["C"], info["c"], info["k"]),
                    rng=np.random.RandomState(info["seed"]), overlap="packed")
    state_at = {int(s): int(d) for s, d in zip(g["state_steps"], g["state_digests"])}
    for t in range(steps):
        sp_state, tm_state = htm.process(xs[t])
        rec = orc.step(xs[t])
        if t % check_every == 0 or t in state_at or t < 50:
            got, want = gpu_record(htm, sp_state, tm_state), oracle_record(rec)
            problems = diff_records(got, want)
            assert not problems, f"{name} step {t}: " + "; ".join(problems) + f"; sc={htm.engine.scalars()[:18]}"
            assert step_digest(**got) == int(g["digests"][t]), f"{name} step {t}: golden digest"
        if t in state_at:
            assert gpu_state_digest(htm) == oracle_state_digest(orc) == state_at[t], f"{name}: learned state, step {t}"
    # the caller's np.random stream stayed in lock-step with the reference's
    a, b = np.random.get_state(), orc.rng.get_state()
    assert np.array_equal(np.random.random_sample(64), orc.rng.random_sample(64)), (a[2], b[2])
    assert htm.engine.check_status() & ~32 == 0
    return htm, orc


@pytest.mark.parametrize("fused,ctas", [("cluster", 8), ("cluster", 1), ("cluster", 3), ("cluster", 16),
                                        ("grid", None), ("grid", 5), ("off", None)])
def test_lockstep_tiny(fused, ctas):
    """Every execution mode of the step: one kernel on a thread-block cluster of
    1..16 CTAs, one cooperative-grid kernel, or one kernel per stage."""
    _lockstep("tiny", fused=fused, fused_ctas=ctas)


@pytest.mark.parametrize("name,ctas,policy", [("tiny", None, "always"), ("tiny", 5, True), ("tiny", 1, "always"),
                                              ("odd", None, True), ("odd", 7, "always"), ("mid", None, "always")])
def test_lockstep_lazy_draws(name, ctas, policy):
    """Lazy rand(L, W+1) (csrc/mt19937.cuh) forced on at small sizes: the matrix is stepped over, the rows of
    growing segments, whole-matrix chunks and the words after the matrix are produced by table jumps.  Any
    wrong stream word shows up in the growth, the jitter or the final np.random position."""
    htm, _ = _lockstep(name, steps=600 if name == "mid" else None, fused="grid", fused_ctas=ctas, lazy_rng=policy,
                       skip_gran=32, skip_min=64, skip_polys=6000)
    assert htm.engine.ctx.skip_polys == 6000


@pytest.mark.parametrize("ctas,threads", [(4, 512), (8, 256), (16, 768)])
def test_lockstep_cluster_kernel_with_smaller_ctas(ctas, threads):
    """The cluster kernel with fewer threads per CTA (the StreamBatch configuration, several
    CTAs per SM) is the same computation."""
    _lockstep("tiny", fused="cluster", fused_ctas=ctas, fused_threads=threads)


def test_cfg2_cluster_kernel_512_threads_1500_steps():
    _lockstep("cfg2", steps=1500, check_every=25, fused="cluster", fused_ctas=4, fused_threads=512)


@pytest.mark.parametrize("fused", ["cluster", "grid", "off"])
def test_lockstep_odd_dims(fused):
    _lockstep("odd", fused=fused)


@pytest.mark.parametrize("name", ["edge", "edge24", "c1", "k1"])
@pytest.mark.parametrize("fused", ["auto", "off"])
def test_lockstep_edge_case_traces(name, fused):
    """Reference traces with empty / full / repeated inputs, one cell per column, one active column."""
    _lockstep(name, fused=fused)


def test_lockstep_many_columns_grid_kernel():
    """16384 columns: the cooperative-grid fused kernel with the grid-wide top-k."""
    import bithtm_b200 as bithtm

    I, C, c, k, seed = 256, 16384, 4, 328, 21
    np.random.seed(seed)
    htm = bithtm.HierarchicalTemporalMemory(I, C, c, k, fused="grid")
    orc = HTMOracle(OracleConfig(I, C, c, k), rng=np.random.RandomState(seed), overlap="packed")
    g = np.random.default_rng(seed)
    base = g.random((5, I)) < 0.2
    for t in range(40):
        x = base[t % 5] ^ (g.random(I) < 0.05)
        sp_state, tm_state = htm.process(x)
        rec = orc.step(x)
        problems = diff_records(gpu_record(htm, sp_state, tm_state), oracle_record(rec))
        assert not problems, f"step {t}: " + "; ".join(problems)
    assert gpu_state_digest(htm) == oracle_state_digest(orc)


@pytest.mark.parametrize("fused", ["cluster", "auto"])
def test_lockstep_many_cells_many_draws(fused):
    """k = 328 active columns x 32 cells: a step draws enough random numbers for many-CTA stream
    production.  fused="auto" must pick the cooperative grid (which has the production phase); an explicit
    fused="cluster" must fall back to one-CTA production instead of leaving planned chunks unproduced."""
    import bithtm_b200 as bithtm

    I, C, c, k, seed = 256, 16384, 32, 328, 23
    np.random.seed(seed)
    htm = bithtm.HierarchicalTemporalMemory(I, C, c, k, fused=fused, max_segments=1 << 16)
    eng = htm.engine
    assert (eng.ctx.fused_mode, eng.ctx.jump_polys > 0) == ((1, False) if fused == "cluster" else (2, True))
    orc = HTMOracle(OracleConfig(I, C, c, k), rng=np.random.RandomState(seed), overlap="packed")
    g = np.random.default_rng(seed)
    base = g.random((4, I)) < 0.2
    for t in range(24):
        x = base[t % 4] ^ (g.random(I) < 0.05)
        sp_state, tm_state = htm.process(x)
        rec = orc.step(x)
        problems = diff_records(gpu_record(htm, sp_state, tm_state), oracle_record(rec))
        assert not problems, f"step {t}: " + "; ".join(problems)
    assert gpu_state_digest(htm) == oracle_state_digest(orc)
    assert np.array_equal(np.random.random_sample(64), orc.rng.random_sample(64))


def test_lockstep_mid():
    _lockstep("mid", check_every=3)


@pytest.mark.parametrize("fused", ["cluster", "grid", "off"])
def test_lockstep_cfg2_1500_steps(fused):
    _lockstep("cfg2", steps=1500, check_every=10, fused=fused)


def _golden_trace(name, **engine_kw):
    """Full-length run compared with the digests recorded from the reference."""
    import bithtm_b200 as bithtm

    info = load_golden(name)
    g = info["g"]
    xs = golden_inputs(info)
    np.random.seed(info["seed"])
    htm = bithtm.HierarchicalTemporalMemory(info["I"], info["C"], info["c"], info["k"], **engine_kw)
    state_at = {int(s): int(d) for s, d in zip(g["state_steps"], g["state_digests"])}
    for t in range(info["steps"]):
        sp_state, tm_state = htm.process(xs[t])
        d = step_digest(**gpu_record(htm, sp_state, tm_state))
        assert d == int(g["digests"][t]), f"{name}: step {t} differs from the reference"
        if t in state_at:
            assert gpu_state_digest(htm) == state_at[t], f"{name}: learned state at step {t}"
    assert tm_state.n_segments == int(g["n_segments_final"])
    assert htm.engine.check_status() & ~32 == 0


def test_cfg1_10k_steps_against_reference_trace():
    """BASELINE configs[0]: example.py defaults, 10 000 steps, bit-exact."""
    _golden_trace("cfg1")


def test_cfg2_10k_steps_against_reference_trace():
    """BASELINE configs[1]: 2048 columns x 1024 inputs, 10 000 steps, bit-exact."""
    _golden_trace("cfg2")


# ------------------------------------------------------------------ other call paths
def test_host_inhibition_mode_matches_argpartition_oracle():
    """Secondary parity mode: any host object in the reference's
    `inhibition=` slot; here np.argpartition itself, same NumPy on both sides."""
    import bithtm_b200 as bithtm

    class ArgpartitionInhibition:  # regularizations.py:24-29, verbatim semantics
        def __init__(self, k):
            self.k = k

        def process(self, x):
            return np.argpartition(x, -self.k)[-self.k:]

    I, C, c, k, seed = 128, 256, 16, 20, 4
    np.random.seed(seed)
    sp = bithtm.SpatialPooler(I, C, k, inhibition=ArgpartitionInhibition(k))
    htm = bithtm.HierarchicalTemporalMemory(I, C, c, k, spatial_pooler=sp)
    orc = HTMOracle(OracleConfig(I, C, c, k), rng=np.random.RandomState(seed), inhibition="argpartition")
    g = np.random.default_rng(seed)
    base = g.random((10, I)) < 0.25
    for t in range(400):
        x = base[t % 10] ^ (g.random(I) < 0.05)
        sp_state, tm_state = htm.process(x)
        rec = orc.step(x)
        problems = diff_records(gpu_record(htm, sp_state, tm_state), oracle_record(rec))
        assert not problems, f"step {t}: " + "; ".join(problems)
    assert gpu_state_digest(htm) == oracle_state_digest(orc)


def test_learning_off_matches_oracle():
    """learning=False: no permanence/segment updates but duty cycles still move
    (networks.py:31-33) and winners/jitter are still drawn."""
    import bithtm_b200 as bithtm

    I, C, c, k, seed = 96, 256, 12, 20, 9
    np.random.seed(seed)
    htm = bithtm.HierarchicalTemporalMemory(I, C, c, k)
    orc = HTMOracle(OracleConfig(I, C, c, k), rng=np.random.RandomState(seed))
    g = np.random.default_rng(seed)
    base = g.random((8, I)) < 0.25
    for t in range(300):
        x = base[t % 8] ^ (g.random(I) < 0.03)
        learning = not (100 <= t < 200)
        sp_state, tm_state = htm.process(x, learning=learning)
        rec = orc.step(x, learning=learning)
        problems = diff_records(gpu_record(htm, sp_state, tm_state), oracle_record(rec))
        assert not problems, f"step {t} (learning={learning}): " + "; ".join(problems)
    assert gpu_state_digest(htm) == oracle_state_digest(orc)


@pytest.mark.parametrize("fused", ["cluster", "off"])
def test_graph_replay_equals_stepwise(fused):
    """bh_step_ring under a CUDA graph (the bench's device-resident path) ends in
    the same learned state as the host-driven path."""
    import bithtm_b200 as bithtm

    info = load_golden("mid")
    steps = 500
    xs = golden_inputs(info, steps)
    np.random.seed(info["seed"])
    a = bithtm.HierarchicalTemporalMemory(info["I"], info["C"], info["c"], info["k"], rng_sync="lazy")
    np.random.seed(info["seed"])
    b = bithtm.HierarchicalTemporalMemory(info["I"], info["C"], info["c"], info["k"], rng_sync="lazy",
                                          ring_len=steps, fused=fused)
    for t in range(steps):
        a.process(xs[t])
    eng = b.engine
    b.temporal_memory._rng.before(eng)
    eng.load_ring(xs)
    per = 20
    handle = eng.graph(per, learning=True)
    for _ in range(steps // per):
        eng.launch_graph(handle, per)
    import torch

    torch.cuda.synchronize()
    assert gpu_state_digest(a) == gpu_state_digest(b)
    assert np.array_equal(a.engine.scalars()[:5], b.engine.scalars()[:5])
    info_g = info["g"]
    state_at = {int(s): int(d) for s, d in zip(info_g["state_steps"], info_g["state_digests"])}
    if steps - 1 in state_at:
        assert gpu_state_digest(b) == state_at[steps - 1]


def test_stale_state_read_raises():
    import bithtm_b200 as bithtm

    np.random.seed(1)
    htm = bithtm.HierarchicalTemporalMemory(64, 128, 8, 10)
    x = np.random.default_rng(0).random(64) < 0.3
    sp1, tm1 = htm.process(x)
    sp1.overlaps  # read in time: cached
    htm.process(x)
    assert sp1.overlaps is not None
    with pytest.raises(RuntimeError):
        tm1.distal_state.prediction


def test_export_shim_feeds_reference_style_reader():
    """The state export in the reference's storage vocabulary (segment_bundle,
    segment_projection.output_edge / output_permanence / invalid_output_edge /
    get_output_edge_target) read the way reference_implementations.py:51-70 reads it
    gives the oracle's synapses."""
    import bithtm_b200 as bithtm

    info = load_golden("tiny")
    xs = golden_inputs(info, 400)
    np.random.seed(info["seed"])
    htm = bithtm.HierarchicalTemporalMemory(info["I"], info["C"], info["c"], info["k"])
    orc = HTMOracle(OracleConfig(info["I"], info["C"], info["c"], info["k"]), rng=np.random.RandomState(info["seed"]))
    for t in range(400):
        htm.process(xs[t])
        orc.step(xs[t])
    dp = htm.temporal_memory.distal_projection
    proj = dp.segment_projection
    segment_cell = dp.segment_bundle[:].squeeze(1).tolist()
    assert segment_cell == orc.seg_owner[:orc.n_seg].tolist()
    assert np.array_equal(dp.bundle_segments, orc.cell_nseg)
    for seg, (synapses, permanences) in enumerate(zip(proj.output_edge[:], proj.output_permanence[:])):
        got = sorted((int(proj.get_output_edge_target(s)), float(p)) for s, p in zip(synapses, permanences)
                     if s != proj.invalid_output_edge)
        v = orc.syn_cell[seg] >= 0
        want = sorted(zip(orc.syn_cell[seg][v].tolist(), orc.syn_perm[seg][v].astype(float).tolist()))
        assert got == want, f"segment {seg}"
        assert int(proj.output_edges[seg, 0]) == len(want)
    last = htm.temporal_memory.last_state
    assert np.array_equal(htm.temporal_memory.flatten_cell(last.active_cell),
                          htm.temporal_memory.flatten_cell(last.active_cell))


def test_example_driver_stream_matches_oracle():
    """example.py's loop (np.random used by the caller between steps, example.py:34,52):
    the interleaved global stream is consumed exactly as the reference consumes it."""
    import bithtm_b200 as bithtm

    I, C, c, seed = 128, 256, 8, 13
    k = 20
    np.random.seed(seed)
    inputs = np.random.rand(7, I) < 0.25
    htm = bithtm.HierarchicalTemporalMemory(I, C, c, k)
    rs = np.random.RandomState(seed)
    inputs_o = rs.rand(7, I) < 0.25
    orc = HTMOracle(OracleConfig(I, C, c, k), rng=rs)
    assert np.array_equal(inputs, inputs_o)
    for t in range(200):
        x = inputs[t % 7] ^ (np.random.rand(I) < 0.05)
        xo = inputs_o[t % 7] ^ (rs.rand(I) < 0.05)
        assert np.array_equal(x, xo), f"caller-visible np.random diverged at step {t}"
        prev_col_pred = orc.cell_prediction.max(axis=1)  # example.py:50
        sp_state, tm_state = htm.process(x)
        rec = orc.step(xo)
        problems = diff_records(gpu_record(htm, sp_state, tm_state), oracle_record(rec))
        assert not problems, f"step {t}: " + "; ".join(problems)
        # the demo's three counts (example.py:55-57), counted on the device and carried by the step summary
        corrects = int(prev_col_pred[rec.active_column].sum())
        assert tm_state.column_metrics == {
            "bursting": int(rec.bursting.sum()), "correct": corrects, "incorrect": int(prev_col_pred.sum()) - corrects,
            "predicted_columns": int(orc.cell_prediction.max(axis=1).sum())}, t


@pytest.mark.parametrize("extra", [[], ["--readback"]])
def test_example_script_prints_the_oracle_counts(extra):
    """example.py (the device-backed copy of the reference's demo, same CLI) run as a script: the per-epoch
    bursting / correct / incorrect totals equal the oracle's on the same seed, with the metrics taken from the
    device counters and -- `--readback` -- computed on the host as the reference does."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    args = ["--epochs", "4", "--input_patterns", "25", "--input_dim", "200", "--column_dim", "512", "--cell_dim", "8",
            "--seed", "5", "--quiet"]
    out = subprocess.run([sys.executable, os.path.join(root, "example.py"), *args, *extra], capture_output=True,
                         text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    rs = np.random.RandomState(5)
    inputs = rs.rand(25, 200) < 0.2
    orc = HTMOracle(OracleConfig(200, 512, 8), rng=rs)
    want = []
    for epoch in range(4):
        tot = np.zeros(3, dtype=np.int64)
        for x0 in inputs:
            prev = orc.cell_prediction.max(axis=1)
            rec = orc.step(x0 ^ (rs.rand(200) < 0.05))
            corrects = int(prev[rec.active_column].sum())
            tot += (int(rec.bursting.sum()), corrects, int(prev.sum()) - corrects)
        want.append(f"epoch {epoch}: bursting {tot[0]}, correct {tot[1]}, incorrect {tot[2]}")
    got = [ln for ln in out.stdout.splitlines() if ln.startswith("epoch")]
    assert got == want


@pytest.mark.gpu
def test_stream_batch_equals_streams_stepped_alone():
    """Independent streams advanced side by side by one CUDA graph (StreamBatch, cfg4's
    mode) end in exactly the state each reaches when stepped alone."""
    import bithtm_b200 as bithtm

    info = load_golden("tiny")
    I, C, c, k = info["I"], info["C"], info["c"], info["k"]
    steps, B = 120, 5
    seeds = [info["seed"] + 17 * i for i in range(B)]
    inputs = [make_inputs(I, info["patterns"], info["density"], info["noise"], steps, s) for s in seeds]

    def build(seed):
        np.random.seed(seed)
        h = bithtm.HierarchicalTemporalMemory(I, C, c, k, rng_sync="lazy", ring_len=steps, max_segments=1 << 12,
                                              fused="cluster", fused_ctas=4, fused_threads=512)
        return h, np.random.get_state()  # every stream continues from its own state

    alone = []
    for s, xs in zip(seeds, inputs):
        h, state = build(s)
        eng = h.engine
        h.temporal_memory._rng.adopt(eng, state)
        eng.load_ring(xs)
        eng.launch_graph(eng.graph(steps, learning=True), steps)
        alone.append(gpu_state_digest(h))
    built = [build(s) for s in seeds]
    nets = [h for h, _ in built]
    batch = bithtm.StreamBatch(nets)
    batch.load_inputs(inputs, rng_states=[st for _, st in built])
    for _ in range(4):
        batch.run(steps // 4)
    batch.check_status()
    assert [gpu_state_digest(h) for h in nets] == alone
    # and they really are different streams
    assert len(set(alone)) == B


@pytest.mark.gpu
@pytest.mark.parametrize("fused", ["cluster", "grid", "off", "shard"])
def test_mixed_learning_and_winner_flags_match_reference_trace(fused):
    """Per-step (learning, return_winner_cell) flags of TemporalMemory.process (networks.py:91):
    inference-only steps draw nothing, the jitter draw is deferred until a later step needs
    it, no growth after a step without winner cells -- against the trace recorded from the
    unmodified reference (tests/golden/mixed.npz), through HierarchicalTemporalMemory.process:
    every flag combination is ONE launch of the fused step kernel (or the per-stage kernels with "off")."""
    import bithtm_b200 as bithtm

    info = load_golden("mixed")
    g = info["g"]
    I, C, c, k, seed, steps = info["I"], info["C"], info["c"], info["k"], info["seed"], info["steps"]
    xs = golden_inputs(info, steps)
    np.random.seed(seed)
    htm = bithtm.HierarchicalTemporalMemory(I, C, c, k, fused=fused, **({"fused_ctas": 12} if fused == "shard" else {}))
    orc = HTMOracle(OracleConfig(I, C, c, k), rng=np.random.RandomState(seed))
    state_at = {int(s): int(d) for s, d in zip(g["state_steps"], g["state_digests"])}
    for t in range(steps):
        lf, wf = bool(g["learning_flags"][t]), bool(g["winner_flags"][t])
        rec = orc.step(xs[t], learning=lf, return_winner_cell=wf)
        sp_state, tm_state = htm.process(xs[t], learning=lf, return_winner_cell=wf)
        ds = tm_state.distal_state
        wc = tm_state.winner_cell
        assert (wc is None) == (not (lf or wf))
        assert (ds.matching_segment_jittered_potential is None) == (not wf)
        got = dict(
            n_segments=tm_state.n_segments, overlaps=sp_state.overlaps, boosted=sp_state.boosted_overlaps,
            active_column=sp_state.active_column, bursting=tm_state.active_column_bursting,
            winner_cell=wc[0] * c + wc[1] if wc is not None else np.zeros(0, dtype=np.int64),
            active_cell=tm_state.active_cell[0] * c + tm_state.active_cell[1],
            matching_segment=ds.matching_segment, matching_activation=ds.matching_segment_activation,
            matching_jit=ds.matching_segment_jittered_potential if wf else np.zeros(0, dtype=np.float32))
        d = diff_records(got, oracle_record(rec))
        assert not d, f"step {t} (learning={lf}, return_winner_cell={wf}): {d}"
        assert step_digest(**got) == int(g["digests"][t]), f"step {t}: differs from the reference trace"
        if t in state_at:
            assert gpu_state_digest(htm) == state_at[t], f"learned state at step {t}"
    # the caller's np.random ends where the reference's does
    rs = np.random.RandomState(seed)
    rs.randn(C, I)
    rs.random_sample(int(np.sum(g["draws"])))
    assert np.array_equal(np.random.random_sample(2000), rs.random_sample(2000))  # same continuation


@pytest.mark.gpu
@pytest.mark.parametrize("fused", ["cluster", "grid", "off"])
def test_degenerate_inputs_match_oracle(fused):
    """Empty input (every overlap 0: the whole top-k is one tie, lowest columns win), full
    input, and the same input repeated -- interleaved with ordinary ones."""
    import bithtm_b200 as bithtm

    I, C, c, k, seed = 96, 320, 16, 12, 21
    g = np.random.default_rng(5)
    xs = []
    for t in range(160):
        r = t % 8
        if r == 3:
            xs.append(np.zeros(I, dtype=bool))
        elif r == 6:
            xs.append(np.ones(I, dtype=bool))
        elif r == 7:
            xs.append(xs[-2].copy())
        else:
            xs.append(g.random(I) < 0.25)
    np.random.seed(seed)
    htm = bithtm.HierarchicalTemporalMemory(I, C, c, k, fused=fused, fused_ctas=5 if fused != "off" else None)
    orc = HTMOracle(OracleConfig(I, C, c, k), rng=np.random.RandomState(seed))
    for t, x in enumerate(xs):
        rec = orc.step(x)
        sp_state, tm_state = htm.process(x)
        d = diff_records(gpu_record(htm, sp_state, tm_state), oracle_record(rec))
        assert not d, f"step {t}: {d}"
    assert gpu_state_digest(htm) == oracle_state_digest(orc)


@pytest.mark.gpu
def test_capacity_overflow_raises():
    """The reference grows its arrays without bound; here an exhausted capacity raises and names it."""
    import bithtm_b200 as bithtm
    from bithtm_b200 import _native as nat

    info = load_golden("tiny")
    xs = golden_inputs(info, 200)
    np.random.seed(info["seed"])
    htm = bithtm.HierarchicalTemporalMemory(info["I"], info["C"], info["c"], info["k"], max_segments=64)
    with pytest.raises(nat.NativeError, match="max_segments"):
        for t in range(200):
            htm.process(xs[t])
    np.random.seed(info["seed"])
    htm = bithtm.HierarchicalTemporalMemory(info["I"], info["C"], info["c"], info["k"], max_synapses_per_segment=32,
                                            max_segments=4096)
    # 32 slots hold a freshly grown segment exactly; later growth on top of surviving synapses does not fit
    with pytest.raises(nat.NativeError, match="max_synapses_per_segment"):
        for t in range(1500):
            htm.process(xs[t % 200])


@pytest.mark.gpu
@pytest.mark.parametrize("I,C,B", [(1024, 2048, 300), (200, 300, 33), (4100, 512, 64), (97, 131, 1), (1024, 2048, 1024),
                                   (1000, 700, 130), (4096, 515, 129), (2048, 4096, 64), (512, 128, 257)])
def test_batched_overlap_equals_loop_of_process(I, C, B):
    """DenseProjection.process_batch == a loop of DenseProjection.process (projections.py:18-21)
    == the dense float64 comparison of the oracle: the auto-dispatched tensor-core path (tcgen05 + TMEM + TMA
    where the shape allows, ragged tiles included), both of its kernels forced, and the popcount kernel."""
    import bithtm_b200 as bithtm

    np.random.seed(9)
    proj = bithtm.projections.DenseProjection(I, C)
    perm = proj.permanence.copy()
    sp = bithtm.SpatialPooler(I, C, max(1, round(0.02 * C)), proximal_projection=proj)
    sp._ensure_engine()
    g = np.random.default_rng(2)
    xs = g.random((B, I)) < 0.3
    got = proj.process_batch(xs)  # int8 tensor-core contraction
    got_popc = proj.process_batch(xs, tensor_core=False)
    want = ((perm >= 0.0)[None, :, :] & xs[:, None, :]).sum(axis=2) if B * C * I < 3e8 else None
    if want is None:
        want = np.stack([((perm >= 0.0) & x).sum(axis=1) for x in xs])
    assert got.dtype == np.int64 and np.array_equal(got, want)
    assert got_popc.dtype == np.int64 and np.array_equal(got_popc, want)
    assert np.array_equal(proj.process_batch(xs, tensor_core="mma"), want)
    if ((I + 31) // 32) % 4 == 0:  # TMA needs 16-byte row pitches
        assert np.array_equal(proj.process_batch(xs, tensor_core="tcgen05"), want)
    else:
        with pytest.raises(Exception):
            proj.process_batch(xs, tensor_core="tcgen05")
    assert np.array_equal(proj.process(xs[min(5, B - 1)]), want[min(5, B - 1)])


def _execution_modes_agree(I, C, c, k, steps, patterns, kw_a, kw_b, min_segments, **common):
    """Two execution modes of the same network, `steps` device-resident timesteps each: bit-identical state
    (permanence, masks, duty cycles, segments, synapses, RNG stream).  A size-independent property for sizes
    and run lengths the oracle cannot follow."""
    import torch

    import bithtm_b200 as bithtm
    from bithtm_b200.projections import DenseProjection

    g = np.random.default_rng(8)
    base = g.random((patterns, I)) < 0.2
    xs = base[np.arange(steps) % patterns] ^ (g.random((steps, I)) < 0.05)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    perm = torch.randn(C, I, dtype=torch.float64, device="cuda", generator=gen) * 0.1

    def run(kw):
        kw = dict(kw)
        per_launch = kw.pop("_per_launch", 0)  # > 0: steps from the device input ring, this many per kernel launch
        flags = kw.pop("_flags", True)          # the `learning` flag word of the launches (BH_STEP_*)
        np.random.seed(12)
        sp = bithtm.SpatialPooler(I, C, k, proximal_projection=DenseProjection(I, C, permanence=perm))
        htm = bithtm.HierarchicalTemporalMemory(I, C, c, k, spatial_pooler=sp, rng_sync="lazy",
                                                **({"ring_len": steps} if per_launch else {}), **common, **kw)
        eng = htm.engine
        htm.temporal_memory._rng.before(eng)
        if per_launch:
            eng.load_ring(xs)
            left = steps
            while left > 0:
                m = min(left, per_launch)
                eng.launch_graph(eng.graph(m, learning=flags), m)
                left -= m
        else:
            for t in range(steps):
                htm.process(eng.pack_input(xs[t]), return_state=False)
        torch.cuda.synchronize()
        eng.check_status()
        return htm

    a = run(kw_a)
    b = run(kw_b)
    ea, eb = a.engine, b.engine
    S = int(ea.scalars()[2])
    assert S == int(eb.scalars()[2]) and S > min_segments
    for name in ("sp_perm", "sp_mask", "duty", "overlaps", "boosted", "col_pred", "col_act", "col_win", "cell_nseg",
                 "cell_maxjit", "cell_npred"):
        assert torch.equal(ea.buf[name], eb.buf[name]), name
    for name in ("seg_owner", "seg_count", "seg_pot", "seg_conn"):
        assert torch.equal(ea.buf[name][:S], eb.buf[name][:S]), name
    E = ea.ctx.syn_capacity
    live = torch.arange(E, device="cuda")[None, :] < ea.buf["seg_count"][:S, None]
    for name in ("syn_cell", "syn_perm"):
        x, y = ea.buf[name][:S * E].view(S, E), eb.buf[name][:S * E].view(S, E)
        assert torch.equal(torch.where(live, x, torch.zeros_like(x)), torch.where(live, y, torch.zeros_like(y))), name
    ka, pa = ea.get_rng_state()
    kb, pb = eb.get_rng_state()
    ra, rb = np.random.RandomState(), np.random.RandomState()
    ra.set_state(("MT19937", ka, pa, 0, 0.0))
    rb.set_state(("MT19937", kb, pb, 0, 0.0))
    assert np.array_equal(ra.random_sample(1000), rb.random_sample(1000))
    return a, b


@pytest.mark.gpu
def test_cfg3_full_size_execution_modes_agree():
    """BASELINE configs[2] at its full size (65536 columns x 16384 inputs, 32 cells, k = 1311), next to the
    oracle lock-step tests below: the whole step as one cooperative kernel and one kernel per stage leave
    bit-identical state on the device after 40 steps."""
    import torch

    free, _ = torch.cuda.mem_get_info()
    if free < 40 << 30:
        pytest.skip("needs ~24 GiB of device memory")
    a, b = _execution_modes_agree(16384, 65536, 32, 1311, 40, 10, dict(fused="grid"), dict(fused="off"), 20000,
                                  max_segments=1 << 17, max_synapses_per_segment=128)
    assert a.engine.ctx.jump_polys > 0 and a.engine.ctx.fused_mode == 2 and b.engine.ctx.fused_mode == 0


@pytest.mark.gpu
@pytest.mark.parametrize("per_launch,team,lazy", [(2, 40, "auto"), (7, 40, "auto"), (50, 64, "auto"), (9, 24, False)])
def test_pipelined_grid_kernel_equals_grid_kernel(per_launch, team, lazy):
    """The two-pipeline kernel (spatial pooler of step s+1 beside the temporal memory of step s, csrc/fused.cuh
    k_step_pipe) against the one-pipeline grid kernel over 700 steps, with 1 / 7 / 50 steps per launch (a
    launch's first step computes its selection on the whole grid, its last one runs no look-ahead): identical
    permanence, masks, duty cycles, overlaps, segments, synapses and stream position."""
    a, b = _execution_modes_agree(1024, 16384, 16, 328, 700, 12,
                                  dict(fused="grid", lazy_rng=lazy, pipeline=team, _per_launch=per_launch),
                                  dict(fused="grid", lazy_rng=lazy, pipeline=0), 4000, max_segments=1 << 16)
    assert a.engine.ctx.pipe_ctas == team and b.engine.ctx.pipe_ctas == 0
    sa, sb = a.engine.scalars(), b.engine.scalars()
    assert np.array_equal(sa[:13], sb[:13])
    k = a.engine.k
    cur = (int(sa[0]) - 1) & 1
    assert np.array_equal(a.engine.buf["active_cols"][cur * k:(cur + 1) * k].cpu().numpy(),
                          b.engine.buf["active_cols"][cur * k:(cur + 1) * k].cpu().numpy())
    assert np.array_equal(a.engine.buf["col_active"].cpu().numpy(), b.engine.buf["col_active"].cpu().numpy())


@pytest.mark.gpu
def test_pipelined_grid_kernel_without_winner_cells():
    """The same with learning on and return_winner_cell off (flag word 3): the jitter draw of every step stays
    pending until the next one (no early draw #1 in the two-pipeline kernel), 300 steps, 6 per launch."""
    _execution_modes_agree(1024, 16384, 16, 328, 300, 12,
                           dict(fused="grid", pipeline=40, _per_launch=6, _flags=3),
                           dict(fused="grid", pipeline=0, _per_launch=1, _flags=3), 2000, max_segments=1 << 16)


@pytest.mark.gpu
@pytest.mark.parametrize("lazy", ["auto", False])
def test_grid_kernel_long_run_equals_per_stage(lazy):
    """700 steps of a 16384-column network (the steady state with its predicted-histogram selection, the
    bookkeeping team, lazy draws and their occasional fall-backs) against one kernel per stage."""
    a, b = _execution_modes_agree(1024, 16384, 16, 328, 700, 12, dict(fused="grid", lazy_rng=lazy), dict(fused="off"),
                                  4000, max_segments=1 << 16)
    assert a.engine.ctx.fused_mode == 2 and (a.engine.ctx.skip_polys > 0) == (lazy == "auto")


def _cfg3_lockstep(fused, steps, patterns=5, **engine_kw):
    """BASELINE configs[2] at its FULL size (65536 columns x 16384 inputs, 32 cells, k = 1311) lock-step
    against the ORACLE (networks.py:146-149): the float64 permanence is drawn once on the device, downloaded
    and handed to both sides; every State field of every step and the learned state are compared.  Few input
    patterns, so that predictions, learning / punished segments and growth of existing segments (not only
    the all-bursting start) happen inside the short run."""
    import torch

    import bithtm_b200 as bithtm
    from bithtm_b200.projections import DenseProjection

    free, _ = torch.cuda.mem_get_info()
    if free < 40 << 30:
        pytest.skip("needs ~24 GiB of device memory")
    I, C, c, k, seed = 16384, 65536, 32, 1311, 12
    g = np.random.default_rng(8)
    base = g.random((patterns, I)) < 0.2
    xs = base[np.arange(steps) % patterns] ^ (g.random((steps, I)) < 0.05)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    perm = torch.randn(C, I, dtype=torch.float64, device="cuda", generator=gen) * 0.1
    host_perm = perm.cpu().numpy()
    np.random.seed(seed)
    sp = bithtm.SpatialPooler(I, C, k, proximal_projection=DenseProjection(I, C, permanence=perm))
    htm = bithtm.HierarchicalTemporalMemory(I, C, c, k, spatial_pooler=sp, fused=fused, max_segments=1 << 17,
                                            max_synapses_per_segment=128, **engine_kw)
    del perm
    orc = HTMOracle(OracleConfig(I, C, c, k), rng=np.random.RandomState(seed), overlap="packed",
                    permanence=host_perm)
    stats = dict(predicted=0, learning=0, matching=0, punished=0, draws=0)
    for t in range(steps):
        sp_state, tm_state = htm.process(xs[t])
        rec = orc.step(xs[t])
        problems = diff_records(gpu_record(htm, sp_state, tm_state), oracle_record(rec))
        assert not problems, f"cfg3 {fused} step {t}: " + "; ".join(problems) + f"; sc={htm.engine.scalars()[:18]}"
        stats["predicted"] += int((~rec.bursting).sum())
        stats["learning"] += len(rec.learning_segment)
        stats["matching"] += len(rec.matching_segment)
        stats["punished"] += len(rec.punished_segment)
        stats["draws"] += rec.draws
    print(f"cfg3 lock-step ({fused}, {steps} steps): {stats}, {rec.n_segments} segments")
    # the run exercised more than the all-bursting start: segments matched and were re-learned / punished
    assert stats["matching"] > 0 and stats["learning"] > steps and stats["draws"] > steps * k * c
    eng = htm.engine
    assert eng.check_status() & ~32 == 0
    # learned state: SP permanence (8 GiB, compared directly), duty cycles, segments, synapses
    got_perm = eng.buf["sp_perm"].cpu().numpy().reshape(C, I)
    assert np.array_equal(got_perm.view(np.uint64), orc.permanence.view(np.uint64)), "SP permanence bits"
    del got_perm
    tmp = htm.temporal_memory.distal_projection
    owner, count, cells, perm_rows = tmp.export_segments()
    assert state_digest(np.zeros(1), htm.spatial_pooler.boosting.duty_cycle, tmp.bundle_segments,
                        canonical_from_rows(owner, cells, perm_rows)) == \
        state_digest(np.zeros(1), orc.duty, orc.cell_nseg, orc.canonical_synapses())
    # the caller's np.random stream ends where the oracle's does
    assert np.array_equal(np.random.random_sample(64), orc.rng.random_sample(64))
    return htm


@pytest.mark.gpu
def test_cfg3_lockstep_oracle_fused_grid():
    """The whole step as one cooperative kernel (long-row overlap, wide SP learning, many-CTA stream
    production, grid-wide top-k at 148 CTAs) against the oracle at cfg3's full size."""
    htm = _cfg3_lockstep("grid", 48, patterns=3)
    assert htm.engine.ctx.fused_mode == 2


@pytest.mark.gpu
def test_cfg3_lockstep_oracle_lazy_draws():
    """The same with every large rand(L, W+1) drawn lazily (jump table of 4096-word steps): rows, chunks and tail."""
    htm = _cfg3_lockstep("grid", 24, patterns=3, lazy_rng="always")
    assert htm.engine.ctx.skip_polys > 0 and htm.engine.ctx.lazy_policy == 1


@pytest.mark.gpu
def test_cfg3_lockstep_oracle_per_stage():
    """One kernel per stage (the fine-grained C entry points) against the oracle at cfg3's full size."""
    htm = _cfg3_lockstep("off", 16, patterns=3)
    assert htm.engine.ctx.fused_mode == 0


# ------------------------------------------------------------------ the plugin surface itself
def test_predictive_projection_plugin_methods_match_oracle():
    """PredictiveProjection.update / .process / .get_jittered_potential_info called with explicit arguments,
    the way the reference's own TemporalMemory.process calls them (networks.py:106-113, 121), against
    HTMOracle.tm_learn / tm_activate; then the network classes take over again (fused path) and stay
    bit-exact."""
    import bithtm_b200 as bithtm

    I, C, c, k, seed = 96, 256, 12, 24, 9
    np.random.seed(seed)
    htm = bithtm.HierarchicalTemporalMemory(I, C, c, k, fused="off")
    orc = HTMOracle(OracleConfig(I, C, c, k), rng=np.random.RandomState(seed))
    pp = htm.temporal_memory.distal_projection
    g = np.random.default_rng(seed)
    base = g.random((6, I)) < 0.25
    prev_state, prev_winners, prev_activation = None, None, np.zeros((C, c), dtype=bool)
    for t in range(260):
        x = base[t % 6] ^ (g.random(I) < 0.05)
        want_jit = t % 7 != 3  # sometimes leave the jitter to the next get_jittered_potential_info
        sp_state = htm.spatial_pooler.process(x)
        ac = orc.sp_inhibit(orc.sp_boost(orc.sp_overlap(x)))
        orc.sp_learn(x, ac)
        orc.sp_duty_update(ac)
        assert np.array_equal(sp_state.active_column, ac)
        # networks.py:99-104 on the host (the oracle's restatement); the caller's own rand(k, c)
        if prev_state is not None:
            mj, _ = pp.get_jittered_potential_info(prev_state)  # networks.py:76
        acp, burst, winners, _ = orc.tm_select(ac)
        if prev_state is not None:
            assert np.array_equal(mj.view(np.uint32), orc.max_jit.view(np.uint32))
        np.random.rand(k, c)  # networks.py:87: keeps the global stream where the reference's would be
        punish = np.ones(C, dtype=bool)
        punish[ac] = False
        learn, _, _, _ = orc.tm_learn(ac, winners)
        pp.update(prev_state, prev_activation.reshape(-1), winners, np.repeat(punish, c), winner_input=prev_winners)
        act_rows = acp | burst[:, None]
        activation = np.zeros((C, c), dtype=bool)
        activation[ac] = act_rows
        rows, cells = np.nonzero(act_rows)
        active_flat = ac[rows] * c + cells
        orc.tm_activate(activation.reshape(-1), want_jitter=want_jit)
        st = pp.process(active_flat, return_jittered_potential_info=want_jit)
        assert np.array_equal(st.segment_potential, orc.seg_potential), t
        assert np.array_equal(st.matching_segment, orc.m_seg), t
        assert np.array_equal(st.matching_segment_activation, orc.m_act), t
        assert np.array_equal(st.prediction, orc.npred.astype(np.float64)), t
        assert np.array_equal(pp.bundle_segments, orc.cell_nseg), t
        if want_jit:
            assert np.array_equal(st.matching_segment_jittered_potential.view(np.uint32), orc.m_jit.view(np.uint32)), t
            assert np.array_equal(st.max_jittered_potential.view(np.uint32), orc.max_jit.view(np.uint32)), t
        else:
            assert st.max_jittered_potential is None
        assert pp.n_segments == orc.n_seg
        orc.cell_activation, orc.prev_winners = activation, winners
        prev_state, prev_winners, prev_activation = st, winners, activation
    assert gpu_state_digest(htm) == oracle_state_digest(orc)
    assert orc.n_seg > 50 and int((~burst).sum()) > 0
    # hand over to the network classes: the fused / staged paths continue from the same device state
    for t in range(260, 330):
        x = base[t % 6] ^ (g.random(I) < 0.05)
        sp_state, tm_state = htm.process(x)
        rec = orc.step(x)
        problems = diff_records(gpu_record(htm, sp_state, tm_state), oracle_record(rec))
        assert not problems, f"after hand-over, step {t}: " + "; ".join(problems)
    assert gpu_state_digest(htm) == oracle_state_digest(orc)
    assert np.array_equal(np.random.random_sample(64), orc.rng.random_sample(64))


def test_stand_alone_projection_allocates_its_own_device_state():
    import bithtm_b200 as bithtm

    pp = bithtm.projections.PredictiveProjection(64 * 4, cell_dim=4, active_columns=6)
    st = pp.process(np.array([0, 5, 9]))
    assert len(st.matching_segment) == 0 and st.prediction.shape == (256,) and not st.prediction.any()
    with pytest.raises(RuntimeError):
        bithtm.projections.PredictiveProjection(64).process(np.array([1]))


@pytest.mark.parametrize("fused", ["cluster", "off"])
def test_explicit_empty_prev_state_starts_a_new_sequence(fused):
    """TemporalMemory.process(sp_state, prev_state=get_empty_state()) (networks.py:91-93 with the state of
    :59-65): predictions, activation, winner cells and the distal state of the previous step are forgotten,
    the learned state is kept."""
    import bithtm_b200 as bithtm

    info = load_golden("tiny")
    I, C, c, k, seed = info["I"], info["C"], info["c"], info["k"], 31
    np.random.seed(seed)
    htm = bithtm.HierarchicalTemporalMemory(I, C, c, k, fused=fused)
    orc = HTMOracle(OracleConfig(I, C, c, k), rng=np.random.RandomState(seed))
    xs = make_inputs(I, 5, 0.25, 0.05, 240, seed)
    for t in range(240):
        if t in (90, 91, 170):  # forget the previous step
            orc.have_prev, orc.prev_winners, orc.jit_pending = False, None, False
            orc.cell_prediction = np.zeros((C, c), dtype=bool)
            orc.cell_activation = np.zeros((C, c), dtype=bool)
            if t == 170:
                htm.reset_sequence()
                sp_state, tm_state = htm.process(xs[t])
            else:
                sp_state = htm.spatial_pooler.process(xs[t])
                sp_state.overlaps, sp_state.boosted_overlaps, sp_state.active_column  # read before the TM step
                tm_state = htm.temporal_memory.process(sp_state, prev_state=htm.temporal_memory.get_empty_state())
        else:
            sp_state, tm_state = htm.process(xs[t])
        rec = orc.step(xs[t])
        problems = diff_records(gpu_record(htm, sp_state, tm_state), oracle_record(rec))
        assert not problems, f"step {t}: " + "; ".join(problems)
    assert gpu_state_digest(htm) == oracle_state_digest(orc)
    with pytest.raises(NotImplementedError):
        htm.temporal_memory.process(sp_state, prev_state=object())


@pytest.mark.parametrize("fused", ["cluster", "grid", "off"])
def test_checkpoint_resume_continues_bit_identically(fused):
    """state_dict() -> a NEW network -> load_state_dict(): the resumed run equals the uninterrupted one
    (every State field each step, learned state, np.random position)."""
    import bithtm_b200 as bithtm

    info = load_golden("tiny")
    I, C, c, k, seed = info["I"], info["C"], info["c"], info["k"], 17
    xs = make_inputs(I, 5, 0.25, 0.05, 260, seed)
    np.random.seed(seed)
    a = bithtm.HierarchicalTemporalMemory(I, C, c, k, fused=fused)
    for t in range(150):
        a.process(xs[t])
    snap = a.state_dict()
    digests = []
    for t in range(150, 260):
        sp_state, tm_state = a.process(xs[t])
        digests.append(step_digest(**gpu_record(a, sp_state, tm_state)))
    final_a, tail_a = gpu_state_digest(a), np.random.random_sample(16)
    np.random.seed(12345)  # an unrelated stream position: load_state_dict must restore the checkpoint's
    b = bithtm.HierarchicalTemporalMemory(I, C, c, k, fused=fused, max_segments=4096)
    b.load_state_dict(snap)
    for t in range(150, 260):
        sp_state, tm_state = b.process(xs[t])
        assert step_digest(**gpu_record(b, sp_state, tm_state)) == digests[t - 150], f"resumed run differs at step {t}"
    assert gpu_state_digest(b) == final_a
    assert np.array_equal(np.random.random_sample(16), tail_a)


def test_import_segments_is_the_inverse_of_export_segments():
    """export_segments() of one network -> import_segments() + permanence / duty cycles into a fresh one:
    both continue identically from an empty previous state."""
    import torch

    import bithtm_b200 as bithtm
    from bithtm_b200.projections import DenseProjection

    info = load_golden("tiny")
    I, C, c, k, seed = info["I"], info["C"], info["c"], info["k"], 41
    xs = make_inputs(I, 5, 0.25, 0.05, 300, seed)
    np.random.seed(seed)
    a = bithtm.HierarchicalTemporalMemory(I, C, c, k)
    for t in range(200):
        a.process(xs[t])
    tmp = a.temporal_memory.distal_projection
    owner, count, cells, perm = tmp.export_segments()
    sp_perm, duty = a.spatial_pooler.proximal_projection.permanence, a.spatial_pooler.boosting.duty_cycle
    sp = bithtm.SpatialPooler(I, C, k, proximal_projection=DenseProjection(I, C, permanence=sp_perm))
    b = bithtm.HierarchicalTemporalMemory(I, C, c, k, spatial_pooler=sp)
    b.engine.buf["duty"].copy_(torch.from_numpy(duty).to(b.engine.device))
    shuffled = np.random.default_rng(3).permutation(cells.shape[1])  # slot order is free
    b.temporal_memory.distal_projection.import_segments(owner, count, cells[:, shuffled], perm[:, shuffled])
    assert gpu_state_digest(a) == gpu_state_digest(b)
    a.reset_sequence()
    state = np.random.get_state()
    da = []
    for t in range(200, 300):
        sp_state, tm_state = a.process(xs[t])
        da.append(step_digest(**gpu_record(a, sp_state, tm_state)))
    np.random.set_state(state)
    for t in range(200, 300):
        sp_state, tm_state = b.process(xs[t])
        assert step_digest(**gpu_record(b, sp_state, tm_state)) == da[t - 200], t
    assert gpu_state_digest(a) == gpu_state_digest(b)


def test_global_inhibition_plugin_orders_negative_values():
    """GlobalInhibition.process on an arbitrary host array (regularizations.py:28-29 accepts any sign)."""
    import bithtm_b200 as bithtm

    C, k = 300, 25
    sp = bithtm.SpatialPooler(200, C, k)
    sp.process(np.zeros(200, dtype=bool))
    g = np.random.default_rng(1)
    for keys in (g.standard_normal(C), -g.random(C), np.where(g.random(C) < 0.5, -0.0, 0.0),
                 np.concatenate([np.full(150, -1.5), np.full(150, 2.0)])):
        assert np.array_equal(sp.inhibition.process(keys), canonical_topk(keys + 0.0, k))
