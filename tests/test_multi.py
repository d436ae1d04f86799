"""Multi-rank path: the top-k candidate exchange on CPU with gloo
(world_size 2, runs anywhere) and the full column-sharded SP + replicated TM on two
GPUs with NCCL (skipped without 2 GPUs)."""

import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _init(rank, world, port, backend):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    dist.init_process_group(backend, rank=rank, world_size=world)


# ----------------------------------------------------------------------------- CPU / gloo
def _gloo_worker(rank, world, port, result):
    import torch
    import torch.distributed as dist

    _init(rank, world, port, "gloo")
    from bithtm_b200._shard import gather_candidates, gather_columns, gather_records
    from oracle.htm_oracle import canonical_topk

    g = np.random.default_rng(5)
    ok = True
    for trial in range(40):
        C, k = 64 * world, int(g.integers(1, 40))
        # few distinct values -> many ties, also across the shard boundary
        keys = g.integers(0, 6 if trial % 2 else 1000, size=C).astype(np.float64)
        lo, hi = rank * C // world, (rank + 1) * C // world
        k_loc = min(k, hi - lo)
        local = canonical_topk(keys[lo:hi], k_loc)  # ascending local positions
        ck, cc = gather_candidates(torch.from_numpy(keys[lo:hi][local]), torch.from_numpy((lo + local).astype(np.int32)))
        ck, cc = ck.numpy(), cc.numpy()
        assert np.all(np.diff(cc) > 0), "gathered candidates must be in ascending column order"
        sel = cc[canonical_topk(ck, k)]  # position tie-break == column tie-break
        ok &= np.array_equal(sel, canonical_topk(keys, k))
        full = gather_columns(torch.from_numpy(keys[lo:hi])).numpy()
        ok &= np.array_equal(full, keys)
        # exchange 2: fixed-size int32 records, delivered in rank order
        send = torch.full((7 + trial,), rank * 1000 + trial, dtype=torch.int32)
        recv = gather_records(send, torch.zeros(world * send.numel(), dtype=torch.int32)).numpy()
        ok &= np.array_equal(recv, np.repeat(np.arange(world) * 1000 + trial, send.numel()))
    flag = torch.tensor([int(ok)])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        result.put(int(flag.item()))
    dist.destroy_process_group()


def test_candidate_exchange_gloo_world2():
    """Per-shard top-min(k, C/world) candidates, gathered in rank order and selected with
    the position tie-break, equal the global canonical top-k (ties included)."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    result = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, result)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert result.get(timeout=10) == 1


# ----------------------------------------------------------------------------- 2 GPUs / NCCL
def _nccl_worker(rank, world, port, result, name, steps, fused="off"):
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    _init(rank, world, port, "nccl")
    import bithtm_b200 as bithtm
    from helpers import golden_inputs, gpu_record, load_golden, step_digest
    from oracle.digest import canonical_from_rows, state_digest

    info = load_golden(name)
    g = info["g"]
    xs = golden_inputs(info, steps)
    np.random.seed(info["seed"])
    kw = dict(fused="shard", fused_ctas=32) if fused == "shard" else {}
    htm = bithtm.HierarchicalTemporalMemory(info["I"], info["C"], info["c"], info["k"], column_shard=True, **kw)
    state_at = {int(s): int(d) for s, d in zip(g["state_steps"], g["state_digests"])}
    bad = None
    for t in range(steps):
        sp_state, tm_state = htm.process(xs[t])
        d = step_digest(**gpu_record(htm, sp_state, tm_state))  # overlaps/boosted reads are collectives
        if d != int(g["digests"][t]) and bad is None:
            bad = f"rank {rank}: step {t} differs from the reference trace"
        if t in state_at:
            sp, tm = htm.spatial_pooler, htm.temporal_memory
            perm = [None] * world
            duty = [None] * world
            dist.all_gather_object(perm, sp.proximal_projection.permanence)
            dist.all_gather_object(duty, sp.boosting.duty_cycle)
            owner, count, cells, pm = tm.distal_projection.export_segments()
            sd = state_digest(np.concatenate(perm), np.concatenate(duty), tm.distal_projection.bundle_segments,
                              canonical_from_rows(owner, cells, pm))
            if sd != state_at[t] and bad is None:
                bad = f"rank {rank}: learned state differs at step {t}"
    flag = torch.tensor([0 if bad else 1], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if bad:
        print(bad, flush=True)
    if rank == 0:
        result.put(int(flag.item()))
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("name,steps,fused,world", [("mid", 1000, "off", 2), ("tiny", 400, "off", 2),
                                                    ("mid", 1000, "shard", 2), ("mid", 1000, "shard", 4),
                                                    ("mid", 1000, "off", 4), ("mid", 1000, "shard", 8)])
def test_column_sharded_two_gpus_match_reference_trace(name, steps, fused, world):
    """Column-sharded SP + segment-sharded TM on 2 / 4 / 8 GPUs -- per-stage kernels with NCCL
    all-gathers ("off"), or one kernel per shard exchanging over NVLink peer memory
    ("shard") -- every rank reproduces the single-network reference trace bit for bit."""
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    result = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, result, name, steps, fused)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    assert result.get(timeout=10) == 1


@pytest.mark.gpu
def test_column_shard_single_process_emulation():
    """The shard entry points on ONE GPU: two engines own one half of the columns
    each; their candidates are concatenated by hand (what the all-gather does)."""
    import torch

    import bithtm_b200 as bithtm
    from bithtm_b200 import _native as nat
    from helpers import golden_inputs, load_golden
    from oracle.htm_oracle import HTMOracle, OracleConfig

    info = load_golden("mid")
    I, C, c, k, seed = info["I"], info["C"], info["c"], info["k"], info["seed"]
    xs = golden_inputs(info, 120)
    shards = []
    for r in range(2):
        np.random.seed(seed)
        shards.append(bithtm.HierarchicalTemporalMemory(I, C, c, k, column_shard=(r, 2), rng_sync="lazy",
                                                        segment_shard=False))  # replicated temporal memory
    orc = HTMOracle(OracleConfig(I, C, c, k), rng=np.random.RandomState(seed))
    for t in range(120):
        rec = orc.step(xs[t])
        cand = []
        for h in shards:
            eng = h.engine
            h.spatial_pooler.boosting._bind(eng)
            words = eng.pack_input(xs[t])
            keys = torch.empty(eng.k_local, dtype=torch.float64, device="cuda")
            cols = torch.empty(eng.k_local, dtype=torch.int32, device="cuda")
            nat.check(nat.lib.bh_sp_shard_local(eng.ref, words.data_ptr(), keys.data_ptr(), cols.data_ptr(), eng.stream))
            cand.append((words, keys, cols))
        all_keys = torch.cat([x[1] for x in cand])
        all_cols = torch.cat([x[2] for x in cand])
        for h, (words, _, _) in zip(shards, cand):
            eng = h.engine
            nat.check(nat.lib.bh_sp_shard_finish(eng.ref, words.data_ptr(), all_keys.data_ptr(), all_cols.data_ptr(),
                                                 int(all_keys.numel()), 1, eng.stream))
            tm = h.temporal_memory
            tm._rng.before(eng)
            nat.check(nat.lib.bh_tm_step(eng.ref, 1, eng.stream))
            eng.epoch += 1
            summary = eng.summary()
            st = tm._finish(summary)
            assert np.array_equal(st._active_column, rec.active_column), f"step {t}"
            wc = st.winner_cell[0] * c + st.winner_cell[1]
            assert np.array_equal(wc, rec.winner_cell), f"step {t}"
            assert st.n_segments == rec.n_segments
        ov = np.concatenate([h.engine.buf["overlaps"].cpu().numpy() for h in shards])
        assert np.array_equal(ov, rec.overlaps)
    perm = np.concatenate([h.spatial_pooler.proximal_projection.permanence for h in shards])
    assert np.array_equal(perm.view(np.uint64), orc.permanence.view(np.uint64))


# ----------------------------------------------------------------------------- segment shards, one GPU
def _emulated_shards(info, world, **kw):
    import bithtm_b200 as bithtm

    shards = []
    for r in range(world):
        np.random.seed(info["seed"])
        shards.append(bithtm.HierarchicalTemporalMemory(info["I"], info["C"], info["c"], info["k"],
                                                        column_shard=(r, world), rng_sync="lazy", **kw))
    return shards


def _emulated_step(shards, x, learning=True, return_winner_cell=True):
    """One timestep of `world` shards living in ONE process: the two all-gathers are
    torch.cat of the per-shard buffers (exactly what NCCL delivers, rank order)."""
    import torch

    from bithtm_b200 import _native as nat

    cand = []
    for h in shards:
        eng = h.engine
        h.spatial_pooler.boosting._bind(eng)
        words = eng.pack_input(x)
        keys = torch.empty(eng.k_local, dtype=torch.float64, device="cuda")
        cols = torch.empty(eng.k_local, dtype=torch.int32, device="cuda")
        nat.check(nat.lib.bh_sp_shard_local(eng.ref, words.data_ptr(), keys.data_ptr(), cols.data_ptr(), eng.stream))
        cand.append((words, keys, cols))
    all_keys = torch.cat([c[1] for c in cand])
    all_cols = torch.cat([c[2] for c in cand])
    for h, (words, _, _) in zip(shards, cand):
        eng = h.engine
        nat.check(nat.lib.bh_sp_shard_finish(eng.ref, words.data_ptr(), all_keys.data_ptr(), all_cols.data_ptr(),
                                             int(all_keys.numel()), int(learning), eng.stream))
        h.temporal_memory._rng.before(eng)
    flags = int(bool(learning)) | (0 if return_winner_cell else 2)
    records = torch.cat([h.engine.tm_shard_pre(flags) for h in shards])  # exchange 2
    states = []
    for h in shards:
        eng = h.engine
        eng.tm_shard_post(records, want_jitter=return_winner_cell)
        states.append(h.temporal_memory._finish(eng.summary(), have_winner=bool(learning or return_winner_cell),
                                                have_jitter=bool(return_winner_cell)))
    return states


@pytest.mark.gpu
@pytest.mark.parametrize("name,world,steps", [("tiny", 2, 600), ("tiny", 4, 400), ("odd", 3, 750), ("mid", 2, 1100),
                                              ("mid", 4, 950)])
def test_segment_shards_single_process_emulation(name, world, steps):
    """Column-sharded SP + segment-sharded TM, every shard in one process: each shard
    reproduces the oracle's step records (winner/active cells, matching segments with
    their jitter draws, segment counts) and together they hold the oracle's learned
    state, including segments recycled across shards."""
    import bithtm_b200 as bithtm  # noqa: F401
    from bithtm_b200 import _native as nat
    from helpers import golden_inputs, load_golden
    from oracle.digest import canonical_from_rows, state_digest
    from oracle.htm_oracle import HTMOracle, OracleConfig

    info = load_golden(name)
    I, C, c, k, seed = info["I"], info["C"], info["c"], info["k"], info["seed"]
    xs = golden_inputs(info, steps)
    shards = _emulated_shards(info, world)
    assert all(h.engine.seg_world == world for h in shards)
    orc = HTMOracle(OracleConfig(I, C, c, k), rng=np.random.RandomState(seed))
    recycled = 0
    for t in range(steps):
        rec = orc.step(xs[t])
        states = _emulated_step(shards, xs[t])
        for r, (h, st) in enumerate(zip(shards, states)):
            where = f"step {t} shard {r}"
            ds = st.distal_state
            assert np.array_equal(st._active_column, rec.active_column), where
            assert np.array_equal(st.winner_cell[0] * c + st.winner_cell[1], rec.winner_cell), where
            assert np.array_equal(st.active_cell[0] * c + st.active_cell[1], rec.active_cell), where
            assert st.n_segments == rec.n_segments, where
            assert np.array_equal(ds.matching_segment, rec.matching_segment), where
            assert np.array_equal(ds.matching_segment_activation, rec.matching_activation), where
            assert np.array_equal(ds.matching_segment_jittered_potential.view(np.uint32),
                                  np.asarray(rec.matching_jit, dtype=np.float32).view(np.uint32)), where
        recycled += int(shards[0].engine.scalars()[nat.SC_NR])
    # learned state: SP rows by column shard, synapse rows by segment shard
    perm = np.concatenate([h.spatial_pooler.proximal_projection.permanence for h in shards])
    duty = np.concatenate([h.spatial_pooler.boosting.duty_cycle for h in shards])
    parts = [h.temporal_memory.distal_projection.export_local_segments() for h in shards]
    for h in shards:
        proj = h.temporal_memory.distal_projection
        owner, count, cells, pm = proj.export_segments(parts=parts)
        got = state_digest(perm, duty, proj.bundle_segments, canonical_from_rows(owner, cells, pm))
        want = state_digest(orc.permanence, orc.duty, orc.cell_nseg, orc.canonical_synapses())
        assert got == want
    assert sum(len(p[0]) for p in parts) == orc.n_seg
    if True:
        assert recycled > 0, "the run never recycled a segment: the cross-shard path was not exercised"


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_segment_shards_mixed_flags_match_reference_trace(world):
    """Per-step (learning, return_winner_cell) flags on segment shards (staged entry points,
    bh_tm_shard_pre flags + bh_tm_shard_post_ex): inference-only steps, deferred jitter draws,
    against the trace recorded from the unmodified reference (tests/golden/mixed.npz)."""
    from helpers import golden_inputs, load_golden
    from oracle.htm_oracle import HTMOracle, OracleConfig

    info = load_golden("mixed")
    g = info["g"]
    I, C, c, k, seed, steps = info["I"], info["C"], info["c"], info["k"], info["seed"], info["steps"]
    xs = golden_inputs(info, steps)
    shards = _emulated_shards(info, world)
    orc = HTMOracle(OracleConfig(I, C, c, k), rng=np.random.RandomState(seed))
    for t in range(steps):
        lf, wf = bool(g["learning_flags"][t]), bool(g["winner_flags"][t])
        rec = orc.step(xs[t], learning=lf, return_winner_cell=wf)
        states = _emulated_step(shards, xs[t], learning=lf, return_winner_cell=wf)
        for r, st in enumerate(states):
            where = f"step {t} shard {r} (learning={lf}, return_winner_cell={wf})"
            ds = st.distal_state
            wc = st.winner_cell
            assert (wc is None) == (not (lf or wf)), where
            assert np.array_equal(st._active_column, rec.active_column), where
            if wc is not None:
                assert np.array_equal(wc[0] * c + wc[1], rec.winner_cell), where
            assert np.array_equal(st.active_cell[0] * c + st.active_cell[1], rec.active_cell), where
            assert st.n_segments == rec.n_segments, where
            assert np.array_equal(ds.matching_segment, rec.matching_segment), where
            assert np.array_equal(ds.matching_segment_activation, rec.matching_activation), where
            assert (ds.matching_segment_jittered_potential is None) == (not wf), where
            if wf:
                assert np.array_equal(ds.matching_segment_jittered_potential.view(np.uint32),
                                      np.asarray(rec.matching_jit, dtype=np.float32).view(np.uint32)), where


# ----------------------------------------------------------------------------- fused sharded step
@pytest.mark.gpu
def test_fused_shard_kernel_single_shard_matches_oracle():
    """fused="shard" with ONE shard: the sharded phase sequence (candidate record, segment
    record, merge) and the in-kernel exchange protocol run against the rank's own region;
    results must equal the oracle's like every other execution mode."""
    import bithtm_b200 as bithtm
    from helpers import diff_records, golden_inputs, gpu_record, gpu_state_digest, load_golden, oracle_record, \
        oracle_state_digest
    from oracle.htm_oracle import HTMOracle, OracleConfig

    info = load_golden("mid")
    I, C, c, k, seed = info["I"], info["C"], info["c"], info["k"], info["seed"]
    steps = 1000
    xs = golden_inputs(info, steps)
    np.random.seed(seed)
    htm = bithtm.HierarchicalTemporalMemory(I, C, c, k, fused="shard", fused_ctas=12)
    assert htm.engine.ctx.fused_mode == 3
    orc = HTMOracle(OracleConfig(I, C, c, k), rng=np.random.RandomState(seed))
    for t in range(steps):
        rec = orc.step(xs[t])
        sp_state, tm_state = htm.process(xs[t])
        d = diff_records(gpu_record(htm, sp_state, tm_state), oracle_record(rec))
        assert not d, f"step {t}: {d}"
    assert gpu_state_digest(htm) == oracle_state_digest(orc)


@pytest.mark.gpu
@pytest.mark.parametrize("name,world,steps,cells", [("tiny", 2, 600, True), ("mid", 2, 1100, True), ("odd", 3, 750, True),
                                                    ("mid", 4, 950, True), ("mid", 2, 1100, False), ("odd", 3, 750, False)])
def test_fused_shard_kernels_concurrent_on_one_gpu(name, world, steps, cells):
    """`world` shards as `world` cooperative kernels running CONCURRENTLY on one GPU (one
    stream each), exchanging their records through each other's regions exactly as they do
    across GPUs over NVLink.  Every shard must reproduce the oracle."""
    import torch

    import bithtm_b200 as bithtm
    from bithtm_b200 import _native as nat
    from helpers import golden_inputs, load_golden
    from oracle.digest import canonical_from_rows, state_digest
    from oracle.htm_oracle import HTMOracle, OracleConfig

    info = load_golden(name)
    I, C, c, k, seed = info["I"], info["C"], info["c"], info["k"], info["seed"]
    xs = golden_inputs(info, steps)
    # cells: the {word, sequence} cell exchange (one-CTA selection / merge); else the copy + fence + flag protocol
    # with grid-wide selection and merge (what very wide networks such as cfg5 use)
    shards = _emulated_shards(info, world, fused="shard", fused_ctas=8, exchange_cells=cells)
    assert all(h.engine.ctx.xch_ll == int(cells) for h in shards)
    regions = [torch.zeros(h.engine.exchange_region_ints(), dtype=torch.int32, device="cuda") for h in shards]
    for h in shards:
        h.engine.set_exchange_regions([r.data_ptr() for r in regions], keepalive=regions)
        h.temporal_memory._rng.before(h.engine)
    streams = [torch.cuda.Stream() for _ in shards]
    orc = HTMOracle(OracleConfig(I, C, c, k), rng=np.random.RandomState(seed))
    torch.cuda.synchronize()
    for t in range(steps):
        rec = orc.step(xs[t])
        words = shards[0].engine.pack_input(xs[t])
        torch.cuda.synchronize()
        for h, st in zip(shards, streams):
            with torch.cuda.stream(st):
                h.process(words, return_state=False)
        torch.cuda.synchronize()
        for r, h in enumerate(shards):
            eng = h.engine
            st = h.temporal_memory._finish(eng.summary())
            where = f"step {t} shard {r}"
            ds = st.distal_state
            assert np.array_equal(st._active_column, rec.active_column), where
            assert np.array_equal(st.winner_cell[0] * c + st.winner_cell[1], rec.winner_cell), where
            assert st.n_segments == rec.n_segments, where
            assert np.array_equal(ds.matching_segment, rec.matching_segment), where
            assert np.array_equal(ds.matching_segment_jittered_potential.view(np.uint32),
                                  np.asarray(rec.matching_jit, dtype=np.float32).view(np.uint32)), where
    perm = np.concatenate([h.spatial_pooler.proximal_projection.permanence for h in shards])
    duty = np.concatenate([h.spatial_pooler.boosting.duty_cycle for h in shards])
    parts = [h.temporal_memory.distal_projection.export_local_segments() for h in shards]
    proj = shards[-1].temporal_memory.distal_projection
    owner, count, cells, pm = proj.export_segments(parts=parts)
    got = state_digest(perm, duty, proj.bundle_segments, canonical_from_rows(owner, cells, pm))
    assert got == state_digest(orc.permanence, orc.duty, orc.cell_nseg, orc.canonical_synapses())


@pytest.mark.gpu
@pytest.mark.parametrize("world,k,ctas,cells,per_launch,pipe", [(2, 655, 60, True, 1, 0), (4, 1100, 36, True, 1, 0),
                                                                 (2, 655, 60, False, 1, 0), (2, 655, 60, True, 8, 24),
                                                                 (4, 1100, 36, True, 5, 14), (2, 655, 60, True, 20, 0)])
def test_sharded_equals_unsharded_at_scale(world, k, ctas, cells, per_launch, pipe):
    """No oracle at this size (validate sharded == unsharded): 32768
    columns x 4096 inputs, many-CTA random-stream production and the grid-wide top-k active.
    One network as a single cooperative kernel vs the same network as two shard kernels
    running concurrently and exchanging through each other's regions.  per_launch > 1: the shards run that
    many steps per kernel launch from their device input rings -- with pipe > 0 as the two-pipeline shard kernel
    (csrc/shard_fused.cuh, k_step_shard_pipe: the selection exchange of step s+1 beside the temporal memory of
    step s, pipe CTAs for the latter)."""
    import torch

    import bithtm_b200 as bithtm
    from bithtm_b200 import _native as nat
    from bithtm_b200.projections import DenseProjection
    from oracle.digest import canonical_from_rows, state_digest

    I, C, c, steps = 4096, 32768, 32, 160 if world == 2 else 100  # world 4: the gathered-candidate merge runs grid-wide too
    g = np.random.default_rng(3)
    base = g.random((20, I)) < 0.2
    xs = base[np.arange(steps) % 20] ^ (g.random((steps, I)) < 0.05)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(99)
    perm = torch.randn(C, I, dtype=torch.float64, device="cuda", generator=gen) * 0.1
    kw = dict(rng_sync="lazy", max_segments=1 << 18, max_synapses_per_segment=128)

    def build(**extra):
        np.random.seed(4)
        rows = perm
        if "column_shard" in extra:
            r, w = extra["column_shard"]
            rows = perm[r * C // w:(r + 1) * C // w]
        sp = bithtm.SpatialPooler(I, C, k, proximal_projection=DenseProjection(I, C, permanence=rows))
        return bithtm.HierarchicalTemporalMemory(I, C, c, k, spatial_pooler=sp, **kw, **extra)

    whole = build(fused="grid")
    assert whole.engine.ctx.jump_polys > 0
    ring = dict(ring_len=steps) if per_launch > 1 else {}
    shards = [build(column_shard=(r, world), fused="shard", fused_ctas=ctas, exchange_cells=cells, pipeline=pipe, **ring)
              for r in range(world)]
    assert all(h.engine.ctx.pipe_ctas == pipe for h in shards)
    regions = [torch.zeros(h.engine.exchange_region_ints(), dtype=torch.int32, device="cuda") for h in shards]
    if per_launch > 1:
        for h in shards:
            h.engine.load_ring(xs)
    for h in shards + [whole]:
        h.temporal_memory._rng.before(h.engine)
    for h in shards:
        h.engine.set_exchange_regions([r.data_ptr() for r in regions], keepalive=regions)
    streams = [torch.cuda.Stream() for _ in shards]
    for t in range(steps):
        words = whole.engine.pack_input(xs[t])
        whole.process(words, return_state=False)
        torch.cuda.synchronize()
        if per_launch > 1:
            if t % per_launch != per_launch - 1:
                continue
            for h, st in zip(shards, streams):  # the last per_launch steps as ONE launch per shard
                with torch.cuda.stream(st):
                    h.engine.launch_graph(h.engine.graph(per_launch, learning=True), per_launch)
        else:
            for h, st in zip(shards, streams):
                with torch.cuda.stream(st):
                    h.process(words, return_state=False)
        torch.cuda.synchronize()
        if t % 20 == 19 or t == steps - 1:
            ref = whole.engine.summary().copy()
            for r, h in enumerate(shards):
                got = h.engine.summary().copy()
                n = 4 + 4 * k  # step, status, segments, winners, active columns, row words (not the RNG key form)
                # the status word is per rank (notes raised by the rank that stores a row reach the others
                # with the next exchange): compared at the end
                got[1] = 0
                want = ref[:n].copy()
                want[1] = 0
                bad = np.flatnonzero(got[:n] != want)
                assert bad.size == 0, f"step {t} shard {r}: summary words {bad[:8]} differ: {got[bad[:8]]} vs {want[bad[:8]]}"
    for h in shards + [whole]:
        h.engine.check_status()
    proj = whole.temporal_memory.distal_projection
    owner, count, cells, pm = proj.export_segments()
    want = state_digest(whole.spatial_pooler.proximal_projection.permanence, whole.spatial_pooler.boosting.duty_cycle,
                        proj.bundle_segments, canonical_from_rows(owner, cells, pm))
    parts = [h.temporal_memory.distal_projection.export_local_segments() for h in shards]
    sproj = shards[0].temporal_memory.distal_projection
    owner, count, cells, pm = sproj.export_segments(parts=parts)
    got = state_digest(np.concatenate([h.spatial_pooler.proximal_projection.permanence for h in shards]),
                       np.concatenate([h.spatial_pooler.boosting.duty_cycle for h in shards]),
                       sproj.bundle_segments, canonical_from_rows(owner, cells, pm))
    assert got == want
    assert int(whole.engine.scalars()[2]) > 5000  # segments were learned
