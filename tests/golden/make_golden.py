#!/usr/bin/env python
"""Generate the golden fixtures under ``tests/golden/`` from the UNMODIFIED
reference (cokwa/bitHTM at ``/root/reference``), and check the NumPy oracle
against it lock-step while doing so.

Run in the build container only (the GPU box has no ``/root/reference``):

    python tests/golden/make_golden.py            # all cases
    python tests/golden/make_golden.py tiny mid   # selected cases

For every case the reference network is built with a deterministic inhibition
object injected through the reference's own ``inhibition=`` constructor slot
(``bithtm/networks.py:16,24``) -- everything else is the
reference's stock code path.  Inputs come from a private ``default_rng`` so the
legacy global ``np.random`` stream is consumed only by the network.

Outputs ``<case>.npz``: per-step 8-byte digests (``oracle/digest.py``), learned
state digests every ``state_every`` steps, and full per-step records for the
first ``full_steps`` steps (small, for debugging a first divergence).
"""

from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle.digest import canonical_from_rows, record_digest, state_digest, step_digest  # noqa: E402
from oracle.htm_oracle import CanonicalGlobalInhibition, HTMOracle, OracleConfig  # noqa: E402

CASES = {
    # name: (input_dim, column_dim, cell_dim, active_columns, steps, patterns, density, noise, seed, state_every, full_steps)
    "tiny": (64, 256, 8, 20, 1500, 12, 0.25, 0.05, 3, 100, 40),
    "odd": (200, 300, 20, 25, 2000, 25, 0.2, 0.05, 7, 250, 20),
    "mid": (256, 512, 32, 22, 3000, 40, 0.2, 0.05, 11, 250, 10),
    "cfg1": (1000, 2048, 32, None, 10000, 100, 0.2, 0.05, 0, 1000, 4),
    "cfg2": (1024, 2048, 32, None, 10000, 100, 0.2, 0.05, 0, 1000, 4),
    # every step draws its (learning, return_winner_cell) flags (networks.py:91) from `mode_schedule`
    "mixed": (64, 256, 8, 20, 1200, 12, 0.25, 0.05, 5, 100, 0),
    # edge cases.  "edge" / "edge24": inputs from `degenerate_inputs` (empty, full and repeated inputs between
    # ordinary ones; "edge" is the configuration of the GPU test test_degenerate_inputs_match_oracle, whose k = 12
    # stays below the matching threshold 15, "edge24" lets segments match and learn); "c1": one cell per column;
    # "k1": a single active column
    "edge": (96, 320, 16, 12, 400, 0, 0.25, 0.0, 21, 100, 0),
    "edge24": (96, 320, 16, 24, 800, 0, 0.25, 0.0, 22, 200, 0),
    "c1": (64, 256, 1, 30, 800, 12, 0.25, 0.05, 4, 200, 0),
    "k1": (64, 256, 8, 1, 600, 12, 0.25, 0.05, 6, 200, 0),
}
DEGENERATE = ("edge", "edge24")


def mode_schedule(steps, seed):
    """Per-step (learning, return_winner_cell): mostly the default (True, True), with runs of
    inference-only steps, learning without the jitter draw, and winners without learning."""
    g = np.random.default_rng(77 + seed)
    r = g.random(steps)
    learning = (r < 0.6) | ((r >= 0.75) & (r < 0.87))
    winner = (r < 0.6) | (r >= 0.87)
    learning[:30] = True  # let some segments form first
    winner[:30] = True
    return learning, winner


def make_inputs(input_dim, patterns, density, noise, steps, seed):
    """example.py:34,52 recipe, but from a private generator."""
    g = np.random.default_rng(1000 + seed)
    base = g.random((patterns, input_dim)) < density
    flips = g.random((steps, input_dim)) < noise
    idx = np.arange(steps) % patterns
    return base[idx] ^ flips


def degenerate_inputs(input_dim, steps, density=0.25, seed=5):
    """Every 8 steps: ..., empty input (all overlaps 0: the whole top-k is one tie), ..., full input, the
    input before it once more.  Same recipe as tests/test_gpu_parity.py::test_degenerate_inputs_match_oracle."""
    g = np.random.default_rng(seed)
    xs = []
    for t in range(steps):
        r = t % 8
        if r == 3:
            xs.append(np.zeros(input_dim, dtype=bool))
        elif r == 6:
            xs.append(np.ones(input_dim, dtype=bool))
        elif r == 7:
            xs.append(xs[-2].copy())
        else:
            xs.append(g.random(input_dim) < density)
    return np.array(xs)


def case_inputs(name, input_dim, patterns, density, noise, steps, seed):
    if name in DEGENERATE:
        return degenerate_inputs(input_dim, steps, density)
    return make_inputs(input_dim, patterns, density, noise, steps, seed)


def reference_record(htm, sp_state, tm_state):
    c = htm.cell_dim
    ds = tm_state.distal_state
    wc = tm_state.winner_cell  # None when neither learning nor return_winner_cell (networks.py:99,125)
    jit = ds.matching_segment_jittered_potential  # None until somebody needs it (projections.py:229-243)
    return dict(
        n_segments=len(htm.temporal_memory.distal_projection.segment_bundle),
        overlaps=sp_state.overlaps,
        boosted=sp_state.boosted_overlaps,
        active_column=sp_state.active_column,
        bursting=tm_state.active_column_bursting,
        winner_cell=wc[0] * c + wc[1] if wc is not None else np.zeros(0, dtype=np.int64),
        active_cell=tm_state.active_cell[0] * c + tm_state.active_cell[1],
        matching_segment=ds.matching_segment,
        matching_activation=ds.matching_segment_activation,
        matching_jit=jit if jit is not None else np.zeros(0, dtype=np.float32),
    )


def reference_state_digest(htm):
    sp, tm = htm.spatial_pooler, htm.temporal_memory
    dp = tm.distal_projection
    proj = dp.segment_projection
    owner = dp.segment_bundle[:].squeeze(1)
    edge = proj.output_edge[:]
    target = proj.get_output_edge_target(edge)
    canon = canonical_from_rows(owner, np.where(edge == proj.invalid_output_edge, -1, target),
                                proj.output_permanence[:])
    return state_digest(sp.proximal_projection.permanence, sp.boosting.duty_cycle,
                        dp.bundle_segments, canon), canon


def run_case(name):
    import bithtm  # the reference

    I, C, c, k, steps, patterns, density, noise, seed, state_every, full_steps = CASES[name]
    xs = case_inputs(name, I, patterns, density, noise, steps, seed)

    np.random.seed(seed)
    k_eff = k if k is not None else round(C * 0.02)
    sp = bithtm.SpatialPooler(I, C, k_eff, inhibition=CanonicalGlobalInhibition(k_eff))
    htm = bithtm.HierarchicalTemporalMemory(I, C, c, active_columns=k_eff, spatial_pooler=sp)
    init_perm = sp.proximal_projection.permanence.copy()

    rs = np.random.RandomState(seed)
    orc = HTMOracle(OracleConfig(I, C, c, k), rng=rs, overlap="packed")
    assert np.array_equal(orc.permanence, init_perm)

    learn_flags, winner_flags = (mode_schedule(steps, seed) if name == "mixed"
                                 else (np.ones(steps, dtype=bool), np.ones(steps, dtype=bool)))
    digests = np.zeros(steps, dtype=np.uint64)
    draws = np.zeros(steps, dtype=np.int64)
    state_steps, state_digests = [], []
    full = {}
    ties = []
    t0 = time.time()
    t_ref = 0.0
    for t in range(steps):
        ta = time.perf_counter()
        lf, wf = bool(learn_flags[t]), bool(winner_flags[t])
        if name == "mixed":  # HierarchicalTemporalMemory.process (networks.py:146-149) with the TM flag exposed
            sp_state = htm.spatial_pooler.process(xs[t], learning=lf)
            tm_state = htm.temporal_memory.process(sp_state, learning=lf, return_winner_cell=wf)
        else:
            sp_state, tm_state = htm.process(xs[t])
        t_ref += time.perf_counter() - ta
        ref = reference_record(htm, sp_state, tm_state)
        rec = orc.step(xs[t], learning=lf, return_winner_cell=wf)
        d_ref = step_digest(**ref)
        d_orc = record_digest(rec)
        if rec.undefined_tie:
            ties.append(t)
        if d_ref != d_orc:
            for f in ref:
                a, b = np.asarray(ref[f]).reshape(-1), np.asarray(getattr(rec, f)).reshape(-1)
                if a.shape != b.shape or not np.array_equal(a, b):
                    print(f"  step {t}: field {f} differs: ref {a[:8]} ... oracle {b[:8]} ...")
            raise SystemExit(f"{name}: oracle diverged from the reference at step {t} (ties so far {ties})")
        digests[t] = d_ref
        draws[t] = rec.draws
        if t < full_steps:
            for f, v in ref.items():
                full[f"s{t}_{f}"] = np.asarray(v).reshape(-1)
        if (t + 1) % state_every == 0 or t == steps - 1:
            sd_ref, _ = reference_state_digest(htm)
            sd_orc = state_digest(orc.permanence, orc.duty, orc.cell_nseg, orc.canonical_synapses())
            if sd_ref != sd_orc:
                raise SystemExit(f"{name}: learned state differs at step {t}")
            state_steps.append(t)
            state_digests.append(sd_ref)
        # the two private/global streams must stay in lock-step
    assert np.random.get_state()[2] == rs.get_state()[2] and np.array_equal(np.random.get_state()[1], rs.get_state()[1])
    wall = time.time() - t0
    out = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(
        out, config=np.array([I, C, c, k_eff, steps, patterns, seed], dtype=np.int64),
        density=np.float64(density), noise=np.float64(noise), learning_flags=learn_flags,
        winner_flags=winner_flags, digests=digests, draws=draws, state_steps=np.array(state_steps, dtype=np.int64),
        state_digests=np.array(state_digests, dtype=np.uint64), undefined_tie_steps=np.array(ties, dtype=np.int64),
        n_segments_final=np.int64(rec.n_segments), reference_steps_per_s=np.float64(steps / t_ref), **full)
    print(f"{name}: {steps} steps ok, S={rec.n_segments}, ties={ties}, ref {steps / t_ref:.1f} steps/s, "
          f"wall {wall:.1f}s, {os.path.getsize(out) / 1024:.1f} KiB")


if __name__ == "__main__":
    for case in (sys.argv[1:] or list(CASES)):
        run_case(case)
