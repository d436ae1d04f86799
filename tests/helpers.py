"""Shared helpers for the parity tests (inputs, golden fixtures, record building)."""

from __future__ import annotations

import os

import numpy as np

from oracle.digest import canonical_from_rows, state_digest, step_digest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = ("overlaps", "boosted", "active_column", "bursting", "winner_cell", "active_cell",
          "matching_segment", "matching_activation", "matching_jit")


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    I, C, c, k, steps, patterns, seed = (int(v) for v in g["config"])
    return dict(g=g, I=I, C=C, c=c, k=k, steps=steps, patterns=patterns, seed=seed,
                density=float(g["density"]), noise=float(g["noise"]))


def make_inputs(input_dim, patterns, density, noise, steps, seed):
    """Same recipe as tests/golden/make_golden.py (example.py:34,52 from a private generator)."""
    g = np.random.default_rng(1000 + seed)
    base = g.random((patterns, input_dim)) < density
    flips = g.random((steps, input_dim)) < noise
    idx = np.arange(steps) % patterns
    return base[idx] ^ flips


def degenerate_inputs(input_dim, steps, density=0.25, seed=5):
    """Same recipe as tests/golden/make_golden.py::degenerate_inputs: every 8 steps an empty input,
    a full input and a repeated input between ordinary ones."""
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


def golden_inputs(info, steps=None):
    if info["patterns"] == 0:  # the degenerate-input traces ("edge", "edge24")
        return degenerate_inputs(info["I"], info["steps"], info["density"])[:steps]
    return make_inputs(info["I"], info["patterns"], info["density"], info["noise"],
                       info["steps"], info["seed"])[:steps]


def gpu_record(htm, sp_state, tm_state):
    """The StepRecord fields, read through the reference-facing State API."""
    c = htm.cell_dim
    ds = tm_state.distal_state
    return dict(
        n_segments=tm_state.n_segments,
        overlaps=sp_state.overlaps,
        boosted=sp_state.boosted_overlaps,
        active_column=sp_state.active_column,
        bursting=tm_state.active_column_bursting,
        winner_cell=tm_state.winner_cell[0] * c + tm_state.winner_cell[1],
        active_cell=tm_state.active_cell[0] * c + tm_state.active_cell[1],
        matching_segment=ds.matching_segment,
        matching_activation=ds.matching_segment_activation,
        matching_jit=ds.matching_segment_jittered_potential,
    )


def oracle_record(rec):
    d = {f: getattr(rec, f) for f in FIELDS}
    d["n_segments"] = rec.n_segments
    return d


def diff_records(a, b):
    """List of human-readable differences between two record dicts."""
    out = []
    if int(a["n_segments"]) != int(b["n_segments"]):
        out.append(f"n_segments {a['n_segments']} vs {b['n_segments']}")
    for f in FIELDS:
        x, y = np.asarray(a[f]).reshape(-1), np.asarray(b[f]).reshape(-1)
        if x.shape != y.shape:
            out.append(f"{f}: length {x.shape[0]} vs {y.shape[0]}; head {x[:6]} vs {y[:6]}")
            continue
        if x.dtype.kind == "f":
            bad = np.flatnonzero(x.view(f"u{x.dtype.itemsize}") != y.astype(x.dtype).view(f"u{x.dtype.itemsize}"))
        else:
            bad = np.flatnonzero(x != y)
        if bad.size:
            i = int(bad[0])
            out.append(f"{f}: {bad.size} mismatches, first at {i}: {x[i]!r} vs {y[i]!r}")
    return out


def gpu_state_digest(htm):
    sp, tm = htm.spatial_pooler, htm.temporal_memory
    owner, count, cells, perm = tm.distal_projection.export_segments()
    canon = canonical_from_rows(owner, cells, perm)
    return state_digest(sp.proximal_projection.permanence, sp.boosting.duty_cycle,
                        tm.distal_projection.bundle_segments, canon)


def oracle_state_digest(orc):
    return state_digest(orc.permanence, orc.duty, orc.cell_nseg, orc.canonical_synapses())


__all__ = ["load_golden", "golden_inputs", "make_inputs", "gpu_record", "oracle_record", "diff_records",
           "gpu_state_digest", "oracle_state_digest", "step_digest", "FIELDS"]
