"""Layout-independent digests of SP+TM step results and of the learned state.

TEST INFRASTRUCTURE ONLY (see ``oracle/htm_oracle.py``).  The same functions are
fed from three sources -- the unmodified reference (``tests/golden/make_golden.py``),
the NumPy oracle and the CUDA path -- so that 10k-step traces can be compared
through 8-byte values committed under ``tests/golden/``.
"""

from __future__ import annotations

import hashlib

import numpy as np

_STEP_FIELDS = (
    ("overlaps", np.int64),
    ("boosted", np.float64),
    ("active_column", np.int64),
    ("bursting", np.uint8),
    ("winner_cell", np.int64),
    ("active_cell", np.int64),
    ("matching_segment", np.int64),
    ("matching_activation", np.int64),
    ("matching_jit", np.float32),
)


def _h(parts) -> int:
    h = hashlib.blake2b(digest_size=8)
    for p in parts:
        a = np.ascontiguousarray(p)
        h.update(np.int64(a.size).tobytes())
        h.update(a.tobytes())
    return int.from_bytes(h.digest(), "little")


def step_digest(n_segments: int, **fields) -> int:
    """Digest of one timestep.  Keyword names are those of ``StepRecord``."""
    parts = [np.int64(n_segments)]
    for name, dt in _STEP_FIELDS:
        parts.append(np.asarray(fields[name]).reshape(-1).astype(dt))
    return _h(parts)


def record_digest(rec) -> int:
    return step_digest(rec.n_segments, **{n: getattr(rec, n) for n, _ in _STEP_FIELDS})


def state_digest(permanence, duty, cell_nseg, canon) -> int:
    """Digest of the learned state: SP permanence (float64 bits), duty cycle
    (float32 bits), segments per cell, and the canonical synapse form
    ``{"owner", "offsets", "cells", "perm_bits"}``."""
    return _h([
        np.asarray(permanence, dtype=np.float64).reshape(-1),
        np.asarray(duty, dtype=np.float32),
        np.asarray(cell_nseg).astype(np.int64),
        np.asarray(canon["owner"]).astype(np.int64),
        np.asarray(canon["offsets"]).astype(np.int64),
        np.asarray(canon["cells"]).astype(np.int64),
        np.asarray(canon["perm_bits"]).astype(np.uint32),
    ])


def canonical_from_rows(owner, syn_cell, syn_perm, invalid=None):
    """Canonical synapse form from row storage.  ``invalid`` is the free-slot
    marker (``None`` -> negative cell ids are free)."""
    owner = np.asarray(owner)
    S = len(owner)
    syn_cell = np.asarray(syn_cell)[:S].astype(np.int64)
    bits = np.ascontiguousarray(np.asarray(syn_perm, dtype=np.float32)[:S]).view(np.uint32)
    valid = (syn_cell >= 0) if invalid is None else (syn_cell != invalid)
    counts = valid.sum(axis=1) if S else np.zeros(0, dtype=np.int64)
    offsets = np.zeros(S + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    rows = np.repeat(np.arange(S), counts)
    cells = syn_cell[valid]
    pb = bits[valid]
    order = np.lexsort((pb, cells, rows))
    return {
        "owner": owner.astype(np.int64),
        "offsets": offsets,
        "cells": cells[order],
        "perm_bits": pb[order],
    }
