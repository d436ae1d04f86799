"""CPU suite (`-m "not gpu"`): the oracle against the golden traces generated from
the unmodified reference, the np.exp restatement, and the C-ABI surface."""

import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import golden_inputs, load_golden, oracle_state_digest
from oracle.digest import record_digest
from oracle.htm_oracle import HTMOracle, OracleConfig, canonical_topk

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name,steps", [("tiny", None), ("odd", None), ("mid", 1500), ("cfg1", 400), ("cfg2", 400),
                                        ("edge", None), ("edge24", None), ("c1", None), ("k1", None)])
def test_oracle_matches_reference_golden(name, steps):
    """Per-step digests and learned-state digests recorded from cokwa/bitHTM
    (tests/golden/make_golden.py) are reproduced by the oracle.  "edge" / "edge24": empty, full and
    repeated inputs (whole top-k tied at overlap 0); "c1": one cell per column; "k1": one active column."""
    info = load_golden(name)
    g = info["g"]
    steps = info["steps"] if steps is None else steps
    xs = golden_inputs(info, steps)
    orc = HTMOracle(OracleConfig(info["I"], info["C"], info["c"], info["k"]),
                    rng=np.random.RandomState(info["seed"]), overlap="packed")
    state_at = {int(s): int(d) for s, d in zip(g["state_steps"], g["state_digests"])}
    for t in range(steps):
        rec = orc.step(xs[t])
        assert record_digest(rec) == int(g["digests"][t]), f"{name}: step {t}"
        assert rec.draws == int(g["draws"][t])
        if t in state_at:
            assert oracle_state_digest(orc) == state_at[t], f"{name}: learned state at step {t}"


def test_oracle_matches_reference_golden_mixed_modes():
    """A trace in which every step has its own (learning, return_winner_cell) flags
    (TemporalMemory.process, networks.py:91-128): inference-only steps draw nothing, the
    jitter draw is deferred until a later step needs it, growth is skipped after a step
    without winner cells."""
    info = load_golden("mixed")
    g = info["g"]
    xs = golden_inputs(info, info["steps"])
    orc = HTMOracle(OracleConfig(info["I"], info["C"], info["c"], info["k"]),
                    rng=np.random.RandomState(info["seed"]), overlap="packed")
    state_at = {int(s): int(d) for s, d in zip(g["state_steps"], g["state_digests"])}
    flags = list(zip(g["learning_flags"], g["winner_flags"]))
    assert len({(bool(a), bool(b)) for a, b in flags}) == 4
    for t in range(info["steps"]):
        lf, wf = bool(flags[t][0]), bool(flags[t][1])
        rec = orc.step(xs[t], learning=lf, return_winner_cell=wf)
        assert record_digest(rec) == int(g["digests"][t]), f"step {t}"
        assert rec.draws == int(g["draws"][t])
        if t in state_at:
            assert oracle_state_digest(orc) == state_at[t], f"learned state at step {t}"


def test_oracle_dense_equals_packed_overlap():
    cfg = OracleConfig(100, 200, 8, 12)
    a = HTMOracle(cfg, rng=np.random.RandomState(5), overlap="dense")
    b = HTMOracle(cfg, rng=np.random.RandomState(5), overlap="packed")
    g = np.random.default_rng(0)
    for _ in range(50):
        x = g.random(100) < 0.3
        assert record_digest(a.step(x)) == record_digest(b.step(x))


def test_canonical_topk_rule():
    keys = np.array([1.0, 3.0, 3.0, 0.5, 3.0, 2.0])
    assert canonical_topk(keys, 2).tolist() == [1, 2]
    assert canonical_topk(keys, 4).tolist() == [1, 2, 4, 5]
    assert canonical_topk(np.zeros(5), 3).tolist() == [0, 1, 2]


def test_np_expf_restatement_matches_numpy():
    """bithtm_b200/csrc/np_expf.h (the sequence the boost kernel runs) == np.exp on float32."""
    so = os.path.join(ROOT, "oracle", "_build", "libnpexp.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(so)
    lib.bh_np_expf_array.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    g = np.random.default_rng(3)
    x = np.concatenate([
        -g.random(2_000_000, dtype=np.float32) * np.float32(15.5),
        np.float32(-14.985366) * np.linspace(0, 1, 100_001, dtype=np.float32),
        np.array([0.0, -0.0, -1e-30, -15.5, -80.0], dtype=np.float32),
    ]).astype(np.float32)
    y = np.empty_like(x)
    lib.bh_np_expf_array(x.ctypes.data, y.ctypes.data, x.size)
    assert np.array_equal(np.exp(x).view(np.uint32), y.view(np.uint32))


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads without a GPU and exports every function the
    header declares; the ctypes struct mirror has the compiled size."""
    from bithtm_b200 import _native as nat

    header = open(os.path.join(ROOT, "include", "bithtm_b200.h")).read()
    declared = set(re.findall(r"^(?:int|size_t)\s+(bh_[a-z0-9_]+)\s*\(", header, flags=re.M))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(nat.lib, name), f"{name} not exported"
    assert declared == set(nat.EXPORTED), declared ^ set(nat.EXPORTED)
    assert nat.lib.bh_ctx_size() == ctypes.sizeof(nat.BhCtx)
    assert nat.lib.bh_abi_version() == nat.ABI_VERSION


def test_layout_without_device():
    """bh_layout(base=NULL) sizes the arena from the context alone (no compute)."""
    from bithtm_b200 import _native as nat

    ctx = nat.BhCtx()
    ctx.input_dim, ctx.input_words, ctx.mask_stride = 1024, 32, 32
    ctx.column_dim, ctx.cell_dim, ctx.active_columns = 2048, 32, 41
    ctx.col_lo, ctx.col_local = 0, 2048
    ctx.seg_capacity, ctx.syn_capacity, ctx.match_capacity, ctx.learn_capacity = 1 << 16, 128, 1 << 16, 1 << 17
    ctx.tm_blocks, ctx.rng_ring_words, ctx.rng_step_words = 148, 1 << 20, 1 << 18
    n = nat.lib.bh_layout(ctypes.byref(ctx), None)
    perm_bytes = 2048 * 1024 * 8
    syn_bytes = (1 << 16) * 128 * 8
    assert perm_bytes + syn_bytes < n < perm_bytes + syn_bytes + (64 << 20)
    assert n % 256 == 0


def test_product_never_imports_oracle():
    """The product path must not route through the oracle."""
    pkg = os.path.join(ROOT, "bithtm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_no_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import bithtm_b200
    from bithtm_b200 import _native as nat

    with pytest.raises(nat.NativeError):
        bithtm_b200.HierarchicalTemporalMemory(64, 128, 8, 10)


def test_mt19937_jump_polynomials():
    """The characteristic polynomial and the jump polynomials the device generator uses
    (bithtm_b200/_mtjump.py) reproduce a simulated MT19937 stream."""
    from bithtm_b200 import _mtjump

    assert len(_mtjump.PHI) == 135 and _mtjump.PHI[-1] == 19937
    assert _mtjump.self_check(seed=11, polys=3)
    key = np.random.RandomState(4).get_state()[1]
    x = _mtjump.raw_stream(key.astype(np.uint32), 3 * 624)
    bg = np.random.MT19937()
    bg.state = {"bit_generator": "MT19937", "state": {"key": key, "pos": 624}}
    raw = bg.random_raw(1248).astype(np.uint32)

    def temper(y):
        y = y ^ (y >> np.uint32(11))
        y = y ^ ((y << np.uint32(7)) & np.uint32(0x9D2C5680))
        y = y ^ ((y << np.uint32(15)) & np.uint32(0xEFC60000))
        return y ^ (y >> np.uint32(18))

    assert np.array_equal(temper(x[624:624 + 1248]), raw)


def test_segment_shard_id_mapping_and_merge():
    """Host mirror of the segment-shard addressing (ids dealt to ranks in blocks of 64; a rank's
    rows compact and in id order) and the merge of per-rank exports back into id order."""
    from bithtm_b200._shard import held_segment_ids, local_row, merge_segment_parts

    for world in (1, 2, 3, 4, 8):
        for S in (0, 1, 63, 64, 65, 500, 64 * world * 3 + 17):
            parts = [held_segment_ids(S, r, world) for r in range(world)]
            allids = np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)
            assert np.array_equal(np.sort(allids), np.arange(S))  # a partition of the ids
            for r, ids in enumerate(parts):
                assert np.all(np.diff(ids) > 0)
                rows = np.array([local_row(int(s), world) for s in ids], dtype=np.int64)
                assert np.array_equal(rows, np.arange(len(ids)))  # compact, in id order
    g = np.random.default_rng(0)
    S, E, world = 300, 8, 3
    count = g.integers(0, E, size=S)
    cells = g.integers(0, 1000, size=(S, E))
    perm = g.random((S, E)).astype(np.float32)
    parts = []
    for r in range(world):
        ids = held_segment_ids(S, r, world)
        parts.append((ids, count[ids], cells[ids], perm[ids]))
    c2, cells2, perm2 = merge_segment_parts(parts[::-1])  # any order of the parts
    assert np.array_equal(c2, count) and np.array_equal(cells2, cells) and np.array_equal(perm2, perm)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's CPU path timed by the driver next to ours)
    runs without a GPU and prints ONE JSON line with the contract's keys."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "30",
                          "--warmup", "3", "--workload", "cfg2"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "steps/s" and d["higher_is_better"] is True
    assert d["steps"] == 30 and d["value"] > 0 and abs(d["value"] * d["ms_per_step"] - 1e3) < 1e-6 * 1e3
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
