"""ctypes binding of ``libbithtm_b200.so`` (``include/bithtm_b200.h``).

There is no CPU fallback: if the shared library has not been built
(``python -c 'import __graft_entry__ as g; g.build()'``) importing this module
raises, and every compute entry point fails without a CUDA device.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BITHTM_B200_LIB") or os.path.join(_HERE, "_lib", "libbithtm_b200.so")  # env: A/B builds

MT_N = 624
ABI_VERSION = 11
R_COUNT = 32  # int64 slots of ctx.rng64 (csrc/mt19937.cuh)

# device scalar block indices (enum in the header)
(SC_STEP, SC_HAVE_PREV, SC_NSEG, SC_NSEG_NEXT, SC_M, SC_W0, SC_W1, SC_L0, SC_L, SC_P, SC_NU, SC_NR,
 SC_STATUS, SC_MT_POS, SC_X_MATCH, SC_X_RECYC_AVAIL, SC_X_RECYC_TOTAL, SC_INPUT_POS, SC_BAR_COUNT, SC_BAR_GEN,
 SC_JIT_PENDING, SC_WNONE0, SC_WNONE1, SC_NGROW, SC_BAR2_COUNT, SC_BAR2_GEN, SC_NPREDCOL, SC_NPREDCOL_PREV, SC_T5_ERR,
 SC_BAR3_COUNT, SC_BAR3_GEN, SC_PIPE_SPLIT) = range(32)
SC_COUNT = 32

ST_SEG_OVERFLOW, ST_SYN_OVERFLOW, ST_MATCH_OVERFLOW, ST_LEARN_OVERFLOW, ST_RAND_OVERFLOW, ST_PRI_TIE = 1, 2, 4, 8, 16, 32
ST_XCH_OVERFLOW = 64
ST_XCH_TIMEOUT = 128
MAX_RANKS = 8
ST_NAMES = {
    ST_SEG_OVERFLOW: "segment capacity exceeded (max_segments)",
    ST_SYN_OVERFLOW: "synapse slots per segment exceeded (max_synapses_per_segment)",
    ST_MATCH_OVERFLOW: "matching-segment list capacity exceeded",
    ST_LEARN_OVERFLOW: "learning-segment list capacity exceeded",
    ST_RAND_OVERFLOW: "a step drew more random numbers than provisioned (rand_capacity)",
    ST_XCH_OVERFLOW: "segment shards: more matching / recyclable segments on a rank than the exchange carries "
                     "(exchange_match_capacity / exchange_recycle_capacity)",
    ST_XCH_TIMEOUT: "fused sharded step: a peer rank's record never arrived (peer not running the same step?)",
    ST_PRI_TIE: "equal growth priorities straddled a selection cut (reference-undefined tie)",
}
ST_FATAL = (ST_SEG_OVERFLOW | ST_SYN_OVERFLOW | ST_MATCH_OVERFLOW | ST_LEARN_OVERFLOW | ST_RAND_OVERFLOW
            | ST_XCH_OVERFLOW | ST_XCH_TIMEOUT)


def summary_ints(k: int) -> int:
    return 4 + 4 * k + MT_N + 1 + 4


_P = C.c_void_p


class BhCtx(C.Structure):
    """Mirror of ``struct bh_ctx``; field order and types must match the header
    (checked against ``bh_ctx_size()`` at load time)."""

    _fields_ = [
        ("input_dim", C.c_int32), ("input_words", C.c_int32), ("mask_stride", C.c_int32),
        ("column_dim", C.c_int32), ("cell_dim", C.c_int32), ("active_columns", C.c_int32),
        ("seg_capacity", C.c_int32), ("syn_capacity", C.c_int32), ("match_capacity", C.c_int32),
        ("learn_capacity", C.c_int32), ("tm_blocks", C.c_int32), ("sm_count", C.c_int32),
        ("rng_ring_words", C.c_int64), ("rng_step_words", C.c_int64), ("col_lo", C.c_int32), ("col_local", C.c_int32),
        ("ring_len", C.c_int32), ("fused_mode", C.c_int32),
        ("seg_rank", C.c_int32), ("seg_world", C.c_int32), ("xm_cap", C.c_int32), ("xr_cap", C.c_int32),
        ("jump_polys", C.c_int32), ("rng_lookahead", C.c_int32), ("device", C.c_int32), ("skip_polys", C.c_int32),
        ("skip_gran", C.c_int32), ("job_cap", C.c_int32), ("lazy_policy", C.c_int32), ("tail_chunks", C.c_int32), ("xch_ll", C.c_int32), ("pipe_ctas", C.c_int32),
        ("skip_min", C.c_int64),
        ("sp_threshold", C.c_double), ("sp_delta_on", C.c_double), ("sp_delta_off", C.c_double),
        ("tm_learn_on", C.c_double), ("tm_learn_off", C.c_double),
        ("tm_punish_on", C.c_double), ("tm_punish_off", C.c_double),
        ("boost_coef", C.c_float), ("duty_momentum", C.c_float), ("duty_increment", C.c_float),
        ("tm_perm_initial", C.c_float), ("tm_perm_threshold", C.c_float), ("epsilon", C.c_float),
        ("tm_learn_can_delete", C.c_int32), ("tm_punish_can_delete", C.c_int32),
        ("seg_activation_threshold", C.c_int32), ("seg_matching_threshold", C.c_int32),
        ("seg_sampling_synapses", C.c_int32), ("fused_ctas", C.c_int32), ("fused_threads", C.c_int32),
        # device pointers
        ("sp_perm", _P), ("sp_mask", _P), ("duty", _P), ("overlaps", _P), ("boosted", _P),
        ("active_cols", _P), ("col_active", _P),
        ("col_pred", _P), ("col_act", _P), ("col_win", _P),
        ("cell_nseg", _P), ("cell_maxjit", _P), ("cell_npred", _P), ("cell_widx", _P),
        ("seg_owner", _P), ("seg_count", _P), ("seg_pot", _P), ("seg_conn", _P),
        ("syn_cell", _P), ("syn_perm", _P),
        ("row_pred", _P), ("row_act", _P), ("row_win", _P), ("row_unacc", _P),
        ("winners", _P), ("unacc", _P),
        ("m_seg", _P), ("m_conn", _P), ("m_jit", _P), ("m_flag", _P),
        ("learn_list", _P), ("punish_list", _P), ("recyc_list", _P),
        ("x_send", _P), ("xk_keys", _P), ("xk_cols", _P), ("blk", _P), ("topk_ws", _P),
        ("mt_key", _P), ("rng_ring", _P), ("mt_jump", _P), ("rng64", _P),
        ("mt_skip", _P), ("rng_jump", _P), ("grow_list", _P),
        ("xpeer", _P * 8),
        ("sc", _P), ("input_ring", _P), ("input_dev", _P), ("input_pinned", _P),
        ("summary_dev", _P), ("summary_pinned", _P),
    ]


# (name, element dtype string) of every DEVICE buffer, for building torch views
DEVICE_BUFFERS = {
    "sp_perm": "float64", "sp_mask": "int32", "duty": "float32", "overlaps": "int32", "boosted": "float64",
    "active_cols": "int32", "col_active": "uint8", "col_pred": "int32", "col_act": "int32", "col_win": "int32",
    "cell_nseg": "int32", "cell_maxjit": "float32", "cell_npred": "int32", "cell_widx": "int32",
    "seg_owner": "int32", "seg_count": "int32", "seg_pot": "int32", "seg_conn": "int32",
    "syn_cell": "int32", "syn_perm": "float32",
    "row_pred": "int32", "row_act": "int32", "row_win": "int32", "row_unacc": "int32",
    "winners": "int32", "unacc": "int32", "m_seg": "int32", "m_conn": "int32", "m_jit": "float32",
    "m_flag": "uint8", "learn_list": "int32", "punish_list": "int32", "recyc_list": "int32",
    "x_send": "int32", "xk_keys": "float64", "xk_cols": "int32", "blk": "int32", "topk_ws": "int32",
    "mt_key": "int32", "rng_ring": "int32", "mt_jump": "int32", "rng64": "int64",
    "mt_skip": "int32", "rng_jump": "int32", "grow_list": "int32", "sc": "int32", "input_ring": "int32", "input_dev": "int32",
    "summary_dev": "int32",
}

_CTXP = C.POINTER(BhCtx)
_SIGNATURES = {
    "bh_layout": (C.c_size_t, [_CTXP, _P]),
    "bh_init": (C.c_int, [_CTXP, _P]),
    "bh_abi_version": (C.c_int, []),
    "bh_ctx_size": (C.c_size_t, []),
    "bh_device_info": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "bh_sp_build_mask": (C.c_int, [_CTXP, _P]),
    "bh_pack_input": (C.c_int, [_CTXP, _P, _P, _P]),
    "bh_pack_inputs": (C.c_int, [_CTXP, _P, C.c_int, C.c_int, _P, _P]),
    "bh_sp_overlap": (C.c_int, [_CTXP, _P, _P]),
    "bh_sp_overlap_batched": (C.c_int, [_CTXP, _P, C.c_int, _P, _P]),
    "bh_sp_overlap_batched_tc": (C.c_int, [_CTXP, _P, C.c_int, _P, _P]),
    "bh_sp_overlap_batched_tc5": (C.c_int, [_CTXP, _P, C.c_int, _P, _P]),
    "bh_sp_overlap_batched_mma": (C.c_int, [_CTXP, _P, C.c_int, _P, _P]),
    "bh_boost": (C.c_int, [_CTXP, _P]),
    "bh_inhibit": (C.c_int, [_CTXP, _P]),
    "bh_set_active_columns": (C.c_int, [_CTXP, _P, _P]),
    "bh_sp_learn": (C.c_int, [_CTXP, _P, _P]),
    "bh_duty_update": (C.c_int, [_CTXP, _P]),
    "bh_sp_step": (C.c_int, [_CTXP, _P, C.c_int, _P]),
    "bh_sp_shard_local": (C.c_int, [_CTXP, _P, _P, _P, _P]),
    "bh_sp_shard_finish": (C.c_int, [_CTXP, _P, _P, _P, C.c_int, C.c_int, _P]),
    "bh_tm_shard_xch_ints": (C.c_size_t, [_CTXP]),
    "bh_xch_region_ints": (C.c_size_t, [_CTXP]),
    "bh_tm_shard_pre": (C.c_int, [_CTXP, C.c_int, _P, _P]),
    "bh_tm_shard_post": (C.c_int, [_CTXP, _P, _P]),
    "bh_tm_shard_post_ex": (C.c_int, [_CTXP, _P, C.c_int, _P]),
    "bh_advance_step": (C.c_int, [_CTXP, _P]),
    "bh_tm_select": (C.c_int, [_CTXP, _P]),
    "bh_tm_learn": (C.c_int, [_CTXP, C.c_int, _P]),
    "bh_tm_activate": (C.c_int, [_CTXP, _P]),
    "bh_tm_step": (C.c_int, [_CTXP, C.c_int, _P]),
    "bh_tm_step_ex": (C.c_int, [_CTXP, C.c_int, C.c_int, C.c_int, _P]),
    "bh_tm_learn_args": (C.c_int, [_CTXP, _P, C.c_int, _P, C.c_int, _P, _P, _P]),
    "bh_tm_activate_cells": (C.c_int, [_CTXP, _P, C.c_int, C.c_int, C.c_int, _P]),
    "bh_tm_fill_jitter": (C.c_int, [_CTXP, _P]),
    "bh_tm_reset": (C.c_int, [_CTXP, _P]),
    "bh_step": (C.c_int, [_CTXP, _P, C.c_int, _P]),
    "bh_step_ring": (C.c_int, [_CTXP, C.c_int, _P]),
    "bh_step_host": (C.c_int, [_CTXP, _P, C.c_int, _P, _P]),
    "bh_host_graph_create": (C.c_int, [_CTXP, C.c_int, _P, C.POINTER(_P)]),
    "bh_step_host_graph": (C.c_int, [_CTXP, _P, _P, _P, _P]),
    "bh_summary": (C.c_int, [_CTXP, _P, _P]),
    "bh_graph_create": (C.c_int, [_CTXP, C.c_int, C.c_int, _P, C.POINTER(_P)]),
    "bh_batch_graph_create": (C.c_int, [C.POINTER(_CTXP), C.c_int, C.c_int, C.c_int, _P, C.POINTER(_P)]),
    "bh_graph_launch": (C.c_int, [_P, _P]),
    "bh_graph_destroy": (C.c_int, [_P]),
    "bh_profile_step": (C.c_int, [_CTXP, _P, C.c_int, _P, C.POINTER(C.c_float), C.POINTER(C.c_char_p), C.c_int]),
    "bh_step_launches": (C.c_int, [_CTXP, C.c_int]),
    "bh_rng_import": (C.c_int, [_CTXP, _P]),
    "bh_rng_export": (C.c_int, [_CTXP, _P]),
    "bh_rng_fill": (C.c_int, [_CTXP, _P, C.c_int64, _P]),
    "bh_test_np_expf": (C.c_int, [_P, _P, C.c_int64, _P]),
}
EXPORTED = tuple(_SIGNATURES)


class NativeError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"bithtm_b200: native library {LIB_PATH} is missing. Build it with "
            "`python -c \"import __graft_entry__ as g; g.build()\"` (needs nvcc); there is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.bh_ctx_size() != C.sizeof(BhCtx):
        raise ImportError(f"bithtm_b200: struct bh_ctx mismatch (library {lib.bh_ctx_size()} bytes, "
                          f"binding {C.sizeof(BhCtx)} bytes); rebuild the library")
    return lib


lib = _load()


def check(rc: int, what: str = "call"):
    if rc == 0:
        return
    if rc <= -1000:
        raise NativeError(f"bithtm_b200: {what} failed with CUDA error {-rc - 1000}")
    names = {-1: "bad argument", -2: "no CUDA device", -3: "unsupported configuration"}
    raise NativeError(f"bithtm_b200: {what} failed: {names.get(rc, rc)}")
