// One network sharded over several GPUs, the whole timestep of a shard as ONE
// cooperative kernel: the phases of fused.cuh with the two exchanges of SURVEY.md 8e
// done INSIDE the kernel over NVLink peer memory -- each rank stores its record
// straight into every peer's receive buffer, publishes a sequence number with a
// system-scope release store and waits for the peers' numbers -- instead of returning to
// the host for an NCCL all-gather.  The same phase functions run behind
// bh_sp_shard_* / bh_tm_shard_* with NCCL (tm_shard.cuh, sp_kernels.cuh).
//
// Exchange region of a rank (int32 units; identical layout on every rank; ctx.xpeer[p]
// = base of rank p's region, mapped into this process):
//   flags [2][8]               kind 0 = top-k candidates, 1 = segment records; entry s =
//                              sequence number (step + 1) of the last record from rank s
//   cand  [2 parity][G][n1]    n1 = 3 * k_loc rounded up to 4: k_loc float64 keys, k_loc columns
//   segs  [2 parity][G][n4]    n4 = bh_tm_shard_xch_ints rounded up to 4
// Records are double-buffered by step parity: a rank can be at most one exchange ahead
// of a peer, because finishing an exchange needs every peer's record of that exchange.
// After it (xch_ll) the cell areas of shard_ll.cuh.  k_step_shard: one pipeline; k_step_shard_pipe (end of this
// file): the selection exchange of step s+1 beside the temporal memory of step s, for launches of several steps.
#pragma once

#include "fused.cuh"
#include "tm_shard.cuh"
#include "shard_ll.cuh"

#define XCH_FLAG_INTS 16
// from this many keys the grid-wide selection (two grid barriers) beats one CTA -- much earlier than in the
// unsharded kernels, because a shard selects min(k, C/G) of only C/G keys and the merge k of G*min(k, C/G):
// large fractions, for which the single-CTA radix select degenerates
#define XCH_TOPK_GRID_MIN 4096
#define XCH_TIMEOUT_CYCLES 6000000000LL  // ~3 s at 2 GHz: a peer that never answers sets BH_ST_XCH_TIMEOUT

__host__ __device__ __forceinline__ int xch_k_loc(const bh_ctx& c) {
  return c.active_columns < c.col_local ? c.active_columns : c.col_local;
}
__host__ __device__ __forceinline__ long long xch_n1(const bh_ctx& c) { return (3LL * xch_k_loc(c) + 3) & ~3LL; }
__host__ __device__ __forceinline__ long long xch_n4(const bh_ctx& c) { return (xch_ints(c) + 3) & ~3LL; }
// the copy + flag areas, then (xch_ll) the cell areas of shard_ll.cuh
__host__ __device__ __forceinline__ long long xch_legacy_ints(const bh_ctx& c) {
  const long long G = c.seg_world > 1 ? c.seg_world : 1;
  return (XCH_FLAG_INTS + 2LL * G * (xch_n1(c) + xch_n4(c)) + 3) & ~3LL;
}
__host__ __device__ __forceinline__ long long xch_region_ints(const bh_ctx& c) {
  return xch_legacy_ints(c) + (c.xch_ll ? ll_area_ints(c) : 0);
}
// ctx.x_send: the packed record of an exchange, then (xch_ll) the append lists of the segment scan
__host__ __device__ __forceinline__ long long xch_send_ints(const bh_ctx& c) {
  const long long n1 = xch_n1(c), n4 = xch_n4(c);
  return (n1 > n4 ? n1 : n4) + (c.xch_ll ? 8 + ll_seg_words(c) : 0);
}
__device__ __forceinline__ int* xch_append_rec(const bh_ctx& c) {
  const long long n1 = xch_n1(c), n4 = xch_n4(c);
  return c.x_send + (n1 > n4 ? n1 : n4);
}
// this rank's receive area of exchange `kind` for the current step parity: record of rank s at + s * n
__device__ __forceinline__ long long xch_recv_off(const bh_ctx& c, int kind, int par) {
  const long long G = c.seg_world;
  return XCH_FLAG_INTS + (kind == 0 ? (long long)par * G * xch_n1(c) : 2 * G * xch_n1(c) + (long long)par * G * xch_n4(c));
}

// All CTAs of the cooperative grid: deliver send[0..n) (n a multiple of 4, 16-byte aligned,
// complete and visible grid-wide) to every rank's receive area and wait for all ranks'.
__device__ void xch_exchange(const bh_ctx& c, int kind, const int* send, long long n, int b, int nb, GridBar& bar) {
  const int G = c.seg_world, me = c.seg_rank;
  const int step = c.sc[BH_SC_STEP];
  const int seq = step + 1;
  const long long off = xch_recv_off(c, kind, step & 1) + (long long)me * n;
  const int4* src = reinterpret_cast<const int4*>(send);
  const long long n4 = n >> 2;
  const long long gid = (long long)b * blockDim.x + threadIdx.x, gsz = (long long)nb * blockDim.x;
#pragma unroll 1
  for (int p = 0; p < G; ++p) {
    int4* dst = reinterpret_cast<int4*>(c.xpeer[p] + off);
#pragma unroll 1
    for (long long i = gid; i < n4; i += gsz) dst[i] = src[i];
  }
  // one system-scope fence per CTA (its threads' stores are ordered before it by the block barrier), one
  // grid barrier, then the sequence numbers; every CTA polls the local flags itself (no second barrier)
  __syncthreads();
  if (threadIdx.x == 0) __threadfence_system();
  grid_barrier(bar, nb);
  if (b == 0 && threadIdx.x < G) {
    int* peer_flag = c.xpeer[threadIdx.x] + kind * 8 + me;
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(peer_flag), "r"(seq) : "memory");
  }
  if (threadIdx.x < G) {
    const int* my_flag = c.xpeer[me] + kind * 8 + threadIdx.x;
    const long long t0 = clock64();
    int v;
    do {
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(my_flag) : "memory");
      if (v < seq && clock64() - t0 > XCH_TIMEOUT_CYCLES) {
        atomicOr(&c.sc[BH_SC_STATUS], BH_ST_XCH_TIMEOUT);
        break;
      }
    } while (v < seq);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(FUSED_THREADS, 1)
    k_step_shard(const __grid_constant__ bh_ctx c, const uint32_t* input_fixed, int n_steps, int flags, int want_summary) {
  extern __shared__ __align__(16) uint32_t s_dyn[];
  // flags: BH_STEP_LEARNING | BH_STEP_NO_WINNER_CELLS, as in fused.cuh
  const int learning = flags & BH_STEP_LEARNING;
  const bool want_jit = !(flags & BH_STEP_NO_WINNER_CELLS);
  const bool want = learning || want_jit;
  const int b = blockIdx.x, nb = gridDim.x;
  const int nw = nb > 1 ? nb - 1 : 1;
  const bool worker = b < nw;
  const bool rng = b == nb - 1;
  GridBar bar = grid_bar_open(reinterpret_cast<unsigned int*>(c.sc + BH_SC_BAR_COUNT));
  GridBar bar2 = grid_bar_open(reinterpret_cast<unsigned int*>(c.sc + BH_SC_BAR2_COUNT));  // the bookkeeping team's own
#define BH_SYNC() grid_barrier(bar, (unsigned)nb)
  // phase timestamps of the last step (CTA 0): ctx.blk row 7, as 64-bit globaltimer ns (tools/cfg3_sharded.py)
  unsigned long long* stamps = reinterpret_cast<unsigned long long*>(c.blk + 7 * BH_BLK_STRIDE);
  int stamp_i = 0;
#define BH_STAMP()                                                            \
  do {                                                                        \
    if (b == 0 && threadIdx.x == 0) {                                         \
      unsigned long long t_;                                                  \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                  \
      stamps[stamp_i] = t_;                                                   \
    }                                                                         \
    ++stamp_i;                                                                \
  } while (0)
  const int G = c.seg_world, me = c.seg_rank;
  const int k = c.active_columns, k_loc = xch_k_loc(c);
  const long long n1 = xch_n1(c), n4 = xch_n4(c);
  int* region = c.xpeer[me];
  const long long gid = (long long)b * blockDim.x + threadIdx.x, gsz = (long long)nb * blockDim.x;
  const int pos0 = c.sc[BH_SC_INPUT_POS];
  for (int step = 0; step < n_steps; ++step) {
    const uint32_t* input =
        input_fixed ? input_fixed : c.input_ring + (long long)((pos0 + step) % c.ring_len) * c.input_words;
    const int par = c.sc[BH_SC_STEP] & 1;
    stamp_i = 0;
    BH_STAMP();
    // P0: overlap + boost of the local columns; draw #1 on the rng CTA
    if (rng && want) {
      ph_fill_jitter(c, s_dyn);  // the previous activation's deferred rand(M) first (no-op otherwise)
      __syncthreads();
      ph_draw(c, 1, 1, nw);
    }
    if (nb == 1) ph_overlap<true, true>(c, input, s_dyn, 0, 1);
    else if (!rng) ph_overlap<true, true>(c, input, s_dyn, b, nb - 1);
    BH_SYNC();
    BH_STAMP();  // 1: overlap
    // P1: global inhibition.  xch_ll: ONE CTA all-gathers the histograms the overlap phase built, derives the
    // threshold bin, all-gathers the columns above it and the members of it, and writes the active-column list
    // (shard_ll.cuh): one phase.  Its verdict (did the predicted binning resolve the threshold?) is identical
    // on every rank; a miss takes the candidate exchange below for this step.
    bool selected = false;
    if (c.xch_ll) {
      int* flag = c.topk_ws + TK3_BASE + TK3_SELECTED;
      if (b == 0) {
        int* llp[BH_MAX_RANKS];
        for (int p = 0; p < (G > 1 ? G : 1); ++p) llp[p] = c.xpeer[p] + xch_legacy_ints(c);
        const bool ok = ph_shard_select_ll(c, llp, s_dyn);
        if (threadIdx.x == 0) *flag = ok ? c.sc[BH_SC_STEP] + 1 : 0;
      }
      if (rng && nb > 1) ph_rng_speculate(c, 1);
      BH_SYNC();
      selected = *flag == c.sc[BH_SC_STEP] + 1;
      if (selected) {
        BH_STAMP();  // 2..4: selection
        BH_STAMP();
        BH_STAMP();
        BH_STAMP();
      }
    }
    if (!selected) {
    // this shard's best k_loc candidates -> record -> exchange 1 -> global top-k on every rank
    int* scratch = reinterpret_cast<int*>(c.row_unacc);  // not in use yet this step
    if (c.col_local >= XCH_TOPK_GRID_MIN) {
      topk_grid(c, reinterpret_cast<const unsigned long long*>(c.boosted), c.col_local, k_loc, scratch, nullptr, nullptr,
                b, nb, bar);
    } else if (b == 0) {
      topk_core(reinterpret_cast<const unsigned long long*>(c.boosted), c.col_local, k_loc, scratch, nullptr, nullptr);
    }
    BH_SYNC();
    BH_STAMP();  // 2: local top-k
    {
      double* rk = reinterpret_cast<double*>(c.x_send);
      int* rc = c.x_send + 2 * k_loc;
#pragma unroll 1
      for (long long i = gid; i < k_loc; i += gsz) {
        const int pos = scratch[i];
        rk[i] = c.boosted[pos];
        rc[i] = c.col_lo + pos;
      }
    }
    BH_SYNC();
    BH_STAMP();  // 3: record
    xch_exchange(c, 0, c.x_send, n1, b, nb, bar);
    BH_STAMP();  // 3: exchange 1
    {
      const int* recv = region + xch_recv_off(c, 0, par);
#pragma unroll 1
      for (long long i = gid; i < (long long)G * k_loc; i += gsz) {
        const int s = (int)(i / k_loc), j = (int)(i - (long long)s * k_loc);
        const int* rec = recv + s * n1;
        const int lo = __ldcv(rec + 2 * j), hi = __ldcv(rec + 2 * j + 1);
        c.xk_keys[i] = __hiloint2double(hi, lo);
        c.xk_cols[i] = __ldcv(rec + 2 * k_loc + j);
      }
    }
    BH_SYNC();
    if ((long long)G * k_loc >= XCH_TOPK_GRID_MIN) {
      if (b == 0) retire_prev_flags(c);
      topk_grid(c, reinterpret_cast<const unsigned long long*>(c.xk_keys), G * k_loc, k,
                c.active_cols + par * k, c.xk_cols, c.col_active, b, nb, bar);
    } else if (b == 0) {
      retire_prev_flags(c);
      topk_core(reinterpret_cast<const unsigned long long*>(c.xk_keys), G * k_loc, k, c.active_cols + par * k, c.xk_cols,
                c.col_active);
    }
    if (rng && nb > 1) ph_rng_speculate(c, 1);
    BH_SYNC();
    BH_STAMP();  // 4: unpack + global top-k
    }  // (!selected)
    // the binning of the next step's histogram after a candidate exchange: from the gathered candidates
    const bool rebin = c.xch_ll && !selected;
    // P2..P7 as in fused.cuh; learning and the segment scan touch only what this rank stores
    // a TEAM of the last CTAs runs the replicated temporal-memory bookkeeping chain while the others learn this
    // shard's spatial-pooler rows (fused.cuh, P2)
    // (a larger team does not shorten the chain: 64 CTAs at 8 shards measured 41 us against 29 with 16 -- its
    // cost is the dependent round trips and the team barriers, which grow with the team)
    const int team = (nb >= 32 && c.sc[BH_SC_M] <= 32768) ? (nb >= 64 ? 16 : 8) : 0;
    if (rebin && b == (team ? nb - team : 0)) tk3_rebin_sharded(c, G > 1 ? G * k_loc : k_loc);
    if (team) {
      const int t0 = nb - team;
      if (b >= t0) {
        // winner bits on the whole team | the drawing CTA forms the ordered winner lists while the others flag the
        // learning segments | it plans draw #2 while they form the learning lists (team barriers in between)
        ph_select_a(c, b - t0, team, want);
        grid_barrier(bar2, (unsigned)team);
        if (rng) ph_select_b(c, 0, 1, want, team);
        else ph_learn_select_a(c, learning, b - t0, team - 1);
        grid_barrier(bar2, (unsigned)team);
        if (rng) ph_draw(c, 2, learning, team - 1, true);
        else ph_learn_select_b(c, learning, b - t0, team - 1);
      } else {
        if (learning) ph_sp_learn<false>(c, input, b, t0);
        ph_duty(c, b, t0);
      }
      BH_SYNC();
      BH_STAMP();  // 5: SP learn + duty | winner bits, lists, learning flags, draw 2, learning lists
      BH_STAMP();
      BH_STAMP();
    } else {
      if (learning) ph_sp_learn<false>(c, input, b, nb);
      ph_duty(c, b, nb);
      if (worker) ph_select_a(c, b, nw, want);
      BH_SYNC();
      BH_STAMP();  // 5: SP learn + duty + winner bits
      if (worker) {
        ph_select_b(c, b, nw, want);
        ph_learn_select_a(c, learning, b, nw);
      }
      BH_SYNC();
      BH_STAMP();  // 6: lists + learning flags
      if (rng) ph_draw(c, 2, learning, nw, true);
      if (worker) ph_learn_select_b(c, learning, b, nw);
      BH_SYNC();
      BH_STAMP();  // 7: learning lists + draw 2
    }
    const bool lazy = c.rng64[R_LAZY] != 0;  // lazy step (fused.cuh): this rank produces only the rows it stores
    if (c.jump_polys > 0 && !lazy) {
      ph_rng_chunks(c, s_dyn, b, nb);
      BH_SYNC();
    }
    BH_STAMP();  // 8: stream chunks
    if (lazy) {
      // stage 1 of the learning pass and the tail jumps, both spread over every CTA (measured and rejected:
      // 48 CTAs of stage 1 next to the jumps on the rest -- no gain on one GPU, and the jump units then need more
      // rounds on a shard: 28 us against 16 at 8 shards)
      if (learning) ph_learn_apply(c, s_dyn, b, nb, 1);
      ph_rng_jumps(c, s_dyn, 0, (int)c.rng64[R_TAIL_CHUNKS], 0, 0, b, nb);
      BH_SYNC();
      ph_rng_lazy_rows(c, s_dyn, b, nb, [&]() { BH_SYNC(); },
                       [&](bool produce_rows) { ph_learn_grow(c, s_dyn, b, nb, produce_rows); });
    } else {
      if (learning) ph_learn_apply(c, s_dyn, b, nb);
      BH_SYNC();
    }
    BH_STAMP();  // 9: learn
    // (the activation words are double-buffered: what ph_post retires is not read by the scan, so it runs at the
    // head of the scan phase instead of in a phase of its own)
    const int ns = lazy ? nb - rng_tail_ctas(c, nb) : nw;  // lazy step: the last CTAs generate the tail instead
    if (lazy) ph_rng_lazy_tail(c, s_dyn, b, nb);
    ph_post(c, b, nb);
    if (b < ns) ph_activate_a(c, b, ns, c.xch_ll ? xch_append_rec(c) : nullptr);
    BH_SYNC();
    BH_STAMP();  // 10: segment scan
    // P8: this rank's record of matching / recyclable segments -> exchange 2 -> merged global lists
    if (c.xch_ll) {
      // the scan appended the record unordered; ONE CTA sorts it, all-gathers and merges (shard_ll.cuh).  A
      // record too long to sort there, or more recyclable segments than the exchange carries (the LOWEST ids
      // must be sent), is first rebuilt in order by all CTAs.
      int* app = xch_append_rec(c);
      const bool repack = app[0] > LL_LOCAL_MATCH_MAX || app[0] > c.xm_cap || app[1] > c.xr_cap ||
                          app[1] > LL_LOCAL_MATCH_MAX;  // uniform in the rank
      if (repack) {
        if (b < ns) ph_shard_pack(c, c.x_send, b, ns);
        BH_SYNC();
      }
      BH_STAMP();  // 11: record
      if (b == 0) {
        int* llp[BH_MAX_RANKS];
        for (int p = 0; p < (G > 1 ? G : 1); ++p) llp[p] = c.xpeer[p] + xch_legacy_ints(c);
        ph_shard_segs_ll(c, llp, s_dyn, repack ? c.x_send : app, repack);
        if (repack && threadIdx.x == 0) {
          app[0] = 0;
          app[1] = 0;
        }
      }
      BH_SYNC();
      BH_STAMP();  // 12: exchange 2 + merge
      BH_STAMP();
    } else {
    if (b < ns) ph_shard_pack(c, c.x_send, b, ns);
    BH_SYNC();
    BH_STAMP();  // 11: record
    xch_exchange(c, 1, c.x_send, n4, b, nb, bar);
    BH_STAMP();  // 12: exchange 2
    ph_shard_merge(c, region + xch_recv_off(c, 1, par), b, nb, (int)n4);
    BH_SYNC();
    BH_STAMP();  // 13: merge
    }
    // P9: draw #3 (a phase of its own only when not covered, see fused.cuh), jitter, predictions
    const int M = c.sc[BH_SC_X_MATCH] < c.match_capacity ? c.sc[BH_SC_X_MATCH] : c.match_capacity;
    const bool ready3 = !want_jit || (long long)M <= c.rng64[R_READY3];
    if (!ready3) {
      if (rng) ph_draw(c, 3, 1, ns);
      BH_SYNC();
    } else if (rng && want_jit) {
      ph_draw3_ready(c, ns);
    }
    if (worker) ph_activate_finish(c, b, nw, ready3, want_jit);
    BH_SYNC();
    BH_STAMP();  // 14: draw 3 + jitter + predictions
  }
  if (want_summary) ph_summary(c, b, nb);
  if (!input_fixed && b == 0 && threadIdx.x == 0) c.sc[BH_SC_INPUT_POS] = pos0 + n_steps;
#undef BH_SYNC
#undef BH_STAMP
}

// ---------------------------------------------------------------------------------------------------------
// The shard's step as a TWO-PIPELINE kernel (see k_step_pipe in fused.cuh): the spatial pooler of step s+1 --
// learn(s) + duty(s) | overlap(s+1) + histogram | one-CTA selection exchange(s+1) into a staging list -- on the
// first CTAs, the temporal memory of step s -- draw 1 | bookkeeping (a sub-team) | learn | scan | one-CTA
// segment exchange | jitter + predictions -- on the last ctx.pipe_ctas CTAs.  The two exchanges of a step then
// overlap each other and the HBM-bound passes instead of adding up.  The teams meet once per step; one CTA
// commits the staged active columns.  A selection whose predicted binning missed is redone after the join by
// the whole grid with the candidate exchange (then the step counter already names that step: the code of
// k_step_shard is used as it is).  Needs the cell exchanges (ctx.xch_ll).
// ---------------------------------------------------------------------------------------------------------
// the candidate exchange of k_step_shard (P1, !selected) for the step the step counter names; all CTAs
__device__ __noinline__ void shard_select_by_candidates(const bh_ctx& c, int b, int nb, GridBar& bar) {
  const int G = c.seg_world;
  const int k = c.active_columns, k_loc = xch_k_loc(c);
  const long long n1 = xch_n1(c);
  const int par = c.sc[BH_SC_STEP] & 1;
  const long long gid = (long long)b * blockDim.x + threadIdx.x, gsz = (long long)nb * blockDim.x;
  int* scratch = reinterpret_cast<int*>(c.row_unacc);  // not in use: the temporal memory is between two steps
  if (c.col_local >= XCH_TOPK_GRID_MIN) {
    topk_grid(c, reinterpret_cast<const unsigned long long*>(c.boosted), c.col_local, k_loc, scratch, nullptr, nullptr, b, nb, bar);
  } else if (b == 0) {
    topk_core(reinterpret_cast<const unsigned long long*>(c.boosted), c.col_local, k_loc, scratch, nullptr, nullptr);
  }
  grid_barrier(bar, nb);
  {
    double* rk = reinterpret_cast<double*>(c.x_send);
    int* rc = c.x_send + 2 * k_loc;
#pragma unroll 1
    for (long long i = gid; i < k_loc; i += gsz) {
      const int pos = scratch[i];
      rk[i] = c.boosted[pos];
      rc[i] = c.col_lo + pos;
    }
  }
  grid_barrier(bar, nb);
  xch_exchange(c, 0, c.x_send, n1, b, nb, bar);
  {
    const int* recv = c.xpeer[c.seg_rank] + xch_recv_off(c, 0, par);
#pragma unroll 1
    for (long long i = gid; i < (long long)G * k_loc; i += gsz) {
      const int s = (int)(i / k_loc), j = (int)(i - (long long)s * k_loc);
      const int* rec = recv + s * n1;
      const int lo = __ldcv(rec + 2 * j), hi = __ldcv(rec + 2 * j + 1);
      c.xk_keys[i] = __hiloint2double(hi, lo);
      c.xk_cols[i] = __ldcv(rec + 2 * k_loc + j);
    }
  }
  grid_barrier(bar, nb);
  if ((long long)G * k_loc >= XCH_TOPK_GRID_MIN) {
    if (b == 0) retire_prev_flags(c);
    topk_grid(c, reinterpret_cast<const unsigned long long*>(c.xk_keys), G * k_loc, k, c.active_cols + par * k, c.xk_cols,
              c.col_active, b, nb, bar);
  } else if (b == 0) {
    retire_prev_flags(c);
    topk_core(reinterpret_cast<const unsigned long long*>(c.xk_keys), G * k_loc, k, c.active_cols + par * k, c.xk_cols,
              c.col_active);
  }
  grid_barrier(bar, nb);
  if (b == 0) tk3_rebin_sharded(c, G > 1 ? G * k_loc : k_loc);  // binning of the next histogram from the candidates
}

__global__ void __launch_bounds__(FUSED_THREADS, 1)
    k_step_shard_pipe(const __grid_constant__ bh_ctx c, const uint32_t* input_fixed, int n_steps, int flags, int want_summary) {
  extern __shared__ __align__(16) uint32_t s_dyn[];
  const int learning = flags & BH_STEP_LEARNING;
  const bool want_jit = !(flags & BH_STEP_NO_WINNER_CELLS);
  const bool want = learning || want_jit;
  const int b = blockIdx.x, nb = gridDim.x;
  const int nt = c.pipe_ctas, ns = nb - nt;   // team sizes
  const bool sp_team = b < ns;
  const int tb = b - ns;                       // index inside the TM team
  const int nw = nt - 1;                       // TM CTAs running the ranged phases
  const bool worker = !sp_team && tb < nw;
  const bool rng = b == nb - 1;                // CTA producing the random draws
  const int team = nt >= 17 ? 16 : (nt >= 9 ? 8 : nt);  // bookkeeping sub-team: the last CTAs of the grid
  GridBar barA = grid_bar_open(reinterpret_cast<unsigned int*>(c.sc + BH_SC_BAR_COUNT));
  GridBar barT = grid_bar_open(reinterpret_cast<unsigned int*>(c.sc + BH_SC_BAR2_COUNT));
  GridBar barS = grid_bar_open(reinterpret_cast<unsigned int*>(c.sc + BH_SC_BAR3_COUNT));
  GridBar barB = grid_bar_open(reinterpret_cast<unsigned int*>(c.blk + 7 * BH_BLK_STRIDE + 600));  // the sub-team's
  unsigned long long* stamps = reinterpret_cast<unsigned long long*>(c.blk + 7 * BH_BLK_STRIDE);
  bool stamp_it = true;
#define BH_STAMP_AT(i)                                             \
  do {                                                             \
    if (threadIdx.x == 0 && (b == 0 || tb == 0) && stamp_it) {     \
      unsigned long long t_;                                       \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));       \
      stamps[(b == 0 ? 0 : 64) + (i)] = t_;                        \
    }                                                              \
  } while (0)
  const int G = c.seg_world;
  const int k = c.active_columns;
  const int step0 = c.sc[BH_SC_STEP];
  const int pos0 = c.sc[BH_SC_INPUT_POS];
  int* stage = c.active_cols + 2 * k;
  int* sel_flag = c.topk_ws + TK3_BASE + TK3_SELECTED;

  // ---- front of the first step on the whole grid: overlap + histogram, selection exchange
  {
    const uint32_t* input = input_fixed ? input_fixed : c.input_ring + (long long)(pos0 % c.ring_len) * c.input_words;
    BH_STAMP_AT(0);
    ph_overlap<true, true>(c, input, s_dyn, b, nb, step0);
    grid_barrier(barA, nb);
    if (b == 0) {
      int* llp[BH_MAX_RANKS];
      for (int p = 0; p < (G > 1 ? G : 1); ++p) llp[p] = c.xpeer[p] + xch_legacy_ints(c);
      const bool ok = ph_shard_select_ll(c, llp, s_dyn);
      if (threadIdx.x == 0) *sel_flag = ok ? step0 + 1 : 0;
    }
    grid_barrier(barA, nb);
    if (*sel_flag != step0 + 1) shard_select_by_candidates(c, b, nb, barA);
    BH_STAMP_AT(1);
  }
  for (int step = 0; step < n_steps; ++step) {
    const int s = step0 + step;
    const bool more = step + 1 < n_steps;
    stamp_it = more || n_steps == 1;
    const uint32_t* input =
        input_fixed ? input_fixed : c.input_ring + (long long)((pos0 + step) % c.ring_len) * c.input_words;
    grid_barrier(barA, nb);  // the active columns of step s are committed; step s-1 is complete
    BH_STAMP_AT(2);
    if (sp_team) {
      if (learning) ph_sp_learn<false>(c, input, b, ns, s);
      ph_duty(c, b, ns);
      BH_STAMP_AT(3);
      if (more) {
        const uint32_t* next = input_fixed ? input_fixed : c.input_ring + (long long)((pos0 + step + 1) % c.ring_len) * c.input_words;
        grid_barrier(barS, ns);
        ph_overlap<true, true>(c, next, s_dyn, b, ns, s + 1);
        grid_barrier(barS, ns);
        BH_STAMP_AT(4);
        if (b == 0) {  // the selection exchange of step s+1, into the staging list
          int* llp[BH_MAX_RANKS];
          for (int p = 0; p < (G > 1 ? G : 1); ++p) llp[p] = c.xpeer[p] + xch_legacy_ints(c);
          const bool ok = ph_shard_select_ll(c, llp, s_dyn, s + 1, stage);
          if (threadIdx.x == 0) *sel_flag = ok ? s + 2 : 0;
        }
        BH_STAMP_AT(5);
      }
    } else {
      // ---- temporal memory of step s
      BH_STAMP_AT(0);
      {  // replicated bookkeeping on the sub-team (as the team of k_step_shard); the other TM CTAs wait.
         // Draw #1 is the drawing CTA's first act there: only the sub-team waits for it.
        const int t0 = nb - team;
        if (b >= t0) {
          if (rng && want) {
            ph_fill_jitter(c, s_dyn);
            __syncthreads();
            ph_draw(c, 1, 1, nw);
          }
          grid_barrier(barB, (unsigned)team);
          BH_STAMP_AT(1);
          if (team > 1) {
            ph_select_a(c, b - t0, team, want);
            grid_barrier(barB, (unsigned)team);
            if (rng) ph_select_b(c, 0, 1, want, team);
            else ph_learn_select_a(c, learning, b - t0, team - 1);
            grid_barrier(barB, (unsigned)team);
            if (rng) ph_draw(c, 2, learning, team - 1, true);
            else ph_learn_select_b(c, learning, b - t0, team - 1);
          }
        }
      }
      grid_barrier(barT, nt);
      BH_STAMP_AT(2);
      const bool lazy = c.rng64[R_LAZY] != 0;
      if (c.jump_polys > 0 && !lazy) {
        ph_rng_chunks(c, s_dyn, tb, nt);
        grid_barrier(barT, nt);
      }
      BH_STAMP_AT(3);
      if (lazy) {
        if (learning) ph_learn_apply(c, s_dyn, tb, nt, 1);
        ph_rng_jumps(c, s_dyn, 0, (int)c.rng64[R_TAIL_CHUNKS], 0, 0, tb, nt);
        grid_barrier(barT, nt);
        ph_rng_lazy_rows(c, s_dyn, tb, nt, [&]() { grid_barrier(barT, nt); },
                         [&](bool produce_rows) { ph_learn_grow(c, s_dyn, tb, nt, produce_rows); });
      } else {
        if (learning) ph_learn_apply(c, s_dyn, tb, nt);
        grid_barrier(barT, nt);
      }
      BH_STAMP_AT(4);
      const int nscan = lazy ? nt - rng_tail_ctas(c, nt) : nw;
      if (lazy) ph_rng_lazy_tail(c, s_dyn, tb, nt);
      ph_post(c, tb, nt);
      if (tb < nscan) ph_activate_a(c, tb, nscan, xch_append_rec(c));
      grid_barrier(barT, nt);
      BH_STAMP_AT(5);
      int* app = xch_append_rec(c);
      const bool repack = app[0] > LL_LOCAL_MATCH_MAX || app[0] > c.xm_cap || app[1] > c.xr_cap ||
                          app[1] > LL_LOCAL_MATCH_MAX;  // uniform in the rank
      if (repack) {
        if (tb < nscan) ph_shard_pack(c, c.x_send, tb, nscan);
        grid_barrier(barT, nt);
      }
      if (tb == 0) {
        int* llp[BH_MAX_RANKS];
        for (int p = 0; p < (G > 1 ? G : 1); ++p) llp[p] = c.xpeer[p] + xch_legacy_ints(c);
        ph_shard_segs_ll(c, llp, s_dyn, repack ? c.x_send : app, repack);
        if (repack && threadIdx.x == 0) {
          app[0] = 0;
          app[1] = 0;
        }
      }
      grid_barrier(barT, nt);
      BH_STAMP_AT(6);
      const int M = c.sc[BH_SC_X_MATCH] < c.match_capacity ? c.sc[BH_SC_X_MATCH] : c.match_capacity;
      const bool ready3 = !want_jit || (long long)M <= c.rng64[R_READY3];
      if (!ready3) {
        if (rng) ph_draw(c, 3, 1, nscan);
        grid_barrier(barT, nt);
      } else if (rng && want_jit) {
        ph_draw3_ready(c, nscan);
      }
      if (worker) ph_activate_finish(c, tb, nw, ready3, want_jit);
      BH_STAMP_AT(7);
    }
    if (more) {
      grid_barrier(barA, nb);
      BH_STAMP_AT(sp_team ? 6 : 8);
      if (*sel_flag == s + 2) {
        if (b == 0) {  // commit the staged selection: flags of step s retired, those of step s+1 set, the list copied
          retire_prev_flags(c);  // (the step counter already says s+1: "previous" is the list of step s)
          int* out = c.active_cols + ((s + 1) & 1) * k;
#pragma unroll 1
          for (int i = threadIdx.x; i < k; i += blockDim.x) {
            const int col = stage[i];
            out[i] = col;
            c.col_active[col] = 1;
          }
        }
      } else {
        shard_select_by_candidates(c, b, nb, barA);  // (the step counter names step s+1 now)
      }
    }
  }
  grid_barrier(barA, nb);
  if (want_summary) ph_summary(c, b, nb);
  if (!input_fixed && b == 0 && threadIdx.x == 0) c.sc[BH_SC_INPUT_POS] = pos0 + n_steps;
#undef BH_STAMP_AT
}
