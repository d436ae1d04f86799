// Temporal-memory phases.  Reference: bithtm/networks.py:91-128 (TemporalMemory.process),
// bithtm/projections.py:194-293 (PredictiveProjection) on top of :27-192 (SparseProjection).
//
// Data layout: a segment is a row of syn_capacity (cell, permanence)
// slots kept COMPACT -- valid synapses are slots [0, seg_count) -- because every
// consumer in the reference is a count over the row, so slot order is free.
// There is no cell->segment forward index: activation is a segment-major scan.
//
// Every list whose ORDER the reference defines (winner cells, unaccounted winners,
// matching segments = index of their jitter draw, learning segments = row of the
// priority matrix, recycled segment ids) is produced by a two-phase ordered
// compaction: phase A counts per CTA over contiguous ranges, phase B derives its
// offset from the per-CTA counts and writes in order.  No atomics decide an order,
// so results are deterministic and equal to the reference's.
//
// Cells are addressed on the device as column * 32 + cell (c <= 32), so column and
// bit come from a shift and a mask; the host converts at the export boundary.
//
// Phase functions take (ctx, b, nb): CTA b of nb cooperating CTAs, any block size
// that is a multiple of 32 (>= 256 for the draw phase).
#pragma once

#include "common.cuh"
#include "mt19937.cuh"

// Totals of the learning bookkeeping, derived identically by every CTA from the
// per-CTA counts written by ph_learn_select_a (projections.py:264-281).
struct LearnTotals {
  int L0, P, n_u, n_r, n_new, L;
  int l_before, p_before, r_before;
  bool seg_overflow, learn_overflow;
};

__device__ __forceinline__ LearnTotals learn_totals(const bh_ctx& c, int learning, int b, int nb, int* s_red) {
  LearnTotals t;
  __shared__ int s_red6[192];
  int before[3], total[3];
  blk_prefix3(BLK(c, BLK_LEARN), BLK(c, BLK_PUNISH), BLK(c, BLK_RECYC), b, nb, s_red6, before, total);
  t.l_before = before[0];
  t.L0 = total[0];
  t.p_before = before[1];
  t.P = total[1];
  t.r_before = before[2];
  int R = total[2];
  if (c.seg_world > 1) {  // segment shards: candidates come merged from the exchange (tm_shard.cuh)
    R = c.sc[BH_SC_X_RECYC_AVAIL];
    t.r_before = 0;
  }
  const int S = c.sc[BH_SC_NSEG];
  t.n_u = (learning && c.sc[BH_SC_HAVE_PREV]) ? c.sc[BH_SC_NU] : 0;
  t.n_r = t.n_u < R ? t.n_u : R;  // projections.py:80-81: recycle first
  t.n_new = t.n_u - t.n_r;        // :90-94: then append
  t.seg_overflow = S + t.n_new > c.seg_capacity;
  if (t.seg_overflow) t.n_new = c.seg_capacity - S;
  t.L = t.L0 + t.n_r + t.n_new;
  t.learn_overflow = t.L > c.learn_capacity;
  if (t.learn_overflow) t.L = c.learn_capacity;
  return t;
}

// ---------------------------------------------------------------------------------
// Random draws.  which = 1: rand(k, c) (networks.py:87); 2: rand(L, W+1)
// (projections.py:120); 3: rand(M) (projections.py:235).  One CTA takes the range of
// the stream (mt19937.cuh); `nw` is the number of CTAs that ran the ranged phases
// (for the per-CTA count arrays).  Draw #2 may leave a parallel production plan that
// ph_rng_chunks (all CTAs) must run before the values are read.
// ---------------------------------------------------------------------------------
// `allow_lazy`: the caller is a cooperative step kernel that runs the lazy phases after draw #2 (mt19937.cuh).
__device__ __noinline__ void ph_draw(const bh_ctx& c, int which, int learning, int nw, bool allow_lazy = false) {
  __shared__ uint32_t x[MT_RING];
  __shared__ long long s_count;
  __shared__ int s_red[32];
  __shared__ int s_row_doubles;
  int m_before = 0, m_total = 0;
  LearnTotals lt;
  lt.L = 0;
  if (which == 3) blk_prefix(BLK(c, BLK_MATCH), 0, nw, s_red, m_before, m_total);
  if (which == 2) lt = learn_totals(c, learning, 0, nw, s_red);
  if (threadIdx.x == 0) {
    int* sc = c.sc;
    const int cur = sc[BH_SC_STEP] & 1;
    s_row_doubles = sc[BH_SC_W0 + (cur ^ 1)] + 1;  // a row of rand(L, W+1)
    long long count = 0;
    if (which == 1) {
      count = (long long)c.active_columns * c.cell_dim;
    } else if (which == 2) {
      // projections.py:191: no growth (and no draw) when the previous step had no winner cells at all
      if (learning && sc[BH_SC_HAVE_PREV] && !sc[BH_SC_WNONE0 + (cur ^ 1)])
        count = (long long)lt.L * (sc[BH_SC_W0 + (cur ^ 1)] + 1);
    } else {
      int M = c.seg_world > 1 ? sc[BH_SC_X_MATCH] : m_total;
      if (M > c.match_capacity) {
        M = c.match_capacity;
        atomicOr(&sc[BH_SC_STATUS], BH_ST_MATCH_OVERFLOW);
      }
      sc[BH_SC_M] = M;
      count = M;
    }
    s_count = count;
  }
  __syncthreads();
  if (which == 1) rng_draw(c, x, s_count, R_OFF1, -1, true, 0, false);
  else if (which == 2) rng_draw(c, x, s_count, R_OFF2, R_N2, false, c.rng_lookahead, true, true, s_row_doubles, allow_lazy);
  else {
    rng_draw(c, x, s_count, R_OFF3, R_N3, false, 0, false);
    if (threadIdx.x == 0) rng_finish_step(c);
  }
}

// Draw #3 without production, when the words exist already (count <= R_READY3, decided
// identically by every CTA): the drawing CTA only does the bookkeeping.
__device__ __noinline__ void ph_draw3_ready(const bh_ctx& c, int nw, int pre_total = -1) {
  __shared__ int s_red[32];
  int m_before, m_total = pre_total;
  if (pre_total < 0) blk_prefix(BLK(c, BLK_MATCH), 0, nw, s_red, m_before, m_total);
  if (threadIdx.x == 0) {
    int M = c.seg_world > 1 ? c.sc[BH_SC_X_MATCH] : m_total;
    if (M > c.match_capacity) {
      M = c.match_capacity;
      atomicOr(&c.sc[BH_SC_STATUS], BH_ST_MATCH_OVERFLOW);
    }
    c.sc[BH_SC_M] = M;
    rng_draw3_commit(c, M);
  }
}

__global__ void __launch_bounds__(MT_THREADS) k_tm_draw(const __grid_constant__ bh_ctx c, int which, int learning) {
  ph_draw(c, which, learning, c.tm_blocks);
}

// chunks of the production plan draw #2 may have left (no-op without one)
__global__ void __launch_bounds__(MT_THREADS, 1) k_rng_chunks(const __grid_constant__ bh_ctx c) {
  extern __shared__ __align__(16) uint32_t s_dyn[];
  ph_rng_chunks(c, s_dyn, blockIdx.x, gridDim.x);
}

// bh_rng_fill: take `count` doubles at the cursor (one CTA; may leave a plan for
// k_rng_chunks), then convert them (any grid).
__global__ void __launch_bounds__(MT_THREADS) k_rng_fill_draw(const __grid_constant__ bh_ctx c, long long count) {
  __shared__ uint32_t x[MT_RING];
  rng_draw(c, x, count, R_OFF1, R_N2, true, 0, true);
}

__global__ void k_rng_fill_copy(const __grid_constant__ bh_ctx c, double* dst) {
  const long long off = c.rng64[R_OFF1], n = c.rng64[R_N2];
#pragma unroll 1
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x)
    dst[q] = rng_uniform(c, off + 2 * q);
}

__global__ void k_rng_import(const __grid_constant__ bh_ctx c) { ph_rng_import(c); }

__global__ void k_rng_export(const __grid_constant__ bh_ctx c) {
  __shared__ int s_out[MT_N + 1];
  rng_export(c, s_out, threadIdx.x, blockDim.x);
  __syncthreads();
#pragma unroll 1
  for (int i = threadIdx.x; i < MT_N; i += blockDim.x) c.mt_key[i] = (uint32_t)s_out[i];
  if (threadIdx.x == 0) c.sc[BH_SC_MT_POS] = s_out[MT_N];
}

// ---------------------------------------------------------------------------------
// (f) bursting + winner cells, phase A: one warp per active column (lane = cell).
// networks.py:95-104 with evaluate_cell_best_matching (:73-82) and
// evaluate_cell_least_used (:84-89).  Also retires the winner words of columns that
// were active last step but are not now.
// ---------------------------------------------------------------------------------
// `want` = learning or return_winner_cell (networks.py:99): without it no winner cells are formed
// and rand(k, c) is not drawn.
__device__ void ph_select_a(const bh_ctx& c, int b, int nb, bool want = true) {
  __shared__ int s_red[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const int k = c.active_columns, cd = c.cell_dim;
  const int cur = c.sc[BH_SC_STEP] & 1;
  const bool have_prev = c.sc[BH_SC_HAVE_PREV] != 0;
  const int* act = c.active_cols + cur * k;
  const int* prev = c.active_cols + (cur ^ 1) * k;
  const long long off1 = c.rng64[R_OFF1];  // draw #1: rand(k, c), row-major
  const Range rg = block_range(k, b, nb);
  int n_win = 0, n_un = 0;
  #pragma unroll 1
  for (int r = rg.begin + warp; r < rg.end; r += warps) {
    const int col = act[r];
    const uint32_t pred = c.col_pred[col];  // previous step's prediction (networks.py:96)
    const bool burst = pred == 0u;          // networks.py:97
    const bool in = lane < cd;
    const int cell = col * 32 + lane;
    // best matching (networks.py:76-81); float32 arithmetic as in the reference
    float mj = in ? c.cell_maxjit[cell] : 0.0f;
    float colmax = warp_max(mj);
    bool col_matching = have_prev && colmax >= (float)c.seg_matching_threshold;
    bool best = have_prev && in && fabsf(__fsub_rn(mj, colmax)) < c.epsilon;
    // least used (networks.py:85-88): f32(f64(count) + u)
    float x = INFINITY;
    if (in && want)
      x = __double2float_rn(__dadd_rn((double)c.cell_nseg[cell], rng_uniform(c, off1 + 2 * ((long long)r * cd + lane))));
    float rowmin = warp_min(x);
    bool least = in && fabsf(__fsub_rn(x, rowmin)) < c.epsilon;
    bool predbit = (pred >> lane) & 1u;
    bool win = want && in && (predbit || (burst && (col_matching ? best : least)));  // networks.py:102
    uint32_t wbits = __ballot_sync(BH_FULL, win);
    // winners no matching segment points at (projections.py:271)
    uint32_t ubits = __ballot_sync(BH_FULL, win && have_prev && mj < c.epsilon);
    if (lane == 0) {
      c.row_pred[r] = pred;
      c.row_win[r] = wbits;
      const uint32_t abits = burst ? low_mask(cd) : pred;  // networks.py:115
      c.row_act[r] = abits;
      c.col_act[(long long)cur * c.column_dim + col] = abits;  // this step's buffer is all-zero here (ph_post)
      c.row_unacc[r] = ubits;
      c.col_win[col] = wbits;
      n_win += __popc(wbits);
      n_un += __popc(ubits);
    }
  }
  #pragma unroll 1
  for (int i = b * blockDim.x + threadIdx.x; i < k; i += nb * blockDim.x) {
    int col = prev[i];
    if (!c.col_active[col]) c.col_win[col] = 0u;
  }
  int tw = block_sum(n_win, s_red);
  int tu = block_sum(n_un, s_red);
  if (threadIdx.x == 0) {
    BLK(c, BLK_WIN)[b] = tw;
    BLK(c, BLK_UNACC)[b] = tu;
  }
}

// Phase B: ordered winner / unaccounted lists (row order of active_column, cells
// ascending: np.where on the [k, c] mask, networks.py:103-104).
// `nb_counts` > 0: ONE CTA (b = 0, nb = 1) forms the whole lists from the counts that nb_counts CTAs left in phase A.
__device__ void ph_select_b(const bh_ctx& c, int b, int nb, bool want = true, int nb_counts = 0) {
  __shared__ int s_red[32];
  const int k = c.active_columns, cd = c.cell_dim, NT = blockDim.x;
  const int cur = c.sc[BH_SC_STEP] & 1;
  const int* act = c.active_cols + cur * k;
  int* wl = c.winners + (long long)cur * k * cd;
  int w_before, w_total, u_before, u_total;
  blk_prefix(BLK(c, BLK_WIN), b, nb_counts > 0 ? nb_counts : nb, s_red, w_before, w_total);
  blk_prefix(BLK(c, BLK_UNACC), b, nb_counts > 0 ? nb_counts : nb, s_red, u_before, u_total);
  const Range rg = block_range(k, b, nb);
  int wbase = w_before, ubase = u_before;
  #pragma unroll 1
  for (int tile = rg.begin; tile < rg.end; tile += NT) {
    const int r = tile + threadIdx.x;
    const bool ok = r < rg.end;
    uint32_t wb = ok ? c.row_win[r] : 0u;
    uint32_t ub = ok ? c.row_unacc[r] : 0u;
    const int cell0 = ok ? act[r] * 32 : 0;
    int tot;
    int p = wbase + block_excl_scan(__popc(wb), s_red, tot);
    wbase += tot;
    while (wb) {
      int bit = __ffs(wb) - 1;
      wb &= wb - 1;
      wl[p++] = cell0 + bit;
    }
    p = ubase + block_excl_scan(__popc(ub), s_red, tot);
    ubase += tot;
    while (ub) {
      int bit = __ffs(ub) - 1;
      ub &= ub - 1;
      c.unacc[p++] = cell0 + bit;
    }
  }
  if (b == 0 && threadIdx.x == 0) {
    c.sc[BH_SC_W0 + cur] = w_total;
    c.sc[BH_SC_NU] = u_total;
    c.sc[BH_SC_WNONE0 + cur] = want ? 0 : 1;  // winner_cell is None (networks.py:125)
  }
}

// ---------------------------------------------------------------------------------
// (f) learning / punished segment selection among the previous matching segments,
// phase A (flags + per-CTA counts, projections.py:264-269), plus the per-CTA count
// of recyclable segments (fewer than `matching threshold` synapses, :80).
// ---------------------------------------------------------------------------------
__device__ void ph_learn_select_a(const bh_ctx& c, int learning, int b, int nb) {
  const int NT = blockDim.x;
  const int M = c.sc[BH_SC_M];
  const Range rg = block_range(M, b, nb);
  int nl = 0, np = 0, nr = 0;
  #pragma unroll 1
  for (int j = rg.begin + threadIdx.x; j < rg.end; j += NT) {
    uint8_t f = 0;
    if (learning) {
      const int s = c.m_seg[j];
      const int owner = c.seg_owner[s];
      const int col = owner >> 5, bit = owner & 31;
      bool is_winner = (c.col_win[col] >> bit) & 1u;                               // :262
      bool seg_active = c.m_conn[j] >= c.seg_activation_threshold;                 // :250
      bool unpredicted = c.cell_npred[owner] == 0;                                 // :266
      bool best = fabsf(__fsub_rn(c.m_jit[j], c.cell_maxjit[owner])) < c.epsilon;  // :267
      if (is_winner && (seg_active || (unpredicted && best))) f |= 1;              // :268
      if (!c.col_active[col]) f |= 2;                                              // :269
    }
    c.m_flag[j] = f;
    nl += f & 1;
    np += (f >> 1) & 1;
  }
  if (learning && c.sc[BH_SC_HAVE_PREV] && c.seg_world <= 1) {
    const int S = c.sc[BH_SC_NSEG], thr = c.seg_matching_threshold;
    const Range sr = block_range(S, b, nb);
    #pragma unroll 1
    for (int s = sr.begin + threadIdx.x; s < sr.end; s += 8 * NT) {  // 8 independent loads in flight per thread
      int v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = s + u * NT < sr.end ? c.seg_count[s + u * NT] : thr;
#pragma unroll
      for (int u = 0; u < 8; ++u) nr += v[u] < thr ? 1 : 0;
    }
  }
  __shared__ int s_red3[96];
  block_sum3(nl, np, nr, s_red3);
  if (threadIdx.x == 0) {
    BLK(c, BLK_LEARN)[b] = nl;
    BLK(c, BLK_PUNISH)[b] = np;
    BLK(c, BLK_RECYC)[b] = nr;
  }
}

// Phase B: ordered learning / punished lists; segments for unaccounted winners:
// recycle the lowest-id short segments, then append new ids (projections.py:79-95,
// 271-281); and the sparse reset of the per-cell results of the previous activation
// (all their readers ran already).
__device__ void ph_learn_select_b(const bh_ctx& c, int learning, int b, int nb) {
  __shared__ int s_red[32];
  const int NT = blockDim.x;
  const int M = c.sc[BH_SC_M];
  const int S = c.sc[BH_SC_NSEG];
  const int thr = c.seg_matching_threshold;
  const LearnTotals lt = learn_totals(c, learning, b, nb, s_red);
  {
    const Range rg = block_range(M, b, nb);
    int lbase = lt.l_before, pbase = lt.p_before;
    #pragma unroll 1
    for (int tile = rg.begin; tile < rg.end; tile += NT) {
      const int j = tile + threadIdx.x;
      const bool ok = j < rg.end;
      const int f = ok ? c.m_flag[j] : 0;
      const int s = ok ? c.m_seg[j] : 0;
      int tot;
      int pos = lbase + block_excl_scan(f & 1, s_red, tot);
      lbase += tot;
      if ((f & 1) && pos < c.learn_capacity) c.learn_list[pos] = s;
      pos = pbase + block_excl_scan((f >> 1) & 1, s_red, tot);
      pbase += tot;
      if (f & 2) c.punish_list[pos] = s;
      if (ok) {
        const int owner = c.seg_owner[s];
        c.cell_maxjit[owner] = 0.0f;
        c.cell_npred[owner] = 0;
        c.col_pred[owner >> 5] = 0u;
      }
    }
  }
  if (lt.n_u > 0 && c.seg_world > 1) {
    // segment shards: the merged ascending list of recyclable ids (every rank identically)
    #pragma unroll 1
    for (int i = b * NT + threadIdx.x; i < lt.n_r; i += nb * NT) {
      const int s = c.recyc_list[i], newo = c.unacc[i];
      atomicSub(&c.cell_nseg[c.seg_owner[s]], 1);
      atomicAdd(&c.cell_nseg[newo], 1);
      c.seg_owner[s] = newo;
      c.seg_count[s] = 0;
      const int pos = lt.L0 + i;
      if (pos < c.learn_capacity) c.learn_list[pos] = s;
    }
    if (b == 0 && threadIdx.x == 0 && lt.n_u > c.sc[BH_SC_X_RECYC_AVAIL] &&
        c.sc[BH_SC_X_RECYC_TOTAL] > c.sc[BH_SC_X_RECYC_AVAIL])
      atomicOr(&c.sc[BH_SC_STATUS], BH_ST_XCH_OVERFLOW);  // a rank had more candidates than it could send
  }
  if (lt.n_u > 0) {
    if (lt.r_before < lt.n_u && c.seg_world <= 1) {
      const Range sr = block_range(S, b, nb);
      int rbase = lt.r_before;
      #pragma unroll 1
      for (int tile = sr.begin; tile < sr.end && rbase < lt.n_u; tile += NT) {
        const int s = tile + threadIdx.x;
        const bool rec = s < sr.end && c.seg_count[s] < thr;
        int tot;
        const int rank = rbase + block_excl_scan(rec ? 1 : 0, s_red, tot);
        rbase += tot;
        if (rec && rank < lt.n_u) {
          const int newo = c.unacc[rank];
          atomicSub(&c.cell_nseg[c.seg_owner[s]], 1);  // :275-276
          atomicAdd(&c.cell_nseg[newo], 1);            // :277
          c.seg_owner[s] = newo;                       // :278
          c.seg_count[s] = 0;                          // :83-85 (row logically emptied)
          const int pos = lt.L0 + rank;
          if (pos < c.learn_capacity) c.learn_list[pos] = s;
        }
      }
    }
    #pragma unroll 1
    for (int i = b * NT + threadIdx.x; i < lt.n_new; i += nb * NT) {
      const int rank = lt.n_r + i, s = S + i;
      const int owner = c.unacc[rank];
      c.seg_owner[s] = owner;  // :280
      c.seg_count[s] = 0;
      atomicAdd(&c.cell_nseg[owner], 1);
      const int pos = lt.L0 + rank;
      if (pos < c.learn_capacity) c.learn_list[pos] = s;
    }
  }
  if (b == 0 && threadIdx.x == 0) {
    c.sc[BH_SC_NPREDCOL_PREV] = c.sc[BH_SC_NPREDCOL];  // the predictions reset above (example.py:50)
    c.sc[BH_SC_NPREDCOL] = 0;
    c.sc[BH_SC_L0] = lt.L0;
    c.sc[BH_SC_L] = lt.L;
    c.sc[BH_SC_P] = lt.P;
    c.sc[BH_SC_NR] = lt.n_r;
    c.sc[BH_SC_NSEG_NEXT] = S + lt.n_new;
    if (lt.seg_overflow) atomicOr(&c.sc[BH_SC_STATUS], BH_ST_SEG_OVERFLOW);
    if (lt.learn_overflow) atomicOr(&c.sc[BH_SC_STATUS], BH_ST_LEARN_OVERFLOW);
  }
}

// ---------------------------------------------------------------------------------
// (f) learning + punishment applied to segment rows.
//  Stage 1, one warp per row: permanence update in float64, stored as float32;
//  synapses whose float64 sum is negative are deleted (projections.py:97-109) and the
//  row is re-compacted in place.
//  Stage 2, one CTA per learning row that must grow (projections.py:111-161):
//  n_add = clip(sample - #synapses to previously active cells, 0, min(sample, W));
//  priorities are float32(rand(L, W+1)); previous winners already on the segment are
//  excluded (bitmap in shared memory); the n_add smallest priorities < 1.0 win.  The
//  cut is found by a 4x8-bit radix select over the float32 bit patterns (positive
//  floats order like unsigned ints); ties at the cut -> lower winner index and the
//  PRI_TIE status bit (np.argsort is undefined there).
//  `s_excl` = dynamic shared memory, ceil(k*c/32) words.
// ---------------------------------------------------------------------------------
__device__ void grow_row(const bh_ctx& c, int row, int s, int n, int n_add, int Wp, const int* prevw, long long off2,
                         bool pr_ok, uint32_t* s_excl, int* s_red, int* s_hist, int* s_rem, uint32_t* s_prefix) {
  const int t = threadIdx.x, NT = blockDim.x, E = c.syn_capacity;
  int* cells = c.syn_cell + (long long)seg_row(c, s) * E;
  float* perms = c.syn_perm + (long long)seg_row(c, s) * E;
  const int excl_words = (Wp + 31) >> 5;
  #pragma unroll 1
  for (int i = t; i < excl_words; i += NT) s_excl[i] = 0u;
  __syncthreads();
  #pragma unroll 1
  for (int slot = t; slot < n; slot += NT) {  // projections.py:117-121
    const int wi = c.cell_widx[cells[slot]];
    if (wi >= 0) atomicOr(&s_excl[wi >> 5], 1u << (wi & 31));
  }
  __syncthreads();
  const long long pr = off2 + 2 * (long long)row * (Wp + 1);  // stream index of this row of rand(L, W+1)
  // candidates: previous winners not yet on the segment with priority < 1.0 (:121-123)
  int nc = 0;
  #pragma unroll 1
  for (int w = t; w < Wp; w += NT) {
    bool ex = (s_excl[w >> 5] >> (w & 31)) & 1u;
    float pri = pr_ok ? __double2float_rn(rng_uniform(c, pr + 2 * w)) : 2.0f;
    nc += (!ex && pri < 1.0f) ? 1 : 0;
  }
  nc = block_sum(nc, s_red);
  uint32_t cut = 0x3f800000u;  // bits of 1.0f: take every candidate
  int ties_wanted = 0;
  if (nc > n_add) {
    if (t == 0) {
      *s_prefix = 0u;
      *s_rem = n_add;
    }
    __syncthreads();
    for (int pass = 3; pass >= 0; --pass) {
      #pragma unroll 1
      for (int i = t; i < 256; i += NT) s_hist[i] = 0;
      __syncthreads();
      const uint32_t prefix = *s_prefix;
      const int shift = pass * 8;
      const uint32_t hi_mask = pass == 3 ? 0u : (0xffffffffu << (shift + 8));
      #pragma unroll 1
      for (int w = t; w < Wp; w += NT) {
        bool ex = (s_excl[w >> 5] >> (w & 31)) & 1u;
        float pri = __double2float_rn(rng_uniform(c, pr + 2 * w));
        uint32_t bits = __float_as_uint(pri);
        if (!ex && pri < 1.0f && (bits & hi_mask) == prefix) atomicAdd(&s_hist[(bits >> shift) & 0xff], 1);
      }
      __syncthreads();
      if (t == 0) {
        int rem = *s_rem, d = 0;
        for (; d < 255; ++d) {
          if (s_hist[d] >= rem) break;
          rem -= s_hist[d];
        }
        *s_rem = rem;
        *s_prefix = prefix | ((uint32_t)d << shift);
      }
      __syncthreads();
    }
    cut = *s_prefix;       // the n_add-th smallest candidate priority
    ties_wanted = *s_rem;  // how many candidates == cut to take (lowest index first)
  }
  // ordered append of the chosen winners
  int base = n, tie_base = 0;
  #pragma unroll 1
  for (int tile = 0; tile < Wp; tile += NT) {
    const int w = tile + t;
    bool cand = false;
    uint32_t bits = 0u;
    if (w < Wp) {
      bool ex = (s_excl[w >> 5] >> (w & 31)) & 1u;
      float pri = pr_ok ? __double2float_rn(rng_uniform(c, pr + 2 * w)) : 2.0f;
      bits = __float_as_uint(pri);
      cand = !ex && pri < 1.0f;
    }
    const bool lt = cand && bits < cut;
    const bool eq = cand && bits == cut && cut != 0x3f800000u;
    int tot_eq, tot;
    const int tie_rank = tie_base + block_excl_scan(eq ? 1 : 0, s_red, tot_eq);
    const bool take = lt || (eq && tie_rank < ties_wanted);
    const int pos = base + block_excl_scan(take ? 1 : 0, s_red, tot);
    if (take) {
      if (pos < E) {
        cells[pos] = prevw[w];
        perms[pos] = c.tm_perm_initial;  // :149
      } else {
        atomicOr(&c.sc[BH_SC_STATUS], BH_ST_SYN_OVERFLOW);
      }
    }
    base += tot;
    tie_base += tot_eq;
  }
  if (t == 0) {
    c.seg_count[s] = base < E ? base : E;  // :161
    if (nc > n_add && tie_base > ties_wanted) atomicOr(&c.sc[BH_SC_STATUS], BH_ST_PRI_TIE);
  }
  __syncthreads();
}

// mode 0: both stages (the priority rows are in the stream ring).  mode 1 (lazy steps, mt19937.cuh): stage 1
// only; growing rows are appended to ctx.grow_list as (row, synapses kept, n_add) for ph_learn_grow.
__device__ void ph_learn_apply(const bh_ctx& c, uint32_t* s_excl, int b, int nb, int mode = 0) {
  __shared__ int s_red[32];
  __shared__ int s_hist[256];
  __shared__ int s_rem;
  __shared__ uint32_t s_prefix;
  __shared__ int s_ngrow;
  __shared__ int s_grow_row[BH_MAX_WARPS], s_grow_n[BH_MAX_WARPS], s_grow_add[BH_MAX_WARPS];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, warps = blockDim.x >> 5;
  const int cd = c.cell_dim, E = c.syn_capacity;
  const int L = c.sc[BH_SC_L], P = c.sc[BH_SC_P];
  const int cur = c.sc[BH_SC_STEP] & 1;
  const int Wp = c.sc[BH_SC_W0 + (cur ^ 1)];
  const int* prevw = c.winners + (long long)(cur ^ 1) * c.active_columns * cd;
  const uint32_t* prev_act = c.col_act + (long long)(cur ^ 1) * c.column_dim;  // cell_activation of the previous step
  const long long off2 = c.rng64[R_OFF2];
  const long long n2 = c.rng64[R_N2];  // doubles draw #2 obtained (capacity-clamped)
  const int sample = c.seg_sampling_synapses;
  const int cap = sample < Wp ? sample : Wp;

  #pragma unroll 1
  for (int base_row = 0; base_row < L + P; base_row += nb * warps) {
    if (t == 0) s_ngrow = 0;
    __syncthreads();
    // ---- stage 1: one warp per row (rows interleaved over CTAs) -------------------
    const int row = base_row + warp * nb + b;
    const int s_row = row < L + P ? (row < L ? c.learn_list[row] : c.punish_list[row - L]) : 0;
    if (row < L + P && seg_held(c, s_row)) {  // segment shards: only the rows stored here
      const bool learn = row < L;
      const int s = s_row;
      const double d_on = learn ? c.tm_learn_on : c.tm_punish_on;
      const double d_off = learn ? c.tm_learn_off : c.tm_punish_off;
      const bool can_delete = learn ? c.tm_learn_can_delete : c.tm_punish_can_delete;
      int* cells = c.syn_cell + (long long)seg_row(c, s) * E;
      float* perms = c.syn_perm + (long long)seg_row(c, s) * E;
      const int n = c.seg_count[s];
      int kept = 0, n_act = 0;
      #pragma unroll 1
      for (int g = 0; g < n; g += 32) {
        const int slot = g + lane;
        const bool valid = slot < n;
        const int cell = valid ? cells[slot] : 0;
        const float p = valid ? perms[slot] : 0.0f;
        const bool act = valid && cell_bit(prev_act, cell);           // previous activation (networks.py:111)
        const double sum = __dadd_rn((double)p, act ? d_on : d_off);  // :102-103
        const bool keep = valid && !(can_delete && sum < 0.0);        // :105-108
        const uint32_t kb = __ballot_sync(BH_FULL, keep);
        const uint32_t ab = __ballot_sync(BH_FULL, keep && act);
        const int pos = kept + __popc(kb & ((1u << lane) - 1u));
        if (keep) {
          cells[pos] = cell;
          perms[pos] = __double2float_rn(sum);
        }
        __syncwarp();
        kept += __popc(kb);
        n_act += __popc(ab);
      }
      if (lane == 0) {
        c.seg_count[s] = kept;
        int n_add = sample - n_act;  // projections.py:114-115
        n_add = n_add < 0 ? 0 : (n_add > cap ? cap : n_add);
        if (learn && n_add > 0) {
          if (mode == 1) {
            const int g = atomicAdd(&c.sc[BH_SC_NGROW], 1);  // learn_capacity entries: g < L always
            c.grow_list[3 * g] = row;
            c.grow_list[3 * g + 1] = kept;
            c.grow_list[3 * g + 2] = n_add;
          } else {
            const int g = atomicAdd(&s_ngrow, 1);
            s_grow_row[g] = row;
            s_grow_n[g] = kept;
            s_grow_add[g] = n_add;
          }
        }
      }
    }
    __syncthreads();
    // ---- stage 2: the whole CTA grows the rows its warps flagged ------------------
    const int ng = s_ngrow;
    if (t == 0 && ng > 0) atomicAdd(&c.sc[BH_SC_NGROW], ng);  // growing rows of the step (lazy / dense policy)
    for (int g = 0; g < ng; ++g) {
      const int grow = s_grow_row[g];
      const bool pr_ok = (long long)(grow + 1) * (Wp + 1) <= n2;
      grow_row(c, grow, c.learn_list[grow], s_grow_n[g], s_grow_add[g], Wp, prevw, off2, pr_ok, s_excl, s_red,
               s_hist, &s_rem, &s_prefix);
    }
    __syncthreads();
  }
}

// Lazy steps, stage 2: grow the rows of ctx.grow_list (their priority rows were produced by jumps).
// `produce_rows`: each row's words are a production job of its own (kind 1) that the growing CTA generates
// first; else the whole matrix was produced in chunks.  smem: max(MT_RING, excl words).
__device__ __noinline__ void ph_learn_grow(const bh_ctx& c, uint32_t* smem, int b, int nb, bool produce_rows) {
  __shared__ int s_red[32];
  __shared__ int s_hist[256];
  __shared__ int s_rem;
  __shared__ uint32_t s_prefix;
  const int G = c.sc[BH_SC_NGROW];
  const int cur = c.sc[BH_SC_STEP] & 1;
  const int Wp = c.sc[BH_SC_W0 + (cur ^ 1)];
  const int* prevw = c.winners + (long long)(cur ^ 1) * c.active_columns * c.cell_dim;
  const long long off2 = c.rng64[R_OFF2], n2 = c.rng64[R_N2];
#pragma unroll 1
  for (int i = b; i < G; i += nb) {
    const int row = c.grow_list[3 * i], n = c.grow_list[3 * i + 1], n_add = c.grow_list[3 * i + 2];
    if (produce_rows) {
      rng_job_generate(c, smem, 1, i, 2 * (Wp + 1), RNG_ROW_SLOT0);
      __syncthreads();
    }
    const bool pr_ok = (long long)(row + 1) * (Wp + 1) <= n2;
    grow_row(c, row, c.learn_list[row], n, n_add, Wp, prevw, off2, pr_ok, smem, s_red, s_hist, &s_rem, &s_prefix);
    __syncthreads();
  }
}

#define LA_THREADS 256
__global__ void __launch_bounds__(LA_THREADS) k_tm_learn_apply(const __grid_constant__ bh_ctx c) {
  extern __shared__ __align__(16) uint32_t s_dyn[];
  ph_learn_apply(c, s_dyn, blockIdx.x, gridDim.x);
}

// ---------------------------------------------------------------------------------
// Commit this step's activation and winner index (after learning read the previous
// ones) and the segment count.  networks.py:118-119; projections.py:117-118.
// ---------------------------------------------------------------------------------
__device__ void ph_post(const bh_ctx& c, int b, int nb) {
  const int k = c.active_columns, cd = c.cell_dim;
  const int cur = c.sc[BH_SC_STEP] & 1;
  const int* prev = c.active_cols + (cur ^ 1) * k;
  const int Wc = c.sc[BH_SC_W0 + cur], Wp = c.sc[BH_SC_W0 + (cur ^ 1)];
  const int* wl_cur = c.winners + (long long)cur * k * cd;
  const int* wl_prev = c.winners + (long long)(cur ^ 1) * k * cd;
  const int gid = b * blockDim.x + threadIdx.x, gsz = nb * blockDim.x;
  // the previous step's activation words (the OTHER buffer) have had their last reader (learning): zero
  // them, so the next step finds its buffer empty and only sets the words of its active columns
  uint32_t* old_act = c.col_act + (long long)(cur ^ 1) * c.column_dim;
  #pragma unroll 1
  for (int r = gid; r < k; r += gsz) old_act[prev[r]] = 0u;
  #pragma unroll 1
  for (int i = gid; i < Wc; i += gsz) c.cell_widx[wl_cur[i]] = i;
  #pragma unroll 1
  for (int i = gid; i < Wp; i += gsz) {
    int cell = wl_prev[i];
    if (!((c.col_win[cell >> 5] >> (cell & 31)) & 1u)) c.cell_widx[cell] = -1;
  }
  if (gid == 0) c.sc[BH_SC_NSEG] = c.sc[BH_SC_NSEG_NEXT];
}

// ---------------------------------------------------------------------------------
// (e) segment activation, phase A: one warp per segment scans its synapses against
// the active-cell bit-words.  potential = synapses to active cells
// (projections.py:175-178), connected-active = those with permanence >= threshold
// (:167-173).
// ---------------------------------------------------------------------------------
#define ACT_BATCH 5  // segments per warp iteration, 64 slots each: 20 independent loads in flight per lane
// `append` (fused sharded step with cell exchanges): the matching and the recyclable segments are also appended
// to the lists at append[8..] (counters append[0], append[1]).
__device__ void ph_activate_a(const bh_ctx& c, int b, int nb, int* append = nullptr) {
  __shared__ int s_red[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const int E = c.syn_capacity;
  const int S = c.sc[BH_SC_NSEG_NEXT];  // == NSEG once ph_post has committed it (it may run alongside)
  const int thr = c.seg_matching_threshold;
  const float pthr = c.tm_perm_threshold;
  const uint32_t* col_act = c.col_act + (long long)(c.sc[BH_SC_STEP] & 1) * c.column_dim;
  const Range rg = block_range(seg_local_count(c, S), b, nb);  // local rows (== segment ids when not sharded)
  int nm = 0, nrec = 0;
  // The synapse counts of a batch are fetched one iteration AHEAD (software pipeline), so the slot loads can be
  // predicated on them -- only live slots travel (rows hold ~32 of 64..128 slots) -- without a third dependent
  // round trip per batch: counts(next) and slots(current) are in flight together.
  int s_first = rg.begin + warp * ACT_BATCH;
  int next_n = (lane < ACT_BATCH && s_first + lane < rg.end) ? c.seg_count[seg_gid(c, s_first + lane)] : 0;
#pragma unroll 1
  for (int s0 = s_first; s0 < rg.end; s0 += warps * ACT_BATCH) {
    const int my_n = next_n;
    {
      const int s1 = s0 + warps * ACT_BATCH;
      next_n = (lane < ACT_BATCH && s1 + lane < rg.end) ? c.seg_count[seg_gid(c, s1 + lane)] : 0;
    }
    int n[ACT_BATCH], cell[ACT_BATCH][2];
    float perm[ACT_BATCH][2];
#pragma unroll
    for (int j = 0; j < ACT_BATCH; ++j) n[j] = __shfl_sync(BH_FULL, my_n, j);
#pragma unroll
    for (int j = 0; j < ACT_BATCH; ++j) {
      const long long base = (long long)(s0 + j) * E + lane;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const bool v = s0 + j < rg.end && lane + 32 * h < n[j];
        cell[j][h] = v ? c.syn_cell[base + 32 * h] : 0;
        perm[j][h] = v ? c.syn_perm[base + 32 * h] : 0.0f;
      }
    }
    if (lane < ACT_BATCH && s0 + lane < rg.end && my_n < thr) {  // recyclable (projections.py:80)
      ++nrec;
      if (append) {
        const int j = atomicAdd(&append[1], 1);
        if (j < c.xr_cap) append[8 + 3 * c.xm_cap + j] = seg_gid(c, s0 + lane);
      }
    }
    uint32_t word[ACT_BATCH][2];
#pragma unroll
    for (int j = 0; j < ACT_BATCH; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) word[j][h] = lane + 32 * h < n[j] ? col_act[cell[j][h] >> 5] : 0u;
#pragma unroll
    for (int j = 0; j < ACT_BATCH; ++j) {
      int pot = 0, conn = 0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const bool a = lane + 32 * h < n[j] && ((word[j][h] >> (cell[j][h] & 31)) & 1u);
        pot += __popc(__ballot_sync(BH_FULL, a));
        conn += __popc(__ballot_sync(BH_FULL, a && perm[j][h] >= pthr));
      }
      if (n[j] > 64) {  // rows longer than two warp-widths
        const int s = s0 + j;
#pragma unroll 1
        for (int g = 64; g < n[j]; g += 32) {
          const int slot = g + lane;
          const bool v = slot < n[j];
          const int cl = v ? c.syn_cell[(long long)s * E + slot] : 0;
          const float pm = v ? c.syn_perm[(long long)s * E + slot] : 0.0f;
          const bool a2 = v && cell_bit(col_act, cl);
          pot += __popc(__ballot_sync(BH_FULL, a2));
          conn += __popc(__ballot_sync(BH_FULL, a2 && pm >= pthr));
        }
      }
      if (lane == 0 && s0 + j < rg.end) {
        const int sid = seg_gid(c, s0 + j);
        c.seg_pot[sid] = pot;
        c.seg_conn[sid] = conn;
        nm += pot >= thr ? 1 : 0;
        if (append && pot >= thr) {  // this rank's matching segments, unordered (shard_ll.cuh)
          const int i = atomicAdd(&append[0], 1);
          if (i < c.xm_cap) {
            append[8 + 3 * i] = sid;
            append[8 + 3 * i + 1] = pot;
            append[8 + 3 * i + 2] = conn;
          }
        }
      }
    }
  }
  int tm = block_sum(nm, s_red);
  if (threadIdx.x == 0) BLK(c, BLK_MATCH)[b] = tm;
  if (c.seg_world > 1) {  // segment shards: per-CTA count of recyclable local rows for ph_shard_pack
    int tr = block_sum(nrec, s_red);
    if (threadIdx.x == 0) BLK(c, BLK_RECYC)[b] = tr;
  }
}

// Phase B: matching list in ascending segment id (np.where, projections.py:247),
// jittered potential f32(f64(potential) + u) with the draw indexed by the rank in
// that list (:234-235), per-cell maximum (:236-237) and active-segment count (:251).
// CTA 0 completes the timestep.
// `ready`: draw #3 was not run as a phase (fused.cuh); its count is the clamped list length.
// `want_jitter` = return_winner_cell (networks.py:121): without it the jitter (and its draw) is left
// to a later step (ph_fill_jitter); this phase then also publishes M.
__device__ void ph_activate_b(const bh_ctx& c, int b, int nb, bool ready = false, bool want_jitter = true,
                              int pre_before = -1, int pre_total = -1) {
  __shared__ int s_red[32];
  const int NT = blockDim.x;
  const int S = c.sc[BH_SC_NSEG];
  const int thr = c.seg_matching_threshold;
  const long long off3 = c.rng64[R_OFF3];
  int m_before = pre_before, m_total = pre_total;
  if (pre_total < 0) blk_prefix(BLK(c, BLK_MATCH), b, nb, s_red, m_before, m_total);  // else: the caller did
  const long long n3 = ready ? (m_total < c.match_capacity ? m_total : c.match_capacity) : c.rng64[R_N3];
  const Range rg = block_range(S, b, nb);
  int base = m_before;
  #pragma unroll 1
  for (int tile = rg.begin; tile < rg.end; tile += NT) {
    const int s = tile + threadIdx.x;
    const bool ok = s < rg.end;
    const int pot = ok ? c.seg_pot[s] : 0;
    const bool match = ok && pot >= thr;
    int tot;
    const int rank = base + block_excl_scan(match ? 1 : 0, s_red, tot);
    base += tot;
    if (match && rank < c.match_capacity) {
      const int conn = c.seg_conn[s];
      const int owner = c.seg_owner[s];
      c.m_seg[rank] = s;
      c.m_conn[rank] = conn;
      if (want_jitter) {
        const double u = rank < n3 ? rng_uniform(c, off3 + 2 * rank) : 0.0;
        const float jit = __double2float_rn(__dadd_rn((double)pot, u));
        c.m_jit[rank] = jit;
        atomicMax(reinterpret_cast<int*>(c.cell_maxjit + owner), __float_as_int(jit));  // jit > 0
      }
      if (conn >= c.seg_activation_threshold) {
        atomicAdd(&c.cell_npred[owner], 1);
        if (atomicOr(&c.col_pred[owner >> 5], 1u << (owner & 31)) == 0u) atomicAdd(&c.sc[BH_SC_NPREDCOL], 1);
      }
    }
  }
  if (b == 0 && threadIdx.x == 0) {
    c.sc[BH_SC_HAVE_PREV] = 1;
    c.sc[BH_SC_STEP] = c.sc[BH_SC_STEP] + 1;
    c.sc[BH_SC_JIT_PENDING] = want_jitter ? 0 : 1;
    if (!want_jitter) {
      if (m_total > c.match_capacity) atomicOr(&c.sc[BH_SC_STATUS], BH_ST_MATCH_OVERFLOW);
      c.sc[BH_SC_M] = m_total < c.match_capacity ? m_total : c.match_capacity;
    }
  }
}

// The deferred jitter of the previous activation (projections.py:229-239 reached through
// get_jittered_potential_info, networks.py:76): rand(M) drawn now, before this step's rand(k, c).
// One CTA; no-op unless BH_SC_JIT_PENDING.
// `x`: MT_RING words of shared memory.
__device__ void ph_fill_jitter(const bh_ctx& c, uint32_t* x) {
  if (!c.sc[BH_SC_JIT_PENDING] || !c.sc[BH_SC_HAVE_PREV]) return;
  const int M = c.sc[BH_SC_M];
  rng_draw(c, x, M, R_OFF3, R_N3, true, 0, false);
  const long long off3 = c.rng64[R_OFF3], n3 = c.rng64[R_N3];
#pragma unroll 1
  for (int j = threadIdx.x; j < M; j += blockDim.x) {
    const int s = c.m_seg[j];
    const double u = j < n3 ? rng_uniform(c, off3 + 2 * j) : 0.0;
    const float jit = __double2float_rn(__dadd_rn((double)c.seg_pot[s], u));
    c.m_jit[j] = jit;
    atomicMax(reinterpret_cast<int*>(c.cell_maxjit + c.seg_owner[s]), __float_as_int(jit));
  }
  __syncthreads();
  if (threadIdx.x == 0) c.sc[BH_SC_JIT_PENDING] = 0;
}
__global__ void __launch_bounds__(MT_THREADS) k_tm_fill_jitter(const __grid_constant__ bh_ctx c) {
  __shared__ uint32_t x[MT_RING];
  ph_fill_jitter(c, x);
}

// ---------------------------------------------------------------------------------
// stand-alone kernels (fine-grained C entry points)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(BH_TM_THREADS) k_tm_select_a(const __grid_constant__ bh_ctx c, int want) {
  ph_select_a(c, blockIdx.x, gridDim.x, want != 0);
}
__global__ void __launch_bounds__(BH_TM_THREADS) k_tm_select_b(const __grid_constant__ bh_ctx c, int want) {
  ph_select_b(c, blockIdx.x, gridDim.x, want != 0);
}
__global__ void __launch_bounds__(BH_TM_THREADS) k_tm_learn_select_a(const __grid_constant__ bh_ctx c, int learning) {
  ph_learn_select_a(c, learning, blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(BH_TM_THREADS) k_tm_learn_select_b(const __grid_constant__ bh_ctx c, int learning) {
  ph_learn_select_b(c, learning, blockIdx.x, gridDim.x);
}
__global__ void k_tm_post(const __grid_constant__ bh_ctx c) { ph_post(c, blockIdx.x, gridDim.x); }
__global__ void __launch_bounds__(BH_TM_THREADS) k_tm_activate_a(const __grid_constant__ bh_ctx c) { ph_activate_a(c, blockIdx.x, gridDim.x); }
__global__ void __launch_bounds__(BH_TM_THREADS) k_tm_activate_b(const __grid_constant__ bh_ctx c, int want_jitter) {
  ph_activate_b(c, blockIdx.x, gridDim.x, false, want_jitter != 0);
}

// ---------------------------------------------------------------------------------
// Stand-alone plugin calls (PredictiveProjection.update / .process with explicit
// arguments, projections.py:245-293) and TemporalMemory.process(prev_state=empty
// state): adopt the caller's lists into the device buffers the phases read.
// ---------------------------------------------------------------------------------
// Forget the previous timestep's context (networks.py:59-65, the empty state): no
// previous predictions, activation, winner cells or distal state.  Learned state
// (segments, synapses, segments per cell, SP) is untouched.  One CTA.
__global__ void __launch_bounds__(1024) k_tm_reset(const __grid_constant__ bh_ctx c) {
  const int k = c.active_columns, cd = c.cell_dim;
  const int last = (c.sc[BH_SC_STEP] & 1) ^ 1;  // buffers of the last completed step
  const int M = c.sc[BH_SC_M];
#pragma unroll 1
  for (int j = threadIdx.x; j < M; j += blockDim.x) {
    const int owner = c.seg_owner[c.m_seg[j]];
    c.cell_maxjit[owner] = 0.0f;
    c.cell_npred[owner] = 0;
    c.col_pred[owner >> 5] = 0u;
  }
  const int Wl = c.sc[BH_SC_W0 + last];
  const int* wl = c.winners + (long long)last * k * cd;
#pragma unroll 1
  for (int i = threadIdx.x; i < Wl; i += blockDim.x) c.cell_widx[wl[i]] = -1;
  if (c.sc[BH_SC_HAVE_PREV]) {
    const int* cols = c.active_cols + last * k;
    uint32_t* act = c.col_act + (long long)last * c.column_dim;
#pragma unroll 1
    for (int r = threadIdx.x; r < k; r += blockDim.x) {
      act[cols[r]] = 0u;
      c.col_win[cols[r]] = 0u;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    c.sc[BH_SC_HAVE_PREV] = 0;
    c.sc[BH_SC_M] = 0;
    c.sc[BH_SC_W0] = 0;
    c.sc[BH_SC_W1] = 0;
    c.sc[BH_SC_WNONE0] = 1;
    c.sc[BH_SC_WNONE1] = 1;
    c.sc[BH_SC_JIT_PENDING] = 0;
    c.sc[BH_SC_NU] = 0;
    c.sc[BH_SC_NPREDCOL] = 0;
  }
}

// PredictiveProjection.update's arguments (projections.py:257): learning_output = ordered winner cells,
// winner_input = ordered previous winner cells (n_prev < 0: None), input_activation as one bit-word per
// column, output_punishment as its per-column complement (1 = the column is active, not punished).
// Grid-stride part (any grid): clear-and-set of the per-column words.  The ordered lists follow in
// k_tm_adopt_lists (one CTA).
__global__ void k_tm_adopt_words(const __grid_constant__ bh_ctx c, const uint32_t* prev_act_words,
                                 const uint8_t* col_active) {
  const int cur = c.sc[BH_SC_STEP] & 1;
  uint32_t* prev_act = c.col_act + (long long)(cur ^ 1) * c.column_dim;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gsz = (long long)gridDim.x * blockDim.x;
#pragma unroll 1
  for (long long j = gid; j < c.column_dim; j += gsz) {
    prev_act[j] = prev_act_words[j];
    c.col_active[j] = col_active[j];
    c.col_win[j] = 0u;
  }
#pragma unroll 1
  for (long long j = gid; j < (long long)c.column_dim * 32; j += gsz) c.cell_widx[j] = -1;
}

__global__ void __launch_bounds__(1024) k_tm_adopt_lists(const __grid_constant__ bh_ctx c, const int* win_cells, int n_win,
                                                         const int* prev_win, int n_prev) {
  __shared__ int s_red[32];
  const int k = c.active_columns, cd = c.cell_dim;
  const int cur = c.sc[BH_SC_STEP] & 1;
  const bool have_prev = c.sc[BH_SC_HAVE_PREV] != 0;
  int* wl = c.winners + (long long)cur * k * cd;
  int* wp = c.winners + (long long)(cur ^ 1) * k * cd;
  int base = 0;
#pragma unroll 1
  for (int tile = 0; tile < n_win; tile += blockDim.x) {
    const int i = tile + threadIdx.x;
    const bool ok = i < n_win;
    const int cell = ok ? win_cells[i] : 0;
    if (ok) {
      wl[i] = cell;
      atomicOr(&c.col_win[cell >> 5], 1u << (cell & 31));
    }
    const bool un = ok && have_prev && c.cell_maxjit[cell] < c.epsilon;  // projections.py:271
    int tot;
    const int pos = base + block_excl_scan(un ? 1 : 0, s_red, tot);
    if (un) c.unacc[pos] = cell;
    base += tot;
  }
#pragma unroll 1
  for (int i = threadIdx.x; i < n_prev; i += blockDim.x) {
    wp[i] = prev_win[i];
    c.cell_widx[prev_win[i]] = i;
  }
  if (threadIdx.x == 0) {
    c.sc[BH_SC_W0 + cur] = n_win;
    c.sc[BH_SC_WNONE0 + cur] = 0;
    c.sc[BH_SC_W0 + (cur ^ 1)] = n_prev > 0 ? n_prev : 0;
    c.sc[BH_SC_WNONE0 + (cur ^ 1)] = n_prev < 0 ? 1 : 0;
    c.sc[BH_SC_NU] = base;
    c.rng64[R_STEP_BASE] = c.rng64[R_CURSOR];  // a stand-alone call starts its own draw budget
  }
}

// PredictiveProjection.process's argument (projections.py:245): the active cells as this step's
// activation words.  Also what k_tm_post would have done for a stand-alone call: reset the per-cell
// results of the previous activation (idempotent) and commit the segment count.
__global__ void k_tm_adopt_active_clear(const __grid_constant__ bh_ctx c, int have_winners) {
  const int cur = c.sc[BH_SC_STEP] & 1;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gsz = (long long)gridDim.x * blockDim.x;
  // both activation buffers: this step's is rebuilt from the argument, the previous step's has had its last
  // reader (learning) and must be empty for the step after this one
#pragma unroll 1
  for (long long j = gid; j < 2LL * c.column_dim; j += gsz) c.col_act[j] = 0u;
  if (gid == 0 && !have_winners) {  // no update() call preceded: winner_cell is None for this step
    c.sc[BH_SC_W0 + cur] = 0;
    c.sc[BH_SC_WNONE0 + cur] = 1;
    c.rng64[R_STEP_BASE] = c.rng64[R_CURSOR];
  }
  const int M = c.sc[BH_SC_M];
#pragma unroll 1
  for (long long j = gid; j < M; j += gsz) {
    const int owner = c.seg_owner[c.m_seg[j]];
    c.cell_maxjit[owner] = 0.0f;
    c.cell_npred[owner] = 0;
    c.col_pred[owner >> 5] = 0u;
  }
  if (gid == 0) {
    c.sc[BH_SC_NSEG] = c.sc[BH_SC_NSEG_NEXT];
    c.sc[BH_SC_NPREDCOL_PREV] = c.sc[BH_SC_NPREDCOL];
    c.sc[BH_SC_NPREDCOL] = 0;
  }
}
// the winner index rotation of ph_post (one CTA): entries of the previous winners out, this step's in
__global__ void __launch_bounds__(1024) k_tm_adopt_widx(const __grid_constant__ bh_ctx c) {
  const int k = c.active_columns, cd = c.cell_dim;
  const int cur = c.sc[BH_SC_STEP] & 1;
  const int Wc = c.sc[BH_SC_W0 + cur], Wp = c.sc[BH_SC_W0 + (cur ^ 1)];
  const int* wl_cur = c.winners + (long long)cur * k * cd;
  const int* wl_prev = c.winners + (long long)(cur ^ 1) * k * cd;
#pragma unroll 1
  for (int i = threadIdx.x; i < Wp; i += blockDim.x) c.cell_widx[wl_prev[i]] = -1;
  __syncthreads();
#pragma unroll 1
  for (int i = threadIdx.x; i < Wc; i += blockDim.x) c.cell_widx[wl_cur[i]] = i;
}
__global__ void k_tm_adopt_active_set(const __grid_constant__ bh_ctx c, const int* cells, int n) {
  const int cur = c.sc[BH_SC_STEP] & 1;
  uint32_t* act = c.col_act + (long long)cur * c.column_dim;
#pragma unroll 1
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    atomicOr(&act[cells[i] >> 5], 1u << (cells[i] & 31));
}

// Step summary for the host (bh_step_host): see include/bithtm_b200.h
__device__ void ph_summary(const bh_ctx& c, int b, int nb) {
  const int k = c.active_columns;
  const int done = c.sc[BH_SC_STEP] - 1;  // the step that just completed
  const int cur = done & 1;
  int* out = c.summary_dev;
  const int gid = b * blockDim.x + threadIdx.x, gsz = nb * blockDim.x;
  if (gid == 0) {
    out[0] = done;
    out[1] = c.sc[BH_SC_STATUS];
    out[2] = c.sc[BH_SC_NSEG];
    out[3] = c.sc[BH_SC_W0 + cur];
  }
  #pragma unroll 1
  for (int i = gid; i < k; i += gsz) {
    out[4 + i] = c.active_cols[cur * k + i];
    out[4 + k + i] = (int)c.row_pred[i];
    out[4 + 2 * k + i] = (int)c.row_act[i];
    out[4 + 3 * k + i] = (int)c.row_win[i];
  }
  rng_export(c, out + 4 + 4 * k, gid, gsz);  // MT19937 state at the stream cursor
  if (gid == 0) {
    int* tail = out + 4 + 4 * k + MT_N + 1;
    tail[0] = c.sc[BH_SC_NPREDCOL_PREV];
    tail[1] = c.sc[BH_SC_NPREDCOL];
    tail[2] = 0;
    tail[3] = 0;
  }
}

__global__ void k_summary(const __grid_constant__ bh_ctx c) { ph_summary(c, blockIdx.x, gridDim.x); }
