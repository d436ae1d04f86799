// Low-latency exchanges of the fused sharded step (shard_fused.cuh) and the two selections built on them.
//
// Transport.  A record travels as 8-byte cells {payload word, sequence number} stored straight into every
// peer's region (st.relaxed.sys.v2: one 8-byte store is one transaction, so a cell is valid as soon as its
// sequence number reads as this step's) -- no system fence, no grid barrier and no separate flag per
// exchange, which is what made the copy + fence + barrier + flag protocol cost 12-14 us.  The receiver spins
// on the cells themselves.  Records are double-buffered by step parity; a rank can be at most one exchange
// ahead of a peer (finishing an exchange needs every peer's record of it), so a cell is never overwritten
// before it was read.  Sequence number = step + 1 (regions start zeroed).
//
// Exchange 1 (global inhibition, regularizations.py:28-29 over column shards), ONE CTA per rank:
//   a. all-gather the 2048-bin histograms the overlap phase built with the predicted binning (sp_kernels.cuh
//      TK3_*): every rank derives the same threshold bin B and how many keys each rank has above / inside it;
//   b. all-gather, per rank, its columns in bins above B (selected for sure) and its (key, column) pairs inside
//      B; every rank ranks the pairs identically (larger key first, ties -> lower column) and writes the
//      global ascending active-column list.
//   A prediction that misses (threshold outside the binned range, > 1024 keys in B) falls back to the
//   candidate exchange of shard_fused.cuh for that step.
// Exchange 2 (matching / recyclable segments, projections.py:247, 80-81 over segment shards), ONE CTA per rank:
//   the scan appended this rank's matching segments unordered; they are sorted by id, all-gathered, and every
//   rank merges the sorted lists by counting (binary searches in shared memory).
#pragma once

#include "sp_kernels.cuh"
#include "tm_shard.cuh"

#define LL_HIST_WORDS (TK2_BINS + 8)
#define LL_MEMBERS TOPK_THREADS      // most keys of the threshold bin over all ranks
#define LL_LOCAL_MATCH_MAX 4096      // most matching segments one rank sorts in shared memory
#define LL_TOTAL_MATCH_MAX 16384     // most matching segments over all ranks merged in shared memory
#define LL_TIMEOUT_CYCLES 6000000000LL

__host__ __device__ __forceinline__ int ll_k_loc(const bh_ctx& c) {
  return c.active_columns < c.col_local ? c.active_columns : c.col_local;
}
__host__ __device__ __forceinline__ long long ll_sel_words(const bh_ctx& c) { return ll_k_loc(c) + 3LL * LL_MEMBERS + 8; }
__host__ __device__ __forceinline__ long long ll_seg_words(const bh_ctx& c) { return 8 + 3LL * c.xm_cap + c.xr_cap; }
// dynamic shared memory (bytes) the two one-CTA phases below need
__host__ __device__ __forceinline__ long long ll_smem_bytes(const bh_ctx& c) {
  const long long sel = 4LL * (TK2_BINS + 2 * LL_MEMBERS + 2 * LL_MEMBERS + 8 + c.active_columns + 2 * ((c.col_local + 31) / 32) + 64);
  const long long seg = 4LL * (LL_TOTAL_MATCH_MAX + LL_LOCAL_MATCH_MAX);
  return sel > seg ? sel : seg;
}
// ints of the LL area of one region: per kind [2 parity][G sources][words] cells of 2 ints
__host__ __device__ __forceinline__ long long ll_area_ints(const bh_ctx& c) {
  const long long G = c.seg_world > 1 ? c.seg_world : 1;
  return 4 * G * (LL_HIST_WORDS + ll_sel_words(c) + ll_seg_words(c));
}
// int offset (inside the LL area) of cell `i` of the record rank `src` sent for exchange `kind` (0 histogram,
// 1 selection, 2 segments) at step parity `par`
__device__ __forceinline__ long long ll_cell(const bh_ctx& c, int kind, int par, int src, long long i) {
  const long long G = c.seg_world > 1 ? c.seg_world : 1;
  const long long w0 = LL_HIST_WORDS, w1 = ll_sel_words(c), w2 = ll_seg_words(c);
  const long long base = kind == 0 ? 0 : (kind == 1 ? 2 * G * w0 : 2 * G * (w0 + w1));
  const long long w = kind == 0 ? w0 : (kind == 1 ? w1 : w2);
  return 2 * (base + ((long long)par * G + src) * w + i);
}

__device__ __forceinline__ void ll_store(int* cell, int v, int seq) {
  asm volatile("st.relaxed.sys.global.v2.s32 [%0], {%1, %2};" ::"l"(cell), "r"(v), "r"(seq) : "memory");
}
// spins until the cell carries `seq`; a peer that never answers sets BH_ST_XCH_TIMEOUT instead of hanging
__device__ __forceinline__ int ll_load(const bh_ctx& c, const int* cell, int seq) {
  int v, s;
  long long t0 = 0;
  for (;;) {
    asm volatile("ld.relaxed.sys.global.v2.s32 {%0, %1}, [%2];" : "=r"(v), "=r"(s) : "l"(cell) : "memory");
    if (s == seq) return v;
    if (t0 == 0) t0 = clock64();
    else if (clock64() - t0 > LL_TIMEOUT_CYCLES) {
      atomicOr(&c.sc[BH_SC_STATUS], BH_ST_XCH_TIMEOUT);
      return 0;
    }
  }
}
// The same cell index in the records of all G source ranks at once (independent loads in flight; spins only on
// the cells that have not arrived): v[g] for g < G.
__device__ __forceinline__ void ll_load_ranks(const bh_ctx& c, const int* mine, int kind, int par, long long i, int seq, int G,
                                              int (&v)[BH_MAX_RANKS]) {
  unsigned pending = (1u << G) - 1u;
  long long t0 = 0;
  while (pending) {
#pragma unroll
    for (int g = 0; g < BH_MAX_RANKS; ++g) {
      if (pending & (1u << g)) {
        int x, sq;
        asm volatile("ld.relaxed.sys.global.v2.s32 {%0, %1}, [%2];" : "=r"(x), "=r"(sq) : "l"(mine + ll_cell(c, kind, par, g, i)) : "memory");
        if (sq == seq) {
          v[g] = x;
          pending &= ~(1u << g);
        }
      }
    }
    if (pending) {
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > LL_TIMEOUT_CYCLES) {
        atomicOr(&c.sc[BH_SC_STATUS], BH_ST_XCH_TIMEOUT);
#pragma unroll
        for (int g = 0; g < BH_MAX_RANKS; ++g)
          if (pending & (1u << g)) v[g] = 0;
        return;
      }
    }
  }
}

// word i of this rank's record -> every rank (its own region included: one code path)
__device__ __forceinline__ void ll_put(const bh_ctx& c, int* const* ll, int kind, int par, long long i, int v, int seq) {
  const int G = c.seg_world > 1 ? c.seg_world : 1;
  const long long off = ll_cell(c, kind, par, c.seg_rank, i);
#pragma unroll 1
  for (int p = 0; p < G; ++p) ll_store(ll[p] + off, v, seq);
}

// globaltimer stamps of the one-CTA phases (ctx.blk row 7 as u64, from index 40 / 52): tools read them
#define LL_STAMP(base, i)                                                                    \
  do {                                                                                       \
    if (threadIdx.x == 0) {                                                                  \
      unsigned long long t_;                                                                 \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                 \
      reinterpret_cast<unsigned long long*>(c.blk + 7 * BH_BLK_STRIDE)[(base) + (i)] = t_;   \
    }                                                                                        \
  } while (0)

// lower bound in a shared-memory ascending int list
__device__ __forceinline__ int smem_lower(const int* list, int n, int key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (list[mid] < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// ------------------------------------------------------------------------------------
// Exchange 1.  ONE CTA (blockDim.x = 1024).  `ll[p]` = LL area of rank p's region.  Returns false (on every
// rank alike) when the step must take the candidate exchange instead; then nothing was written but the flag.
// smem: 2048 + 64 + 3 * LL_MEMBERS + 2 * LL_MEMBERS (u64 keys) + active_columns + k_loc + 64 ints.
// ------------------------------------------------------------------------------------
// `step_ov` / `out_ov` (two-pipeline kernel): the step this selection belongs to when it runs ahead of the step
// counter, and a staging list instead of active_cols -- then no column flag is touched here.
__device__ __noinline__ bool ph_shard_select_ll(const bh_ctx& c, int* const* ll, uint32_t* smem, int step_ov = -1,
                                                int* out_ov = nullptr) {
  __shared__ int s_scan[32];
  __shared__ int s_bin, s_rem, s_ncand, s_cnt[4 * BH_MAX_RANKS];
  __shared__ unsigned long long s_kth_key, s_gmax;
  const int t = threadIdx.x, NT = blockDim.x, lane = t & 31;
  const int G = c.seg_world > 1 ? c.seg_world : 1, me = c.seg_rank;
  const int k = c.active_columns, k_loc = ll_k_loc(c), n = c.col_local;
  const int step = step_ov >= 0 ? step_ov : c.sc[BH_SC_STEP], par = step & 1, seq = step + 1;
  int* ws3 = c.topk_ws + TK3_BASE;
  int* ws = c.topk_ws + TK2_BASE;
  int* ghist = ws3 + TK3_HIST + par * TK2_BINS;
  int* s_h = reinterpret_cast<int*>(smem);                       // [2048] summed histogram, descending bin order
  unsigned long long* s_mkey = reinterpret_cast<unsigned long long*>(s_h + TK2_BINS);  // [LL_MEMBERS]
  int* s_mcol = reinterpret_cast<int*>(s_mkey + LL_MEMBERS);      // [LL_MEMBERS] global column of a member
  int* s_msel = s_mcol + LL_MEMBERS;                              // [LL_MEMBERS] selected?  then: exclusive prefix
  int* s_above = s_msel + LL_MEMBERS + 1;                         // [k] columns above the bin, rank by rank
  const unsigned long long* keys = reinterpret_cast<const unsigned long long*>(c.boosted);
  unsigned long long* w64 = reinterpret_cast<unsigned long long*>(ws);
  const bool have_hist = ws3[TK3_READY] == seq;

  LL_STAMP(40, 0);
  if (!out_ov) retire_prev_flags(c);  // (independent of the selection; the new flags are set at the end)
  // a. histogram (+ the largest local key) -> everybody
  const unsigned long long lmax = w64[0];
#pragma unroll 1
  for (int i = t; i < TK2_BINS + 3; i += NT) {
    int v;
    if (i < TK2_BINS) v = have_hist ? ghist[i] : 0;
    else v = i == TK2_BINS ? (int)(unsigned)(lmax & 0xffffffffull) : (i == TK2_BINS + 1 ? (int)(unsigned)(lmax >> 32) : (have_hist ? 1 : 0));
    ll_put(c, ll, 0, par, i, v, seq);
  }
  LL_STAMP(40, 1);
  int* mine = ll[me];
  __shared__ int s_meta[3 * BH_MAX_RANKS];
  if (t < 3 * G) s_meta[t] = ll_load(c, mine + ll_cell(c, 0, par, t / 3, TK2_BINS + t % 3), seq);
  // every thread owns two bins (descending order: slot j holds bin 2047 - j) of every rank's histogram and keeps
  // them in registers: the per-rank counts above / inside the threshold bin need no second read
  int hv[2][BH_MAX_RANKS];
  int my_sum = 0;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int bin = TK2_BINS - 1 - (2 * t + i);
#pragma unroll
    for (int g = 0; g < BH_MAX_RANKS; ++g) hv[i][g] = 0;
    if (bin >= 1) ll_load_ranks(c, mine, 0, par, bin, seq, G, hv[i]);
    int sum = 0;
#pragma unroll
    for (int g = 0; g < BH_MAX_RANKS; ++g) sum += hv[i][g];
    s_h[2 * t + i] = sum;
    my_sum += sum;
  }
  if (t < 4 * BH_MAX_RANKS) s_cnt[t] = 0;
  if (t == 0) s_ncand = -1;
  __syncthreads();
  LL_STAMP(40, 2);
  if (t == 0) {
    unsigned long long gm = 0ull;
    int all_hist = 1;
    for (int g = 0; g < G; ++g) {
      const unsigned long long m = ((unsigned long long)(unsigned)s_meta[3 * g + 1] << 32) | (unsigned)s_meta[3 * g];
      all_hist &= s_meta[3 * g + 2];
      gm = m > gm ? m : gm;
    }
    s_gmax = gm;
    if (!all_hist) s_ncand = -2;
  }
  __syncthreads();
  bool ok = s_ncand == -1;
  if (ok) {
    int total;
    const int before = block_excl_scan(my_sum, s_scan, total);
    if (before < k && before + my_sum >= k) {
      int r = k - before;
      for (int i = 0; i < 2; ++i) {
        const int h = s_h[2 * t + i];
        if (r > 0 && h >= r) {
          s_bin = TK2_BINS - 1 - (2 * t + i);
          s_rem = r;
          s_ncand = h;
          r = -1;
        } else if (r > 0) {
          r -= h;
        }
      }
    }
    __syncthreads();
    ok = s_ncand >= 0 && s_ncand <= LL_MEMBERS;
  }
  if (!ok) return false;  // identical decision on every rank (same gathered data)
  const int bin = s_bin, rem = s_rem;
  // per rank: keys above the bin (a_g) and inside it (m_g), from the registers
  {
    int a[BH_MAX_RANKS];
#pragma unroll
    for (int g = 0; g < BH_MAX_RANKS; ++g) a[g] = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int b2 = TK2_BINS - 1 - (2 * t + i);
#pragma unroll
      for (int g = 0; g < BH_MAX_RANKS; ++g) {
        if (b2 > bin) a[g] += hv[i][g];
        if (b2 == bin) s_cnt[BH_MAX_RANKS + g] = hv[i][g];
      }
    }
#pragma unroll
    for (int g = 0; g < BH_MAX_RANKS; ++g) {
      const int w = warp_sum(a[g]);
      if (lane == 0 && w) atomicAdd(&s_cnt[g], w);
    }
  }
  __syncthreads();
  LL_STAMP(40, 3);
  // b. this rank's columns above the bin and its members of the bin, in ascending column order -> everybody.
  // Pass 1 reads the keys coalesced (a warp covers 32 consecutive columns: its ballots ARE the class masks of
  // that group); pass 2: thread t owns consecutive mask words, one block scan per list orders everything, and
  // only the members' keys are read a second time.
  const Tk3Binning binning = tk3_binning(ws3);
  {
    uint32_t* s_ma = reinterpret_cast<uint32_t*>(s_above + k);  // [n / 32] columns above the bin
    const int n_words = (n + 31) >> 5;
    uint32_t* s_mm = s_ma + n_words;                             // [n / 32] members of the bin
    const int n32 = (n + 31) & ~31;
#pragma unroll 1
    for (int j0 = t; j0 < n32; j0 += 8 * NT) {  // 8 independent key loads in flight per thread
      unsigned long long kv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) kv[u] = j0 + u * NT < n ? keys[j0 + u * NT] : 0ull;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = j0 + u * NT;
        if (j < n32) {  // (uniform in the warp)
          const int kb = j < n ? tk3_bin(binning, kv[u]) : 0;
          const uint32_t a = __ballot_sync(BH_FULL, kb > bin), m = __ballot_sync(BH_FULL, j < n && kb == bin);
          if (lane == 0) {
            s_ma[j >> 5] = a;
            s_mm[j >> 5] = m;
          }
        }
      }
    }
    __syncthreads();
    const int wpt = (n_words + NT - 1) / NT;
    const int w0 = t * wpt, w1 = w0 + wpt < n_words ? w0 + wpt : n_words;
    int na = 0, nm = 0;
    for (int w = w0; w < w1; ++w) {
      na += __popc(s_ma[w]);
      nm += __popc(s_mm[w]);
    }
    int tot_a, tot_m;
    int pa = block_excl_scan(na, s_scan, tot_a);
    int pm = block_excl_scan(nm, s_scan, tot_m);
    for (int w = w0; w < w1; ++w) {
      uint32_t a = s_ma[w], m = s_mm[w];
      while (a) {
        const int i = __ffs(a) - 1;
        a &= a - 1;
        if (pa < k_loc) ll_put(c, ll, 1, par, pa, c.col_lo + 32 * w + i, seq);
        ++pa;
      }
      while (m) {
        const int i = __ffs(m) - 1;
        m &= m - 1;
        const int j = 32 * w + i;
        if (pm < LL_MEMBERS) {
          const unsigned long long key = keys[j];
          const long long wd = k_loc + 3LL * pm;
          ll_put(c, ll, 1, par, wd, (int)(unsigned)(key & 0xffffffffull), seq);
          ll_put(c, ll, 1, par, wd + 1, (int)(unsigned)(key >> 32), seq);
          ll_put(c, ll, 1, par, wd + 2, c.col_lo + j, seq);
        }
        ++pm;
      }
    }
  }
  LL_STAMP(40, 4);
  // gather: offsets of every rank's part
  if (t == 0) {
    int oa = 0, om = 0;
    for (int g = 0; g < G; ++g) {
      s_cnt[2 * BH_MAX_RANKS + g] = oa;
      s_cnt[3 * BH_MAX_RANKS + g] = om;
      oa += s_cnt[g];
      om += s_cnt[BH_MAX_RANKS + g];
    }
  }
  __syncthreads();
  int n_above = 0, n_mem = 0;
  for (int g = 0; g < G; ++g) {
    n_above += s_cnt[g];
    n_mem += s_cnt[BH_MAX_RANKS + g];
  }
  // (flat over all ranks' entries: the loads of different ranks are in flight together)
#pragma unroll 1
  for (int e = t; e < n_above; e += NT) {
    int g = 0;
    while (g + 1 < G && e >= s_cnt[2 * BH_MAX_RANKS + g + 1]) ++g;
    s_above[e] = ll_load(c, mine + ll_cell(c, 1, par, g, e - s_cnt[2 * BH_MAX_RANKS + g]), seq);
  }
#pragma unroll 1
  for (int e = t; e < n_mem; e += NT) {
    int g = 0;
    while (g + 1 < G && e >= s_cnt[3 * BH_MAX_RANKS + g + 1]) ++g;
    const long long w = k_loc + 3LL * (e - s_cnt[3 * BH_MAX_RANKS + g]);
    const unsigned lo = (unsigned)ll_load(c, mine + ll_cell(c, 1, par, g, w), seq);
    const unsigned hi = (unsigned)ll_load(c, mine + ll_cell(c, 1, par, g, w + 1), seq);
    s_mkey[e] = ((unsigned long long)hi << 32) | lo;
    s_mcol[e] = ll_load(c, mine + ll_cell(c, 1, par, g, w + 2), seq);
  }
  __syncthreads();
  LL_STAMP(40, 5);
  // rank the members: larger key first, ties -> lower column; the first `rem` are selected
  if (t < n_mem) {
    const unsigned long long mk = s_mkey[t];
    const int mc = s_mcol[t];
    int ahead = 0;
#pragma unroll 2
    for (int i = 0; i < n_mem; ++i) {
      const unsigned long long ok2 = s_mkey[i];
      ahead += (ok2 > mk || (ok2 == mk && s_mcol[i] < mc)) ? 1 : 0;
    }
    s_msel[t] = ahead < rem ? 1 : 0;
    if (ahead == rem - 1) s_kth_key = mk;
  }
  __syncthreads();
  {  // exclusive prefix of the selected flags over the member array (n_mem <= 1024 = one tile)
    int tot;
    const int f = t < n_mem ? s_msel[t] : 0;
    const int pre = block_excl_scan(f, s_scan, tot);
    __syncthreads();
    if (t < n_mem) s_msel[t] = pre | (f << 30);
    if (t == 0) s_msel[n_mem] = tot;
  }
  __syncthreads();
  int* out = out_ov ? out_ov : c.active_cols + par * k;
  const bool set_flags = out_ov == nullptr;
  // rank offsets in the final list: everything of the ranks before
  // (selected members of rank g = prefix at the end of its part - prefix at its start)
  auto pre_at = [&](int idx) { return idx < n_mem ? (s_msel[idx] & 0x3fffffff) : s_msel[n_mem]; };
#pragma unroll 1
  for (int g = 0; g < G; ++g) {
    const int a = s_cnt[g], m = s_cnt[BH_MAX_RANKS + g], oa = s_cnt[2 * BH_MAX_RANKS + g], om = s_cnt[3 * BH_MAX_RANKS + g];
    const int off = oa + pre_at(om);  // columns above the bin + selected members of the ranks before g
#pragma unroll 1
    for (int i = t; i < a; i += NT) {  // a column above the bin: after the selected members of g with lower columns
      const int col = s_above[oa + i];
      const int lb = smem_lower(s_mcol + om, m, col);
      const int pos = off + i + (pre_at(om + lb) - pre_at(om));
      if (pos < k) {
        out[pos] = col;
        if (set_flags) c.col_active[col] = 1;
      }
    }
#pragma unroll 1
    for (int i = t; i < m; i += NT) {
      if (!(s_msel[om + i] >> 30)) continue;
      const int col = s_mcol[om + i];
      const int pos = off + smem_lower(s_above + oa, a, col) + (pre_at(om + i) - pre_at(om));
      if (pos < k) {
        out[pos] = col;
        if (set_flags) c.col_active[col] = 1;
      }
    }
  }
  // leave the workspaces clean; binning of the next step from this step's k-th and largest key
  int* oh = ws3 + TK3_HIST + (par ^ 1) * TK2_BINS;
#pragma unroll 1
  for (int i = t; i < TK2_BINS; i += NT) oh[i] = 0;
  if (t == 0) {
    w64[0] = 0ull;
    w64[1] = 0ull;
    ws[TK2_VALID] = 0;
    tk3_set_binning(ws3, s_kth_key, s_gmax > s_kth_key ? s_gmax : s_kth_key);
  }
  (void)n_above;
  LL_STAMP(40, 6);
  return true;
}

// Ascending positions of n DISTINCT ids (all < S, held in shared memory) without comparing every pair: ids are
// bucketed by value into blockDim.x buckets (segment ids of a rank are spread evenly over [0, S): a few per
// bucket), grouped by bucket with shared-memory counters, and ranked inside their bucket.  emit(i, pos) is called
// once per entry.  s_grp: n ints, s_cnt: 2 * blockDim.x ints, s_scan: 32 ints of scratch.  One CTA, all threads.
// (Counting smaller ids over the whole list -- the first version -- is n^2 / 32 warp instructions: 5.5 us for
// the 657 matching segments of a rank at 2 shards, more for its recyclable ids.)
template <typename F>
__device__ __forceinline__ void ll_sort_emit(const int* s_id, int n, int S, int* s_grp, int* s_cnt, int* s_scan, F emit) {
  const int t = threadIdx.x, NT = blockDim.x;
  int* cnt = s_cnt;
  int* start = s_cnt + NT;
  const long long scale = S > 0 ? S : 1;
  cnt[t] = 0;
  __syncthreads();
#pragma unroll 1
  for (int i = t; i < n; i += NT) {
    const long long b = (long long)s_id[i] * NT / scale;
    atomicAdd(&cnt[b < NT ? (int)b : NT - 1], 1);
  }
  __syncthreads();
  int tot;
  const int st = block_excl_scan(cnt[t], s_scan, tot);
  __syncthreads();
  start[t] = st;
  cnt[t] = 0;
  __syncthreads();
#pragma unroll 1
  for (int i = t; i < n; i += NT) {
    const int id = s_id[i];
    const long long b0 = (long long)id * NT / scale;
    const int b = b0 < NT ? (int)b0 : NT - 1;
    s_grp[start[b] + atomicAdd(&cnt[b], 1)] = id;
  }
  __syncthreads();
#pragma unroll 1
  for (int i = t; i < n; i += NT) {
    const int id = s_id[i];
    const long long b0 = (long long)id * NT / scale;
    const int b = b0 < NT ? (int)b0 : NT - 1;
    int pos = start[b];
    const int e = start[b] + cnt[b];
    for (int j = start[b]; j < e; ++j) pos += s_grp[j] < id ? 1 : 0;
    emit(i, pos);
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------
// Exchange 2.  ONE CTA.  The scan (ph_activate_a) appended this rank's matching segments as (id, potential,
// connected) triples at rec[8 + 3 i] with the counter rec[0], and its recyclable segment ids at
// rec[8 + 3 xm_cap + j] with the counter rec[1] (entries beyond the capacities are not stored) -- unordered.
// `sorted`: the record was rebuilt in ascending id order (ph_shard_pack; rec = the packed record, header
// [count, recyclable sent, recyclable true count, status]).
// smem: LL_TOTAL_MATCH_MAX + LL_LOCAL_MATCH_MAX ints.
// ------------------------------------------------------------------------------------
__device__ __noinline__ void ph_shard_segs_ll(const bh_ctx& c, int* const* ll, uint32_t* smem, int* rec, bool sorted) {
  __shared__ int s_n[4 * BH_MAX_RANKS + 4];
  const int t = threadIdx.x, NT = blockDim.x;
  const int G = c.seg_world > 1 ? c.seg_world : 1, me = c.seg_rank;
  const int step = c.sc[BH_SC_STEP], par = step & 1, seq = step + 1;
  int* s_all = reinterpret_cast<int*>(smem);            // [LL_TOTAL_MATCH_MAX] ids of all ranks, rank by rank
  int* s_key = s_all + LL_TOTAL_MATCH_MAX;              // [LL_LOCAL_MATCH_MAX] local ids (unsorted)
  int status = c.sc[BH_SC_STATUS];
  int n, nr, nr_total;
  const int* trip;
  const int* rids;
  if (sorted) {
    n = rec[0];
    nr = rec[1];
    nr_total = rec[2];
    status |= rec[3];
    trip = nullptr;
    rids = rec + 4 + 3 * c.xm_cap;
  } else {
    n = rec[0];
    nr_total = rec[1];
    if (n > c.xm_cap || n > LL_LOCAL_MATCH_MAX) {
      status |= BH_ST_XCH_OVERFLOW;
      n = c.xm_cap < LL_LOCAL_MATCH_MAX ? c.xm_cap : LL_LOCAL_MATCH_MAX;
    }
    nr = nr_total < c.xr_cap ? nr_total : c.xr_cap;
    trip = rec + 8;
    rids = rec + 8 + 3 * c.xm_cap;
  }
  LL_STAMP(52, 0);
  // header
  if (t < 4) ll_put(c, ll, 2, par, t, t == 0 ? n : (t == 1 ? nr : (t == 2 ? nr_total : status)), seq);
  if (sorted) {  // packed record: ids / potentials / connected counts as three arrays
#pragma unroll 1
    for (int i = t; i < n; i += NT) {
      ll_put(c, ll, 2, par, 8 + 3LL * i, rec[4 + i], seq);
      ll_put(c, ll, 2, par, 8 + 3LL * i + 1, rec[4 + c.xm_cap + i], seq);
      ll_put(c, ll, 2, par, 8 + 3LL * i + 2, rec[4 + 2 * c.xm_cap + i], seq);
    }
#pragma unroll 1
    for (int i = t; i < nr; i += NT) ll_put(c, ll, 2, par, 8 + 3LL * c.xm_cap + i, rids[i], seq);
  } else {
    // sort by id (ids are distinct): entry i goes to position #{ids < id_i}  (scratch: the area the gathered ids
    // will occupy afterwards)
    __shared__ int s_sort_scan[32];
    const int S_ids = c.sc[BH_SC_NSEG_NEXT];
#pragma unroll 1
    for (int i = t; i < n; i += NT) s_key[i] = trip[3 * i];
    __syncthreads();
    ll_sort_emit(s_key, n, S_ids, s_all, s_all + LL_LOCAL_MATCH_MAX, s_sort_scan, [&](int i, int pos) {
      ll_put(c, ll, 2, par, 8 + 3LL * pos, s_key[i], seq);
      ll_put(c, ll, 2, par, 8 + 3LL * pos + 1, trip[3 * i + 1], seq);
      ll_put(c, ll, 2, par, 8 + 3LL * pos + 2, trip[3 * i + 2], seq);
    });
#pragma unroll 1
    for (int i = t; i < nr; i += NT) s_key[i] = rids[i];
    __syncthreads();
    ll_sort_emit(s_key, nr, S_ids, s_all, s_all + LL_LOCAL_MATCH_MAX, s_sort_scan,
                 [&](int i, int pos) { ll_put(c, ll, 2, par, 8 + 3LL * c.xm_cap + pos, s_key[i], seq); });
  }
  LL_STAMP(52, 1);
  // gather the headers
  int* mine = ll[me];
  if (t < 4 * G) s_n[(t & 3) * BH_MAX_RANKS + (t >> 2)] = ll_load(c, mine + ll_cell(c, 2, par, t >> 2, t & 3), seq);
  __syncthreads();
  int M = 0, R = 0, RT = 0, st = 0;
  for (int g = 0; g < G; ++g) {
    M += s_n[g];
    R += s_n[BH_MAX_RANKS + g];
    RT += s_n[2 * BH_MAX_RANKS + g];
    st |= s_n[3 * BH_MAX_RANKS + g];
  }
  LL_STAMP(52, 2);
  if (M > LL_TOTAL_MATCH_MAX) st |= BH_ST_XCH_OVERFLOW;
  // ids of all ranks into shared memory (rank by rank, each ascending), then merge by counting -- flat over all
  // entries, so that the loads of different ranks are in flight together
  __shared__ int s_off[2 * BH_MAX_RANKS + 2];
  if (t == 0) {
    int o = 0, o2 = 0;
    for (int g = 0; g < G; ++g) {
      s_off[g] = o;
      s_off[BH_MAX_RANKS + 1 + g] = o2;
      o += s_n[g];
      o2 += s_n[BH_MAX_RANKS + g];
    }
    s_off[G] = o;
    s_off[BH_MAX_RANKS + 1 + G] = o2;
  }
  __syncthreads();
  const int Mc = M < LL_TOTAL_MATCH_MAX ? M : LL_TOTAL_MATCH_MAX;
#pragma unroll 1
  for (int e = t; e < Mc; e += NT) {
    int g = 0;
    while (g + 1 < G && e >= s_off[g + 1]) ++g;
    s_all[e] = ll_load(c, mine + ll_cell(c, 2, par, g, 8 + 3LL * (e - s_off[g])), seq);
  }
  __syncthreads();
#pragma unroll 1
  for (int e = t; e < Mc; e += NT) {
    int g = 0;
    while (g + 1 < G && e >= s_off[g + 1]) ++g;
    const int i = e - s_off[g], id = s_all[e];
    const int pot = ll_load(c, mine + ll_cell(c, 2, par, g, 8 + 3LL * i + 1), seq);
    const int conn = ll_load(c, mine + ll_cell(c, 2, par, g, 8 + 3LL * i + 2), seq);
    int pos = i;
    for (int o = 0; o < G; ++o) {
      const int lo = s_off[o], hi = s_off[o + 1] < Mc ? s_off[o + 1] : Mc;
      if (o != g && hi > lo) pos += smem_lower(s_all + lo, hi - lo, id);
    }
    c.seg_pot[id] = pot;  // every rank knows the potentials of the matching segments
    c.seg_conn[id] = conn;
    if (pos < c.match_capacity) {
      c.m_seg[pos] = id;
      c.m_conn[pos] = conn;
    }
  }
  __syncthreads();
  // recyclable ids: the same merge (few; usually none)
  const int* r_off = s_off + BH_MAX_RANKS + 1;
  const int Rc = R < LL_TOTAL_MATCH_MAX ? R : LL_TOTAL_MATCH_MAX;
#pragma unroll 1
  for (int e = t; e < Rc; e += NT) {
    int g = 0;
    while (g + 1 < G && e >= r_off[g + 1]) ++g;
    s_all[e] = ll_load(c, mine + ll_cell(c, 2, par, g, 8 + 3LL * c.xm_cap + (e - r_off[g])), seq);
  }
  __syncthreads();
#pragma unroll 1
  for (int e = t; e < Rc; e += NT) {
    int g = 0;
    while (g + 1 < G && e >= r_off[g + 1]) ++g;
    const int id = s_all[e];
    int pos = e - r_off[g];
    for (int o = 0; o < G; ++o) {
      const int lo = r_off[o], hi = r_off[o + 1] < Rc ? r_off[o + 1] : Rc;
      if (o != g && hi > lo) pos += smem_lower(s_all + lo, hi - lo, id);
    }
    c.recyc_list[pos] = id;
  }
  LL_STAMP(52, 3);
  if (t == 0) {
    c.sc[BH_SC_X_MATCH] = M;
    c.sc[BH_SC_X_RECYC_AVAIL] = R;
    c.sc[BH_SC_X_RECYC_TOTAL] = RT;
    if (st) atomicOr(&c.sc[BH_SC_STATUS], st);
    if (!sorted) {  // counters of the scan's append lists, for the next step
      rec[0] = 0;
      rec[1] = 0;
    }
  }
}
