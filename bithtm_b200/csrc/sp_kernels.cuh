// Spatial-pooler kernels: overlap (bit-packed popcount GEMV), boosting, canonical
// top-k inhibition, float64 permanence learning with mask re-pack, duty-cycle EMA.
// Reference: bithtm/projections.py:6-24, bithtm/regularizations.py:4-29,
// bithtm/networks.py:26-35.
#pragma once

#include "common.cuh"
#include "np_expf.h"

#define SP_THREADS 256

// ---------------------------------------------------------------------------------
// Connected mask from the float64 permanence (one warp per 32 consecutive inputs).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(SP_THREADS) k_sp_build_mask(const bh_ctx c) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = SP_THREADS / 32;
  const long long n_words = (long long)c.column_dim * c.mask_stride;
  for (long long w = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); w < n_words;
       w += (long long)gridDim.x * warps_per_block) {
    int row = (int)(w / c.mask_stride);
    int wi = (int)(w - (long long)row * c.mask_stride);
    int i = wi * 32 + lane;
    bool on = false;
    if (i < c.input_dim) on = c.sp_perm[(long long)row * c.input_dim + i] >= c.sp_threshold;
    uint32_t bits = __ballot_sync(BH_FULL, on);
    if (lane == 0) c.sp_mask[w] = bits;
  }
}

// bool bytes -> packed words
__global__ void k_pack_input(const bh_ctx c, const uint8_t* __restrict__ src, uint32_t* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= c.input_words) return;
  int i = w * 32 + lane;
  bool on = i < c.input_dim && src[i] != 0;
  uint32_t bits = __ballot_sync(BH_FULL, on);
  if (lane == 0) dst[w] = bits;
}

// ---------------------------------------------------------------------------------
// (a) overlap + (c) boost.  G lanes cooperate on one mask row with 128-bit loads
// (G = min(32, mask_stride/4) rounded to a power of two), so a warp covers 32/G
// rows and every load instruction is a full coalesced 16 B per lane.
// projections.py:18-21, regularizations.py:15-17.
// ---------------------------------------------------------------------------------
template <bool BOOST>
__global__ void __launch_bounds__(SP_THREADS) k_sp_overlap(const bh_ctx c, const uint32_t* __restrict__ input,
                                                          int group) {
  extern __shared__ __align__(16) uint32_t s_in[];  // mask_stride words (zero padded)
  for (int i = threadIdx.x; i < c.mask_stride; i += blockDim.x) s_in[i] = i < c.input_words ? input[i] : 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int sub = lane % group;               // lane within its row group
  const int rows_per_warp = 32 / group;
  const int warp_global = blockIdx.x * (SP_THREADS / 32) + (threadIdx.x >> 5);
  const int n_warps = gridDim.x * (SP_THREADS / 32);
  const int vec_per_row = c.mask_stride / 4;
  const uint4* __restrict__ mask4 = reinterpret_cast<const uint4*>(c.sp_mask);
  const uint4* s_in4 = reinterpret_cast<const uint4*>(s_in);
  for (int base = warp_global * rows_per_warp; base < c.column_dim; base += n_warps * rows_per_warp) {
    int row = base + lane / group;
    int acc = 0;
    if (row < c.column_dim) {
      const uint4* mrow = mask4 + (long long)row * vec_per_row;
      for (int v = sub; v < vec_per_row; v += group) {
        uint4 m = __ldg(mrow + v);
        uint4 x = s_in4[v];
        acc += __popc(m.x & x.x) + __popc(m.y & x.y) + __popc(m.z & x.z) + __popc(m.w & x.w);
      }
    }
    for (int o = group >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(BH_FULL, acc, o);
    if (sub == 0 && row < c.column_dim) {
      c.overlaps[row] = acc;
      if (BOOST) {
        float f = bh_np_expf(__fmul_rn(c.boost_coef, c.duty[row]));
        c.boosted[row] = __dmul_rn((double)f, (double)acc);
      }
    }
  }
}

__global__ void k_boost(const bh_ctx c) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= c.column_dim) return;
  float f = bh_np_expf(__fmul_rn(c.boost_coef, c.duty[j]));
  c.boosted[j] = __dmul_rn((double)f, (double)c.overlaps[j]);
}

// ---------------------------------------------------------------------------------
// (b) canonical global inhibition: k largest keys, ties -> lower column index,
// output ascending.  Keys are non-negative doubles, so their bit patterns order
// like unsigned integers: MSB-first radix select (8 x 8 bits) for the k-th key,
// then an ordered compaction.  Single CTA of 1024 threads.
// regularizations.py:28-29 (np.argpartition's tie-break/order are undefined).
// ---------------------------------------------------------------------------------
#define TOPK_THREADS 1024

__global__ void __launch_bounds__(TOPK_THREADS) k_topk(const bh_ctx c) {
  __shared__ int hist[256];
  __shared__ int s_scan[32];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_remaining;
  const int t = threadIdx.x;
  const int C = c.column_dim, k = c.active_columns;
  const int cur = c.sc[BH_SC_STEP] & 1;
  int* out = c.active_cols + cur * k;
  const int* prev = c.active_cols + (cur ^ 1) * k;
  const unsigned long long* keys = reinterpret_cast<const unsigned long long*>(c.boosted);

  // retire the previous step's column flags
  for (int i = t; i < k; i += TOPK_THREADS) c.col_active[prev[i]] = 0;
  if (t == 0) { s_prefix = 0ull; s_remaining = k; }
  __syncthreads();

  for (int pass = 7; pass >= 0; --pass) {
    for (int i = t; i < 256; i += TOPK_THREADS) hist[i] = 0;
    __syncthreads();
    const unsigned long long prefix = s_prefix;
    const int shift = pass * 8;
    const unsigned long long hi_mask = pass == 7 ? 0ull : (~0ull << (shift + 8));
    for (int j = t; j < C; j += TOPK_THREADS) {
      unsigned long long key = keys[j];
      if ((key & hi_mask) == prefix) atomicAdd(&hist[(int)((key >> shift) & 0xff)], 1);
    }
    __syncthreads();
    if (t == 0) {
      int rem = s_remaining, d = 255;
      for (; d > 0; --d) {
        if (hist[d] >= rem) break;
        rem -= hist[d];
      }
      s_remaining = rem;  // still needed from bin d (d == 0 takes whatever is left)
      s_prefix = prefix | ((unsigned long long)d << shift);
    }
    __syncthreads();
  }
  const unsigned long long kth = s_prefix;  // value of the k-th largest key
  const int ties_wanted = s_remaining;      // how many keys == kth to take (lowest index first)
  __syncthreads();

  // ordered compaction over column index, tiles of 1024
  int base_sel = 0, base_tie = 0;
  for (int tile = 0; tile < C; tile += TOPK_THREADS) {
    int j = tile + t;
    unsigned long long key = j < C ? keys[j] : 0ull;
    bool gt = j < C && key > kth;
    bool eq = j < C && key == kth;
    int tie_total, sel_total;
    int tie_rank = base_tie + block_excl_scan(eq ? 1 : 0, s_scan, tie_total);
    bool take = gt || (eq && tie_rank < ties_wanted);
    int pos = base_sel + block_excl_scan(take ? 1 : 0, s_scan, sel_total);
    if (take && pos < k) {
      out[pos] = j;
      c.col_active[j] = 1;
    }
    base_sel += sel_total;
    base_tie += tie_total;
  }
}

// host-inhibition mode: adopt an explicit ordered list
__global__ void k_set_active(const bh_ctx c, const int32_t* __restrict__ cols) {
  const int k = c.active_columns;
  const int cur = c.sc[BH_SC_STEP] & 1;
  int* out = c.active_cols + cur * k;
  const int* prev = c.active_cols + (cur ^ 1) * k;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    c.col_active[prev[i]] = 0;
    c.col_active[out[i]] = 0;  // in case bh_inhibit already ran this step
  }
  __syncthreads();
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    int j = cols[i];
    out[i] = j;
    c.col_active[j] = 1;
  }
}

// ---------------------------------------------------------------------------------
// (d) SP learning: permanence[active] += input ? d_on : d_off in float64 (no clip),
// and the connected-mask rows of the touched columns are re-packed in the same
// pass.  One CTA per active column; a warp handles 32 consecutive inputs so the
// 256-byte permanence segment is coalesced and the ballot is the mask word.
// projections.py:23-24.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(SP_THREADS) k_sp_learn(const bh_ctx c, const uint32_t* __restrict__ input) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = c.active_columns;
  const int cur = c.sc[BH_SC_STEP] & 1;
  const int* act = c.active_cols + cur * k;
  for (int r = blockIdx.x; r < k; r += gridDim.x) {
    const int col = act[r];
    double* prow = c.sp_perm + (long long)col * c.input_dim;
    uint32_t* mrow = c.sp_mask + (long long)col * c.mask_stride;
    for (int w = warp; w < c.input_words; w += SP_THREADS / 32) {
      int i = w * 32 + lane;
      uint32_t xin = __ldg(input + w);
      bool on = false;
      if (i < c.input_dim) {
        double p = __dadd_rn(prow[i], ((xin >> lane) & 1u) ? c.sp_delta_on : c.sp_delta_off);
        prow[i] = p;
        on = p >= c.sp_threshold;
      }
      uint32_t bits = __ballot_sync(BH_FULL, on);
      if (lane == 0) mrow[w] = bits;
    }
  }
}

// (c) duty-cycle EMA: two separately rounded float32 operations.
// regularizations.py:19-21; runs even when learning is off (networks.py:33).
__global__ void k_duty_update(const bh_ctx c) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= c.column_dim) return;
  float d = __fmul_rn(c.duty[j], c.duty_momentum);
  if (c.col_active[j]) d = __fadd_rn(d, c.duty_increment);
  c.duty[j] = d;
}
