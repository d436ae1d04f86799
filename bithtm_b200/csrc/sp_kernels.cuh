// Spatial-pooler phases: overlap (bit-packed popcount GEMV), boosting, canonical
// top-k inhibition, float64 permanence learning with mask re-pack, duty-cycle EMA.
// Reference: bithtm/projections.py:6-24, bithtm/regularizations.py:4-29,
// bithtm/networks.py:26-35.
#pragma once

#include "common.cuh"
#include "np_expf.h"

#define SP_THREADS 256

// workspace of topk_grid inside ctx.topk_ws (ints; after the 8192 ints of topk_multi)
#define TK2_BASE 8192
#define TK2_VALID 4        // 1 when u64[0] = max key, u64[1] = ~min key were published by ph_overlap<true>
#define TK2_CALL 5         // call counter: parity selects one of two histogram / candidate buffers
#define TK2_PAR0 16
#define TK2_BINS 2048      // 11-bit digit below the keys' common prefix
#define TK2_PSIZE 5136     // per parity: ghist[2048], count (+7 pad), cand_idx[1024], cand_key u64[1024]
#define TK2_TAB (TK2_PAR0 + 2 * TK2_PSIZE)  // [256 CTAs]: keys of a CTA's range in bins above the k-th key's
#define TK2_MAX_CTAS 256
#define TK2_INTS (TK2_TAB + TK2_MAX_CTAS)

// Predicted histogram (ctx.topk_ws + TK3_BASE): the overlap phase bins the boosted keys while it produces them,
// with a binning centred on the PREVIOUS step's k-th key, so that the selection that follows starts from a
// finished global histogram (no pass of its own, one grid barrier less).  Any monotone binning gives an exact
// selection; a prediction that misses (threshold below the binned range, > 1024 keys in the threshold bin)
// only costs the general path.
#define TK3_BASE 24576
#define TK3_VALID 0        // a binning is set (else: the exponent bits, which always work and rarely resolve)
#define TK3_SHIFT 1
#define TK3_BASE64 1       // u64 index: base key bits (ints 2, 3)
#define TK3_READY 4        // step + 1 of the step whose keys hist[step & 1] holds
#define TK3_SELECTED 5     // sharded step: step + 1 when the histogram exchange selected the active columns
#define TK3_HIST 8         // [2][TK2_BINS]
#define TK3_CNT (TK3_HIST + 2 * TK2_BINS)         // [2][8]: members of the threshold bin gathered so far
#define TK3_IDX (TK3_CNT + 16)                    // [2][1024] their positions
#define TK3_KEY (TK3_IDX + 2 * 1024)              // [2][1024] u64 keys (8-byte aligned)
#define TK3_TAB (TK3_KEY + 4 * 1024)              // [256] per CTA: keys of its range in bins above the threshold bin
#define TK3_INTS (TK3_TAB + 256)

struct Tk3Binning {
  unsigned long long base;
  int shift;
};
__device__ __forceinline__ Tk3Binning tk3_binning(const int* ws3) {
  Tk3Binning g;
  if (ws3[TK3_VALID]) {
    g.base = reinterpret_cast<const unsigned long long*>(ws3)[TK3_BASE64];
    g.shift = ws3[TK3_SHIFT];
  } else {
    g.base = 0ull;
    g.shift = 52;
  }
  return g;
}
// bin 0 = below the binned range (not counted), 1..2046 in range, 2047 = everything above
__device__ __forceinline__ int tk3_bin(const Tk3Binning& g, unsigned long long key) {
  if (key < g.base) return 0;
  const unsigned long long d = (key - g.base) >> g.shift;
  return d >= (unsigned long long)(TK2_BINS - 2) ? TK2_BINS - 1 : (int)d + 1;
}
// binning for the next step from this step's k-th and largest selected key: the k-th key lands in the middle
// bin, the selected keys span at most 256 bins
__device__ __forceinline__ void tk3_set_binning(int* ws3, unsigned long long kth, unsigned long long mx) {
  const unsigned long long range = mx > kth ? mx - kth : 1ull;
  int s = 0;
  while (s < 62 && (range >> s) > 256ull) ++s;
  const unsigned long long below = 1023ull << s;
  reinterpret_cast<unsigned long long*>(ws3)[TK3_BASE64] = kth > below ? kth - below : 0ull;
  ws3[TK3_SHIFT] = s;
  ws3[TK3_VALID] = 1;
}

__device__ __noinline__ unsigned long long block_reduce_u64(unsigned long long v, bool want_max,
                                                               unsigned long long* sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long n = __shfl_xor_sync(BH_FULL, v, o);
    v = want_max ? (n > v ? n : v) : (n < v ? n : v);
  }
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  unsigned long long r = sm[lane < nw ? lane : 0];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long n = __shfl_xor_sync(BH_FULL, r, o);
    r = want_max ? (n > r ? n : r) : (n < r ? n : r);
  }
  return r;
}


// ---------------------------------------------------------------------------------
// Connected mask from the float64 permanence (one warp per 32 consecutive inputs).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(SP_THREADS) k_sp_build_mask(const __grid_constant__ bh_ctx c) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = SP_THREADS / 32;
  const long long n_words = (long long)c.col_local * c.mask_stride;
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
__global__ void k_pack_input(const __grid_constant__ bh_ctx c, const uint8_t* __restrict__ src, uint32_t* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= c.input_words) return;
  int i = w * 32 + lane;
  bool on = i < c.input_dim && src[i] != 0;
  uint32_t bits = __ballot_sync(BH_FULL, on);
  if (lane == 0) dst[w] = bits;
}

// n_inputs bool vectors [n_inputs][I] (one byte per bit) -> packed rows of `pitch` words (>= input_words; the
// padding words are written as zero): the operand layout of the batched overlaps
__global__ void k_pack_inputs(const __grid_constant__ bh_ctx c, const uint8_t* __restrict__ src, int n_inputs, int pitch,
                              uint32_t* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const long long n_words = (long long)n_inputs * pitch;
  for (long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_words;
       w += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int row = (int)(w / pitch), wi = (int)(w - (long long)row * pitch);
    const int i = wi * 32 + lane;
    const bool on = i < c.input_dim && src[(long long)row * c.input_dim + i] != 0;
    const uint32_t bits = __ballot_sync(BH_FULL, on);
    if (lane == 0) dst[w] = bits;
  }
}

// ---------------------------------------------------------------------------------
// (a) overlap + (c) boost.  G lanes cooperate on one mask row with 128-bit loads
// (G = largest power of two <= min(32, mask_stride/4)), so a warp covers 32/G rows
// and every load instruction is a full coalesced 16 B per lane.  `s_in` = dynamic
// shared memory of mask_stride words (16-byte aligned).
// projections.py:18-21, regularizations.py:15-17.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ int overlap_group(int mask_stride) {
  int vec = mask_stride / 4, g = 1;
  while (g * 2 <= vec && g < 32) g *= 2;
  return g;
}

template <bool BOOST, bool HIST = false>
__device__ __forceinline__ void ph_overlap(const bh_ctx& c, const uint32_t* input, uint32_t* s_in, int b, int nb,
                                           int step_ov = -1) {  // step_ov: the step this pass belongs to (pipelined kernel)
  __shared__ unsigned long long s_range[2];  // max key, max ~key of this CTA
  // HIST: this CTA's keys in the predicted binning (see TK3_*): TK2_BINS ints of dynamic shared memory after
  // the staged input
  int* s_khist = reinterpret_cast<int*>(s_in + c.mask_stride);
  int* ws3 = c.topk_ws + TK3_BASE;
  Tk3Binning binning;
  if (HIST) {
    binning = tk3_binning(ws3);
#pragma unroll 1
    for (int i = threadIdx.x; i < TK2_BINS; i += blockDim.x) s_khist[i] = 0;
  }
  #pragma unroll 1
  for (int i = threadIdx.x; i < c.mask_stride; i += blockDim.x) s_in[i] = i < c.input_words ? input[i] : 0u;
  if (threadIdx.x < 2) s_range[threadIdx.x] = 0ull;
  __syncthreads();
  const int group = overlap_group(c.mask_stride);
  const int lane = threadIdx.x & 31;
  const int sub = lane % group;  // lane within its row group
  const int rows_per_warp = 32 / group;
  const int warps = blockDim.x >> 5;
  const int warp_global = b * warps + (threadIdx.x >> 5);
  const int n_warps = nb * warps;
  const int vec_per_row = c.mask_stride / 4;
  const uint4* mask4 = reinterpret_cast<const uint4*>(c.sp_mask);
  const uint4* s_in4 = reinterpret_cast<const uint4*>(s_in);
  const int n_rows = c.col_local;  // this rank's columns (all of them when not sharded)
  unsigned long long key_min = ~0ull, key_max = 0ull;  // of the boosted keys this thread wrote (for topk_grid)
  // per-row epilogue: overlap count, boosted key (NumPy-exact exp, exact float64 product), key range
  auto finish = [&](int row, int acc) {
    c.overlaps[row] = acc;
    if (BOOST) {
      float f = bh_np_expf(__fmul_rn(c.boost_coef, c.duty[row]));
      const double bo = __dmul_rn((double)f, (double)acc);
      c.boosted[row] = bo;
      const unsigned long long key = (unsigned long long)__double_as_longlong(bo);
      key_min = key < key_min ? key : key_min;
      key_max = key > key_max ? key : key_max;
      if (HIST) {
        const int bin = tk3_bin(binning, key);
        if (bin) atomicAdd(&s_khist[bin], 1);
      }
    }
  };
  auto popc4 = [](const uint4& m, const uint4& x) {
    return __popc(m.x & x.x) + __popc(m.y & x.y) + __popc(m.z & x.z) + __popc(m.w & x.w);
  };
  const int stride_rows = n_warps * rows_per_warp;
  if (vec_per_row == 4 * group && group == 32) {
    // Long rows (HBM-bound sizes): every lane owns exactly four 16-byte vectors of a row.  A warp takes
    // R consecutive rows at a time: the next row's four loads are issued before the current row is
    // reduced (software pipeline), and the per-row epilogue (exp, boost, stores, key range) runs once
    // per R rows with lane i finishing row i instead of once per row on a single lane.
    int R = 1;
    while (R < 32 && (long long)2 * R * n_warps <= n_rows) R *= 2;
    const int n_groups = (n_rows + R - 1) / R;
#pragma unroll 1
    for (int g = warp_global; g < n_groups; g += n_warps) {
      const int row0 = g * R;
      const int rows_here = n_rows - row0 < R ? n_rows - row0 : R;
      int my_acc = 0;
      uint4 cur[4];
      {
        const uint4* mrow = mask4 + (long long)row0 * vec_per_row + lane;
#pragma unroll
        for (int j = 0; j < 4; ++j) cur[j] = mrow[j * 32];
      }
#pragma unroll 1
      for (int i = 0; i < rows_here; ++i) {
        uint4 nxt[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) nxt[j] = make_uint4(0u, 0u, 0u, 0u);
        if (i + 1 < rows_here) {
          const uint4* mrow = mask4 + (long long)(row0 + i + 1) * vec_per_row + lane;
#pragma unroll
          for (int j = 0; j < 4; ++j) nxt[j] = mrow[j * 32];
        }
        int acc = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc += popc4(cur[j], s_in4[lane + j * 32]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(BH_FULL, acc, o);
        if (lane == i) my_acc = acc;
#pragma unroll
        for (int j = 0; j < 4; ++j) cur[j] = nxt[j];
      }
      if (lane < rows_here) finish(row0 + lane, my_acc);
    }
  } else {
#pragma unroll 1
    for (int base = warp_global * rows_per_warp; base < n_rows; base += stride_rows) {
      int row = base + lane / group;
      int acc = 0;
      if (row < n_rows) {
        const uint4* mrow = mask4 + (long long)row * vec_per_row;
        int v = sub;
#pragma unroll 1
        for (; v + 3 * group < vec_per_row; v += 4 * group) {  // 4 independent 16-byte loads in flight
          const uint4 m0 = mrow[v], m1 = mrow[v + group], m2 = mrow[v + 2 * group], m3 = mrow[v + 3 * group];
          acc += popc4(m0, s_in4[v]) + popc4(m1, s_in4[v + group]) + popc4(m2, s_in4[v + 2 * group]) +
                 popc4(m3, s_in4[v + 3 * group]);
        }
#pragma unroll 1
        for (; v < vec_per_row; v += group) acc += popc4(mrow[v], s_in4[v]);
      }
      for (int o = group >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(BH_FULL, acc, o);
      if (sub == 0 && row < n_rows) finish(row, acc);
    }
  }
  if (BOOST) {  // range of the keys, for the top-k that follows (saves it a pass): warp -> CTA -> one red per CTA
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long a = __shfl_xor_sync(BH_FULL, key_min, o), z = __shfl_xor_sync(BH_FULL, key_max, o);
      key_min = a < key_min ? a : key_min;
      key_max = z > key_max ? z : key_max;
    }
    if (lane == 0 && key_max >= key_min) {
      atomicMax(&s_range[0], key_max);
      atomicMax(&s_range[1], ~key_min);
    }
    __syncthreads();
    if (threadIdx.x == 0 && (s_range[0] | s_range[1])) {
      unsigned long long* w64 = reinterpret_cast<unsigned long long*>(c.topk_ws + TK2_BASE);
      atomicMax(&w64[0], s_range[0]);
      atomicMax(&w64[1], s_range[1]);
      c.topk_ws[TK2_BASE + TK2_VALID] = 1;
    }
    if (HIST) {  // (the barrier above also completed this CTA's shared histogram)
      const int step = step_ov >= 0 ? step_ov : c.sc[BH_SC_STEP];
      int* ghist = ws3 + TK3_HIST + (step & 1) * TK2_BINS;
#pragma unroll 1
      for (int i = threadIdx.x; i < TK2_BINS; i += blockDim.x) {
        const int h = s_khist[i];
        if (h) atomicAdd(&ghist[i], h);
      }
      if (b == 0 && threadIdx.x == 0) ws3[TK3_READY] = step + 1;
    }
  }
}

template <bool BOOST>
__global__ void __launch_bounds__(SP_THREADS) k_sp_overlap(const __grid_constant__ bh_ctx c, const uint32_t* input) {
  extern __shared__ __align__(16) uint32_t s_dyn[];
  ph_overlap<BOOST>(c, input, s_dyn, blockIdx.x, gridDim.x);
}

// ---------------------------------------------------------------------------------
// (a') overlaps of many input vectors against the one connected mask: a 32 x 32 tile of
// (input, column) pairs per CTA, mask and input words staged in shared memory 32 words at
// a time (+1 padding: conflict-free), output coalesced along the columns.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_sp_overlap_batched(const __grid_constant__ bh_ctx c, const uint32_t* __restrict__ inputs,
                                                            int n_inputs, int32_t* __restrict__ out) {
  __shared__ uint32_t s_m[32][33], s_x[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // tx: column in tile, ty: input in tile
  const int col0 = blockIdx.x * 32, in0 = blockIdx.y * 32;
  int acc = 0;
  for (int w0 = 0; w0 < c.input_words; w0 += 32) {
    const int w = w0 + tx;
    const int col = col0 + ty, inp = in0 + ty;
    s_m[ty][tx] = (col < c.col_local && w < c.input_words) ? c.sp_mask[(long long)col * c.mask_stride + w] : 0u;
    s_x[ty][tx] = (inp < n_inputs && w < c.input_words) ? inputs[(long long)inp * c.input_words + w] : 0u;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; ++j) acc += __popc(s_m[tx][j] & s_x[ty][j]);
    __syncthreads();
  }
  if (col0 + tx < c.col_local && in0 + ty < n_inputs) out[(long long)(in0 + ty) * c.col_local + col0 + tx] = acc;
}

// ---------------------------------------------------------------------------------
// (a'') the same batched overlap as an int8 TENSOR-CORE contraction:
//   out[b][j] = sum_i x[b][i] * m[j][i],  x, m in {0,1}  ==  popcount(x[b] & m[j]).
// A = inputs (row-major, K = input bits), B = connected mask (one mask row per output
// column, K-contiguous: the "col" operand).  Both stay bit-packed in HBM and in shared
// memory and are widened to bytes in registers right before the MMA, so operand traffic is
// 1 bit per element.  mma.sync.m16n8k32.u8.u8.s32, exact in int32.
//
// Fragment layout (PTX ISA, m16n8k32 .u8): g = lane / 4, t = lane % 4;
//   A reg0 = (row g, k 4t..4t+3)  reg1 = (row g+8, same k)  reg2/3 = same rows, k + 16
//   B reg0 = (k 4t..4t+3, col g)  reg1 = k + 16;  C = (row g, col 2t, 2t+1), (row g+8, ...).
// A dot product does not care which bit sits at which k as long as A and B agree, so the
// assignment is chosen for the cheapest widening: byte j of the "low" register of thread t
// is bit 8j + 2t + 1 of the 32-bit word, byte j of the "high" register is bit 8j + 2t:
//   A (values 0 / 1):    y = w >> 2t;          hi = y & 0x01010101;  lo = (y >> 1) & 0x01010101
//   B (values 0 / 128):  y = w * 2^(6 - 2t);   lo = y & 0x80808080;  hi = (y * 2) & 0x80808080
// (four cheap integer ops per word; the compiler turns the power-of-two multiplies into shifts -- forcing
// them onto the FMA pipe as real IMADs measured slower, 829 vs 747 us at 256 x 65536 x 16384), and
// the accumulators are divided by 128 at the end (exact; K * 128 < 2^31 for K < 2^24 bits).
//
// CTA tile 64 inputs x 128 columns, 4 warps of 32 x 64 (2 x 8 MMA tiles, 64 accumulators),
// K staged 512 bits at a time through two shared-memory buffers filled by cp.async (the next
// chunk travels while this one is multiplied); shared rows padded to 20 words so the 128-bit
// fragment reads (8 rows x 16 B) are conflict-free.
// ---------------------------------------------------------------------------------
#define OTC_M 64
#define OTC_N 128
#define OTC_KW 16
#define OTC_LD 20
#define OTC_THREADS 128

__device__ __forceinline__ void otc_mma(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 4-byte asynchronous copy global -> shared; !ok copies nothing and writes zero.
__device__ __forceinline__ void otc_cp4(uint32_t* dst_smem, const uint32_t* src, bool ok) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  const int n = ok ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}

__global__ void __launch_bounds__(OTC_THREADS, 4) k_sp_overlap_batched_tc(const __grid_constant__ bh_ctx c,
                                                                          const uint32_t* __restrict__ inputs, int n_inputs,
                                                                          int32_t* __restrict__ out) {
  __shared__ __align__(16) uint32_t s_a[2][OTC_M][OTC_LD], s_b[2][OTC_N][OTC_LD];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = (warp & 1) * 32, wn = (warp >> 1) * 64;
  const int m0 = blockIdx.y * OTC_M, n0 = blockIdx.x * OTC_N;
  const int a_shift = 2 * t;
  const uint32_t b_mul = 1u << (6 - 2 * t);
  int acc[2][8][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[i][j][r] = 0;

  // stage one K chunk (OTC_KW words of every row of the tile) into buffer `buf`: a warp covers
  // two rows per instruction (lanes 0-15 / 16-31), asynchronously, out-of-range -> zero
  auto stage = [&](int buf, int w0) {
    const int w = w0 + (lane & 15), half = lane >> 4;
    const bool w_ok = w < c.input_words;
#pragma unroll
    for (int r = 2 * warp + half; r < OTC_M; r += OTC_THREADS / 16) {
      const bool ok = w_ok && m0 + r < n_inputs;
      otc_cp4(&s_a[buf][r][lane & 15], ok ? inputs + (long long)(m0 + r) * c.input_words + w : inputs, ok);
    }
#pragma unroll
    for (int r = 2 * warp + half; r < OTC_N; r += OTC_THREADS / 16) {
      const bool ok = w_ok && n0 + r < c.col_local;
      otc_cp4(&s_b[buf][r][lane & 15], ok ? c.sp_mask + (long long)(n0 + r) * c.mask_stride + w : c.sp_mask, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int n_chunks = (c.input_words + OTC_KW - 1) / OTC_KW;
  stage(0, 0);
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < n_chunks) {
      stage(buf ^ 1, (ch + 1) * OTC_KW);  // buffer buf^1 was released by the barrier ending chunk ch-1
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
#pragma unroll 1
    for (int kw = 0; kw < OTC_KW; kw += 4) {
      uint32_t a[4][2][4];  // [word of the group of 4][m tile][fragment register]
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const uint4 r_lo = *reinterpret_cast<const uint4*>(&s_a[buf][wm + i * 16 + g][kw]);
        const uint4 r_hi = *reinterpret_cast<const uint4*>(&s_a[buf][wm + i * 16 + g + 8][kw]);
        const uint32_t lo[4] = {r_lo.x, r_lo.y, r_lo.z, r_lo.w}, hi[4] = {r_hi.x, r_hi.y, r_hi.z, r_hi.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t y0 = lo[q] >> a_shift, y1 = hi[q] >> a_shift;
          a[q][i][0] = (y0 >> 1) & 0x01010101u;
          a[q][i][1] = (y1 >> 1) & 0x01010101u;
          a[q][i][2] = y0 & 0x01010101u;
          a[q][i][3] = y1 & 0x01010101u;
        }
      }
      // 4 column tiles at a time: 8 independent accumulator tiles between two MMAs on the same one
#pragma unroll
      for (int j0 = 0; j0 < 8; j0 += 4) {
        uint32_t wb[4][4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const uint4 r_b = *reinterpret_cast<const uint4*>(&s_b[buf][wn + (j0 + jj) * 8 + g][kw]);
          wb[jj][0] = r_b.x, wb[jj][1] = r_b.y, wb[jj][2] = r_b.z, wb[jj][3] = r_b.w;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const uint32_t y = wb[jj][q] * b_mul;
            const uint32_t b0 = y & 0x80808080u, b1 = (y * 2u) & 0x80808080u;
            otc_mma(acc[0][j0 + jj], a[q][0], b0, b1);
            otc_mma(acc[1][j0 + jj], a[q][1], b0, b1);
          }
      }
    }
    __syncthreads();
  }

  const bool pair_ok = (c.col_local & 1) == 0;  // 8-byte stores need an even row pitch (col is always even)
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = n0 + wn + j * 8 + 2 * t;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = m0 + wm + i * 16 + g + 8 * h;
        if (row >= n_inputs || col >= c.col_local) continue;
        int32_t* o = out + (long long)row * c.col_local + col;
        const int v0 = acc[i][j][2 * h] >> 7, v1 = acc[i][j][2 * h + 1] >> 7;
        if (pair_ok) {
          *reinterpret_cast<int2*>(o) = make_int2(v0, v1);
        } else {
          o[0] = v0;
          if (col + 1 < c.col_local) o[1] = v1;
        }
      }
    }
}

__global__ void k_boost(const __grid_constant__ bh_ctx c) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= c.col_local) return;
  float f = bh_np_expf(__fmul_rn(c.boost_coef, c.duty[j]));
  c.boosted[j] = __dmul_rn((double)f, (double)c.overlaps[j]);
}

// ---------------------------------------------------------------------------------
// (b) canonical global inhibition: k largest keys, ties -> lower column index,
// output ascending.  Keys are non-negative doubles, so their bit patterns order
// like unsigned integers.  Single CTA (any block size that is a multiple of 32):
//   1. block min/max -> the leading bits all keys share are skipped;
//   2. 8-bit MSB-first radix passes (warp-aggregated shared-memory histogram,
//      warp-parallel bin scan) until the bin holding the k-th key has at most
//      blockDim.x members (usually one pass);
//   3. those candidates are ranked by (key desc, index asc) by counting, which
//      yields the exact k-th (key, index) pair;
//   4. ordered compaction over the column index.
// If all 64 bits are consumed and the bin is still larger (many equal keys), the
// lowest indices among the equal keys are taken by an ordered tie count.
// regularizations.py:28-29 (np.argpartition's tie-break/order are undefined).
// ---------------------------------------------------------------------------------
#define TOPK_THREADS 1024

// Shared-memory scratch of the top-k variants (one instance per kernel: they never run at
// the same time inside a CTA).
struct TopkScratch {
  unsigned long long cand_key[TOPK_THREADS];
  int cand_idx[TOPK_THREADS];
  int hist[2048];
};
__device__ __forceinline__ TopkScratch& topk_scratch() {
  __shared__ TopkScratch s;
  return s;
}

// keys[0..n): select the k largest (ties -> lower position); positions ascending are
// written as out[i] = map ? map[pos] : pos, and flags[that value] = 1 when flags != null.
__device__ __noinline__ void topk_core(const unsigned long long* keys, const int C, const int k, int* out, const int* map,
                          uint8_t* flags) {
  TopkScratch& sm = topk_scratch();
  int* hist = sm.hist;
  unsigned long long* cand_key = sm.cand_key;
  int* cand_idx = sm.cand_idx;
  __shared__ int s_scan[32];
  __shared__ unsigned long long s_u64[32];
  __shared__ int s_bin, s_rem, s_ncand, s_kth_idx;
  __shared__ unsigned long long s_kth_key;
  const int t = threadIdx.x, lane = t & 31, NT = blockDim.x;

  // 1. shared leading bits
  unsigned long long mn = ~0ull, mx = 0ull;
  #pragma unroll 1
  for (int j = t; j < C; j += NT) {
    unsigned long long key = keys[j];
    mn = key < mn ? key : mn;
    mx = key > mx ? key : mx;
  }
  mn = block_reduce_u64(mn, false, s_u64);
  mx = block_reduce_u64(mx, true, s_u64);
  int consumed = (mn == mx) ? 64 : __clzll((long long)(mn ^ mx));  // leading bits equal for all candidates
  unsigned long long prefix =
      (consumed == 0) ? 0ull : (consumed == 64 ? mx : (mx >> (64 - consumed)) << (64 - consumed));
  int rem = k;    // how many still to take from the current candidate set
  int ncand = C;  // size of the current candidate set (keys matching `prefix` on the consumed bits)

  // 2. radix passes
  while (consumed < 64 && ncand > NT) {
    const int shift = (64 - consumed - 8) > 0 ? (64 - consumed - 8) : 0;
    const int width = 64 - consumed - shift;  // 1..8 bits
    const unsigned long long hi_mask = consumed == 0 ? 0ull : (~0ull << (64 - consumed));
    #pragma unroll 1
    for (int i = t; i < 256; i += NT) hist[i] = 0;
    __syncthreads();
    #pragma unroll 1
    for (int base = 0; base < C; base += NT) {
      const int j = base + t;
      const unsigned long long key = j < C ? keys[j] : 0ull;
      const bool in = j < C && (key & hi_mask) == prefix;
      const int d = (int)((key >> shift) & ((1u << width) - 1u));
      // warp-aggregated histogram: one shared atomic per distinct digit per warp
      const unsigned peers = __match_any_sync(BH_FULL, in ? d : -1);
      if (in && lane == (__ffs(peers) - 1)) atomicAdd(&hist[d], __popc(peers));
    }
    __syncthreads();
    if (t < 32) {
      // bins in descending order, 8 per lane: lane 0 owns bins 255..248
      int local[8], sum = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        local[i] = hist[255 - (lane * 8 + i)];
        sum += local[i];
      }
      int incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(BH_FULL, incl, o);
        if (lane >= o) incl += n;
      }
      const int before = incl - sum;  // keys in bins above this lane's bins
      if (before < rem && incl >= rem) {
        int r = rem - before;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (r > 0 && local[i] >= r) {
            s_bin = 255 - (lane * 8 + i);
            s_rem = r;
            s_ncand = local[i];
            r = -1;
          } else if (r > 0) {
            r -= local[i];
          }
        }
      }
    }
    __syncthreads();
    prefix |= (unsigned long long)s_bin << shift;
    rem = s_rem;
    ncand = s_ncand;
    consumed += width;
    __syncthreads();
  }

  // 3. exact k-th (key, index)
  bool tie_mode = false;  // many identical keys: take the `rem` lowest indices among them
  unsigned long long kth_key = prefix;
  int kth_idx = 0x7fffffff;
  if (ncand > NT) {
    tie_mode = true;  // consumed == 64: every candidate equals prefix
  } else {
    const unsigned long long hi_mask = consumed == 0 ? 0ull : (consumed == 64 ? ~0ull : (~0ull << (64 - consumed)));
    if (t == 0) s_ncand = 0;
    __syncthreads();
    #pragma unroll 1
    for (int j = t; j < C; j += NT) {
      const unsigned long long key = keys[j];
      if ((key & hi_mask) == prefix) {
        const int p = atomicAdd(&s_ncand, 1);
        cand_key[p] = key;
        cand_idx[p] = j;
      }
    }
    __syncthreads();
    const int n = s_ncand;
    if (t < n) {
      const unsigned long long mk = cand_key[t];
      const int mi = cand_idx[t];
      int ahead = 0;
#pragma unroll 2
      for (int i = 0; i < n; ++i) {
        const unsigned long long ok = cand_key[i];
        ahead += (ok > mk || (ok == mk && cand_idx[i] < mi)) ? 1 : 0;
      }
      if (ahead == rem - 1) {
        s_kth_key = mk;
        s_kth_idx = mi;
      }
    }
    __syncthreads();
    kth_key = s_kth_key;
    kth_idx = s_kth_idx;
  }

  // 4. ordered compaction over column index
  int base_sel = 0, base_tie = 0;
  #pragma unroll 1
  for (int tile = 0; tile < C; tile += NT) {
    const int j = tile + t;
    const unsigned long long key = j < C ? keys[j] : 0ull;
    const bool gt = j < C && key > kth_key;
    const bool eq = j < C && key == kth_key;
    bool take;
    if (tie_mode) {
      int tie_total;
      const int tie_rank = base_tie + block_excl_scan(eq ? 1 : 0, s_scan, tie_total);
      base_tie += tie_total;
      take = gt || (eq && tie_rank < rem);
    } else {
      take = gt || (eq && j <= kth_idx);
    }
    int sel_total;
    const int pos = base_sel + block_excl_scan(take ? 1 : 0, s_scan, sel_total);
    if (take && pos < k) {
      const int col = map ? map[j] : j;
      out[pos] = col;
      if (flags) flags[col] = 1;
    }
    base_sel += sel_total;
  }
}

// ---------------------------------------------------------------------------------
// (b) for small column counts (C <= 4 keys per thread, one CTA): the same canonical
// selection with every key held in registers.  One 11-bit histogram pass over the bits
// below the keys' common prefix finds the bin of the k-th key (a handful of members at
// these sizes), its members are ranked by counting, and the ordered compaction is a
// single block scan because thread order is index order.  The key range comes from the
// producer (ph_overlap<true>, ctx.topk_ws) when `ws` says so.  Falls back to topk_core.
// ---------------------------------------------------------------------------------
#define TOPK_SMALL_KPT 4
__device__ __noinline__ void topk_small(const unsigned long long* keys, const int C, const int k, int* out, const int* map,
                           uint8_t* flags, int* ws) {
  TopkScratch& sm = topk_scratch();
  int* hist = sm.hist;
  unsigned long long* cand_key = sm.cand_key;
  int* cand_idx = sm.cand_idx;
  __shared__ int s_scan[32];
  __shared__ unsigned long long s_u64[32];
  __shared__ int s_bin, s_rem, s_ncand, s_kth_idx, s_cnt;
  __shared__ unsigned long long s_kth_key;
  const int t = threadIdx.x, NT = blockDim.x;
  const int kpt = (C + NT - 1) / NT;  // keys per thread, <= TOPK_SMALL_KPT (caller checks)
  unsigned long long key[TOPK_SMALL_KPT];
  unsigned long long mn = ~0ull, mx = 0ull;
#pragma unroll
  for (int i = 0; i < TOPK_SMALL_KPT; ++i) {
    const int j = t * kpt + i;
    const bool in = i < kpt && j < C;
    key[i] = in ? keys[j] : 0ull;
    if (in) {
      mn = key[i] < mn ? key[i] : mn;
      mx = key[i] > mx ? key[i] : mx;
    }
  }
  unsigned long long* w64 = reinterpret_cast<unsigned long long*>(ws);
  if (ws && ws[TK2_VALID]) {  // published by the overlap phase; consumed (reset) here
    mx = w64[0];
    mn = ~w64[1];
    __syncthreads();
    if (t == 0) {
      w64[0] = 0ull;
      w64[1] = 0ull;
      ws[TK2_VALID] = 0;
    }
  } else {
    mn = block_reduce_u64(mn, false, s_u64);
    mx = block_reduce_u64(mx, true, s_u64);
  }
  const int consumed = (mn == mx) ? 64 : __clzll((long long)(mn ^ mx));
  if (consumed == 64) {  // all keys equal: the general algorithm handles the tie rule
    topk_core(keys, C, k, out, map, flags);
    return;
  }
  const int shift = (64 - consumed - 11) > 0 ? (64 - consumed - 11) : 0;
  const int width = 64 - consumed - shift;
  const unsigned dmask = (1u << width) - 1u;
#pragma unroll 1
  for (int i = t; i < 2048; i += NT) hist[i] = 0;
  if (t == 0) s_cnt = 0;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < TOPK_SMALL_KPT; ++i)
    if (i < kpt && t * kpt + i < C) atomicAdd(&hist[(int)((key[i] >> shift) & dmask)], 1);
  __syncthreads();
  {  // bins in descending order, 2048 / NT per thread; ordered block scan finds the bin of the k-th key
    const int per = (2048 + NT - 1) / NT;
    int sum = 0;
    for (int i = 0; i < per; ++i) {
      const int bin = 2047 - (t * per + i);
      sum += bin >= 0 ? hist[bin] : 0;
    }
    int total;
    const int before = block_excl_scan(sum, s_scan, total);
    if (before < k && before + sum >= k) {
      int r = k - before;
      for (int i = 0; i < per; ++i) {
        const int bin = 2047 - (t * per + i);
        const int h = bin >= 0 ? hist[bin] : 0;
        if (r > 0 && h >= r) {
          s_bin = bin;
          s_rem = r;
          s_ncand = h;
          r = -1;
        } else if (r > 0) {
          r -= h;
        }
      }
    }
  }
  __syncthreads();
  if (s_ncand > 256) {
    topk_core(keys, C, k, out, map, flags);
    return;
  }
  const int bin = s_bin, rem = s_rem;
#pragma unroll
  for (int i = 0; i < TOPK_SMALL_KPT; ++i)
    if (i < kpt && t * kpt + i < C && (int)((key[i] >> shift) & dmask) == bin) {
      const int p = atomicAdd(&s_cnt, 1);
      cand_key[p] = key[i];
      cand_idx[p] = t * kpt + i;
    }
  __syncthreads();
  const int nc = s_cnt;
  if (t < nc) {
    const unsigned long long mk = cand_key[t];
    const int mi = cand_idx[t];
    int ahead = 0;
    for (int i = 0; i < nc; ++i) {
      const unsigned long long ok = cand_key[i];
      ahead += (ok > mk || (ok == mk && cand_idx[i] < mi)) ? 1 : 0;
    }
    if (ahead == rem - 1) {
      s_kth_key = mk;
      s_kth_idx = mi;
    }
  }
  __syncthreads();
  const unsigned long long kth_key = s_kth_key;
  const int kth_idx = s_kth_idx;
  int n_take = 0;
  bool take[TOPK_SMALL_KPT];
#pragma unroll
  for (int i = 0; i < TOPK_SMALL_KPT; ++i) {
    const int j = t * kpt + i;
    take[i] = i < kpt && j < C && (key[i] > kth_key || (key[i] == kth_key && j <= kth_idx));
    n_take += take[i] ? 1 : 0;
  }
  int total;
  int pos = block_excl_scan(n_take, s_scan, total);
#pragma unroll
  for (int i = 0; i < TOPK_SMALL_KPT; ++i)
    if (take[i] && pos < k) {
      const int j = t * kpt + i;
      const int col = map ? map[j] : j;
      out[pos++] = col;
      if (flags) flags[col] = 1;
    }
}

// retire the previous step's column flags (single CTA, before the new ones are set)
__device__ __forceinline__ void retire_prev_flags(const bh_ctx& c) {
  const int k = c.active_columns;
  const int* prev = c.active_cols + ((c.sc[BH_SC_STEP] & 1) ^ 1) * k;
#pragma unroll 1
  for (int i = threadIdx.x; i < k; i += blockDim.x) c.col_active[prev[i]] = 0;
  __syncthreads();
}

// GlobalInhibition.process on this rank's keys = all columns (not sharded)
__device__ void ph_topk(const bh_ctx& c) {
  retire_prev_flags(c);
  const int k = c.active_columns;
  if (c.column_dim <= TOPK_SMALL_KPT * (int)blockDim.x)
    topk_small(reinterpret_cast<const unsigned long long*>(c.boosted), c.column_dim, k,
               c.active_cols + (c.sc[BH_SC_STEP] & 1) * k, nullptr, c.col_active, c.topk_ws + TK2_BASE);
  else
    topk_core(reinterpret_cast<const unsigned long long*>(c.boosted), c.column_dim, k,
              c.active_cols + (c.sc[BH_SC_STEP] & 1) * k, nullptr, c.col_active);
}

__global__ void __launch_bounds__(TOPK_THREADS) k_topk(const __grid_constant__ bh_ctx c) { ph_topk(c); }

// Column shard, exchange 1: this shard's best min(k, col_local) candidates as (key,
// global column) pairs in ascending column order.  `scratch` holds k_loc ints.
__global__ void __launch_bounds__(TOPK_THREADS)
    k_topk_shard_local(const __grid_constant__ bh_ctx c, int* scratch, double* cand_keys, int32_t* cand_cols) {
  const int k_loc = c.active_columns < c.col_local ? c.active_columns : c.col_local;
  topk_core(reinterpret_cast<const unsigned long long*>(c.boosted), c.col_local, k_loc, scratch, nullptr, nullptr);
  __syncthreads();
  for (int i = threadIdx.x; i < k_loc; i += blockDim.x) {
    const int pos = scratch[i];
    cand_keys[i] = c.boosted[pos];
    cand_cols[i] = c.col_lo + pos;
  }
}

// Global top-k among the gathered candidates (rank order = ascending column, so the
// position tie-break is the column tie-break); identical on every rank.
__global__ void __launch_bounds__(TOPK_THREADS)
    k_topk_shard_merge(const __grid_constant__ bh_ctx c, const double* cand_keys, const int32_t* cand_cols, int n) {
  retire_prev_flags(c);
  const int k = c.active_columns;
  topk_core(reinterpret_cast<const unsigned long long*>(cand_keys), n, k, c.active_cols + (c.sc[BH_SC_STEP] & 1) * k,
            cand_cols, c.col_active);
}

// ---------------------------------------------------------------------------------
// (b) for large column counts: the same selection on a cooperative grid of nb CTAs
// (grid barriers between stages).  Workspace `ws` (ctx.topk_ws, zero between calls):
//   u64[0] = max key, u64[1] = max ~key (=> min), u64[2] = k-th key, u64[3..3+1024) =
//   candidate keys; then ints: [0] cand count, [1] k-th index, [2] tie mode, [3] rem,
//   [8..8+8*256) per-pass histograms, [.. +1024) candidate indices, [.. +1024) per-CTA
//   tie counts.  Per-CTA selected counts use ctx.blk row 6.
// ---------------------------------------------------------------------------------
#define TKW_U64_CAND 4
#define TKW_INT_BASE ((TKW_U64_CAND + TOPK_THREADS) * 2)
#define TKW_HIST 8
#define TKW_CIDX (TKW_HIST + 8 * 256)
#define TKW_TIES (TKW_CIDX + TOPK_THREADS)
#define TKW_INTS (TKW_INT_BASE + TKW_TIES + BH_BLK_STRIDE)
#define BLK_TOPK 6

__device__ __noinline__ void topk_multi(const bh_ctx& c, const unsigned long long* keys, const int n, const int k, int* out,
                           const int* map, uint8_t* flags, int b, int nb, GridBar& bar) {
  int* hist = topk_scratch().hist;
  __shared__ int s_scan[32];
  __shared__ unsigned long long s_u64[32];
  __shared__ int s_bin, s_rem, s_ncand;
  unsigned long long* w64 = reinterpret_cast<unsigned long long*>(c.topk_ws);
  int* wi = c.topk_ws + TKW_INT_BASE;
  const int t = threadIdx.x, lane = t & 31, NT = blockDim.x;
  const Range rg = block_range(n, b, nb);

  // stage 1: global min / max
  unsigned long long mn = ~0ull, mx = 0ull;
#pragma unroll 1
  for (int j = rg.begin + t; j < rg.end; j += NT) {
    const unsigned long long key = keys[j];
    mn = key < mn ? key : mn;
    mx = key > mx ? key : mx;
  }
  mn = block_reduce_u64(mn, false, s_u64);
  mx = block_reduce_u64(mx, true, s_u64);
  if (t == 0 && rg.begin < rg.end) {
    atomicMax(&w64[0], mx);
    atomicMax(&w64[1], ~mn);
  }
  grid_barrier(bar, nb);
  mx = w64[0];
  mn = ~w64[1];
  int consumed = (mn == mx) ? 64 : __clzll((long long)(mn ^ mx));
  unsigned long long prefix =
      (consumed == 0) ? 0ull : (consumed == 64 ? mx : (mx >> (64 - consumed)) << (64 - consumed));
  int rem = k, ncand = n, pass = 0;

  // stage 2: radix passes over global histograms until <= TOPK_THREADS candidates remain
#pragma unroll 1
  while (consumed < 64 && ncand > TOPK_THREADS && pass < 8) {
    const int shift = (64 - consumed - 8) > 0 ? (64 - consumed - 8) : 0;
    const int width = 64 - consumed - shift;
    const unsigned long long hi_mask = consumed == 0 ? 0ull : (~0ull << (64 - consumed));
    int* ghist = wi + TKW_HIST + pass * 256;
#pragma unroll 1
    for (int i = t; i < 256; i += NT) hist[i] = 0;
    __syncthreads();
#pragma unroll 1
    for (int base = rg.begin; base < rg.end; base += NT) {
      const int j = base + t;
      const unsigned long long key = j < rg.end ? keys[j] : 0ull;
      const bool in = j < rg.end && (key & hi_mask) == prefix;
      const int d = (int)((key >> shift) & ((1u << width) - 1u));
      const unsigned peers = __match_any_sync(BH_FULL, in ? d : -1);
      if (in && lane == (__ffs(peers) - 1)) atomicAdd(&hist[d], __popc(peers));
    }
    __syncthreads();
#pragma unroll 1
    for (int i = t; i < 256; i += NT)
      if (hist[i]) atomicAdd(&ghist[i], hist[i]);
    grid_barrier(bar, nb);
    if (t < 32) {
      int local[8], sum = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        local[i] = ghist[255 - (lane * 8 + i)];
        sum += local[i];
      }
      int incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(BH_FULL, incl, o);
        if (lane >= o) incl += v;
      }
      const int before = incl - sum;
      if (before < rem && incl >= rem) {
        int r = rem - before;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (r > 0 && local[i] >= r) {
            s_bin = 255 - (lane * 8 + i);
            s_rem = r;
            s_ncand = local[i];
            r = -1;
          } else if (r > 0) {
            r -= local[i];
          }
        }
      }
    }
    __syncthreads();
    prefix |= (unsigned long long)s_bin << shift;
    rem = s_rem;
    ncand = s_ncand;
    consumed += width;
    ++pass;
    __syncthreads();
  }

  // stage 3: gather the candidates (unordered), rank them in CTA 0
  const bool tie_mode = ncand > TOPK_THREADS;  // > 1024 identical keys at the cut
  if (!tie_mode) {
    const unsigned long long hi_mask = consumed == 0 ? 0ull : (consumed == 64 ? ~0ull : (~0ull << (64 - consumed)));
#pragma unroll 1
    for (int j = rg.begin + t; j < rg.end; j += NT) {
      const unsigned long long key = keys[j];
      if ((key & hi_mask) == prefix) {
        const int p = atomicAdd(&wi[0], 1);
        w64[TKW_U64_CAND + p] = key;
        wi[TKW_CIDX + p] = j;
      }
    }
  }
  grid_barrier(bar, nb);
  if (!tie_mode && b == 0) {
    const int nc = wi[0];
    if (t < nc) {
      const unsigned long long mk = w64[TKW_U64_CAND + t];
      const int mi = wi[TKW_CIDX + t];
      int ahead = 0;
#pragma unroll 2
      for (int i = 0; i < nc; ++i) {
        const unsigned long long ok = w64[TKW_U64_CAND + i];
        ahead += (ok > mk || (ok == mk && wi[TKW_CIDX + i] < mi)) ? 1 : 0;
      }
      if (ahead == rem - 1) {
        w64[2] = mk;
        wi[1] = mi;
      }
    }
  }
  if (tie_mode && b == 0 && t == 0) {
    w64[2] = prefix;
    wi[1] = 0x7fffffff;
  }
  grid_barrier(bar, nb);
  const unsigned long long kth_key = w64[2];
  const int kth_idx = wi[1];

  // stage 4: per-CTA counts over contiguous ranges, then the ordered write
  int n_sel = 0, n_tie = 0;
#pragma unroll 1
  for (int j = rg.begin + t; j < rg.end; j += NT) {
    const unsigned long long key = keys[j];
    n_tie += key == kth_key ? 1 : 0;
    n_sel += (key > kth_key || (!tie_mode && key == kth_key && j <= kth_idx)) ? 1 : 0;
  }
  n_sel = block_sum(n_sel, s_scan);
  n_tie = block_sum(n_tie, s_scan);
  if (t == 0) {
    BLK(c, BLK_TOPK)[b] = n_sel;
    wi[TKW_TIES + b] = n_tie;
  }
  grid_barrier(bar, nb);
  int tie_before = 0, tie_all = 0, gt_before = 0, gt_all = 0;
  blk_prefix(BLK(c, BLK_TOPK), b, nb, s_scan, gt_before, gt_all);
  if (tie_mode) blk_prefix(wi + TKW_TIES, b, nb, s_scan, tie_before, tie_all);
  // in tie mode the selected set = all keys > kth plus the first `rem` equal keys by index;
  // the output position of an element is (#greater before it) + (#taken ties before it)
  int base_sel = gt_before + (tie_mode ? (tie_before < rem ? tie_before : rem) : 0), base_tie = tie_before;
#pragma unroll 1
  for (int tile = rg.begin; tile < rg.end; tile += NT) {
    const int j = tile + t;
    const unsigned long long key = j < rg.end ? keys[j] : 0ull;
    const bool gt = j < rg.end && key > kth_key;
    const bool eq = j < rg.end && key == kth_key;
    bool take;
    if (tie_mode) {
      int tie_total;
      const int tie_rank = base_tie + block_excl_scan(eq ? 1 : 0, s_scan, tie_total);
      base_tie += tie_total;
      take = gt || (eq && tie_rank < rem);
    } else {
      take = gt || (eq && j <= kth_idx);
    }
    int sel_total;
    const int pos = base_sel + block_excl_scan(take ? 1 : 0, s_scan, sel_total);
    if (take && pos < k) {
      const int col = map ? map[j] : j;
      out[pos] = col;
      if (flags) flags[col] = 1;
    }
    base_sel += sel_total;
  }
  // leave the workspace zeroed for the next call (all readers are past the last barrier)
  grid_barrier(bar, nb);
  if (b == 0) {
#pragma unroll 1
    for (int i = t; i < TKW_HIST + 8 * 256; i += NT) wi[i] = 0;
    if (t < 3) w64[t] = 0ull;
  }
}

// ---------------------------------------------------------------------------------
// (b) grid-wide selection with TWO grid barriers (topk_multi needs seven): the key range
// comes from the producer (ph_overlap<true>), one 8-bit histogram pass finds the bin of
// the k-th key, every CTA publishes how many of its keys lie in higher bins, the bin's
// members (<= 1024) are gathered, and then every CTA -- redundantly, from the same data
// -- ranks them, derives the exact k-th (key, index) and its own output offset.  Falls
// back to topk_multi for degenerate inputs (all keys equal, > 1024 keys in the bin).
// Workspace: ctx.topk_ws + TK2_BASE (see the TK2_* layout above).
// ---------------------------------------------------------------------------------
__device__ __noinline__ void topk_grid(const bh_ctx& c, const unsigned long long* keys, const int n, const int k, int* out,
                          const int* map, uint8_t* flags, int b, int nb, GridBar& bar) {
  TopkScratch& sm = topk_scratch();
  int* hist = sm.hist;
  unsigned long long* cand_key = sm.cand_key;
  int* cand_idx = sm.cand_idx;
  __shared__ int s_scan[32];
  __shared__ unsigned long long s_u64[32];
  __shared__ int s_bin, s_rem, s_ncand, s_kth_idx;
  __shared__ unsigned long long s_kth_key;
  int* ws = c.topk_ws + TK2_BASE;
  unsigned long long* w64 = reinterpret_cast<unsigned long long*>(ws);
  const int t = threadIdx.x, NT = blockDim.x;
  const Range rg = block_range(n, b, nb);
  const bool valid = ws[TK2_VALID] != 0;
  const int par = ws[TK2_CALL] & 1;
#ifdef BH_TOPK_STAMPS
  unsigned long long* tk_stamps = reinterpret_cast<unsigned long long*>(c.blk + 7 * BH_BLK_STRIDE) + 40;
  int tk_i = 0;
#define TK_STAMP()                                              \
  do {                                                          \
    if (b == 0 && t == 0) {                                     \
      unsigned long long t_;                                    \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));    \
      tk_stamps[tk_i] = t_;                                     \
    }                                                           \
    ++tk_i;                                                     \
  } while (0)
#else
#define TK_STAMP() do {} while (0)
#endif
  TK_STAMP();
  if (nb > TK2_MAX_CTAS) {
    topk_multi(c, keys, n, k, out, map, flags, b, nb, bar);
    return;
  }
  if (!valid) {  // nobody published the key range: one extra pass and barrier
    unsigned long long mn = ~0ull, mx = 0ull;
#pragma unroll 1
    for (int j = rg.begin + t; j < rg.end; j += NT) {
      const unsigned long long key = keys[j];
      mn = key < mn ? key : mn;
      mx = key > mx ? key : mx;
    }
    mn = block_reduce_u64(mn, false, s_u64);
    mx = block_reduce_u64(mx, true, s_u64);
    if (t == 0 && rg.begin < rg.end) {
      atomicMax(&w64[0], mx);
      atomicMax(&w64[1], ~mn);
    }
    grid_barrier(bar, nb);
  }
  const unsigned long long mx = w64[0], mn = ~w64[1];
  const int consumed = (mn == mx) ? 64 : __clzll((long long)(mn ^ mx));
  const int shift = (64 - consumed - 11) > 0 ? (64 - consumed - 11) : 0;
  const int width = 64 - consumed - shift;
  const unsigned dmask = (1u << width) - 1u;
  int* ghist = ws + TK2_PAR0 + par * TK2_PSIZE;
  int* gcount = ghist + TK2_BINS;
  int* gidx = ghist + TK2_BINS + 8;
  unsigned long long* gkey = reinterpret_cast<unsigned long long*>(ghist + TK2_BINS + 8 + TOPK_THREADS);
  int* tab = ws + TK2_TAB;
  const bool degenerate = consumed == 64;  // all keys equal (uniform over the grid)

  // stage 1: local histogram (kept in shared memory for stage 2) -> global histogram
  if (!degenerate) {
#pragma unroll 1
    for (int i = t; i < TK2_BINS; i += NT) hist[i] = 0;
    __syncthreads();
#pragma unroll 1
    for (int j = rg.begin + t; j < rg.end; j += NT) atomicAdd(&hist[(int)((keys[j] >> shift) & dmask)], 1);
    __syncthreads();
#pragma unroll 1
    for (int i = t; i < TK2_BINS; i += NT) {
      const int h = hist[i];
      if (h) atomicAdd(&ghist[i], h);
    }
  }
  TK_STAMP();
  grid_barrier(bar, nb);
  TK_STAMP();
  // stage 2: the bin of the k-th key (every CTA, from the same global histogram: bins in descending
  // order, an ordered block scan); this CTA's keys in higher bins; gather the bin's members
  if (!degenerate) {
    const int per = (TK2_BINS + NT - 1) / NT;
    int sum = 0;
    for (int i = 0; i < per; ++i) {
      const int bin = TK2_BINS - 1 - (t * per + i);
      sum += bin >= 0 ? ghist[bin] : 0;
    }
    int total;
    const int before = block_excl_scan(sum, s_scan, total);
    if (before < k && before + sum >= k) {
      int r = k - before;
      for (int i = 0; i < per; ++i) {
        const int bin = TK2_BINS - 1 - (t * per + i);
        const int h = bin >= 0 ? ghist[bin] : 0;
        if (r > 0 && h >= r) {
          s_bin = bin;
          s_rem = r;
          s_ncand = h;
          r = -1;
        } else if (r > 0) {
          r -= h;
        }
      }
    }
  }
  __syncthreads();
  const bool fallback = degenerate || s_ncand > TOPK_THREADS;  // uniform over the grid
  const int bin = s_bin, rem = s_rem;
  if (!fallback) {
    int above = 0;  // this CTA's keys in bins above `bin` (its local histogram is still in shared memory)
#pragma unroll 1
    for (int i = bin + 1 + t; i < TK2_BINS; i += NT) above += hist[i];
    above = block_sum(above, s_scan);
    if (t == 0) tab[b] = above;
#pragma unroll 1
    for (int j = rg.begin + t; j < rg.end; j += NT) {
      const unsigned long long key = keys[j];
      if ((int)((key >> shift) & dmask) == bin) {
        const int p = atomicAdd(gcount, 1);
        gkey[p] = key;
        gidx[p] = j;
      }
    }
  }
  TK_STAMP();
  grid_barrier(bar, nb);
  TK_STAMP();
  if (b == 0) {  // leave the OTHER buffer and the key range clean for the next call; advance the parity
    int* oh = ws + TK2_PAR0 + (par ^ 1) * TK2_PSIZE;
#pragma unroll 1
    for (int i = t; i < TK2_BINS + 8; i += NT) oh[i] = 0;
    if (t == 0) {
      w64[0] = 0ull;
      w64[1] = 0ull;
      ws[TK2_VALID] = 0;
      ws[TK2_CALL] = par ^ 1;
    }
  }
  if (fallback) {  // degenerate input: the general algorithm (own workspace)
    topk_multi(c, keys, n, k, out, map, flags, b, nb, bar);
    return;
  }
  // stage 3 (every CTA): exact k-th (key, index); own output offset; ordered write
  const int nc = *gcount;
#pragma unroll 1
  for (int i = t; i < nc; i += NT) {
    cand_key[i] = gkey[i];
    cand_idx[i] = gidx[i];
  }
  __syncthreads();
  if (t < nc) {
    const unsigned long long mk = cand_key[t];
    const int mi = cand_idx[t];
    int ahead = 0;
#pragma unroll 2
    for (int i = 0; i < nc; ++i) {
      const unsigned long long ok = cand_key[i];
      ahead += (ok > mk || (ok == mk && cand_idx[i] < mi)) ? 1 : 0;
    }
    if (ahead == rem - 1) {
      s_kth_key = mk;
      s_kth_idx = mi;
    }
  }
  __syncthreads();
  const unsigned long long kth_key = s_kth_key;
  const int kth_idx = s_kth_idx;
  TK_STAMP();
  int before = 0;
  if (t < nc) {  // selected members of the bin that precede this CTA's range
    const unsigned long long mk = cand_key[t];
    const int mi = cand_idx[t];
    before = (mi < rg.begin && (mk > kth_key || (mk == kth_key && mi <= kth_idx))) ? 1 : 0;
  }
#pragma unroll 1
  for (int i = t; i < b; i += NT) before += tab[i];  // keys of earlier ranges in higher bins
  int base_sel = block_sum(before, s_scan);
#pragma unroll 1
  for (int tile = rg.begin; tile < rg.end; tile += NT) {
    const int j = tile + t;
    const unsigned long long key = j < rg.end ? keys[j] : 0ull;
    const bool take = j < rg.end && (key > kth_key || (key == kth_key && j <= kth_idx));
    int sel_total;
    const int pos = base_sel + block_excl_scan(take ? 1 : 0, s_scan, sel_total);
    if (take && pos < k) {
      const int col = map ? map[j] : j;
      out[pos] = col;
      if (flags) flags[col] = 1;
    }
    base_sel += sel_total;
  }
  TK_STAMP();
  if (b == 0 && t == 0) {
#ifdef BH_TOPK_STAMPS
    tk_stamps[15] = (unsigned long long)nc | ((unsigned long long)valid << 32);
#endif
  }
#undef TK_STAMP
}

// ---------------------------------------------------------------------------------
// (b) grid-wide selection from the histogram the overlap phase left (TK3_*): ONE grid barrier.  Every CTA
// finds the bin of the k-th key from the finished global histogram, counts its own keys in higher bins and
// gathers the members of that bin; after the barrier every CTA ranks the members, derives its output offset
// and writes its part of the ordered list (as topk_grid stage 3).  Falls back to topk_grid when the
// prediction missed.  keys[j] is column j (unsharded networks).
// ---------------------------------------------------------------------------------
__device__ __noinline__ void topk_grid_hist(const bh_ctx& c, const unsigned long long* keys, const int n, const int k, int* out,
                               uint8_t* flags, int b, int nb, GridBar& bar, int step_ov = -1) {
  TopkScratch& sm = topk_scratch();
  int* hist = sm.hist;
  unsigned long long* cand_key = sm.cand_key;
  int* cand_idx = sm.cand_idx;
  __shared__ int s_scan[32];
  __shared__ int s_bin, s_rem, s_ncand, s_kth_idx;
  __shared__ unsigned long long s_kth_key;
  int* ws3 = c.topk_ws + TK3_BASE;
  int* ws = c.topk_ws + TK2_BASE;
  const int t = threadIdx.x, NT = blockDim.x;
  const int step = step_ov >= 0 ? step_ov : c.sc[BH_SC_STEP], par = step & 1;
  int* ghist = ws3 + TK3_HIST + par * TK2_BINS;
  int* gcount = ws3 + TK3_CNT + par * 8;
  int* gidx = ws3 + TK3_IDX + par * 1024;
  unsigned long long* gkey = reinterpret_cast<unsigned long long*>(ws3 + TK3_KEY) + par * 1024;
  int* tab = ws3 + TK3_TAB;
  const Range rg = block_range(n, b, nb);
  bool ok = ws3[TK3_READY] == step + 1 && nb <= TK2_MAX_CTAS;
  if (t == 0) s_ncand = -1;
  __syncthreads();
  if (ok) {  // bins in descending order: the bin in which the cumulative count reaches k
    const int per = (TK2_BINS + NT - 1) / NT;
    int sum = 0;
    for (int i = 0; i < per; ++i) {
      const int bin = TK2_BINS - 1 - (t * per + i);
      const int h = bin >= 1 ? ghist[bin] : 0;
      hist[t * per + i] = h;  // (descending order; reused below)
      sum += h;
    }
    int total;
    const int before = block_excl_scan(sum, s_scan, total);
    if (before < k && before + sum >= k) {
      int r = k - before;
      for (int i = 0; i < per; ++i) {
        const int h = hist[t * per + i];
        if (r > 0 && h >= r) {
          s_bin = TK2_BINS - 1 - (t * per + i);
          s_rem = r;
          s_ncand = h;
          r = -1;
        } else if (r > 0) {
          r -= h;
        }
      }
    }
    __syncthreads();
    ok = s_ncand >= 0 && s_ncand <= TOPK_THREADS;  // threshold inside the binned range, few keys in its bin
  }
  if (!ok) {  // uniform over the grid: the general selection (own histogram pass)
    if (b == 0 && t == 0) ws3[TK3_READY] = 0;
    topk_grid(c, keys, n, k, out, nullptr, flags, b, nb, bar);
    return;
  }
  const int bin = s_bin, rem = s_rem;
  const Tk3Binning binning = tk3_binning(ws3);
  int above = 0;
#pragma unroll 1
  for (int j = rg.begin + t; j < rg.end; j += NT) {
    const unsigned long long key = keys[j];
    const int kb = tk3_bin(binning, key);
    above += kb > bin ? 1 : 0;
    if (kb == bin) {
      const int p = atomicAdd(gcount, 1);
      gkey[p] = key;
      gidx[p] = j;
    }
  }
  above = block_sum(above, s_scan);
  if (t == 0) tab[b] = above;
  grid_barrier(bar, nb);
  const int nc = *gcount;
#pragma unroll 1
  for (int i = t; i < nc; i += NT) {
    cand_key[i] = gkey[i];
    cand_idx[i] = gidx[i];
  }
  __syncthreads();
  if (t < nc) {
    const unsigned long long mk = cand_key[t];
    const int mi = cand_idx[t];
    int ahead = 0;
#pragma unroll 2
    for (int i = 0; i < nc; ++i) {
      const unsigned long long okey = cand_key[i];
      ahead += (okey > mk || (okey == mk && cand_idx[i] < mi)) ? 1 : 0;
    }
    if (ahead == rem - 1) {
      s_kth_key = mk;
      s_kth_idx = mi;
    }
  }
  __syncthreads();
  const unsigned long long kth_key = s_kth_key;
  const int kth_idx = s_kth_idx;
  int before = 0;
  if (t < nc) {  // selected members of the bin that precede this CTA's range
    const unsigned long long mk = cand_key[t];
    const int mi = cand_idx[t];
    before = (mi < rg.begin && (mk > kth_key || (mk == kth_key && mi <= kth_idx))) ? 1 : 0;
  }
#pragma unroll 1
  for (int i = t; i < b; i += NT) before += tab[i];
  int base_sel = block_sum(before, s_scan);
#pragma unroll 1
  for (int tile = rg.begin; tile < rg.end; tile += NT) {
    const int j = tile + t;
    const unsigned long long key = j < rg.end ? keys[j] : 0ull;
    const bool take = j < rg.end && (key > kth_key || (key == kth_key && j <= kth_idx));
    int sel_total;
    const int pos = base_sel + block_excl_scan(take ? 1 : 0, s_scan, sel_total);
    if (take && pos < k) {
      out[pos] = j;
      if (flags) flags[j] = 1;
    }
    base_sel += sel_total;
  }
  if (b == nb - 1) {  // (the last CTA: after its own reads) leave the workspaces clean for the next call
    __syncthreads();
    int* oh = ws3 + TK3_HIST + (par ^ 1) * TK2_BINS;
#pragma unroll 1
    for (int i = t; i < TK2_BINS; i += NT) oh[i] = 0;
    if (t == 0) {
      ws3[TK3_CNT + (par ^ 1) * 8] = 0;
      unsigned long long* w64 = reinterpret_cast<unsigned long long*>(ws);
      const unsigned long long mx = w64[0];
      w64[0] = 0ull;
      w64[1] = 0ull;
      ws[TK2_VALID] = 0;
      tk3_set_binning(ws3, kth_key, mx > kth_key ? mx : kth_key);
    }
  }
}

// After a selection that did not set the binning itself (topk_grid fallback): from the selected keys.
// One CTA, in a later phase (the selection's output must be complete).
__device__ __forceinline__ void tk3_rebin_from_selection(const bh_ctx& c, const int* act_ov = nullptr, int step_ov = -1) {
  __shared__ unsigned long long s_u64b[32];
  int* ws3 = c.topk_ws + TK3_BASE;
  const int step = step_ov >= 0 ? step_ov : c.sc[BH_SC_STEP];
  if (ws3[TK3_READY] == step + 1) return;  // topk_grid_hist succeeded and set it
  const int k = c.active_columns;
  const int* act = act_ov ? act_ov : c.active_cols + (step & 1) * k;
  const unsigned long long* keys = reinterpret_cast<const unsigned long long*>(c.boosted);
  unsigned long long mn = ~0ull, mx = 0ull;
#pragma unroll 1
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const unsigned long long key = keys[act[i]];
    mn = key < mn ? key : mn;
    mx = key > mx ? key : mx;
  }
  mn = block_reduce_u64(mn, false, s_u64b);
  mx = block_reduce_u64(mx, true, s_u64b);
  if (threadIdx.x == 0) tk3_set_binning(ws3, mn, mx);
  // the histogram of the NEXT step (other parity) must start empty
  int* oh = ws3 + TK3_HIST + ((step & 1) ^ 1) * TK2_BINS;
#pragma unroll 1
  for (int i = threadIdx.x; i < TK2_BINS; i += blockDim.x) oh[i] = 0;
  int* mine = ws3 + TK3_HIST + (step & 1) * TK2_BINS;  // and this step's was not consumed
#pragma unroll 1
  for (int i = threadIdx.x; i < TK2_BINS; i += blockDim.x) mine[i] = 0;
  if (threadIdx.x < 2) ws3[TK3_CNT + threadIdx.x * 8] = 0;  // (a member count of two steps ago may be left)
}

// The same for a column shard after a candidate exchange: the gathered candidates (xk_keys / xk_cols, n of
// them, identical on every rank) contain every selected column.  One CTA.
__device__ __forceinline__ void tk3_rebin_sharded(const bh_ctx& c, int n) {
  __shared__ unsigned long long s_u64c[32];
  int* ws3 = c.topk_ws + TK3_BASE;
  const unsigned long long* keys = reinterpret_cast<const unsigned long long*>(c.xk_keys);
  unsigned long long mn = ~0ull, mx = 0ull;
#pragma unroll 1
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned long long key = keys[i];
    mx = key > mx ? key : mx;
    if (c.col_active[c.xk_cols[i]]) mn = key < mn ? key : mn;
  }
  mn = block_reduce_u64(mn, false, s_u64c);
  mx = block_reduce_u64(mx, true, s_u64c);
  if (threadIdx.x == 0) tk3_set_binning(ws3, mn, mx > mn ? mx : mn);
#pragma unroll 1
  for (int i = threadIdx.x; i < 2 * TK2_BINS; i += blockDim.x) ws3[TK3_HIST + i] = 0;
  if (threadIdx.x == 0) ws3[TK3_READY] = 0;
}

// stand-alone cooperative kernels built on topk_multi (grid = one CTA per SM)
__global__ void __launch_bounds__(TOPK_THREADS, 1) k_topk_multi(const __grid_constant__ bh_ctx c) {
  GridBar bar = grid_bar_open(reinterpret_cast<unsigned int*>(c.sc + BH_SC_BAR_COUNT));
  if (blockIdx.x == 0) retire_prev_flags(c);
  const int k = c.active_columns;
  topk_grid(c, reinterpret_cast<const unsigned long long*>(c.boosted), c.column_dim, k,
            c.active_cols + (c.sc[BH_SC_STEP] & 1) * k, nullptr, c.col_active, blockIdx.x, gridDim.x, bar);
}

__global__ void __launch_bounds__(TOPK_THREADS, 1)
    k_topk_shard_local_multi(const __grid_constant__ bh_ctx c, int* scratch, double* cand_keys, int32_t* cand_cols) {
  GridBar bar = grid_bar_open(reinterpret_cast<unsigned int*>(c.sc + BH_SC_BAR_COUNT));
  const int k_loc = c.active_columns < c.col_local ? c.active_columns : c.col_local;
  topk_grid(c, reinterpret_cast<const unsigned long long*>(c.boosted), c.col_local, k_loc, scratch, nullptr, nullptr,
            blockIdx.x, gridDim.x, bar);
  grid_barrier(bar, gridDim.x);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k_loc; i += gridDim.x * blockDim.x) {
    const int pos = scratch[i];
    cand_keys[i] = c.boosted[pos];
    cand_cols[i] = c.col_lo + pos;
  }
}

// host-inhibition mode: adopt an explicit ordered list
__global__ void k_set_active(const __grid_constant__ bh_ctx c, const int32_t* __restrict__ cols) {
  const int k = c.active_columns;
  const int cur = c.sc[BH_SC_STEP] & 1;
  int* out = c.active_cols + cur * k;
  const int* prev = c.active_cols + (cur ^ 1) * k;
  #pragma unroll 1
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    c.col_active[prev[i]] = 0;
    c.col_active[out[i]] = 0;  // in case bh_inhibit already ran this step
  }
  __syncthreads();
  #pragma unroll 1
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
__device__ __forceinline__ void sp_learn_wide(const bh_ctx& c, const uint32_t* input, int b, int nb, int step_ov = -1) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const int k = c.active_columns, I = c.input_dim, words = c.input_words;
  const int cur = (step_ov >= 0 ? step_ov : c.sc[BH_SC_STEP]) & 1;
  const int* act = c.active_cols + cur * k;
  const double d_on = c.sp_delta_on, d_off = c.sp_delta_off, thr = c.sp_threshold;
  // work unit = a quarter / half of a long row, so that k rows spread evenly over any number of CTAs
  // (only when a CTA would otherwise get few rows: streaming whole rows is faster -- 63 vs 75 us at cfg3)
  const long long rows_here = (long long)k * c.col_local / c.column_dim;  // active columns of this shard, on average
  const int parts = rows_here >= 4LL * nb ? 1 : (words >= 512 ? 4 : (words >= 256 ? 2 : 1));
  const int part_words = ((words + parts - 1) / parts + 3) & ~3;
#pragma unroll 1
  for (int u = b; u < k * parts; u += nb) {
    const int r = u / parts, part = u - r * parts;
    const int col = act[r] - c.col_lo;  // local row; columns of other shards are skipped
    if (col < 0 || col >= c.col_local) continue;
    double* prow = c.sp_perm + (long long)col * I;
    uint32_t* mrow = c.sp_mask + (long long)col * c.mask_stride;
    const int w_end = (part + 1) * part_words < words ? (part + 1) * part_words : words;
#pragma unroll 1
    for (int w0 = part * part_words + warp * 4; w0 < w_end; w0 += warps * 4) {  // 4 x 256 B of permanence in flight per warp
      double p[4];
      uint32_t xin[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int w = w0 + j, i = w * 32 + lane;
        xin[j] = w < words ? input[w] : 0u;
        p[j] = (w < words && i < I) ? prow[i] : -1.0;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int w = w0 + j, i = w * 32 + lane;
        bool on = false;
        if (w < words && i < I) {
          const double q = __dadd_rn(p[j], ((xin[j] >> lane) & 1u) ? d_on : d_off);
          prow[i] = q;
          on = q >= thr;
        }
        const uint32_t bits = __ballot_sync(BH_FULL, on);
        if (lane == 0 && w < words) mrow[w] = bits;
      }
    }
  }
}

// Short rows (a warp covers 4 x 32 inputs per iteration, so a row needs only a few warps): the warps of
// a CTA form teams that work on different rows at the same time.
__device__ __forceinline__ void sp_learn_grouped(const bh_ctx& c, const uint32_t* input, int b, int nb, int step_ov = -1) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const int k = c.active_columns, I = c.input_dim, words = c.input_words;
  const int cur = (step_ov >= 0 ? step_ov : c.sc[BH_SC_STEP]) & 1;
  const int* act = c.active_cols + cur * k;
  const double d_on = c.sp_delta_on, d_off = c.sp_delta_off, thr = c.sp_threshold;
  const int wpr = (words + 3) / 4;  // warps that have work on one row
  const int groups = warps / wpr, grp = warp / wpr, wig = warp - grp * wpr;
  if (grp >= groups) return;  // leftover warps
#pragma unroll 1
  for (int r = b * groups + grp; r < k; r += nb * groups) {
    const int col = act[r] - c.col_lo;  // local row; columns of other shards are skipped
    if (col < 0 || col >= c.col_local) continue;
    double* prow = c.sp_perm + (long long)col * I;
    uint32_t* mrow = c.sp_mask + (long long)col * c.mask_stride;
    const int w0 = wig * 4;
    double p[4];
    uint32_t xin[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int w = w0 + j, i = w * 32 + lane;
      xin[j] = w < words ? input[w] : 0u;
      p[j] = (w < words && i < I) ? prow[i] : -1.0;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int w = w0 + j, i = w * 32 + lane;
      bool on = false;
      if (w < words && i < I) {
        const double q = __dadd_rn(p[j], ((xin[j] >> lane) & 1u) ? d_on : d_off);
        prow[i] = q;
        on = q >= thr;
      }
      const uint32_t bits = __ballot_sync(BH_FULL, on);
      if (lane == 0 && w < words) mrow[w] = bits;
    }
  }
}

// SHORT_ROWS: also compile the team variant (kernels that serve small networks); the HBM-bound grid
// kernel keeps only the wide loop (the extra code costs it registers and 5 us per step at cfg3).
template <bool SHORT_ROWS = true>
__device__ __forceinline__ void ph_sp_learn(const bh_ctx& c, const uint32_t* input, int b, int nb, int step_ov = -1) {
  if (SHORT_ROWS && (c.input_words + 3) / 4 * 2 <= (int)(blockDim.x >> 5)) sp_learn_grouped(c, input, b, nb, step_ov);
  else sp_learn_wide(c, input, b, nb, step_ov);
}

__global__ void __launch_bounds__(SP_THREADS) k_sp_learn(const __grid_constant__ bh_ctx c, const uint32_t* input) {
  ph_sp_learn<true>(c, input, blockIdx.x, gridDim.x);
}

// (c) duty-cycle EMA: two separately rounded float32 operations.
// regularizations.py:19-21; runs even when learning is off (networks.py:33).
__device__ __forceinline__ void ph_duty(const bh_ctx& c, int b, int nb) {
  #pragma unroll 1
  for (int j = b * blockDim.x + threadIdx.x; j < c.col_local; j += nb * blockDim.x) {
    float d = __fmul_rn(c.duty[j], c.duty_momentum);
    if (c.col_active[c.col_lo + j]) d = __fadd_rn(d, c.duty_increment);
    c.duty[j] = d;
  }
}

__global__ void k_duty_update(const __grid_constant__ bh_ctx c) { ph_duty(c, blockIdx.x, gridDim.x); }
