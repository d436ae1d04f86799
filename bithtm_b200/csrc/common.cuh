// Shared device helpers for the bithtm_b200 kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bithtm_b200.h"

#define BH_FULL 0xffffffffu
#define BH_TM_THREADS 256            // block size of the ranged TM kernels
#define BH_TM_WARPS (BH_TM_THREADS / 32)
#define BH_BLK_STRIDE 1024           // ints per row of ctx.blk

// rows of ctx.blk (per-CTA counts used for ordered, deterministic compaction)
enum { BLK_WIN = 0, BLK_UNACC, BLK_LEARN, BLK_PUNISH, BLK_MATCH, BLK_RECYC, BLK_ROWS = 8 };

struct Range {
  int begin, end;
};

// Contiguous share of [0, n) owned by block b of nb (ascending by block: the
// concatenation over blocks is the original order).
__device__ __forceinline__ Range block_range(int n, int b, int nb) {
  int chunk = (n + nb - 1) / nb;
  long long lo = (long long)b * chunk;
  Range r;
  r.begin = lo < n ? (int)lo : n;
  r.end = (r.begin + chunk < n) ? r.begin + chunk : n;
  return r;
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(BH_FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(BH_FULL, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(BH_FULL, v, o));
  return v;
}

// Block-wide sum; every thread gets the total.  `sm` holds >= 32 ints.
__device__ __forceinline__ int block_sum(int v, int* sm) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect sm from a previous use
  if (lane == 0) sm[w] = v;
  __syncthreads();
  int t = 0;
  for (int i = 0; i < nw; ++i) t += sm[i];
  return t;
}

// Block-wide exclusive prefix of a per-thread count (thread order); `total`
// receives the block sum.  `sm` holds >= 32 ints.  All threads must call.
__device__ __forceinline__ int block_excl_scan(int v, int* sm, int& total) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(BH_FULL, inc, o);
    if (lane >= o) inc += n;
  }
  __syncthreads();
  if (lane == 31) sm[w] = inc;
  __syncthreads();
  int off = 0, tot = 0;
  for (int i = 0; i < nw; ++i) {
    int s = sm[i];
    if (i < w) off += s;
    tot += s;
  }
  total = tot;
  return off + inc - v;
}

// Sum of counts[0..b) and of counts[0..nb) (nb <= 1024), computed by the whole
// block.  `sm` holds >= 32 ints.
__device__ __forceinline__ void blk_prefix(const int* counts, int b, int nb, int* sm, int& before, int& total) {
  int pre = 0, all = 0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    int v = counts[i];
    all += v;
    if (i < b) pre += v;
  }
  before = block_sum(pre, sm);
  total = block_sum(all, sm);
}

__device__ __forceinline__ bool cell_bit(const uint32_t* col_words, int cell, int c) {
  int col = cell / c;
  return (__ldg(col_words + col) >> (cell - col * c)) & 1u;
}

__device__ __forceinline__ uint32_t low_mask(int c) { return c >= 32 ? 0xffffffffu : ((1u << c) - 1u); }
