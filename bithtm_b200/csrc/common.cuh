// Shared device helpers for the bithtm_b200 kernels (sm_100a).
//
// Every stage of the timestep is a `__device__` phase function `ph_*(ctx, b, nb, ...)`
// executed by CTA `b` of `nb` cooperating CTAs with blockDim.x threads each.  The
// same phase code runs (i) as its own kernel behind the fine-grained C entry points
// and (ii) inside the fused step kernel, where phases are separated by a cluster
// barrier (small networks) or a grid barrier (large ones).  Buffers written during a
// step are never read through the non-coherent path (no __ldg / ld.global.nc).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bithtm_b200.h"

#define BH_FULL 0xffffffffu
#define BH_TM_THREADS 256  // block size of the stand-alone ranged TM kernels
#define BH_BLK_STRIDE 1024  // ints per row of ctx.blk
#define BH_MAX_WARPS 32

// rows of ctx.blk (per-CTA counts used for ordered, deterministic compaction)
enum { BLK_WIN = 0, BLK_UNACC, BLK_LEARN, BLK_PUNISH, BLK_MATCH, BLK_RECYC, BLK_ROWS = 8 };
#define BLK(c, row) ((c).blk + (row)*BH_BLK_STRIDE)

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

// Block-wide sum; every thread gets the total.  `sm` holds >= 32 ints.  The
// cross-warp stage is a redundant warp shuffle reduction in every warp (small code,
// no third barrier).
__device__ __noinline__ int block_sum(int v, int* sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect sm from a previous use
  if (lane == 0) sm[w] = v;
  __syncthreads();
  return warp_sum(lane < nw ? sm[lane] : 0);
}

// Three block-wide sums with one pair of barriers; thread 0's results are valid (all threads' are).
// `sm` holds >= 96 ints.
__device__ __noinline__ void block_sum3(int& a, int& b, int& c, int* sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  a = warp_sum(a);
  b = warp_sum(b);
  c = warp_sum(c);
  __syncthreads();  // protect sm from a previous use
  if (lane == 0) {
    sm[w] = a;
    sm[32 + w] = b;
    sm[64 + w] = c;
  }
  __syncthreads();
  a = warp_sum(lane < nw ? sm[lane] : 0);
  b = warp_sum(lane < nw ? sm[32 + lane] : 0);
  c = warp_sum(lane < nw ? sm[64 + lane] : 0);
}

// Block-wide exclusive prefix of a per-thread count (thread order); `total`
// receives the block sum.  `sm` holds >= 32 ints.  All threads must call.
__device__ __noinline__ int block_excl_scan(int v, int* sm, int& total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(BH_FULL, inc, o);
    if (lane >= o) inc += n;
  }
  __syncthreads();
  if (lane == 31) sm[w] = inc;
  __syncthreads();
  int ws = lane < nw ? sm[lane] : 0;  // per-warp sums, scanned redundantly by every warp
  int winc = ws;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(BH_FULL, winc, o);
    if (lane >= o) winc += n;
  }
  total = __shfl_sync(BH_FULL, winc, 31);
  const int off = __shfl_sync(BH_FULL, winc - ws, w);
  return off + inc - v;
}

// Sum of counts[0..b) and of counts[0..nb) (nb <= 1024), computed by the whole
// block.  `sm` holds >= 32 ints.
__device__ __noinline__ void blk_prefix(const int* counts, int b, int nb, int* sm, int& before, int& total) {
  int pre = 0, all = 0;
  #pragma unroll 1
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    int v = counts[i];
    all += v;
    if (i < b) pre += v;
  }
  before = block_sum(pre, sm);
  total = block_sum(all, sm);
}

// The same for three count arrays at once (one pass, one pair of barriers).  `sm` holds >= 192 ints.
__device__ __noinline__ void blk_prefix3(const int* c0, const int* c1, const int* c2, int b, int nb, int* sm,
                                         int (&before)[3], int (&total)[3]) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int v[6] = {0, 0, 0, 0, 0, 0};
  #pragma unroll 1
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    const int a0 = c0[i], a1 = c1[i], a2 = c2[i];
    v[1] += a0;
    v[3] += a1;
    v[5] += a2;
    if (i < b) {
      v[0] += a0;
      v[2] += a1;
      v[4] += a2;
    }
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) v[j] = warp_sum(v[j]);
  __syncthreads();  // protect sm from a previous use
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < 6; ++j) sm[j * 32 + w] = v[j];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 6; ++j) v[j] = warp_sum(lane < nw ? sm[j * 32 + lane] : 0);
  before[0] = v[0];
  total[0] = v[1];
  before[1] = v[2];
  total[1] = v[3];
  before[2] = v[4];
  total[2] = v[5];
}

// device cell id = column * 32 + cell-in-column (c <= 32)
__device__ __forceinline__ bool cell_bit(const uint32_t* col_words, int cell) {
  return (col_words[cell >> 5] >> (cell & 31)) & 1u;
}

// ------------------------------------------------------------------------------------
// segment shards: ids are dealt to ranks in blocks of 64 (block b -> rank b % world);
// a rank stores the synapse rows of its segments compactly, in ascending id order.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ bool seg_held(const bh_ctx& c, int s) {
  return c.seg_world <= 1 || ((s >> 6) % c.seg_world) == c.seg_rank;
}
__device__ __forceinline__ int seg_row(const bh_ctx& c, int s) {  // local row of a held segment
  return c.seg_world <= 1 ? s : ((((s >> 6) / c.seg_world) << 6) | (s & 63));
}
__device__ __forceinline__ int seg_gid(const bh_ctx& c, int row) {  // segment id of a local row
  return c.seg_world <= 1 ? row : ((((row >> 6) * c.seg_world + c.seg_rank) << 6) | (row & 63));
}
__device__ __forceinline__ int seg_local_count(const bh_ctx& c, int S) {  // held ids below S
  if (c.seg_world <= 1) return S;
  const int nblk = S >> 6, rem = S & 63, q = nblk / c.seg_world, r = nblk % c.seg_world;
  return (q + (c.seg_rank < r ? 1 : 0)) * 64 + (c.seg_rank == r ? rem : 0);
}

__device__ __forceinline__ uint32_t low_mask(int c) { return c >= 32 ? 0xffffffffu : ((1u << c) - 1u); }

// ------------------------------------------------------------------------------------
// barriers between phases of the fused step kernel
// ------------------------------------------------------------------------------------
// All CTAs of the (single) thread-block cluster; release/acquire at cluster scope
// makes global-memory writes of earlier phases visible (and invalidates L1).
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// All CTAs of a cooperative (co-resident) grid: self-resetting counter + generation
// word in global memory.  `bar[0]` = arrivals, `bar[1]` = generation.
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int n_ctas) {
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int gen;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(bar + 1) : "memory");
    __threadfence();
    unsigned int prev = atomicAdd(bar, 1u);
    if (prev == n_ctas - 1) {
      bar[0] = 0u;
      __threadfence();
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar + 1) : "memory");
    } else {
      unsigned int now;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(now) : "l"(bar + 1) : "memory");
      } while (now == gen);
    }
    __threadfence();
  }
  __syncthreads();
}

// The same barrier with the generation word cached by the CTA: a kernel opens the barrier once (one load of the
// generation, before its first arrival -- nobody can complete a barrier this CTA has not arrived at) and every
// wait then costs ONE round trip (the arrival) + the spin, instead of load + arrival + spin.
struct GridBar {
  unsigned int* w;   // [0] arrivals, [1] generation
  unsigned int gen;  // the generation this CTA is in (thread 0)
};
__device__ __forceinline__ GridBar grid_bar_open(unsigned int* words) {
  GridBar g;
  g.w = words;
  g.gen = 0u;
  if (threadIdx.x == 0) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g.gen) : "l"(words + 1) : "memory");
  return g;
}
__device__ __forceinline__ void grid_barrier(GridBar& g, unsigned int n_ctas) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(g.w, 1u);
    if (prev == n_ctas - 1) {
      g.w[0] = 0u;
      __threadfence();
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(g.w + 1) : "memory");
    } else {
      unsigned int now;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(now) : "l"(g.w + 1) : "memory");
      } while (now == g.gen);
    }
    g.gen += 1u;
    __threadfence();
  }
  __syncthreads();
}

