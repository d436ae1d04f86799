// PROTOTYPE (stand-alone program; not part of libbithtm_b200.so, nothing in the product path includes it).
// Run on a B200 at the very end of round 1: results BIT-IDENTICAL to the popcount reference kernel at
// 1024 x 2048 x 1024 (18.9 us) and at 256 x 65536 x 16384 (638 us = 862 int8 TOP/s; the shipped
// mma.sync kernel: 12.5 us / 747 us).  Tried with the last GPU seconds, both without effect: 16 instead
// of 8 producer warps (648 us), three stages of register look-ahead for the global words (657 us) --
// Reading the SASS afterwards explains both: `fence.proxy.async.shared::cta` is MEMBAR.ALL.CTA +
// FENCE.VIEW.ASYNC, and the MEMBAR also drains the thread's outstanding global loads, so every stage pays a
// full L2 / HBM round trip whatever the look-ahead (1.25 us per 128x256x128 stage measured).  Variant 1 below
// (producer groups that own ring slots, loads issued after the fence) is the fix; it compiles but has NOT
// been run yet.  Further: loads by dedicated warps or TMA into a raw-word ring; 128-byte-swizzled tiles; a
// dedicated epilogue warpgroup with a double-buffered accumulator; tile order against the 3.46-wave tail.
// Shared-memory budget: 48 KB written + 48 KB read per stage against 128 B/clk -> floor ~750 clk per stage.
//
// Shared-mask batched overlap (DenseProjection.process for many inputs against one connected mask,
// bitHTM projections.py:18-21) as a tcgen05 int8 contraction:
//     out[b][j] = sum_i x[b][i] * m[j][i]        x, m in {0, 1}
// The shipped kernel (k_sp_overlap_batched_tc, mma.sync m16n8k32) reaches 64 % of the mma.sync int8
// peak (1.15 POP/s); tcgen05.mma kind::i8 has 4x that ceiling.  tcgen05 reads its operands from
// shared memory through descriptors, so the bit-packed rows must be widened to bytes IN shared
// memory (8x the bit traffic, by the CUDA cores) -- the design question this prototype answers is
// whether that widening keeps up with the tensor pipe.
//
// Structure (one persistent CTA per SM, 9 warps):
//   warps 0-7  producers: global bit words -> {0,1} bytes in the UMMA canonical K-major layout
//              (no swizzle: 8-row x 16-byte core matrices; in 16-byte units the tile is
//              ((8, rows/8), 2) : ((1, SBO), LBO), cute/atom/mma_traits_sm100.hpp make_umma_desc),
//              4-stage ring, full/empty mbarriers; they are also the epilogue (tcgen05.ld -> global)
//   warp 8     allocates TMEM (256 columns) and issues tcgen05.mma.cta_group::1.kind::i8,
//              M = 128 inputs, N = 256 columns, K = 32 bytes per instruction, accumulator in TMEM;
//              tcgen05.commit releases the stage / publishes the accumulator.
// A dot product does not care which bit sits at which k as long as both operands agree: byte j of
// register s (s = 0..7) of a 32-bit word is bit 8j + s, i.e. (w >> s) & 0x01010101 -- two integer
// ops per 4 bytes.
// Every mbarrier wait is bounded (1 s of globaltimer) and raises an error flag instead of hanging.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/_build/overlap_tcgen05 \
//        tools/experiments/overlap_tcgen05.cu
//   timeout 60 tools/_build/overlap_tcgen05 [B] [C] [I] [variant]
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int TILE_M = 128;                       // inputs per tile  (UMMA M)
constexpr int TILE_N = 256;                       // columns per tile (UMMA N)
constexpr int BLOCK_KW = 4;                       // 32-bit words of K per stage (128 int8 of K)
constexpr int STAGES = 4;
constexpr int CHUNKS = BLOCK_KW * 2;              // 16-byte K chunks per row and stage
constexpr int A_LBO = TILE_M * 16;                // bytes between K chunks (chunk-major tiles)
constexpr int B_LBO = TILE_N * 16;
constexpr int SBO = 128;                          // bytes between 8-row groups
constexpr int A_STAGE = CHUNKS * A_LBO;           // 16 KiB
constexpr int B_STAGE = CHUNKS * B_LBO;           // 32 KiB
constexpr int STAGE_BYTES = A_STAGE + B_STAGE;
#ifndef PRODUCER_WARPS_N
#define PRODUCER_WARPS_N 8
#endif
constexpr int PRODUCER_WARPS = PRODUCER_WARPS_N;  // multiple of 4 (TMEM lane quarters), divides 1536 / 32
constexpr int COLS_PER_WARP = TILE_N / (PRODUCER_WARPS / 4);
constexpr int PRODUCERS = PRODUCER_WARPS * 32;
constexpr int THREADS = PRODUCERS + 32;
constexpr int ITEMS = (TILE_M + TILE_N) * BLOCK_KW / PRODUCERS;  // (row, word) pairs per producer thread: 6
constexpr int TMEM_COLS = 256;
constexpr int SPIN_CAP = 1 << 26;
static_assert((TILE_M + TILE_N) * BLOCK_KW % PRODUCERS == 0, "items must divide");

// ---- PTX helpers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint64_t* b, uint32_t parity, int* err) {
  const uint32_t a = smem_u32(b);
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int it = 0; it < SPIN_CAP; ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    if (ok) return true;
    if ((it & 255) == 255) {  // every 256 polls: somebody else failed?  wall-clock bound: 1 s
      if (*(volatile int*)err) return false;
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 1000000000ull) break;
    }
  }
  atomicExch(err, 1);
  return false;
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // cute::UMMA::SmemDescriptor: start [0,14), LBO [16,30), SBO [32,46) (all >> 4), version [46,48) = 1,
  // base_offset 0, lbo_mode 0, layout_type [61,64) = 0 (SWIZZLE_NONE)
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// cute::UMMA::InstrDescriptor for kind::i8: c_format [4,6) = 2 (S32), a_format [7,10) = 0 (u8),
// b_format [10,13) = 0 (u8), a_major [15] = 0 (K), b_major [16] = 0 (K), n_dim [17,23) = N >> 3,
// m_dim [24,29) = M >> 4
constexpr uint32_t IDESC = (2u << 4) | ((uint32_t)(TILE_N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate) {
  const uint32_t zero = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(IDESC), "r"(accumulate), "r"(zero)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// ---- the kernel ----------------------------------------------------------------------------------
// mask [C][mask_stride] words, inputs [B][words] words, out [B][C] int32
__global__ void __launch_bounds__(THREADS, 1)
    k_overlap_tcgen05(const uint32_t* __restrict__ mask, int mask_stride, const uint32_t* __restrict__ inputs, int words,
                      int B, int C, int32_t* __restrict__ out, int* err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], acc_full, acc_empty;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool producer = warp < PRODUCER_WARPS;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], PRODUCER_WARPS);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_full, 1);
    mbar_init(&acc_empty, PRODUCER_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == PRODUCER_WARPS) {  // the MMA warp owns the TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  const int tiles_m = (B + TILE_M - 1) / TILE_M, tiles_n = (C + TILE_N - 1) / TILE_N;
  const int n_tiles = tiles_m * tiles_n;
  const int k_stages = (words + BLOCK_KW - 1) / BLOCK_KW;

  if (producer) {
    // item i of this thread: pair index p = i * PRODUCERS + tid -> word w = p & 3, row r = p >> 2
    // (r < TILE_M: input row, else mask row r - TILE_M)
    auto fetch = [&](int m0, int n0, int kw0, uint32_t (&wd)[ITEMS]) {
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        const int p = i * PRODUCERS + tid, w = kw0 + (p & 3), r = p >> 2;
        uint32_t v = 0;
        if (w < words) {
          if (r < TILE_M) {
            if (m0 + r < B) v = __ldg(inputs + (long long)(m0 + r) * words + w);
          } else if (n0 + r - TILE_M < C) {
            v = __ldg(mask + (long long)(n0 + r - TILE_M) * mask_stride + w);
          }
        }
        wd[i] = v;
      }
    };
    int it = 0;  // stages produced so far (ring position)
    int tile_iter = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_iter) {
      const int m0 = (tile % tiles_m) * TILE_M, n0 = (tile / tiles_m) * TILE_N;
      // register ring: the words of the next three stages are in flight while this one is widened (a stage
      // is shorter than one L2 / HBM round trip: with one stage of look-ahead the ring ran at 1.25 us per
      // stage = the load latency)
      uint32_t cur[ITEMS], n1[ITEMS], n2[ITEMS], n3[ITEMS];
      fetch(m0, n0, 0, cur);
      if (k_stages > 1) fetch(m0, n0, BLOCK_KW, n1);
      if (k_stages > 2) fetch(m0, n0, 2 * BLOCK_KW, n2);
      for (int ks = 0; ks < k_stages; ++ks, ++it) {
        if (ks + 3 < k_stages) fetch(m0, n0, (ks + 3) * BLOCK_KW, n3);
        const int s = it % STAGES;
        if (!mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1, err)) return;
        uint8_t* a_st = smem + s * STAGE_BYTES;
        uint8_t* b_st = a_st + A_STAGE;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
          const int p = i * PRODUCERS + tid, w = p & 3, r = p >> 2;
          const uint32_t x = cur[i];
          uint4 lo, hi;
          lo.x = x & 0x01010101u, lo.y = (x >> 1) & 0x01010101u, lo.z = (x >> 2) & 0x01010101u, lo.w = (x >> 3) & 0x01010101u;
          hi.x = (x >> 4) & 0x01010101u, hi.y = (x >> 5) & 0x01010101u, hi.z = (x >> 6) & 0x01010101u, hi.w = (x >> 7) & 0x01010101u;
          uint8_t* base = r < TILE_M ? a_st + (2 * w) * A_LBO + r * 16 : b_st + (2 * w) * B_LBO + (r - TILE_M) * 16;
          const int lbo = r < TILE_M ? A_LBO : B_LBO;
          *reinterpret_cast<uint4*>(base) = lo;
          *reinterpret_cast<uint4*>(base + lbo) = hi;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to UMMA
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[s]);
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) cur[i] = n1[i], n1[i] = n2[i], n2[i] = n3[i];
      }
      // epilogue: accumulator of this tile TMEM -> registers -> global.  Warp w reads TMEM lanes
      // 32 * (w % 4) .. + 31 (its quarter) and the column group w / 4.
      if (!mbar_wait(&acc_full, tile_iter & 1, err)) return;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int q = warp & 3, half = warp >> 2;
      const int row = m0 + 32 * q + lane;
#pragma unroll 1
      for (int cblk = 0; cblk < COLS_PER_WARP / 32; ++cblk) {
        const int col0 = half * COLS_PER_WARP + cblk * 32;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)col0, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < B) {
          int32_t* o = out + (long long)row * C + n0 + col0;
          if ((C & 3) == 0 && n0 + col0 + 32 <= C) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              reinterpret_cast<int4*>(o)[j] = make_int4((int)v[4 * j], (int)v[4 * j + 1], (int)v[4 * j + 2], (int)v[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + col0 + j < C) o[j] = (int)v[j];
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty);
    }
  } else {
    // ===== MMA issuer: one elected lane =====
    int it = 0, tile_iter = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_iter) {
      if (tile_iter > 0) {  // the epilogue of the previous tile must have drained the accumulator
        if (!mbar_wait(&acc_empty, (tile_iter - 1) & 1, err)) break;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      for (int ks = 0; ks < k_stages; ++ks, ++it) {
        const int s = it % STAGES;
        if (!mbar_wait(&full_bar[s], (it / STAGES) & 1, err)) goto done;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES), b_addr = a_addr + A_STAGE;
#pragma unroll
          for (int j = 0; j < BLOCK_KW; ++j) {  // one instruction per 32 bytes of K = two 16-byte chunks
            const uint64_t da = umma_desc(a_addr + 2 * j * A_LBO, A_LBO, SBO);
            const uint64_t db = umma_desc(b_addr + 2 * j * B_LBO, B_LBO, SBO);
            umma_i8(tmem_base, da, db, (ks | j) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);                      // stage free once these MMAs have read it
          if (ks == k_stages - 1) umma_commit(&acc_full);  // accumulator complete
        }
        __syncwarp();
      }
    }
  done:;
  }
  __syncthreads();
  if (warp == PRODUCER_WARPS)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
}

// ---- variant 1 (NOT YET RUN): producer groups that own ring slots ----------------------------------
// Finding from the SASS of variant 0: `fence.proxy.async.shared::cta` is MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, and
// the MEMBAR drains the thread's outstanding global loads too -- so the register look-ahead of variant 0 is
// defeated and every stage pays a full L2 / HBM round trip (matches the measured 1.25 us per stage, and
// explains why neither more producer warps nor deeper look-ahead changed anything).
// Here 16 producer warps form 4 groups; group g owns ring slot g, i.e. every 4th stage, widens it alone
// (12 words per thread) and issues the loads of its NEXT stage only AFTER its fence + arrive.  Each group
// still pays load latency + fence per stage, but four groups do so concurrently.
constexpr int G_WARPS = 16, G_GROUPS = 4, G_GROUP_THREADS = G_WARPS / G_GROUPS * 32;
constexpr int G_THREADS = G_WARPS * 32 + 32;
constexpr int G_ITEMS = (TILE_M + TILE_N) * BLOCK_KW / G_GROUP_THREADS;  // 12
constexpr int G_COLS_PER_WARP = TILE_N / (G_WARPS / 4);                 // 64
static_assert(G_GROUPS == STAGES, "a group owns one ring slot");

__global__ void __launch_bounds__(G_THREADS, 1)
    k_overlap_tcgen05_grouped(const uint32_t* __restrict__ mask, int mask_stride, const uint32_t* __restrict__ inputs,
                              int words, int B, int C, int32_t* __restrict__ out, int* err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], acc_full, acc_empty;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool producer = warp < G_WARPS;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], G_WARPS / G_GROUPS);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_full, 1);
    mbar_init(&acc_empty, G_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == G_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  const int tiles_m = (B + TILE_M - 1) / TILE_M, tiles_n = (C + TILE_N - 1) / TILE_N;
  const int n_tiles = tiles_m * tiles_n;
  const int k_stages = (words + BLOCK_KW - 1) / BLOCK_KW;

  if (producer) {
    const int grp = warp / (G_WARPS / G_GROUPS), gt = tid - grp * G_GROUP_THREADS;  // thread index inside the group
    // item i of this thread: pair index p = i * G_GROUP_THREADS + gt -> word w = p & 3, row r = p >> 2
    auto fetch = [&](int m0, int n0, int kw0, uint32_t (&wd)[G_ITEMS]) {
#pragma unroll
      for (int i = 0; i < G_ITEMS; ++i) {
        const int p = i * G_GROUP_THREADS + gt, w = kw0 + (p & 3), r = p >> 2;
        uint32_t v = 0;
        if (w < words) {
          if (r < TILE_M) {
            if (m0 + r < B) v = __ldg(inputs + (long long)(m0 + r) * words + w);
          } else if (n0 + r - TILE_M < C) {
            v = __ldg(mask + (long long)(n0 + r - TILE_M) * mask_stride + w);
          }
        }
        wd[i] = v;
      }
    };
    uint8_t* a_st = smem + grp * STAGE_BYTES;  // this group's ring slot
    uint8_t* b_st = a_st + A_STAGE;
    int it_base = 0, tile_iter = 0;  // it_base: stages of all earlier tiles of this CTA
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_iter, it_base += k_stages) {
      const int m0 = (tile % tiles_m) * TILE_M, n0 = (tile / tiles_m) * TILE_N;
      int ks = (grp - it_base) & (G_GROUPS - 1);  // first stage of this tile whose ring position is this group's slot
      uint32_t cur[G_ITEMS];
      if (ks < k_stages) fetch(m0, n0, ks * BLOCK_KW, cur);
      for (; ks < k_stages; ks += G_GROUPS) {
        const int it = it_base + ks;
        if (!mbar_wait(&empty_bar[grp], ((it / STAGES) & 1) ^ 1, err)) return;
#pragma unroll
        for (int i = 0; i < G_ITEMS; ++i) {
          const int p = i * G_GROUP_THREADS + gt, w = p & 3, r = p >> 2;
          const uint32_t x = cur[i];
          uint4 lo, hi;
          lo.x = x & 0x01010101u, lo.y = (x >> 1) & 0x01010101u, lo.z = (x >> 2) & 0x01010101u, lo.w = (x >> 3) & 0x01010101u;
          hi.x = (x >> 4) & 0x01010101u, hi.y = (x >> 5) & 0x01010101u, hi.z = (x >> 6) & 0x01010101u, hi.w = (x >> 7) & 0x01010101u;
          uint8_t* base = r < TILE_M ? a_st + (2 * w) * A_LBO + r * 16 : b_st + (2 * w) * B_LBO + (r - TILE_M) * 16;
          const int lbo = r < TILE_M ? A_LBO : B_LBO;
          *reinterpret_cast<uint4*>(base) = lo;
          *reinterpret_cast<uint4*>(base + lbo) = hi;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[grp]);
        // only now the loads of this group's next stage: nothing of them is in flight at the fence above,
        // and they travel while the group waits for its slot to be consumed
        if (ks + G_GROUPS < k_stages) fetch(m0, n0, (ks + G_GROUPS) * BLOCK_KW, cur);
      }
      if (!mbar_wait(&acc_full, tile_iter & 1, err)) return;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int q = warp & 3, cgrp = warp >> 2;
      const int row = m0 + 32 * q + lane;
#pragma unroll 1
      for (int cblk = 0; cblk < G_COLS_PER_WARP / 32; ++cblk) {
        const int col0 = cgrp * G_COLS_PER_WARP + cblk * 32;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)col0, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < B) {
          int32_t* o = out + (long long)row * C + n0 + col0;
          if ((C & 3) == 0 && n0 + col0 + 32 <= C) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              reinterpret_cast<int4*>(o)[j] = make_int4((int)v[4 * j], (int)v[4 * j + 1], (int)v[4 * j + 2], (int)v[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + col0 + j < C) o[j] = (int)v[j];
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty);
    }
  } else {
    int it = 0, tile_iter = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_iter) {
      if (tile_iter > 0) {
        if (!mbar_wait(&acc_empty, (tile_iter - 1) & 1, err)) break;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      for (int ks = 0; ks < k_stages; ++ks, ++it) {
        const int s = it % STAGES;
        if (!mbar_wait(&full_bar[s], (it / STAGES) & 1, err)) goto done;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES), b_addr = a_addr + A_STAGE;
#pragma unroll
          for (int j = 0; j < BLOCK_KW; ++j) {
            const uint64_t da = umma_desc(a_addr + 2 * j * A_LBO, A_LBO, SBO);
            const uint64_t db = umma_desc(b_addr + 2 * j * B_LBO, B_LBO, SBO);
            umma_i8(tmem_base, da, db, (ks | j) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
          if (ks == k_stages - 1) umma_commit(&acc_full);
        }
        __syncwarp();
      }
    }
  done:;
  }
  __syncthreads();
  if (warp == G_WARPS)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
}

// reference: AND + popcount, one thread per output
__global__ void k_overlap_ref(const uint32_t* mask, int mask_stride, const uint32_t* inputs, int words, int B, int C, int32_t* out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)B * C) return;
  const int b = (int)(i / C), c = (int)(i % C);
  int acc = 0;
  for (int w = 0; w < words; ++w) acc += __popc(mask[(long long)c * mask_stride + w] & inputs[(long long)b * words + w]);
  out[i] = acc;
}

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e_ = (x);                                                      \
    if (e_ != cudaSuccess) {                                                   \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 2;                                                                \
    }                                                                          \
  } while (0)

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 1024, C = argc > 2 ? atoi(argv[2]) : 2048, I = argc > 3 ? atoi(argv[3]) : 1024;
  const int variant = argc > 4 ? atoi(argv[4]) : 0;  // 0: validated kernel; 1: producer groups (not yet run)
  const int words = (I + 31) / 32, stride = (words + 3) & ~3;
  std::vector<uint32_t> h_mask((size_t)C * stride, 0), h_in((size_t)B * words, 0);
  uint64_t s = 0x9E3779B97F4A7C15ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 16); };
  for (int c = 0; c < C; ++c)
    for (int w = 0; w < words; ++w) h_mask[(size_t)c * stride + w] = rnd();
  for (auto& v : h_in) v = rnd() & rnd();
  if (I % 32) {
    const uint32_t m = (1u << (I % 32)) - 1;
    for (int c = 0; c < C; ++c) h_mask[(size_t)c * stride + words - 1] &= m;
    for (int b = 0; b < B; ++b) h_in[(size_t)b * words + words - 1] &= m;
  }
  uint32_t *d_mask, *d_in;
  int32_t *d_out, *d_ref;
  int* d_err;
  CK(cudaMalloc(&d_mask, h_mask.size() * 4));
  CK(cudaMalloc(&d_in, h_in.size() * 4));
  CK(cudaMalloc(&d_out, (size_t)B * C * 4));
  CK(cudaMalloc(&d_ref, (size_t)B * C * 4));
  CK(cudaMalloc(&d_err, 4));
  CK(cudaMemcpy(d_mask, h_mask.data(), h_mask.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_in, h_in.data(), h_in.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_out, 0xFF, (size_t)B * C * 4));
  CK(cudaMemset(d_err, 0, 4));
  const long long n = (long long)B * C;
  k_overlap_ref<<<(unsigned)((n + 255) / 256), 256>>>(d_mask, stride, d_in, words, B, C, d_ref);
  CK(cudaDeviceSynchronize());

  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int smem_bytes = STAGES * STAGE_BYTES + 1024;
  CK(cudaFuncSetAttribute(k_overlap_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  CK(cudaFuncSetAttribute(k_overlap_tcgen05_grouped, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const int tiles = ((B + TILE_M - 1) / TILE_M) * ((C + TILE_N - 1) / TILE_N);
  const int grid_dim = tiles < sms ? tiles : sms;
  auto launch = [&]() {
    if (variant == 1)
      k_overlap_tcgen05_grouped<<<grid_dim, G_THREADS, smem_bytes>>>(d_mask, stride, d_in, words, B, C, d_out, d_err);
    else
      k_overlap_tcgen05<<<grid_dim, THREADS, smem_bytes>>>(d_mask, stride, d_in, words, B, C, d_out, d_err);
  };
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  launch();
  CK(cudaDeviceSynchronize());
  int h_err = 0;
  CK(cudaMemcpy(&h_err, d_err, 4, cudaMemcpyDeviceToHost));
  if (h_err) {
    printf("FAIL: a barrier wait hit its iteration cap (pipeline protocol error)\n");
    return 1;
  }
  std::vector<int32_t> h_out((size_t)n), h_ref((size_t)n);
  CK(cudaMemcpy(h_out.data(), d_out, n * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h_ref.data(), d_ref, n * 4, cudaMemcpyDeviceToHost));
  long long bad = 0;
  for (long long i = 0; i < n; ++i)
    if (h_out[i] != h_ref[i]) {
      if (bad < 8) printf("  mismatch at input %lld column %lld: got %d want %d\n", i / C, i % C, h_out[i], h_ref[i]);
      ++bad;
    }
  if (bad) {
    printf("FAIL: %lld of %lld outputs differ\n", bad, n);
    return 1;
  }
  const int reps = 20;
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; ++r) launch();
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double us = ms * 1e3 / reps;
  printf("{\"workload\": \"%d inputs x %d columns x %d bits\", \"variant\": %d, \"bit_identical\": true, \"us\": %.2f, \"int8_tops\": %.1f}\n", B, C,
         I, variant, us, 2.0 * B * C * (double)words * 32 / us / 1e6);
  return 0;
}
