// Shared-mask batched overlap on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
//     out[b][j] = sum_i x[b][i] * m[j][i],   x, m in {0, 1}      ( == popcount(x[b] & m[j]) )
//
// i.e. DenseProjection.process (bithtm projections.py:18-21) for MANY inputs against ONE connected mask,
// as a [B x I] . [I x C] int8 contraction with int32 accumulation (exact).  Both operands stay bit-packed in
// HBM; tcgen05.mma reads byte operands from shared memory through descriptors, so the bits are widened to
// {0, 1} bytes in shared memory on the way in -- by CUDA cores, 8x the bit traffic, which is what bounds
// the kernel: 48 KB of shared-memory writes per 128 x 256 x 128 stage (384 clk at 128 B/clk) against 732 clk
// of tensor-pipe time for the same stage.
//
// One persistent CTA per SM, 18 warps:
//   warp 17     TMA loader: cp.async.bulk.tensor.2d of the RAW packed words of a stage (128 input rows +
//               256 mask rows x 4 words = 6 KB) into an 8-slot ring; out-of-range rows / words arrive as zeros.
//               (Round 1's prototype loaded the words with ordinary global loads from the widening warps: the
//               proxy fence each stage needs then drained those loads, one full memory round trip per stage.)
//   warps 0-11  widen, one row of the stage per thread: raw words -> {0, 1} bytes in the UMMA canonical K-major
//               layout (no swizzle: 8-row x 16-byte core matrices, chunk-major tiles), 3-slot ring of 48 KB
//               stages, full / empty mbarriers; fence.proxy.async.
//   warps 0-15  epilogue (tcgen05.ld -> global), a TMEM lane quarter and a column group each.
//   warp 16     allocates TMEM (256 columns) and issues tcgen05.mma.cta_group::1.kind::i8, M = 128, N = 256,
//               K = 32 bytes per instruction, accumulator in TMEM; tcgen05.commit frees the stage / publishes
//               the accumulator.
// A dot product does not care which bit sits at which k as long as both operands agree: byte j of register s
// (s = 0..7) of a 32-bit word is bit 8j + s, i.e. (w >> s) & 0x01010101 -- two integer ops per 4 bytes.
// Every mbarrier wait is bounded (1 s of globaltimer) and raises an error flag instead of hanging.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace t5 {

constexpr int TILE_M = 128;                  // inputs per tile  (UMMA M)
constexpr int TILE_N = 256;                  // columns per tile (UMMA N)
constexpr int KW = 4;                        // 32-bit words of K per stage (128 int8 of K)
constexpr int STAGES = 3;                    // widened stages in flight
constexpr int RAW_STAGES = 8;                // raw stages in flight (TMA look-ahead)
constexpr int CHUNKS = KW * 2;               // 16-byte K chunks per row and stage
constexpr int A_LBO = TILE_M * 16;           // bytes between K chunks (chunk-major tiles)
constexpr int B_LBO = TILE_N * 16;
constexpr int SBO = 128;                     // bytes between 8-row groups
constexpr int A_STAGE = CHUNKS * A_LBO;      // 16 KiB
constexpr int B_STAGE = CHUNKS * B_LBO;      // 32 KiB
constexpr int STAGE_BYTES = A_STAGE + B_STAGE;
constexpr int A_RAW = TILE_M * KW * 4;       // 2 KiB
constexpr int B_RAW = TILE_N * KW * 4;       // 4 KiB
constexpr int RAW_BYTES = A_RAW + B_RAW;
constexpr int WIDEN_WARPS = 16;              // multiple of 4 (TMEM lane quarters)
constexpr int WIDENERS = WIDEN_WARPS * 32;
constexpr int THREADS = WIDENERS + 64;       // + MMA warp + TMA warp
constexpr int ROW_WARPS = (TILE_M + TILE_N) / 32;         // warps that widen: one row of the stage per thread
constexpr int COLS_PER_WARP = TILE_N / (WIDEN_WARPS / 4);
constexpr int TMEM_COLS = 256;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + RAW_STAGES * RAW_BYTES + 1024;
constexpr int SPIN_CAP = 1 << 26;
static_assert(KW == 4 && ROW_WARPS <= WIDEN_WARPS, "a row's words are one 16-byte load");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint64_t* b, uint32_t parity, int* err) {
  const uint32_t a = smem_u32(b);
  unsigned long long t0 = 0;
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
      if (t0 == 0) t0 = t1;
      else if (t1 - t0 > 1000000000ull) break;
    }
  }
  atomicExch(err, 1);
  return false;
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // UMMA shared-memory descriptor: start [0,14), LBO [16,30), SBO [32,46) (all >> 4), version [46,48) = 1,
  // base_offset 0, lbo_mode 0, layout_type [61,64) = 0 (no swizzle)
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// UMMA instruction descriptor for kind::i8: c_format [4,6) = 2 (S32), a_format [7,10) = 0 (u8),
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
// one box of a 2-D tensor map -> shared memory; completion is counted in bytes on `bar`
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// map_in:   uint32 [B][words]       box {KW, TILE_M}
// map_mask: uint32 [C][mask_stride] box {KW, TILE_N}
// out [B][C] int32; err: device flag raised when a barrier wait times out
__global__ void __launch_bounds__(THREADS, 1)
    k_sp_overlap_batched_t5(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_mask, int words,
                            int B, int C, int32_t* __restrict__ out, int* err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], raw_full[RAW_STAGES], raw_empty[RAW_STAGES],
      acc_full, acc_empty;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* raw_base = smem + STAGES * STAGE_BYTES;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], ROW_WARPS);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < RAW_STAGES; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&raw_empty[s], ROW_WARPS);
    }
    mbar_init(&acc_full, 1);
    mbar_init(&acc_empty, WIDEN_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WIDEN_WARPS) {  // the MMA warp owns the TMEM allocation
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
  const int k_stages = (words + KW - 1) / KW;

  if (warp < WIDEN_WARPS) {
    // ===== widening warps (and epilogue) =====
    int it = 0, tile_iter = 0;  // it: stages so far (both ring positions derive from it)
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_iter) {
      const int m0 = (tile % tiles_m) * TILE_M, n0 = (tile / tiles_m) * TILE_N;
      for (int ks = 0; ks < k_stages; ++ks, ++it) {
        if (warp >= ROW_WARPS) continue;  // (warps 12-15 only take part in the epilogue)
        const int rs = it % RAW_STAGES, s = it % STAGES;
        if (!mbar_wait(&raw_full[rs], (it / RAW_STAGES) & 1, err)) return;
        // One ROW of the stage per thread (rows 0..127 of the input tile, then 256 mask rows: warps 0-11): its
        // KW = 4 words arrive as one 16-byte load, and consecutive lanes store to consecutive rows of a core
        // matrix, i.e. to consecutive 16-byte slots -- every shared-memory access of the stage is conflict-free.
        // (A (row, word) pair per thread with the word index fastest -- the first version -- put the four words
        // of a row 2 * LBO apart on the same banks: 16 wavefronts per STS.128 instead of 4, 2400 clk per stage.)
        const uint4 x4 = reinterpret_cast<const uint4*>(raw_base + rs * RAW_BYTES)[tid];
        if (!mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1, err)) return;
        uint8_t* a_st = smem + s * STAGE_BYTES;
        const int r = tid;
        uint8_t* base = r < TILE_M ? a_st + r * 16 : a_st + A_STAGE + (r - TILE_M) * 16;
        const int lbo = r < TILE_M ? A_LBO : B_LBO;
        const uint32_t xs[KW] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int w = 0; w < KW; ++w) {
          const uint32_t x = xs[w];
          uint4 lo, hi;
          lo.x = x & 0x01010101u, lo.y = (x >> 1) & 0x01010101u, lo.z = (x >> 2) & 0x01010101u, lo.w = (x >> 3) & 0x01010101u;
          hi.x = (x >> 4) & 0x01010101u, hi.y = (x >> 5) & 0x01010101u, hi.z = (x >> 6) & 0x01010101u, hi.w = (x >> 7) & 0x01010101u;
          *reinterpret_cast<uint4*>(base + (2 * w) * lbo) = lo;
          *reinterpret_cast<uint4*>(base + (2 * w + 1) * lbo) = hi;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to UMMA
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&full_bar[s]);
          mbar_arrive(&raw_empty[rs]);
        }
      }
      // epilogue: accumulator of this tile TMEM -> registers -> global.  Warp w reads TMEM lanes
      // 32 * (w % 4) .. + 31 (its quarter) and the column group w / 4.
      if (!mbar_wait(&acc_full, tile_iter & 1, err)) return;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int q = warp & 3, cgrp = warp >> 2;
      const int row = m0 + 32 * q + lane;
#pragma unroll 1
      for (int cblk = 0; cblk < COLS_PER_WARP / 32; ++cblk) {
        const int col0 = cgrp * COLS_PER_WARP + cblk * 32;
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
  } else if (warp == WIDEN_WARPS) {
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
          for (int j = 0; j < KW; ++j) {  // one instruction per 32 bytes of K = two 16-byte chunks
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
  } else {
    // ===== TMA loader: one elected lane keeps RAW_STAGES stages of packed words in flight =====
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = (tile % tiles_m) * TILE_M, n0 = (tile / tiles_m) * TILE_N;
        for (int ks = 0; ks < k_stages; ++ks, ++it) {
          const int rs = it % RAW_STAGES;
          if (!mbar_wait(&raw_empty[rs], ((it / RAW_STAGES) & 1) ^ 1, err)) break;
          uint8_t* dst = raw_base + rs * RAW_BYTES;
          mbar_expect_tx(&raw_full[rs], RAW_BYTES);
          tma_load_2d(dst, &map_in, ks * KW, m0, &raw_full[rs]);
          tma_load_2d(dst + A_RAW, &map_mask, ks * KW, n0, &raw_full[rs]);
        }
      }
    }
  }
  __syncthreads();
  if (warp == WIDEN_WARPS)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
}

}  // namespace t5
