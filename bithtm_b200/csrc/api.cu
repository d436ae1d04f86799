// C ABI of bithtm_b200 (see include/bithtm_b200.h).  Host-side launch code only:
// the library is stateless, every entry point is a sequence of kernel launches on
// the caller's stream over the caller's buffers.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "sp_kernels.cuh"
#include "tm_kernels.cuh"
#include "fused.cuh"
#include "tm_shard.cuh"
#include "shard_fused.cuh"
#include "overlap_tcgen05.cuh"

#define CU_RET(expr)                                   \
  do {                                                 \
    cudaError_t e__ = (expr);                          \
    if (e__ != cudaSuccess) return -(1000 + (int)e__); \
  } while (0)

// Optional per-launch CUDA-event timing (bh_profile_step): every kernel launch site
// ends with LAUNCHED("name"), which records an event on the launching stream.
#define BH_PROF_MAX 48
struct ProfileRec {
  cudaEvent_t ev[BH_PROF_MAX + 1];
  const char* name[BH_PROF_MAX];
  int n;
  cudaStream_t st;
};
static thread_local ProfileRec* g_prof = nullptr;

#define LAUNCHED(nm)                                            \
  do {                                                          \
    CU_RET(cudaGetLastError());                                 \
    if (g_prof && g_prof->n < BH_PROF_MAX) {                    \
      g_prof->name[g_prof->n] = nm;                             \
      cudaEventRecord(g_prof->ev[++g_prof->n], g_prof->st);     \
    }                                                           \
  } while (0)
#define LAUNCH_CHECK() LAUNCHED("misc")

// Every entry point makes ctx->device current for the duration of the call (kernel launches and
// function attributes apply to the CURRENT device, whatever stream is passed) and restores the caller's.
struct DevGuard {
  int prev = -1;
  bool switched = false;
  explicit DevGuard(const bh_ctx* x) {
    if (x && x->device >= 0 && cudaGetDevice(&prev) == cudaSuccess && prev != x->device)
      switched = cudaSetDevice(x->device) == cudaSuccess;
  }
  ~DevGuard() {
    if (switched) cudaSetDevice(prev);
  }
};
#define BH_MAX_DEVICES 64
static inline int current_device_slot() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= BH_MAX_DEVICES) d = 0;
  return d;
}

static inline cudaStream_t S_(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------
// layout
// ------------------------------------------------------------------------------------
namespace {
struct Carver {
  char* base;
  size_t off;
  template <typename T>
  void take(T*& ptr, size_t count) {
    off = (off + 255) & ~size_t(255);
    ptr = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
  }
};
}  // namespace

extern "C" size_t bh_layout(bh_ctx* x, void* base) {
  Carver cv{reinterpret_cast<char*>(base), 0};
  const size_t C = x->column_dim, I = x->input_dim, c = x->cell_dim, k = x->active_columns;
  const size_t N = C * 32 /* device cell id = column * 32 + cell */, S = x->seg_capacity, E = x->syn_capacity, M = x->match_capacity;
  const size_t CL = x->col_local;  // columns owned by this rank (== C when not sharded)
  cv.take(x->sp_perm, CL * I);
  cv.take(x->sp_mask, CL * (size_t)x->mask_stride);
  cv.take(x->duty, CL);
  cv.take(x->overlaps, CL);
  cv.take(x->boosted, CL);
  cv.take(x->active_cols, 3 * k);  // [2] ping-pong by step parity + the staged selection of the step ahead (k_step_pipe)
  cv.take(x->col_active, C);
  cv.take(x->col_pred, C);
  cv.take(x->col_act, 2 * C);  // ping-pong by step parity
  cv.take(x->col_win, C);
  cv.take(x->cell_nseg, N);
  cv.take(x->cell_maxjit, N);
  cv.take(x->cell_npred, N);
  cv.take(x->cell_widx, N);
  cv.take(x->seg_owner, S);
  cv.take(x->seg_count, S);
  cv.take(x->seg_pot, S);
  cv.take(x->seg_conn, S);
  // segment shards: rows of the 64-id blocks dealt to this rank (round-robin)
  const size_t W = x->seg_world > 1 ? x->seg_world : 1;
  const size_t rows = W > 1 ? ((S + 63) / 64 + W - 1) / W * 64 : S;
  cv.take(x->syn_cell, rows * E);
  cv.take(x->syn_perm, rows * E);
  cv.take(x->row_pred, k);
  cv.take(x->row_act, k);
  cv.take(x->row_win, k);
  cv.take(x->row_unacc, k);
  cv.take(x->winners, 2 * k * c);
  cv.take(x->unacc, k * c);
  cv.take(x->m_seg, M);
  cv.take(x->m_conn, M);
  cv.take(x->m_jit, M);
  cv.take(x->m_flag, M);
  cv.take(x->learn_list, (size_t)x->learn_capacity);
  cv.take(x->punish_list, M);
  cv.take(x->recyc_list, W > 1 ? W * (size_t)x->xr_cap : 0);
  if (x->fused_mode == 3) {  // fused sharded step: record staging and the gathered top-k candidates
    cv.take(x->x_send, (size_t)xch_send_ints(*x));
    cv.take(x->xk_keys, W * (size_t)xch_k_loc(*x));
    cv.take(x->xk_cols, W * (size_t)xch_k_loc(*x));
  } else {
    x->x_send = nullptr;
    x->xk_keys = nullptr;
    x->xk_cols = nullptr;
  }
  cv.take(x->blk, (size_t)BLK_ROWS * BH_BLK_STRIDE);
  cv.take(x->topk_ws, (size_t)BH_TOPK_WS_INTS);
  cv.take(x->mt_key, (size_t)BH_MT_N);
  cv.take(x->rng_ring, (size_t)x->rng_ring_words);
  cv.take(x->mt_jump, (size_t)x->jump_polys * BH_MT_N);
  cv.take(x->rng64, (size_t)R_COUNT);
  cv.take(x->mt_skip, (size_t)x->skip_polys * BH_MT_N);
  cv.take(x->rng_jump, x->skip_polys > 0 ? (size_t)x->job_cap * RNG_JOB_STRIDE : 0);
  cv.take(x->grow_list, x->skip_polys > 0 ? (size_t)x->learn_capacity * 3 : 0);
  cv.take(x->sc, (size_t)BH_SC_COUNT);
  cv.take(x->input_ring, (size_t)x->ring_len * x->input_words);
  cv.take(x->input_dev, (size_t)x->mask_stride);
  cv.take(x->summary_dev, (size_t)BH_SUMMARY_INTS(k));
  return (cv.off + 255) & ~size_t(255);
}

static int check_ctx(const bh_ctx* x) {
  if (!x) return BH_E_BADARG;
  if (x->cell_dim < 1 || x->cell_dim > 32) return BH_E_UNSUPPORTED;
  if (x->input_dim < 1 || x->column_dim < 1 || x->active_columns < 1 || x->active_columns > x->column_dim)
    return BH_E_BADARG;
  if (x->input_words != (x->input_dim + 31) / 32 || x->mask_stride % 4 != 0 || x->mask_stride < x->input_words)
    return BH_E_BADARG;
  if (x->tm_blocks < 1 || x->tm_blocks > BH_BLK_STRIDE) return BH_E_BADARG;
  if (x->col_local < 1 || x->col_lo < 0 || x->col_lo + x->col_local > x->column_dim) return BH_E_BADARG;
  if (x->fused_mode < 0 || x->fused_mode > 3) return BH_E_BADARG;
  if (x->col_local != x->column_dim && x->fused_mode && x->fused_mode != 3) return BH_E_UNSUPPORTED;
  if (x->fused_mode == 3) {  // one kernel per shard, exchanges in-kernel: SP by column and TM by segment, same ranks
    const int W = x->seg_world > 1 ? x->seg_world : 1;
    if (W > BH_MAX_RANKS || x->xm_cap < 1 || x->xr_cap < 1) return BH_E_BADARG;
    if ((long long)x->col_local * W != x->column_dim || x->col_lo != x->seg_rank * x->col_local) return BH_E_BADARG;
  }
  if (x->syn_capacity < 32 || x->syn_capacity % 32 != 0) return BH_E_BADARG;
  if (x->rng_ring_words < (1 << 20) || x->rng_ring_words > (1LL << 31) ||
      (x->rng_ring_words & (x->rng_ring_words - 1)))
    return BH_E_BADARG;
  // the ring spans a step's draws (absolute stream index -> slot): twice, or -- networks whose rand(L, W+1) is
  // only ever stepped over (lazy draws) -- once plus a margin
  if (x->rng_step_words < 2 * BH_MT_N) return BH_E_BADARG;
  if (2 * x->rng_step_words > x->rng_ring_words &&
      !(x->skip_polys > 0 && x->lazy_policy == 1 && x->rng_step_words + (1 << 24) <= x->rng_ring_words))
    return BH_E_BADARG;
  // one round of chunks must cover a whole step (plus lookahead) when the stream is produced by many CTAs
  if (x->jump_polys > 0 && (long long)x->jump_polys * RNG_CHUNK < x->rng_step_words + x->rng_step_words / 2 + RNG_CHUNK)
    return BH_E_BADARG;
  if (x->jump_polys < 0 || x->rng_lookahead < 0) return BH_E_BADARG;
  if (x->skip_polys < 0) return BH_E_BADARG;
  if (x->skip_polys > 0) {  // lazy draws (mt19937.cuh): table granularity, and a job slot for every chunk of the largest step
    if (x->skip_gran < 32 || (x->skip_gran & (x->skip_gran - 1)) || x->skip_min < 2LL * x->skip_gran) return BH_E_BADARG;
    if (x->job_cap < RNG_ROW_SLOT0 + x->rng_step_words / RNG_LAZY_CHUNK + 2) return BH_E_BADARG;
  }
  // the cluster kernel has no phase for the chunks a many-CTA production plan leaves (fused.cuh)
  if (x->fused_mode == 1 && x->jump_polys > 0) return BH_E_UNSUPPORTED;
  if (x->pipe_ctas != 0) {  // two-pipeline kernels: both teams non-empty; grid mode with its grid-wide selection,
                            // or the sharded step with cell exchanges
    if (x->pipe_ctas < 2 || x->pipe_ctas > x->fused_ctas - 1 || x->fused_ctas > TK2_MAX_CTAS) return BH_E_UNSUPPORTED;
    if (!((x->fused_mode == 2 && x->column_dim >= 16384) || (x->fused_mode == 3 && x->xch_ll))) return BH_E_UNSUPPORTED;
  }
  if (x->seg_world > 1) {
    if (x->seg_rank < 0 || x->seg_rank >= x->seg_world || x->xm_cap < 1 || x->xr_cap < 1) return BH_E_BADARG;
    if (x->fused_mode && x->fused_mode != 3) return BH_E_UNSUPPORTED;  // the exchange sits between kernels
  }
  return 0;
}

extern "C" int bh_abi_version(void) { return BH_ABI_VERSION; }
extern "C" size_t bh_ctx_size(void) { return sizeof(bh_ctx); }

extern "C" int bh_device_info(int device, int* sm_count, int* cc_major, int* cc_minor) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return BH_E_NODEVICE;
  cudaDeviceProp p;
  CU_RET(cudaGetDeviceProperties(&p, device));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return 0;
}

__global__ void k_fill_i32(int32_t* p, long long n, int32_t v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = v;
}

extern "C" int bh_init(const bh_ctx* x, void* stream) {
  DevGuard dev_guard_(x);
  int rc = check_ctx(x);
  if (rc) return rc;
  long long N = (long long)x->column_dim * 32;
  k_fill_i32<<<cdiv(N, 256) < 1184 ? cdiv(N, 256) : 1184, 256, 0, S_(stream)>>>(x->cell_widx, N, -1);
  LAUNCH_CHECK();
  k_rng_import<<<1, 256, 0, S_(stream)>>>(*x);  // a defined (all-zero key) stream until the caller imports one
  LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------
// spatial pooler
// ------------------------------------------------------------------------------------
static int overlap_group_host(const bh_ctx* x) {
  int vec = x->mask_stride / 4, g = 1;
  while (g * 2 <= vec && g < 32) g *= 2;
  return g;
}

static int sp_grid(const bh_ctx* x, int rows_per_block) {
  int want = cdiv(x->col_local, rows_per_block);
  int cap = (x->sm_count > 0 ? x->sm_count : 148) * 8;
  return want < cap ? want : cap;
}

extern "C" int bh_sp_build_mask(const bh_ctx* x, void* stream) {
  DevGuard dev_guard_(x);
  int rc = check_ctx(x);
  if (rc) return rc;
  long long words = (long long)x->col_local * x->mask_stride;
  int grid = cdiv(words, SP_THREADS / 32);
  int cap = (x->sm_count > 0 ? x->sm_count : 148) * 16;
  k_sp_build_mask<<<grid < cap ? grid : cap, SP_THREADS, 0, S_(stream)>>>(*x);
  LAUNCH_CHECK();
  return 0;
}

extern "C" int bh_pack_input(const bh_ctx* x, const uint8_t* bool_dev, uint32_t* words_dev, void* stream) {
  DevGuard dev_guard_(x);
  k_pack_input<<<cdiv((long long)x->input_words * 32, 256), 256, 0, S_(stream)>>>(*x, bool_dev, words_dev);
  LAUNCH_CHECK();
  return 0;
}

extern "C" int bh_pack_inputs(const bh_ctx* x, const uint8_t* bool_dev, int n_inputs, int pitch_words, uint32_t* words_dev,
                              void* stream) {
  DevGuard dev_guard_(x);
  if (!x || !bool_dev || !words_dev || n_inputs < 0 || pitch_words < x->input_words) return BH_E_BADARG;
  if (n_inputs == 0) return 0;
  const long long warps = (long long)n_inputs * pitch_words;
  const int cap = (x->sm_count > 0 ? x->sm_count : 148) * 16;
  const int grid = cdiv(warps * 32, 256);
  k_pack_inputs<<<grid < cap ? grid : cap, 256, 0, S_(stream)>>>(*x, bool_dev, n_inputs, pitch_words, words_dev);
  LAUNCH_CHECK();
  return 0;
}

template <bool BOOST>
static int launch_overlap(const bh_ctx* x, const uint32_t* in, cudaStream_t st) {
  int g = overlap_group_host(x);
  int rows_per_block = (SP_THREADS / 32) * (32 / g);
  size_t smem = (size_t)x->mask_stride * 4;
  k_sp_overlap<BOOST><<<sp_grid(x, rows_per_block), SP_THREADS, smem, st>>>(*x, in);
  LAUNCHED(BOOST ? "sp_overlap_boost" : "sp_overlap");
  return 0;
}

extern "C" int bh_sp_overlap(const bh_ctx* x, const uint32_t* in, void* stream) {
  DevGuard dev_guard_(x);
  return launch_overlap<false>(x, in, S_(stream));
}

extern "C" int bh_sp_overlap_batched(const bh_ctx* x, const uint32_t* inputs_dev, int n_inputs, int32_t* overlaps_out,
                                     void* stream) {
  DevGuard dev_guard_(x);
  int rc = check_ctx(x);
  if (rc) return rc;
  if (!inputs_dev || !overlaps_out || n_inputs < 0) return BH_E_BADARG;
  if (n_inputs == 0) return 0;
  dim3 grid(cdiv(x->col_local, 32), cdiv(n_inputs, 32));
  if (grid.y > 65535) return BH_E_BADARG;
  k_sp_overlap_batched<<<grid, 1024, 0, S_(stream)>>>(*x, inputs_dev, n_inputs, overlaps_out);
  LAUNCHED("sp_overlap_batched");
  return 0;
}

extern "C" int bh_sp_overlap_batched_mma(const bh_ctx* x, const uint32_t* inputs_dev, int n_inputs, int32_t* overlaps_out,
                                         void* stream);

// cuTensorMapEncodeTiled through the runtime (no link-time dependency on libcuda: the library must load
// without a driver for the build / symbol checks)
typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tmap_encode_fn tmap_encoder() {
  static tmap_encode_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && p)
      fn = reinterpret_cast<tmap_encode_fn>(p);
  }
  return fn;
}
// uint32 [rows][pitch_words] (extent `words` of every row is data), box {t5::KW words, box_rows rows}
static int tmap_words_2d(CUtensorMap* m, const uint32_t* base, int rows, int words, int pitch_words, int box_rows) {
  tmap_encode_fn enc = tmap_encoder();
  if (!enc) return BH_E_UNSUPPORTED;
  const cuuint64_t dims[2] = {(cuuint64_t)words, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch_words * 4};
  const cuuint32_t box[2] = {(cuuint32_t)t5::KW, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint32_t*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : BH_E_BADARG;
}

static bool t5_eligible(const bh_ctx* x, const uint32_t* inputs_dev, int n_inputs) {
  // TMA: 16-byte row pitches and bases; worth it from a few tiles of work on
  // (measured: 568 vs 747 us at 256 x 65536 x 16384, but 18.7 vs 12.5 us at 1024 x 2048 x 1024 -- a few short tiles
  // per SM do not amortise the pipeline fill and the epilogue)
  const long long tiles = (long long)cdiv(n_inputs, t5::TILE_M) * cdiv(x->col_local, t5::TILE_N);
  const long long k_stages = cdiv(x->input_words, t5::KW);
  const int sms = x->sm_count > 0 ? x->sm_count : 148;
  return x->input_words % 4 == 0 && (reinterpret_cast<uintptr_t>(inputs_dev) & 15) == 0 && tiles * k_stages >= 64LL * sms;
}

extern "C" int bh_sp_overlap_batched_tc5(const bh_ctx* x, const uint32_t* inputs_dev, int n_inputs, int32_t* overlaps_out,
                                         void* stream) {
  DevGuard dev_guard_(x);
  int rc = check_ctx(x);
  if (rc) return rc;
  if (!inputs_dev || !overlaps_out || n_inputs < 0) return BH_E_BADARG;
  if (n_inputs == 0) return 0;
  if (x->input_words % 4 != 0 || (reinterpret_cast<uintptr_t>(inputs_dev) & 15)) return BH_E_UNSUPPORTED;
  CUtensorMap map_in, map_mask;
  if ((rc = tmap_words_2d(&map_in, inputs_dev, n_inputs, x->input_words, x->input_words, t5::TILE_M))) return rc;
  if ((rc = tmap_words_2d(&map_mask, x->sp_mask, x->col_local, x->input_words, x->mask_stride, t5::TILE_N))) return rc;
  static bool attr_done[BH_MAX_DEVICES] = {};
  const int d = current_device_slot();
  if (!attr_done[d]) {
    CU_RET(cudaFuncSetAttribute(t5::k_sp_overlap_batched_t5, cudaFuncAttributeMaxDynamicSharedMemorySize, t5::SMEM_BYTES));
    attr_done[d] = true;
  }
  const int tiles = cdiv(n_inputs, t5::TILE_M) * cdiv(x->col_local, t5::TILE_N);
  const int sms = x->sm_count > 0 ? x->sm_count : 148;
  t5::k_sp_overlap_batched_t5<<<tiles < sms ? tiles : sms, t5::THREADS, t5::SMEM_BYTES, S_(stream)>>>(
      map_in, map_mask, x->input_words, n_inputs, x->col_local, overlaps_out, x->sc + BH_SC_T5_ERR);
  LAUNCHED("sp_overlap_batched_tc5");
  return 0;
}

extern "C" int bh_sp_overlap_batched_tc(const bh_ctx* x, const uint32_t* inputs_dev, int n_inputs, int32_t* overlaps_out,
                                        void* stream) {
  DevGuard dev_guard_(x);
  int rc = check_ctx(x);
  if (rc) return rc;
  if (!inputs_dev || !overlaps_out || n_inputs < 0) return BH_E_BADARG;
  if (n_inputs == 0) return 0;
  if (t5_eligible(x, inputs_dev, n_inputs) && tmap_encoder())  // tcgen05 + TMEM + TMA; else the mma.sync kernel
    return bh_sp_overlap_batched_tc5(x, inputs_dev, n_inputs, overlaps_out, stream);
  return bh_sp_overlap_batched_mma(x, inputs_dev, n_inputs, overlaps_out, stream);
}

extern "C" int bh_sp_overlap_batched_mma(const bh_ctx* x, const uint32_t* inputs_dev, int n_inputs, int32_t* overlaps_out,
                                         void* stream) {
  DevGuard dev_guard_(x);
  int rc = check_ctx(x);
  if (rc) return rc;
  if (!inputs_dev || !overlaps_out || n_inputs < 0) return BH_E_BADARG;
  if (n_inputs == 0) return 0;
  if ((long long)x->input_words * 32 >= (1ll << 24)) return BH_E_UNSUPPORTED;  // accumulators hold 128 * overlap
  dim3 grid(cdiv(x->col_local, OTC_N), cdiv(n_inputs, OTC_M));
  if (grid.y > 65535) return BH_E_BADARG;
  k_sp_overlap_batched_tc<<<grid, OTC_THREADS, 0, S_(stream)>>>(*x, inputs_dev, n_inputs, overlaps_out);
  LAUNCHED("sp_overlap_batched_tc");
  return 0;
}

extern "C" int bh_boost(const bh_ctx* x, void* stream) {
  DevGuard dev_guard_(x);
  k_boost<<<cdiv(x->col_local, 256), 256, 0, S_(stream)>>>(*x);
  LAUNCH_CHECK();
  return 0;
}

// cooperative launch of a <<<sm_count, TOPK_THREADS>>> kernel taking (ctx, extra args...)
template <typename... Args>
static int launch_coop(void (*kern)(const bh_ctx, Args...), const bh_ctx* x, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(x->sm_count > 0 ? x->sm_count : 148);
  cfg.blockDim = dim3(TOPK_THREADS);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CU_RET(cudaLaunchKernelEx(&cfg, kern, *x, args...));
  return 0;
}

#define BH_TOPK_MULTI_MIN 16384  // columns from which the grid-wide top-k pays for its barriers

extern "C" int bh_inhibit(const bh_ctx* x, void* stream) {
  DevGuard dev_guard_(x);
  if (x->column_dim >= BH_TOPK_MULTI_MIN) {
    int rc = launch_coop(k_topk_multi, x, S_(stream));
    if (rc) return rc;
    LAUNCHED("topk_multi");
    return 0;
  }
  k_topk<<<1, TOPK_THREADS, 0, S_(stream)>>>(*x);
  LAUNCHED("topk");
  return 0;
}

extern "C" int bh_set_active_columns(const bh_ctx* x, const int32_t* cols_dev, void* stream) {
  DevGuard dev_guard_(x);
  k_set_active<<<1, 1024, 0, S_(stream)>>>(*x, cols_dev);
  LAUNCH_CHECK();
  return 0;
}

extern "C" int bh_sp_learn(const bh_ctx* x, const uint32_t* in, void* stream) {
  DevGuard dev_guard_(x);
  int cap = (x->sm_count > 0 ? x->sm_count : 148) * 8;
  int grid = x->active_columns < cap ? x->active_columns : cap;
  k_sp_learn<<<grid, SP_THREADS, 0, S_(stream)>>>(*x, in);
  LAUNCHED("sp_learn");
  return 0;
}

extern "C" int bh_duty_update(const bh_ctx* x, void* stream) {
  DevGuard dev_guard_(x);
  k_duty_update<<<cdiv(x->col_local, 256), 256, 0, S_(stream)>>>(*x);
  LAUNCHED("duty_update");
  return 0;
}

static int sp_step(const bh_ctx* x, const uint32_t* in, int learning, cudaStream_t st) {
  int rc;
  if ((rc = launch_overlap<true>(x, in, st))) return rc;
  if ((rc = bh_inhibit(x, st))) return rc;
  if (learning && (rc = bh_sp_learn(x, in, st))) return rc;
  return bh_duty_update(x, st);
}

// ---- column shard: exchange 1 (top-k candidates) ---------------------------------------
extern "C" int bh_sp_shard_local(const bh_ctx* x, const uint32_t* in, double* cand_keys, int32_t* cand_cols,
                                 void* stream) {
  DevGuard dev_guard_(x);
  int rc = check_ctx(x);
  if (rc) return rc;
  if (!cand_keys || !cand_cols) return BH_E_BADARG;
  cudaStream_t st = S_(stream);
  if ((rc = launch_overlap<true>(x, in, st))) return rc;
  // scratch for the selected local positions: the (not yet used) current active-column buffer
  int* scratch = x->row_unacc ? reinterpret_cast<int*>(x->row_unacc) : nullptr;
  if (!scratch) return BH_E_BADARG;
  if (x->col_local >= BH_TOPK_MULTI_MIN) {
    if ((rc = launch_coop(k_topk_shard_local_multi, x, st, scratch, cand_keys, cand_cols))) return rc;
  } else {
    k_topk_shard_local<<<1, TOPK_THREADS, 0, st>>>(*x, scratch, cand_keys, cand_cols);
  }
  LAUNCHED("topk_shard_local");
  return 0;
}

extern "C" int bh_sp_shard_finish(const bh_ctx* x, const uint32_t* in, const double* cand_keys,
                                  const int32_t* cand_cols, int n, int learning, void* stream) {
  DevGuard dev_guard_(x);
  int rc = check_ctx(x);
  if (rc) return rc;
  if (!cand_keys || !cand_cols || n < x->active_columns) return BH_E_BADARG;
  cudaStream_t st = S_(stream);
  k_topk_shard_merge<<<1, TOPK_THREADS, 0, st>>>(*x, cand_keys, cand_cols, n);
  LAUNCHED("topk_shard_merge");
  if (learning && (rc = bh_sp_learn(x, in, st))) return rc;
  return bh_duty_update(x, st);
}

extern "C" int bh_sp_step(const bh_ctx* x, const uint32_t* in, int learning, void* stream) {
  DevGuard dev_guard_(x);
  int rc = check_ctx(x);
  if (rc) return rc;
  return sp_step(x, in, learning, S_(stream));
}

__global__ void k_advance_step(const __grid_constant__ bh_ctx c) { c.sc[BH_SC_STEP] = c.sc[BH_SC_STEP] + 1; }

extern "C" int bh_advance_step(const bh_ctx* x, void* stream) {
  DevGuard dev_guard_(x);
  k_advance_step<<<1, 1, 0, S_(stream)>>>(*x);
  LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------
// temporal memory
// ------------------------------------------------------------------------------------
static int tm_post_and_scan(const bh_ctx* x, cudaStream_t st);

static int tm_select(const bh_ctx* x, int want, cudaStream_t st) {
  if (want) {
    k_tm_fill_jitter<<<1, MT_THREADS, 0, st>>>(*x);  // no-op unless the last activation deferred its draw
    LAUNCHED("tm_fill_jitter");
    k_tm_draw<<<1, MT_THREADS, 0, st>>>(*x, 1, 1);
    LAUNCHED("tm_draw1");
  }
  k_tm_select_a<<<x->tm_blocks, BH_TM_THREADS, 0, st>>>(*x, want);
  LAUNCHED("tm_select_a");
  k_tm_select_b<<<x->tm_blocks, BH_TM_THREADS, 0, st>>>(*x, want);
  LAUNCHED("tm_select_b");
  return 0;
}

extern "C" int bh_tm_select(const bh_ctx* x, void* stream) {
  DevGuard dev_guard_(x);
  return tm_select(x, 1, S_(stream));
}

static int learn_apply_smem(const bh_ctx* x) {
  long long bits = (long long)x->active_columns * x->cell_dim;
  return (int)(((bits + 31) / 32) * 4);
}

// function attributes are per device: remembered per device ordinal of the calling thread's current device
static int prepare_chunks() {
  static bool done[BH_MAX_DEVICES] = {};
  const int d = current_device_slot();
  if (!done[d]) {
    CU_RET(cudaFuncSetAttribute(k_rng_chunks, cudaFuncAttributeMaxDynamicSharedMemorySize, RNG_CHUNK_SMEM));
    done[d] = true;
  }
  return 0;
}

extern "C" int bh_tm_learn(const bh_ctx* x, int learning, void* stream) {
  DevGuard dev_guard_(x);
  cudaStream_t st = S_(stream);
  k_tm_learn_select_a<<<x->tm_blocks, BH_TM_THREADS, 0, st>>>(*x, learning);
  LAUNCHED("tm_learn_select_a");
  k_tm_learn_select_b<<<x->tm_blocks, BH_TM_THREADS, 0, st>>>(*x, learning);
  LAUNCHED("tm_learn_select_b");
  k_tm_draw<<<1, MT_THREADS, 0, st>>>(*x, 2, learning);
  LAUNCHED("tm_draw2");
  if (x->jump_polys > 0) {  // the chunks of the plan draw #2 may have left
    int prc = prepare_chunks();
    if (prc) return prc;
    int cap = (x->sm_count > 0 ? x->sm_count : 148);
    k_rng_chunks<<<x->jump_polys < cap ? x->jump_polys : cap, MT_THREADS, RNG_CHUNK_SMEM, st>>>(*x);
    LAUNCHED("rng_chunks");
  }
  if (learning) {
    int smem = learn_apply_smem(x);
    if (smem > 200 * 1024) return BH_E_UNSUPPORTED;
    if (smem > 40 * 1024)
      CU_RET(cudaFuncSetAttribute(k_tm_learn_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int grid = (x->sm_count > 0 ? x->sm_count : 148) * 2;
    k_tm_learn_apply<<<grid, LA_THREADS, smem, st>>>(*x);
    LAUNCHED("tm_learn_apply");
  }
  return 0;
}

extern "C" int bh_tm_activate(const bh_ctx* x, void* stream) {
  DevGuard dev_guard_(x);
  if (x->seg_world > 1) return BH_E_UNSUPPORTED;  // use bh_tm_shard_pre / bh_tm_shard_post
  cudaStream_t st = S_(stream);
  int k = x->active_columns, kc = k * x->cell_dim;
  k_tm_post<<<cdiv(kc, 256) < 256 ? cdiv(kc, 256) : 256, 256, 0, st>>>(*x);
  LAUNCHED("tm_post");
  k_tm_activate_a<<<x->tm_blocks, BH_TM_THREADS, 0, st>>>(*x);
  LAUNCHED("tm_activate_a");
  k_tm_draw<<<1, MT_THREADS, 0, st>>>(*x, 3, 1);
  LAUNCHED("tm_draw3");
  k_tm_activate_b<<<x->tm_blocks, BH_TM_THREADS, 0, st>>>(*x, 1);
  LAUNCHED("tm_activate_b");
  return 0;
}

// ---- stand-alone plugin calls with explicit arguments -----------------------------------
extern "C" int bh_tm_reset(const bh_ctx* x, void* stream) {
  DevGuard dev_guard_(x);
  int rc = check_ctx(x);
  if (rc) return rc;
  k_tm_reset<<<1, 1024, 0, S_(stream)>>>(*x);
  LAUNCHED("tm_reset");
  return 0;
}

extern "C" int bh_tm_fill_jitter(const bh_ctx* x, void* stream) {
  DevGuard dev_guard_(x);
  int rc = check_ctx(x);
  if (rc) return rc;
  k_tm_fill_jitter<<<1, MT_THREADS, 0, S_(stream)>>>(*x);
  LAUNCHED("tm_fill_jitter");
  return 0;
}

extern "C" int bh_tm_learn_args(const bh_ctx* x, const int32_t* winner_cells_dev, int n_winners,
                                const int32_t* prev_winner_cells_dev, int n_prev_winners,
                                const uint32_t* prev_activation_words_dev, const uint8_t* column_active_dev,
                                void* stream) {
  DevGuard dev_guard_(x);
  int rc = check_ctx(x);
  if (rc) return rc;
  const int kc = x->active_columns * x->cell_dim;
  if (x->seg_world > 1 || n_winners < 0 || n_winners > kc || n_prev_winners > kc || !prev_activation_words_dev ||
      !column_active_dev || (n_winners > 0 && !winner_cells_dev) || (n_prev_winners > 0 && !prev_winner_cells_dev))
    return BH_E_BADARG;
  cudaStream_t st = S_(stream);
  k_tm_fill_jitter<<<1, MT_THREADS, 0, st>>>(*x);  // get_jittered_potential_info(prev_state), projections.py:263
  LAUNCHED("tm_fill_jitter");
  const int cap = (x->sm_count > 0 ? x->sm_count : 148) * 8;
  const int grid = cdiv((long long)x->column_dim * 32, 256);
  k_tm_adopt_words<<<grid < cap ? grid : cap, 256, 0, st>>>(*x, prev_activation_words_dev, column_active_dev);
  LAUNCHED("tm_adopt_words");
  k_tm_adopt_lists<<<1, 1024, 0, st>>>(*x, winner_cells_dev, n_winners, prev_winner_cells_dev, n_prev_winners);
  LAUNCHED("tm_adopt_lists");
  return bh_tm_learn(x, 1, stream);
}

extern "C" int bh_tm_activate_cells(const bh_ctx* x, const int32_t* active_cells_dev, int n_active, int want_jitter,
                                    int have_winners, void* stream) {
  DevGuard dev_guard_(x);
  int rc = check_ctx(x);
  if (rc) return rc;
  if (x->seg_world > 1 || n_active < 0 || (n_active > 0 && !active_cells_dev)) return BH_E_BADARG;
  cudaStream_t st = S_(stream);
  const int cap = (x->sm_count > 0 ? x->sm_count : 148) * 8;
  const int grid = cdiv(2LL * x->column_dim, 256);
  k_tm_adopt_active_clear<<<grid < cap ? grid : cap, 256, 0, st>>>(*x, have_winners);
  LAUNCHED("tm_adopt_active_clear");
  if (n_active > 0) {
    const int g2 = cdiv(n_active, 256);
    k_tm_adopt_active_set<<<g2 < cap ? g2 : cap, 256, 0, st>>>(*x, active_cells_dev, n_active);
    LAUNCHED("tm_adopt_active_set");
  }
  k_tm_adopt_widx<<<1, 1024, 0, st>>>(*x);
  LAUNCHED("tm_adopt_widx");
  k_tm_activate_a<<<x->tm_blocks, BH_TM_THREADS, 0, st>>>(*x);
  LAUNCHED("tm_activate_a");
  if (want_jitter) {
    k_tm_draw<<<1, MT_THREADS, 0, st>>>(*x, 3, 1);
    LAUNCHED("tm_draw3");
  }
  k_tm_activate_b<<<x->tm_blocks, BH_TM_THREADS, 0, st>>>(*x, want_jitter ? 1 : 0);
  LAUNCHED("tm_activate_b");
  return 0;
}

extern "C" int bh_tm_step_ex(const bh_ctx* x, int learning, int want_winner, int want_jitter, void* stream) {
  DevGuard dev_guard_(x);
  int rc = check_ctx(x);
  if (rc) return rc;
  if (x->seg_world > 1) return BH_E_UNSUPPORTED;
  cudaStream_t st = S_(stream);
  if ((rc = tm_select(x, (learning || want_winner) ? 1 : 0, st))) return rc;
  if ((rc = bh_tm_learn(x, learning, stream))) return rc;
  if (want_jitter) return bh_tm_activate(x, stream);
  if ((rc = tm_post_and_scan(x, st))) return rc;
  k_tm_activate_b<<<x->tm_blocks, BH_TM_THREADS, 0, st>>>(*x, 0);
  LAUNCHED("tm_activate_b");
  return 0;
}

// ---- segment shards: exchange 2 ----------------------------------------------------------
extern "C" size_t bh_tm_shard_xch_ints(const bh_ctx* x) { return x ? (size_t)xch_ints(*x) : 0; }
extern "C" size_t bh_xch_region_ints(const bh_ctx* x) { return x ? (size_t)xch_region_ints(*x) : 0; }

static int tm_post_and_scan(const bh_ctx* x, cudaStream_t st) {
  int k = x->active_columns, kc = k * x->cell_dim;
  k_tm_post<<<cdiv(kc, 256) < 256 ? cdiv(kc, 256) : 256, 256, 0, st>>>(*x);
  LAUNCHED("tm_post");
  k_tm_activate_a<<<x->tm_blocks, BH_TM_THREADS, 0, st>>>(*x);
  LAUNCHED("tm_activate_a");
  return 0;
}

extern "C" int bh_tm_shard_pre(const bh_ctx* x, int learning, int32_t* send_dev, void* stream) {
  DevGuard dev_guard_(x);
  int rc = check_ctx(x);
  if (rc) return rc;
  if (x->seg_world <= 1 || !send_dev) return BH_E_BADARG;
  cudaStream_t st = S_(stream);
  if (learning < 0 || learning > 3) return BH_E_BADARG;
  const int learn = learning & BH_STEP_LEARNING, winners = !(learning & BH_STEP_NO_WINNER_CELLS);
  if ((rc = tm_select(x, (learn || winners) ? 1 : 0, st))) return rc;
  if ((rc = bh_tm_learn(x, learn, stream))) return rc;
  if ((rc = tm_post_and_scan(x, st))) return rc;
  k_tm_shard_pack<<<x->tm_blocks, BH_TM_THREADS, 0, st>>>(*x, send_dev);
  LAUNCHED("tm_shard_pack");
  return 0;
}

extern "C" int bh_tm_shard_post_ex(const bh_ctx* x, const int32_t* recv_dev, int want_jitter, void* stream) {
  DevGuard dev_guard_(x);
  int rc = check_ctx(x);
  if (rc) return rc;
  if (x->seg_world <= 1 || !recv_dev) return BH_E_BADARG;
  cudaStream_t st = S_(stream);
  k_tm_shard_merge<<<x->tm_blocks, BH_TM_THREADS, 0, st>>>(*x, recv_dev);
  LAUNCHED("tm_shard_merge");
  if (want_jitter) {
    k_tm_draw<<<1, MT_THREADS, 0, st>>>(*x, 3, 1);
    LAUNCHED("tm_draw3");
  }
  k_tm_activate_finish<<<x->tm_blocks, BH_TM_THREADS, 0, st>>>(*x, want_jitter ? 1 : 0);
  LAUNCHED("tm_activate_finish");
  return 0;
}

extern "C" int bh_tm_shard_post(const bh_ctx* x, const int32_t* recv_dev, void* stream) {
  return bh_tm_shard_post_ex(x, recv_dev, 1, stream);
}

extern "C" int bh_tm_step(const bh_ctx* x, int learning, void* stream) {
  DevGuard dev_guard_(x);
  int rc = check_ctx(x);
  if (rc) return rc;
  if ((rc = bh_tm_select(x, stream))) return rc;
  if ((rc = bh_tm_learn(x, learning, stream))) return rc;
  return bh_tm_activate(x, stream);
}

// ------------------------------------------------------------------------------------
// whole step
// ------------------------------------------------------------------------------------
static int fused_smem(const bh_ctx* x) {
  int a = x->mask_stride * 4 + (x->fused_mode >= 2 ? TK2_BINS * 4 : 0), b = learn_apply_smem(x);
  int m = a > b ? a : b;
  if (x->fused_mode >= 2 && x->jump_polys > 0 && m < RNG_CHUNK_SMEM) m = RNG_CHUNK_SMEM;
  if (m < MT_RING * 4) m = MT_RING * 4;  // the deferred-jitter draw of P0 stages the stream in dynamic shared memory
  if (x->fused_mode >= 2 && x->skip_polys > 0 && m < RNG_LAZY_SMEM_WORDS * 4) m = RNG_LAZY_SMEM_WORDS * 4;
  if (x->fused_mode == 3 && x->xch_ll && m < ll_smem_bytes(*x)) m = (int)(ll_smem_bytes(*x) < (1 << 20) ? ll_smem_bytes(*x) : (1 << 20));
  return m;
}

// One-time function attributes (per device): non-portable cluster sizes, dynamic smem.
static int prepare_fused(int mode) {
  static bool done_all[BH_MAX_DEVICES][4] = {};
  if (mode < 1 || mode > 3) return BH_E_BADARG;
  bool* done = done_all[current_device_slot()];
  if (done[mode]) return 0;
  if (mode == 1) {
    CU_RET(cudaFuncSetAttribute(k_step_fused<1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CU_RET(cudaFuncSetAttribute(k_step_fused<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  } else if (mode == 2) {
    CU_RET(cudaFuncSetAttribute(k_step_fused<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CU_RET(cudaFuncSetAttribute(k_step_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  } else {
    CU_RET(cudaFuncSetAttribute(k_step_shard, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CU_RET(cudaFuncSetAttribute(k_step_shard_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  }
  done[mode] = true;
  return 0;
}

// One launch of the fused kernel: n_steps consecutive steps (input_fixed == NULL ->
// inputs from the device ring), optionally followed by the host summary.
static int launch_fused(const bh_ctx* x, const uint32_t* input_fixed, int n_steps, int learning, int want_summary,
                        cudaStream_t st) {
  const int nb = x->fused_ctas;
  if (nb < 1) return BH_E_BADARG;
  if (learning < 0 || learning > 3) return BH_E_UNSUPPORTED;
  const int smem = fused_smem(x);
  if (smem > 160 * 1024) return BH_E_UNSUPPORTED;
  int prc = prepare_fused(x->fused_mode);
  if (prc) return prc;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(nb);
  cfg.blockDim = dim3(FUSED_THREADS);
  if (x->fused_mode == 1 && x->fused_threads) {  // cluster kernel with smaller CTAs (several per SM)
    if (x->fused_threads < 256 || x->fused_threads > FUSED_THREADS || x->fused_threads % 32) return BH_E_BADARG;
    cfg.blockDim = dim3(x->fused_threads);
  }
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (x->fused_mode == 1) {
    if (nb > 16) return BH_E_BADARG;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = nb;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    CU_RET(cudaLaunchKernelEx(&cfg, k_step_fused<1>, *x, input_fixed, n_steps, learning, want_summary));
  } else if (x->fused_mode == 2) {
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    if (x->pipe_ctas > 0 && n_steps > 1)  // (a single step has nothing to look ahead to)
      CU_RET(cudaLaunchKernelEx(&cfg, k_step_pipe, *x, input_fixed, n_steps, learning, want_summary));
    else
      CU_RET(cudaLaunchKernelEx(&cfg, k_step_fused<2>, *x, input_fixed, n_steps, learning, want_summary));
  } else {
    const int W = x->seg_world > 1 ? x->seg_world : 1;
    for (int r = 0; r < W; ++r)
      if (!x->xpeer[r]) return BH_E_BADARG;
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    if (x->pipe_ctas > 0 && n_steps > 1)
      CU_RET(cudaLaunchKernelEx(&cfg, k_step_shard_pipe, *x, input_fixed, n_steps, learning, want_summary));
    else
      CU_RET(cudaLaunchKernelEx(&cfg, k_step_shard, *x, input_fixed, n_steps, learning, want_summary));
  }
  LAUNCHED(x->fused_mode == 1 ? "step_fused_cluster" : (x->fused_mode == 2 ? (x->pipe_ctas > 0 && n_steps > 1 ? "step_pipe" : "step_fused_grid")
                                                               : (x->pipe_ctas > 0 && n_steps > 1 ? "step_shard_pipe" : "step_shard")));
  return 0;
}

extern "C" int bh_step(const bh_ctx* x, const uint32_t* in, int learning, void* stream) {
  DevGuard dev_guard_(x);
  int rc = check_ctx(x);
  if (rc) return rc;
  if (learning < 0 || learning > 3) return BH_E_UNSUPPORTED;
  if (x->fused_mode) return launch_fused(x, in, 1, learning, 0, S_(stream));
  const int learn = learning & BH_STEP_LEARNING, winners = !(learning & BH_STEP_NO_WINNER_CELLS);
  if ((rc = sp_step(x, in, learn, S_(stream)))) return rc;
  return bh_tm_step_ex(x, learn, winners, winners, stream);
}

extern "C" int bh_step_launches(const bh_ctx* x, int learning) {
  DevGuard dev_guard_(x);
  if (x && x->fused_mode) return 1;
  // overlap+boost, topk, [sp_learn], duty | draw1, select a/b | learn-select a/b, draw2, [chunks], [apply] |
  // post, activate a, draw3, activate b
  return (learning ? 15 : 13) + (x && x->jump_polys > 0 ? 1 : 0);
}

__global__ void k_ring_fetch(const __grid_constant__ bh_ctx c) {
  // copy the next ring row to input_dev and advance the cursor (single CTA)
  int pos = c.sc[BH_SC_INPUT_POS];
  const uint32_t* src = c.input_ring + (long long)(pos % c.ring_len) * c.input_words;
  for (int i = threadIdx.x; i < c.input_words; i += blockDim.x) c.input_dev[i] = src[i];
  __syncthreads();
  if (threadIdx.x == 0) c.sc[BH_SC_INPUT_POS] = pos + 1;
}

extern "C" int bh_step_ring(const bh_ctx* x, int learning, void* stream) {
  DevGuard dev_guard_(x);
  if (!x || x->ring_len <= 0) return BH_E_BADARG;
  if (x->fused_mode) {
    int rc = check_ctx(x);
    return rc ? rc : launch_fused(x, nullptr, 1, learning, 0, S_(stream));
  }
  k_ring_fetch<<<1, 256, 0, S_(stream)>>>(*x);
  LAUNCHED("ring_fetch");
  return bh_step(x, x->input_dev, learning, stream);
}

// pack bool bytes on the host: bit i of word i/32 (little-endian bit order)
static void pack_host(const bh_ctx* x, const uint8_t* input_bool_host) {
  for (int w = 0; w < x->input_words; ++w) {
    uint32_t bits = 0;
    int lim = x->input_dim - w * 32;
    if (lim > 32) lim = 32;
    const uint8_t* p = input_bool_host + w * 32;
    if (lim == 32) {
      // 8 bool bytes (0 / 1, np.bool_) -> 8 bits with one multiply: byte i lands on bit 56 + i
      for (int g = 0; g < 4; ++g) {
        uint64_t v;
        memcpy(&v, p + 8 * g, 8);
        v = (v | (v >> 1) | (v >> 2) | (v >> 3) | (v >> 4) | (v >> 5) | (v >> 6) | (v >> 7)) & 0x0101010101010101ULL;  // != 0
        bits |= (uint32_t)((v * 0x0102040810204080ULL) >> 56) << (8 * g);
      }
    } else {
      for (int b = 0; b < lim; ++b) bits |= (uint32_t)(p[b] != 0) << b;
    }
    x->input_pinned[w] = bits;
  }
}

static int step_host_enqueue(const bh_ctx* x, int learning, cudaStream_t st) {
  int rc;
  CU_RET(cudaMemcpyAsync(x->input_dev, x->input_pinned, (size_t)x->input_words * 4, cudaMemcpyHostToDevice, st));
  if (x->fused_mode) {
    if ((rc = launch_fused(x, x->input_dev, 1, learning, 1, st))) return rc;  // step + summary gather
  } else {
    if ((rc = bh_step(x, x->input_dev, learning, st))) return rc;
    int k = x->active_columns;
    int n = k > BH_MT_N + 1 ? k : BH_MT_N + 1;
    k_summary<<<cdiv(n, 256) < 64 ? cdiv(n, 256) : 64, 256, 0, st>>>(*x);
    LAUNCHED("summary");
  }
  CU_RET(cudaMemcpyAsync(x->summary_pinned, x->summary_dev, (size_t)BH_SUMMARY_INTS(x->active_columns) * 4,
                         cudaMemcpyDeviceToHost, st));
  return 0;
}

extern "C" int bh_host_graph_create(const bh_ctx* x, int learning, void* stream, void** out) {
  DevGuard dev_guard_(x);
  if (!out) return BH_E_BADARG;
  int rc = check_ctx(x);
  if (rc) return rc;
  if (!x->input_pinned || !x->summary_pinned) return BH_E_BADARG;
  const bool zero_copy = x->fused_mode == 1 || x->fused_mode == 2;
  if (x->fused_mode && (rc = prepare_fused(x->fused_mode))) return rc;
  cudaStream_t st = S_(stream);
  cudaGraph_t graph = nullptr;
  CU_RET(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  if (zero_copy) {  // the kernel alone: pinned host input in, pinned host summary + flag out
    bh_ctx y = *x;
    y.summary_dev = x->summary_pinned;
    rc = launch_fused(&y, x->input_pinned, 1, learning, 2, st);
  } else {
    rc = step_host_enqueue(x, learning, st);
  }
  cudaError_t e = cudaStreamEndCapture(st, &graph);
  if (rc) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess) return -(1000 + (int)e);
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return -(1000 + (int)e);
  *out = exec;
  return 0;
}

extern "C" int bh_step_host_graph(const bh_ctx* x, void* graph_exec, const uint8_t* input_bool_host,
                                  int32_t* summary_host, void* stream) {
  DevGuard dev_guard_(x);
  if (!x || !graph_exec || !input_bool_host) return BH_E_BADARG;
  cudaStream_t st = S_(stream);
  pack_host(x, input_bool_host);
  const bool zero_copy = x->fused_mode == 1 || x->fused_mode == 2;
  // Spin on host memory the device writes last -- the kernel's completion flag (zero-copy step) or
  // word 0 of the summary the D2H node delivers -- instead of sleeping in the driver.
  volatile int32_t* flag = zero_copy ? x->summary_pinned + BH_SUMMARY_INTS(x->active_columns) : x->summary_pinned;
  flag[0] = -1;
  CU_RET(cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(graph_exec), st));
  long spins = 0;
  for (; flag[0] == -1 && spins < 40000000; ++spins) {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  if (!zero_copy || flag[0] == -1) CU_RET(cudaStreamSynchronize(st));  // copy nodes / a kernel that faulted
  if (summary_host) memcpy(summary_host, x->summary_pinned, (size_t)BH_SUMMARY_INTS(x->active_columns) * 4);
  return 0;
}

extern "C" int bh_step_host(const bh_ctx* x, const uint8_t* input_bool_host, int learning, int32_t* summary_host,
                            void* stream) {
  DevGuard dev_guard_(x);
  int rc = check_ctx(x);
  if (rc) return rc;
  if (!input_bool_host || !x->input_pinned || !x->summary_pinned) return BH_E_BADARG;
  cudaStream_t st = S_(stream);
  pack_host(x, input_bool_host);
  if ((rc = step_host_enqueue(x, learning, st))) return rc;
  CU_RET(cudaStreamSynchronize(st));
  if (summary_host) memcpy(summary_host, x->summary_pinned, (size_t)BH_SUMMARY_INTS(x->active_columns) * 4);
  return 0;
}

extern "C" int bh_summary(const bh_ctx* x, int32_t* summary_host, void* stream) {
  DevGuard dev_guard_(x);
  if (!x || !x->summary_pinned) return BH_E_BADARG;
  cudaStream_t st = S_(stream);
  int k = x->active_columns;
  int n = k > BH_MT_N + 1 ? k : BH_MT_N + 1;
  k_summary<<<cdiv(n, 256) < 64 ? cdiv(n, 256) : 64, 256, 0, st>>>(*x);
  LAUNCH_CHECK();
  size_t bytes = (size_t)BH_SUMMARY_INTS(k) * 4;
  CU_RET(cudaMemcpyAsync(x->summary_pinned, x->summary_dev, bytes, cudaMemcpyDeviceToHost, st));
  CU_RET(cudaStreamSynchronize(st));
  if (summary_host) memcpy(summary_host, x->summary_pinned, bytes);
  return 0;
}

// One step with a CUDA event after every launch; per-launch milliseconds and names.
extern "C" int bh_profile_step(const bh_ctx* x, const uint32_t* in, int learning, void* stream, float* ms_out,
                               const char** names_out, int max_out) {
  DevGuard dev_guard_(x);
  int rc = check_ctx(x);
  if (rc) return rc;
  ProfileRec rec;
  rec.n = 0;
  rec.st = S_(stream);
  for (int i = 0; i <= BH_PROF_MAX; ++i) CU_RET(cudaEventCreate(&rec.ev[i]));
  CU_RET(cudaEventRecord(rec.ev[0], rec.st));
  g_prof = &rec;
  rc = bh_step(x, in, learning, stream);
  g_prof = nullptr;
  cudaError_t e = cudaStreamSynchronize(rec.st);
  int n = 0;
  if (rc == 0 && e == cudaSuccess) {
    for (; n < rec.n && n < max_out; ++n) {
      cudaEventElapsedTime(&ms_out[n], rec.ev[n], rec.ev[n + 1]);
      if (names_out) names_out[n] = rec.name[n];
    }
  }
  for (int i = 0; i <= BH_PROF_MAX; ++i) cudaEventDestroy(rec.ev[i]);
  if (rc) return rc;
  if (e != cudaSuccess) return -(1000 + (int)e);
  return n;
}

// ------------------------------------------------------------------------------------
// CUDA graphs over bh_step_ring
// ------------------------------------------------------------------------------------
extern "C" int bh_graph_create(const bh_ctx* x, int steps_per_graph, int learning, void* stream, void** out) {
  DevGuard dev_guard_(x);
  if (!out || steps_per_graph < 1) return BH_E_BADARG;
  int rc = check_ctx(x);
  if (rc) return rc;
  cudaStream_t st = S_(stream);
  cudaGraph_t graph = nullptr;
  if (x->fused_mode && (rc = prepare_fused(x->fused_mode))) return rc;  // not inside the capture
  CU_RET(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  if (x->fused_mode) {
    rc = launch_fused(x, nullptr, steps_per_graph, learning, 0, st);  // all steps in one launch
  } else {
    for (int i = 0; i < steps_per_graph && rc == 0; ++i) rc = bh_step_ring(x, learning, stream);
  }
  cudaError_t e = cudaStreamEndCapture(st, &graph);
  if (rc) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess) return -(1000 + (int)e);
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return -(1000 + (int)e);
  *out = exec;
  return 0;
}

// n independent networks in one graph: fork from `stream` onto side streams, one fused launch per
// network, join.  The side streams exist only during the capture.
extern "C" int bh_batch_graph_create(const bh_ctx* const* ctxs, int n, int steps_per_graph, int learning,
                                     void* stream, void** out) {
  if (!out || !ctxs || n < 1 || steps_per_graph < 1) return BH_E_BADARG;
  DevGuard dev_guard_(ctxs[0]);
  for (int i = 0; i < n; ++i) {
    if (ctxs[i] && ctxs[i]->device != ctxs[0]->device) return BH_E_BADARG;  // one graph, one device
    int rc = check_ctx(ctxs[i]);
    if (rc) return rc;
    if (!ctxs[i]->fused_mode || ctxs[i]->ring_len <= 0) return BH_E_UNSUPPORTED;
    if ((rc = prepare_fused(ctxs[i]->fused_mode))) return rc;
  }
  cudaStream_t st = S_(stream);
  // one capture stream per concurrent kernel the device can hold (148 SMs / the smallest useful cluster)
  const int NS = n < 64 ? n : 64;
  cudaStream_t side[64];
  cudaEvent_t fork, join[64];
  for (int i = 0; i < NS; ++i) {
    CU_RET(cudaStreamCreateWithFlags(&side[i], cudaStreamNonBlocking));
    CU_RET(cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming));
  }
  CU_RET(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
  cudaGraph_t graph = nullptr;
  int rc = 0;
  CU_RET(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  cudaError_t e = cudaEventRecord(fork, st);
  for (int i = 0; i < NS && e == cudaSuccess; ++i) e = cudaStreamWaitEvent(side[i], fork, 0);
  for (int i = 0; i < n && rc == 0 && e == cudaSuccess; ++i)
    rc = launch_fused(ctxs[i], nullptr, steps_per_graph, learning, 0, side[i % NS]);
  for (int i = 0; i < NS && e == cudaSuccess; ++i) {
    e = cudaEventRecord(join[i], side[i]);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, join[i], 0);
  }
  cudaError_t e2 = cudaStreamEndCapture(st, &graph);
  for (int i = 0; i < NS; ++i) {
    cudaStreamDestroy(side[i]);
    cudaEventDestroy(join[i]);
  }
  cudaEventDestroy(fork);
  if (rc || e != cudaSuccess || e2 != cudaSuccess) {
    if (graph) cudaGraphDestroy(graph);
    return rc ? rc : -(1000 + (int)(e != cudaSuccess ? e : e2));
  }
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return -(1000 + (int)e);
  *out = exec;
  return 0;
}

extern "C" int bh_graph_launch(void* graph_exec, void* stream) {
  CU_RET(cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(graph_exec), S_(stream)));
  return 0;
}

extern "C" int bh_graph_destroy(void* graph_exec) {
  if (graph_exec) CU_RET(cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(graph_exec)));
  return 0;
}

// ------------------------------------------------------------------------------------
// randomness + test hooks
// ------------------------------------------------------------------------------------
extern "C" int bh_rng_import(const bh_ctx* x, void* stream) {
  DevGuard dev_guard_(x);
  int rc = check_ctx(x);
  if (rc) return rc;
  k_rng_import<<<1, 256, 0, S_(stream)>>>(*x);
  LAUNCH_CHECK();
  return 0;
}

extern "C" int bh_rng_export(const bh_ctx* x, void* stream) {
  DevGuard dev_guard_(x);
  int rc = check_ctx(x);
  if (rc) return rc;
  k_rng_export<<<1, 256, 0, S_(stream)>>>(*x);
  LAUNCH_CHECK();
  return 0;
}

extern "C" int bh_rng_fill(const bh_ctx* x, double* dst_dev, int64_t count, void* stream) {
  DevGuard dev_guard_(x);
  if (!x || !dst_dev || count < 0 || 2 * count > x->rng_step_words) return BH_E_BADARG;
  cudaStream_t st = S_(stream);
  k_rng_fill_draw<<<1, MT_THREADS, 0, st>>>(*x, (long long)count);
  LAUNCH_CHECK();
  if (x->jump_polys > 0) {
    int prc = prepare_chunks();
    if (prc) return prc;
    int cap = (x->sm_count > 0 ? x->sm_count : 148);
    k_rng_chunks<<<x->jump_polys < cap ? x->jump_polys : cap, MT_THREADS, RNG_CHUNK_SMEM, st>>>(*x);
    LAUNCH_CHECK();
  }
  int grid = cdiv(count > 0 ? count : 1, 256);
  k_rng_fill_copy<<<grid < 1184 ? grid : 1184, 256, 0, st>>>(*x, dst_dev);
  LAUNCH_CHECK();
  return 0;
}

__global__ void k_test_np_expf(const float* x, float* y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = bh_np_expf(x[i]);
}

extern "C" int bh_test_np_expf(const float* x_dev, float* y_dev, int64_t n, void* stream) {
  k_test_np_expf<<<1184, 256, 0, S_(stream)>>>(x_dev, y_dev, (long long)n);
  LAUNCH_CHECK();
  return 0;
}
