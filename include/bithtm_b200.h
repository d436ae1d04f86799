/*
 * bithtm_b200 -- C ABI of the B200 (sm_100a) implementation of the bitHTM
 * spatial-pooler + temporal-memory timestep.
 *
 * The reference (cokwa/bitHTM) has NO FFI: its plugin boundary is duck-typed
 * Python constructor injection (bithtm/networks.py:14-24, 48-55, 132-144).  The
 * entry points below are what a ctypes binding of that boundary calls; each one
 * cites the reference method it replaces.  The Python classes in
 * bithtm_b200/{networks,projections,regularizations}.py are that binding.
 *
 * Conventions
 *  - The library is STATELESS: every call takes a `bh_ctx` (plain struct of
 *    sizes, constants and DEVICE pointers owned by the caller).  Nothing is
 *    allocated per call.  `bh_layout` tells the caller how to carve one arena.
 *  - Every call is asynchronous on `stream` (a cudaStream_t passed as void*),
 *    returns 0 or a negative error (-cudaError or BH_E_*), never throws.
 *  - Capacity overflows (segments, synapses/segment, matching list, random
 *    buffer) are recorded in the device status word sc[BH_SC_STATUS] and
 *    checked lazily by the caller.
 *  - No CPU fallback exists: without a CUDA device every compute call fails.
 */
#ifndef BITHTM_B200_H
#define BITHTM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BH_ABI_VERSION 11
#define BH_MT_N 624
#define BH_SUMMARY_INTS(k) (4 + 4 * (k) + BH_MT_N + 1 + 4)
#define BH_TOPK_WS_INTS 81920

/* error codes (negative); CUDA errors are returned as -(1000 + cudaError_t) */
#define BH_E_BADARG (-1)
#define BH_E_NODEVICE (-2)
#define BH_E_UNSUPPORTED (-3)

/* device scalar block: int32 sc[BH_SC_COUNT] */
enum {
  BH_SC_STEP = 0,      /* timesteps completed; parity selects ping-pong buffers   */
  BH_SC_HAVE_PREV,     /* 0 until the first activation (prev distal state None)    */
  BH_SC_NSEG,          /* segments allocated (reference: len(segment_bundle))      */
  BH_SC_NSEG_NEXT,     /* staged by learn-select, committed before learn-apply     */
  BH_SC_M,             /* matching segments of the last activation                 */
  BH_SC_W0,            /* winner-cell count, buffer 0                              */
  BH_SC_W1,            /* winner-cell count, buffer 1                              */
  BH_SC_L0,            /* learning segments chosen among matching ones             */
  BH_SC_L,             /* rows of the learning list (L0 + recycled + new)          */
  BH_SC_P,             /* punished segments                                        */
  BH_SC_NU,            /* winners with no matching segment (unaccounted)           */
  BH_SC_NR,            /* recycled segments this step                              */
  BH_SC_STATUS,        /* BH_ST_* bits                                             */
  BH_SC_MT_POS,        /* position inside mt_key (0..624) of an imported/exported   */
                       /* MT19937 state (bh_rng_import / bh_rng_export)             */
  BH_SC_X_MATCH,       /* segment shards: matching segments over all ranks (merged)  */
  BH_SC_X_RECYC_AVAIL, /* recyclable segment ids received (merged, ascending)       */
  BH_SC_X_RECYC_TOTAL, /* recyclable segments over all ranks (true count)           */
  BH_SC_INPUT_POS,     /* cursor into the device input ring (bh_step_ring)         */
  BH_SC_BAR_COUNT,     /* grid barrier of the fused kernel: arrivals               */
  BH_SC_BAR_GEN,       /*                                   generation             */
  BH_SC_JIT_PENDING,   /* the last activation did not draw its jitter (rand(M)); a  */
                       /* later step that needs it draws it first (projections.py:  */
                       /* 229-243 run lazily from networks.py:76)                   */
  BH_SC_WNONE0,        /* winner_cell of buffer 0 / 1 is None (a step with neither  */
  BH_SC_WNONE1,        /* learning nor return_winner_cell): no growth, no rand(L,W+1) */
  BH_SC_NGROW,         /* learning segments of this step that grow synapses (n_add > 0,  */
                       /* projections.py:114-115): the rows of rand(L, W+1) that are read */
  BH_SC_BAR2_COUNT,    /* barrier of the CTA team that does the temporal-memory         */
  BH_SC_BAR2_GEN,      /* bookkeeping while the other CTAs learn the spatial pooler     */
  BH_SC_NPREDCOL,      /* columns with a predicted cell after the last activation        */
  BH_SC_NPREDCOL_PREV, /* ... before this step's learning (example.py:50, the demo's metrics) */
  BH_SC_T5_ERR,        /* a barrier wait of the tcgen05 batched overlap timed out          */
  BH_SC_BAR3_COUNT,    /* barrier of the spatial-pooler team of the two-pipeline step kernel */
  BH_SC_BAR3_GEN,
  BH_SC_PIPE_SPLIT,    /* CTAs of the temporal-memory team in the current step (two-pipeline kernel)  */
  BH_SC_COUNT = 32
};

/* status bits */
#define BH_ST_SEG_OVERFLOW 1   /* more segments than seg_capacity                  */
#define BH_ST_SYN_OVERFLOW 2   /* a segment needed more than syn_capacity slots    */
#define BH_ST_MATCH_OVERFLOW 4 /* more matching segments than match_capacity       */
#define BH_ST_LEARN_OVERFLOW 8 /* more learning rows than learn_capacity           */
#define BH_ST_RAND_OVERFLOW 16 /* a step drew more uniforms than the stream ring holds */
#define BH_ST_PRI_TIE 32       /* equal priorities straddled a growth cut (the     */
                               /* reference's np.argsort is undefined there)       */
#define BH_ST_XCH_OVERFLOW 64  /* segment shards: a rank had more matching or      */
                               /* recyclable segments than the exchange carries    */
#define BH_ST_XCH_TIMEOUT 128  /* fused sharded step: a peer's record never arrived */
#define BH_MAX_RANKS 8

typedef struct bh_ctx {
  /* ---- sizes --------------------------------------------------------------- */
  int32_t input_dim;       /* I                                                    */
  int32_t input_words;     /* ceil(I / 32)                                         */
  int32_t mask_stride;     /* uint32 words per connected-mask row, multiple of 4   */
  int32_t column_dim;      /* C                                                    */
  int32_t cell_dim;        /* c, 1..32.  CAP: a column's cells are the bits of ONE 32-bit word in   */
                           /* col_pred / col_act / col_win, the row words of the step summary and the */
                           /* warp ballots that form winner cells (lane = cell); the reference accepts  */
                           /* any cell_dim (networks.py:48-55), its default and every BASELINE config  */
                           /* use 32.  bh_layout / check_ctx return BH_E_UNSUPPORTED beyond 32.          */
  int32_t active_columns;  /* k                                                    */
  int32_t seg_capacity;    /* S_cap                                                */
  int32_t syn_capacity;    /* E_cap: synapse slots per segment, multiple of 32     */
  int32_t match_capacity;  /* M_cap                                                */
  int32_t learn_capacity;  /* L_cap                                                */
  int32_t tm_blocks;       /* NB: CTAs of the ranged TM kernels (<= 1024)          */
  int32_t sm_count;        /* SMs of the device (grid sizing)                      */
  int64_t rng_ring_words;  /* words in rng_ring: a power of two >= 2^20 and              */
                           /* >= 2 * rng_step_words                                     */
  int64_t rng_step_words;  /* most stream words (2 per float64) one timestep may draw  */
  int32_t col_lo;          /* column shard: first global column owned by this rank  */
  int32_t col_local;       /* columns owned (C when not sharded); the SP buffers    */
                           /* sp_perm/sp_mask/duty/overlaps/boosted are local-sized */
  int32_t ring_len;        /* rows in input_ring (0 = none)                        */
  int32_t fused_mode;      /* bh_step*: 0 = one kernel per stage, 1 = one kernel on a  */
                           /* thread-block cluster, 2 = one cooperative-grid kernel,  */
                           /* 3 = one cooperative-grid kernel per SHARD with the two  */
                           /* exchanges done in-kernel over peer memory (xpeer)       */
  int32_t seg_rank;        /* segment shard: this rank holds the synapse rows of the   */
  int32_t seg_world;       /* segments whose 64-id block b has b % seg_world ==       */
                           /* seg_rank (1 = all); syn_cell/syn_perm hold local rows   */
  int32_t xm_cap;          /* exchange capacity per rank: matching segments           */
  int32_t xr_cap;          /*                             recyclable segment ids      */
  int32_t jump_polys;      /* rows of mt_jump (0 = the stream is produced by one CTA)  */
  int32_t rng_lookahead;   /* stream words draw #2 produces beyond its own need (the   */
                           /* following rand(M) and next step's rand(k, c))            */
  int32_t device;          /* CUDA device ordinal the buffers live on: every entry point */
                           /* makes it current for the duration of the call (-1 = keep  */
                           /* the calling thread's current device)                     */
  int32_t skip_polys;      /* rows of mt_skip (0 = every draw is materialised): row b - 1 =  */
                           /* t^(b * skip_gran) mod phi, b = 1..skip_polys             */
  int32_t skip_gran;       /* stream words per jump-table step (a power of two)        */
  int32_t job_cap;         /* slots of rng_jump: production jobs of one lazy step       */
  int32_t lazy_policy;     /* 0: a step is lazy once few rows grew in the last two steps  */
                           /* (dense production is cheaper while most rows grow); 1: always */
  int32_t tail_chunks;     /* chunks the words after a skipped matrix are produced in (one jump */
                           /* each; 0 = 4)                                                     */
  int32_t xch_ll;          /* fused_mode 3: exchanges as 8-byte {word, sequence} cells stored into */
                           /* the peers' regions (csrc/shard_ll.cuh); 0 = copy + fence + flag    */
  int32_t pipe_ctas;  /* fused_mode 2: > 0 = two-pipeline kernel, CTAs of its temporal-memory team */
  int64_t skip_min;        /* a rand(L, W+1) of at least this many stream words may be  */
                           /* drawn lazily (fused_mode >= 2): only the rows of growing  */
                           /* segments are produced, by jumps (csrc/mt19937.cuh)        */

  /* ---- constants, evaluated on the host with the reference's expressions ----- */
  double sp_threshold;     /* projections.py:19  permanence >= threshold           */
  double sp_delta_on;      /* projections.py:24  1.0*(inc+dec)-dec                 */
  double sp_delta_off;     /* projections.py:24  0.0*(inc+dec)-dec                 */
  double tm_learn_on;      /* projections.py:102 True *(da-di)+di, learn           */
  double tm_learn_off;     /*                    False*(da-di)+di, learn           */
  double tm_punish_on;     /*                    punish                            */
  double tm_punish_off;
  float boost_coef;        /* regularizations.py:16  f32(-(intensity/density))     */
  float duty_momentum;     /* regularizations.py:20  f32(momentum)                 */
  float duty_increment;    /* regularizations.py:21  f32(1.0 - momentum)           */
  float tm_perm_initial;   /* projections.py:149 f32(0.21)                         */
  float tm_perm_threshold; /* projections.py:171 f32(0.5)                          */
  float epsilon;           /* networks.py:91     f32(1e-8)                         */
  int32_t tm_learn_can_delete;  /* projections.py:105 min(da,di) < 0               */
  int32_t tm_punish_can_delete;
  int32_t seg_activation_threshold;
  int32_t seg_matching_threshold;
  int32_t seg_sampling_synapses;
  int32_t fused_ctas;      /* CTAs of the fused kernel (cluster size <= 16, or grid)  */
  int32_t fused_threads;   /* threads per CTA of the cluster kernel (fused_mode 1): a   */
                           /* multiple of 32 in [256, 1024]; 0 = 1024.  Fewer threads  */
                           /* let several CTAs share an SM (independent streams)       */

  /* ---- spatial pooler (DenseProjection / ExponentialBoosting / inhibition) --- */
  double* sp_perm;         /* [C][I] float64 permanence, row-major                 */
  uint32_t* sp_mask;       /* [C][mask_stride] connected bits (perm >= threshold)  */
  float* duty;             /* [C]                                                  */
  int32_t* overlaps;       /* [C]                                                  */
  double* boosted;         /* [C]                                                  */
  int32_t* active_cols;    /* [2][k] ping-pong by step parity, reference order     */
  uint8_t* col_active;     /* [C] 1 for the current active columns                 */

  /* ---- temporal memory: per column (bit b = cell b) and per cell.  Cells are    */
  /* addressed on the device as column * 32 + cell; N32 = 32 * C.                  */
  uint32_t* col_pred;      /* [C] cell_prediction        networks.py:122           */
  uint32_t* col_act;       /* [2][C] cell_activation     networks.py:118-119; buffer  */
                           /* step & 1 = this step's, the other = the previous step's   */
  uint32_t* col_win;       /* [C] winner cells of the current step                 */
  int32_t* cell_nseg;      /* [N32] bundle_segments        projections.py:227        */
  float* cell_maxjit;      /* [N32] max_jittered_potential projections.py:236-237    */
  int32_t* cell_npred;     /* [N32] prediction (active segments per cell)  :251      */
  int32_t* cell_widx;      /* [N32] index in the previous winner list or -1          */

  /* ---- segments (rows kept compact: valid synapses are slots [0, count)) ----- */
  int32_t* seg_owner;      /* [S_cap] segment_bundle     projections.py:226        */
  int32_t* seg_count;      /* [S_cap] output_edges       projections.py:42         */
  int32_t* seg_pot;        /* [S_cap] segment_potential  projections.py:246        */
  int32_t* seg_conn;       /* [S_cap] connected-active count                       */
  int32_t* syn_cell;       /* [rows][E_cap] presynaptic cell (column * 32 + cell);  */
  float* syn_perm;         /* [rows][E_cap] float32 permanence; rows = S_cap, or    */
                           /* the locally held share of it (segment shards)         */

  /* ---- per-step lists ---------------------------------------------------------- */
  uint32_t* row_pred;      /* [k] prev prediction bits of each active column       */
  uint32_t* row_act;       /* [k] active cells   (networks.py:115)                 */
  uint32_t* row_win;       /* [k] winner cells   (networks.py:102)                 */
  uint32_t* row_unacc;     /* [k] winners without a matching segment               */
  int32_t* winners;        /* [2][k*c] ordered flat winner cells, ping-pong        */
  int32_t* unacc;          /* [k*c] ordered unaccounted winners                    */
  int32_t* m_seg;          /* [M_cap] matching_segment, ascending id               */
  int32_t* m_conn;         /* [M_cap] matching_segment_activation                  */
  float* m_jit;            /* [M_cap] matching_segment_jittered_potential          */
  uint8_t* m_flag;         /* [M_cap] bit0 learn, bit1 punish                      */
  int32_t* learn_list;     /* [L_cap] learning_segment (projections.py:281 order)  */
  int32_t* punish_list;    /* [M_cap] punished_segment                             */
  int32_t* recyc_list;     /* [seg_world * xr_cap] segment shards: merged recyclable ids */
  int32_t* x_send;         /* [max record] this rank's exchange record (fused_mode 3)    */
  double* xk_keys;         /* [seg_world * k_loc] gathered top-k candidate keys          */
  int32_t* xk_cols;        /* [seg_world * k_loc] and their global columns               */
  int32_t* blk;            /* [8][1024] per-CTA counts for ordered compaction      */
  int32_t* topk_ws;        /* [BH_TOPK_WS_INTS] workspace of the multi-CTA top-k   */

  /* ---- randomness: legacy MT19937 stream of np.random, produced on the device -- */
  uint32_t* mt_key;        /* [624] staging of an imported / exported state         */
  uint32_t* rng_ring;      /* [rng_ring_words] raw stream words by absolute index   */
  uint32_t* mt_jump;       /* [jump_polys][624] jump polynomials (caller-filled,    */
                           /* bithtm_b200/_mtjump.py) for multi-CTA production      */
  long long* rng64;        /* [32] producer / consumer cursors (csrc/mt19937.cuh)   */
  uint32_t* mt_skip;       /* [skip_polys][624] jump table (caller-filled, _mtjump.skip_table) */
  uint32_t* rng_jump;      /* [job_cap][640] jump results of a lazy step (zero between uses)   */
  int32_t* grow_list;      /* [learn_capacity][3] (row, synapses kept, n_add) of growing rows */

  /* ---- fused sharded step: exchange regions of all ranks, mapped into this process ---- */
  int32_t* xpeer[BH_MAX_RANKS]; /* xpeer[r] = base of rank r's region of bh_xch_region_ints() */
                           /* int32 (zero-filled by the caller; xpeer[seg_rank] is local)  */

  /* ---- scalars, input ring, host staging ---------------------------------------- */
  int32_t* sc;             /* [BH_SC_COUNT]                                        */
  uint32_t* input_ring;    /* [ring_len][input_words] packed inputs (bh_step_ring) */
  uint32_t* input_dev;     /* [input_words] staging for bh_step_host               */
  uint32_t* input_pinned;  /* HOST pinned [input_words]                            */
  int32_t* summary_dev;    /* [BH_SUMMARY_INTS(k)] step summary (see bh_step_host) */
  int32_t* summary_pinned; /* HOST pinned [BH_SUMMARY_INTS(k) + 4]; the extra words  */
                           /* are the completion flag of the zero-copy host step   */
} bh_ctx;

/* Arena layout: fills every DEVICE pointer of `ctx` with `base + offset`, given
 * the sizes already set in ctx.  With base == NULL only returns the byte count.
 * All sub-buffers are 256-byte aligned.  The caller zero-fills the arena and
 * then calls bh_init. */
size_t bh_layout(bh_ctx* ctx, void* base);

/* Set the non-zero initial values (cell_widx = -1, ...).  Arena must be zeroed. */
int bh_init(const bh_ctx* ctx, void* stream);

int bh_abi_version(void);
/* sizeof(bh_ctx) as compiled, so a binding can verify its struct mirror */
size_t bh_ctx_size(void);
int bh_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);

/* ---- spatial pooler ------------------------------------------------------------- */
/* Build the connected mask from sp_perm (after uploading the permanence drawn by
 * DenseProjection.__init__, projections.py:16). */
int bh_sp_build_mask(const bh_ctx* ctx, void* stream);
/* Pack a device bool[I] (one byte per bit) into input words. */
int bh_pack_input(const bh_ctx* ctx, const uint8_t* bool_dev, uint32_t* words_dev, void* stream);
/* Pack n_inputs device bool vectors [n_inputs][I] into rows of pitch_words (>= input_words) words, zero padded:
 * the input operand of the batched overlaps below. */
int bh_pack_inputs(const bh_ctx* ctx, const uint8_t* bool_dev, int n_inputs, int pitch_words, uint32_t* words_dev,
                   void* stream);
/* DenseProjection.process (projections.py:18-21) -> ctx->overlaps */
int bh_sp_overlap(const bh_ctx* ctx, const uint32_t* input_words_dev, void* stream);
/* DenseProjection.process for n_inputs input vectors against the ONE connected mask of this
 * network (a loop of projections.py:18-21 over inputs that share the projection, e.g. inference
 * over many streams with SP learning off): inputs_dev [n_inputs][input_words] packed words,
 * overlaps_out [n_inputs][col_local] int32.  Bit-packed AND + popcount on the integer pipe. */
int bh_sp_overlap_batched(const bh_ctx* ctx, const uint32_t* inputs_dev, int n_inputs, int32_t* overlaps_out,
                          void* stream);
/* Same contract and same results as bh_sp_overlap_batched, computed as an int8 tensor-core
 * contraction [n_inputs x I] . [I x C] with int32 accumulation (exact); the operands stay bit-packed
 * in HBM.  bh_sp_overlap_batched_tc picks the kernel:
 *  - bh_sp_overlap_batched_tc5: tcgen05.mma kind::i8 (128 x 256 x 32 per instruction), accumulator in TMEM,
 *    packed words brought in by TMA and widened to bytes in shared memory (csrc/overlap_tcgen05.cuh).
 *    Needs input_words % 4 == 0 and a 16-byte aligned inputs_dev (TMA pitches), else BH_E_UNSUPPORTED; a
 *    pipeline fault raises sc[BH_SC_T5_ERR] instead of hanging.
 *  - bh_sp_overlap_batched_mma: mma.sync m16n8k32 u8, operands widened in registers: small or ragged
 *    shapes.  BH_E_UNSUPPORTED for 2^24 or more input bits (use bh_sp_overlap_batched). */
int bh_sp_overlap_batched_tc(const bh_ctx* ctx, const uint32_t* inputs_dev, int n_inputs, int32_t* overlaps_out,
                             void* stream);
int bh_sp_overlap_batched_tc5(const bh_ctx* ctx, const uint32_t* inputs_dev, int n_inputs, int32_t* overlaps_out,
                              void* stream);
int bh_sp_overlap_batched_mma(const bh_ctx* ctx, const uint32_t* inputs_dev, int n_inputs, int32_t* overlaps_out,
                              void* stream);
/* ExponentialBoosting.process (regularizations.py:15-17) -> ctx->boosted */
int bh_boost(const bh_ctx* ctx, void* stream);
/* GlobalInhibition.process (regularizations.py:28-29) with the canonical rule
 * (larger first, ties -> lower column, ascending output) -> active_cols[cur] */
int bh_inhibit(const bh_ctx* ctx, void* stream);
/* Host-inhibition mode: take an explicit ORDERED active_column list (device). */
int bh_set_active_columns(const bh_ctx* ctx, const int32_t* cols_dev, void* stream);
/* DenseProjection.update (projections.py:23-24) on active_cols[cur] */
int bh_sp_learn(const bh_ctx* ctx, const uint32_t* input_words_dev, void* stream);
/* ExponentialBoosting.update (regularizations.py:19-21) */
int bh_duty_update(const bh_ctx* ctx, void* stream);
/* SpatialPooler.process (networks.py:26-35) = the five calls above */
int bh_sp_step(const bh_ctx* ctx, const uint32_t* input_words_dev, int learning, void* stream);

/* ---- column-sharded spatial pooler (exchange 1) -----------------------
 * Each rank owns columns [col_lo, col_lo + col_local): their permanence rows, mask rows
 * and duty cycles.  bh_sp_shard_local computes the local overlaps / boosted keys and
 * this shard's best min(k, col_local) candidates (canonical order) as (key, global
 * column) pairs, ascending column.  The caller all-gathers the pairs in rank order
 * (NCCL) and passes the n = world * min(k, col_local) gathered pairs to
 * bh_sp_shard_finish, which selects the global top-k on every rank identically, then
 * learns / updates duty cycles for the local columns.  The temporal memory that
 * follows is replicated. */
int bh_sp_shard_local(const bh_ctx* ctx, const uint32_t* input_words_dev, double* cand_keys_out,
                      int32_t* cand_cols_out, void* stream);
int bh_sp_shard_finish(const bh_ctx* ctx, const uint32_t* input_words_dev, const double* cand_keys,
                       const int32_t* cand_cols, int n, int learning, void* stream);

/* ---- segment-sharded temporal memory (exchange 2) ------------------------
 * The synapse rows are distributed over seg_world ranks by segment id (blocks of 64 ids,
 * round-robin); the small per-cell / per-column state and the segment owners are
 * replicated, so bursting, winner selection and the learning bookkeeping run
 * identically on every rank from the same inputs and the same MT19937 stream.  Per
 * timestep ONE exchange is needed: after the segment scan each rank contributes its
 * matching segments (id, potential, connected-active count; ascending id) and its
 * lowest recyclable segment ids; the caller all-gathers these fixed-size records in rank
 * order and every rank merges them into the global ascending lists the reference's
 * orderings are defined on (projections.py:80-81, 235, 247, 281).
 *
 * bh_tm_shard_xch_ints: int32 per rank record.  Record layout: [0] matching count,
 * [1] recyclable ids sent, [2] recyclable count (true), [3] status bits; id[xm_cap],
 * potential[xm_cap], connected[xm_cap]; recyclable id[xr_cap].
 * bh_tm_shard_pre : networks.py:95-119 + learning + the local segment scan; fills
 *                   send_dev with this rank's record.
 * bh_tm_shard_post: merges the seg_world gathered records (recv_dev, rank order), draws
 *                   rand(M), computes jitter / predictions; completes the timestep.
 * `learning` of bh_tm_shard_pre carries the BH_STEP_* flags; a step with
 * BH_STEP_NO_WINNER_CELLS ends with bh_tm_shard_post_ex(..., want_jitter = 0) (the jitter
 * draw stays pending, networks.py:121). */
size_t bh_tm_shard_xch_ints(const bh_ctx* ctx);
/* fused_mode 3: size (int32) of one rank's exchange region.  Every rank allocates one in
 * memory its peers can map (CUDA IPC / symmetric memory), zero-fills it and passes all
 * bases in ctx->xpeer; bh_step / bh_step_ring / bh_graph_* then run the shard's whole
 * timestep as one kernel that exchanges the two records over NVLink itself. */
size_t bh_xch_region_ints(const bh_ctx* ctx);
int bh_tm_shard_pre(const bh_ctx* ctx, int learning, int32_t* send_dev, void* stream);
int bh_tm_shard_post(const bh_ctx* ctx, const int32_t* recv_dev, void* stream);
int bh_tm_shard_post_ex(const bh_ctx* ctx, const int32_t* recv_dev, int want_jitter, void* stream);

/* Complete a timestep when no temporal memory follows (stand-alone SpatialPooler):
 * sc[BH_SC_STEP] += 1 so the ping-pong buffers rotate. */
int bh_advance_step(const bh_ctx* ctx, void* stream);

/* ---- temporal memory -------------------------------------------------------------- */
/* networks.py:95-104: bursting + winner cells (consumes rand(k, c)) */
int bh_tm_select(const bh_ctx* ctx, void* stream);
/* PredictiveProjection.update (projections.py:257-293) (consumes rand(L, W+1)) */
int bh_tm_learn(const bh_ctx* ctx, int learning, void* stream);
/* networks.py:115-128 + PredictiveProjection.process (projections.py:245-255)
 * (consumes rand(M)); completes the timestep (sc[BH_SC_STEP] += 1) */
int bh_tm_activate(const bh_ctx* ctx, void* stream);
/* TemporalMemory.process (networks.py:91-128) = the three calls above */
int bh_tm_step(const bh_ctx* ctx, int learning, void* stream);
/* The same with TemporalMemory.process's return_winner_cell flag (networks.py:91):
 * want_winner = 0 with learning = 0 is the inference-only step (no winner cells, no random
 * draw at all); want_jitter = return_winner_cell (networks.py:121): when 0 the rand(M) draw of
 * the activation is deferred to the next step that needs it, exactly as the reference's lazy
 * fill_jittered_potential_info does.  Per-stage kernels; not for segment shards. */
int bh_tm_step_ex(const bh_ctx* ctx, int learning, int want_winner, int want_jitter, void* stream);

/* ---- the distal projection as a stand-alone plugin, with explicit arguments --------------
 * For callers that drive PredictiveProjection themselves, the way the reference's own
 * TemporalMemory.process does (networks.py:106-113, 121): lists are device int32 arrays of cells
 * addressed as column * 32 + cell; not for segment shards.
 * bh_tm_learn_args    : PredictiveProjection.update (projections.py:257-293).  winner_cells =
 *                       learning_output (ordered), prev_winner_cells = winner_input (n_prev < 0: None),
 *                       prev_activation_words [C] = input_activation as one bit-word per column,
 *                       column_active [C] = 1 where output_punishment is False.  prev_state is the
 *                       activation this context computed last.  Consumes rand(L, W+1).
 * bh_tm_activate_cells: PredictiveProjection.process (projections.py:245-255) on an explicit list of
 *                       active cells; want_jitter = return_jittered_potential_info (consumes rand(M));
 *                       have_winners = a bh_tm_learn_args call of this timestep supplied the winner
 *                       cells.  Completes the timestep (sc[BH_SC_STEP] += 1).
 * bh_tm_fill_jitter   : PredictiveProjection.fill_jittered_potential_info (projections.py:229-239) for
 *                       an activation that deferred it; no-op otherwise.
 * bh_tm_reset         : TemporalMemory.process(prev_state = get_empty_state()) (networks.py:59-65,
 *                       91-93): forget the previous timestep's predictions, activation, winner cells and
 *                       distal state; the learned state is untouched. */
int bh_tm_learn_args(const bh_ctx* ctx, const int32_t* winner_cells_dev, int n_winners,
                     const int32_t* prev_winner_cells_dev, int n_prev_winners,
                     const uint32_t* prev_activation_words_dev, const uint8_t* column_active_dev, void* stream);
int bh_tm_activate_cells(const bh_ctx* ctx, const int32_t* active_cells_dev, int n_active, int want_jitter,
                         int have_winners, void* stream);
int bh_tm_fill_jitter(const bh_ctx* ctx, void* stream);
int bh_tm_reset(const bh_ctx* ctx, void* stream);

/* ---- whole timestep: HierarchicalTemporalMemory.process (networks.py:146-149) ------
 * `learning` of bh_step / bh_step_ring / bh_step_host* / bh_graph_* is a flag word: BH_STEP_LEARNING (1) =
 * the `learning` argument; BH_STEP_NO_WINNER_CELLS (2) = TemporalMemory.process(return_winner_cell=False)
 * (networks.py:91, 99, 121).  0 and 1 are the reference's defaults; 2 is the inference-only step (no winner
 * cells, no random draw, only the duty cycles change; one fused kernel like the learning step); 3 learns
 * without drawing the jitter of the activation (it is drawn by the next step that needs it).  Segment shards
 * (fused_mode 3) support 0 and 1. */
#define BH_STEP_LEARNING 1
#define BH_STEP_NO_WINNER_CELLS 2
int bh_step(const bh_ctx* ctx, const uint32_t* input_words_dev, int learning, void* stream);
/* Same, input taken from input_ring[sc[BH_SC_INPUT_POS]++ % ring_len] (no host
 * involvement; CUDA-graph friendly). */
int bh_step_ring(const bh_ctx* ctx, int learning, void* stream);
/* End-to-end call with HOST buffers: packs `input_bool_host` (I bytes), copies it
 * to the device, runs the step, copies the summary back and synchronises.
 * summary_host (BH_SUMMARY_INTS(k) int32): [0]=step index, [1]=status, [2]=n_segments,
 * [3]=winner count, then active_column[k], row_pred[k], row_act[k], row_win[k], then
 * the MT19937 state after the step (624 key words + position) so the caller can
 * keep np.random in lock-step, then [predicted columns before this step, predicted columns
 * after it, 0, 0]: with row_pred this gives example.py:55-57's bursting / correct / incorrect
 * column counts without reading cell_prediction back. */
int bh_step_host(const bh_ctx* ctx, const uint8_t* input_bool_host, int learning,
                 int32_t* summary_host, void* stream);

/* The same end-to-end step as ONE CUDA graph launch.  With a fused step kernel the
 * graph is that kernel alone: it reads the packed input straight from ctx->input_pinned
 * and stores the summary straight into ctx->summary_pinned over PCIe (the pinned buffers
 * must be device-accessible, as cudaHostAlloc memory is), followed by a completion flag
 * the host spins on -- no copy nodes, no driver synchronisation on the critical path.
 * Without a fused kernel: H2D copy node, the per-stage kernels, D2H copy node.
 * bh_host_graph_create captures it once per (ctx, learning); bh_step_host_graph packs the
 * input, launches and waits. */
int bh_host_graph_create(const bh_ctx* ctx, int learning, void* stream, void** graph_exec_out);
int bh_step_host_graph(const bh_ctx* ctx, void* graph_exec, const uint8_t* input_bool_host,
                       int32_t* summary_host, void* stream);

/* Copy the summary of the last completed step to the host and synchronise (used
 * after bh_tm_step / bh_step when the caller did not go through bh_step_host). */
int bh_summary(const bh_ctx* ctx, int32_t* summary_host, void* stream);

/* ---- CUDA graphs over bh_step_ring -------------------------------------------------- */
int bh_graph_create(const bh_ctx* ctx, int steps_per_graph, int learning, void* stream, void** graph_exec_out);
int bh_graph_launch(void* graph_exec, void* stream);
int bh_graph_destroy(void* graph_exec);
/* Independent streams: ONE graph that advances n independent networks
 * (ctxs[i], each with its own arena and device input ring, fused_mode != 0) by
 * steps_per_graph timesteps, with no dependency between them, so their kernels run
 * side by side on the SMs a single small network leaves idle.  Launch with
 * bh_graph_launch, free with bh_graph_destroy. */
int bh_batch_graph_create(const bh_ctx* const* ctxs, int n, int steps_per_graph, int learning, void* stream,
                          void** graph_exec_out);
/* One bh_step with a CUDA event recorded on `stream` after every kernel launch.
 * Returns the number of launches n (>= 0) and fills ms_out[0..n) / names_out[0..n)
 * (static strings); synchronises the stream. */
int bh_profile_step(const bh_ctx* ctx, const uint32_t* input_words_dev, int learning, void* stream,
                    float* ms_out, const char** names_out, int max_out);
/* number of kernel launches one bh_step issues (for bench accounting) */
int bh_step_launches(const bh_ctx* ctx, int learning);

/* ---- randomness ------------------------------------------------------------------------ */
/* Adopt the MT19937 state the caller wrote to ctx->mt_key / sc[BH_SC_MT_POS]
 * (np.random.get_state()): the device stream continues from it. */
int bh_rng_import(const bh_ctx* ctx, void* stream);
/* Write the state at the device's stream cursor to ctx->mt_key / sc[BH_SC_MT_POS]. */
int bh_rng_export(const bh_ctx* ctx, void* stream);
/* Take the next `count` float64 uniforms of the stream (np.random.random_sample)
 * into dst_dev; 2 * count <= rng_step_words. */
int bh_rng_fill(const bh_ctx* ctx, double* dst_dev, int64_t count, void* stream);

/* ---- test hooks ------------------------------------------------------------------------- */
/* y[i] = NumPy's float32 SIMD exp(x[i]) (the sequence used by bh_boost). */
int bh_test_np_expf(const float* x_dev, float* y_dev, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BITHTM_B200_H */
