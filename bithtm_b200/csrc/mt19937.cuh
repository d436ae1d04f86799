// Device-side MT19937 producing np.random's legacy float64 stream.
//
// The reference draws every random number from the global legacy np.random stream
// (networks.py:87, projections.py:120,235).  How many are drawn per step (L*(W+1), M)
// is decided on the device, but the stream itself does not depend on the data, so the
// generator is a PRODUCER that appends raw (untempered) stream words to a ring in
// global memory, and the draws are CONSUMERS that take a range of it:
//
//   rng_ring[a & (rng_ring_words-1)] = stream word with absolute index a
//   rng64[R_PRODUCED] = words [.., PRODUCED) are in the ring
//   rng64[R_CURSOR]   = first word no draw has taken yet
//
// Absolute index 0 is key[0] of the last state the host imported (bh_rng_import).  A
// random_sample() double is built from two consecutive tempered words.  Any 624
// consecutive stream words plus an offset are a valid np.random state, which is how
// the state at the cursor is exported back (ph_rng_export / the step summary).
//
// Producers.
//  * serial (one CTA): after two classic 227-word waves (x[n] = x[n-227] ^
//    twist(x[n-624], x[n-623])) the recurrence expanded three times,
//        x[n] = x[n-681] ^ f(n-1078) ^ f(n-851) ^ f(n-624),  f(m) = twist(x[m], x[m+1]),
//    gives 623 independent words per barrier.
//  * parallel (any number of CTAs): chunk p of RNG_CHUNK words starts at
//    PLAN_BASE + p * RNG_CHUNK.  Its first 624 words are obtained from the RNG_WINDOW
//    words before PLAN_BASE with the jump polynomial g_p = t^(RNG_WINDOW + p*RNG_CHUNK)
//    mod phi(t) (bithtm_b200/_mtjump.py): x[m + D] = XOR_{i: g[i]=1} x[m + i]; the rest of
//    the chunk follows with the serial recurrence inside the CTA.
//    Once the ring holds enough history the jump is replaced by the sparse recurrence of
//    phi(t)^(2^s) = sum over phi's 135 exponents e of t^(e << s):
//        x[n] = XOR_{e in PHI, e < 19937} x[n - (19937 << s) + (e << s)],
//    whose nearest operand lies 623 << s words back, i.e. before PLAN_BASE for a large
//    enough s: 134 coalesced loads per word instead of ~10^4 shared-memory XORs.
#pragma once

#include "common.cuh"

#define MT_N 624
#define MT_M 397
#define MT_THREADS 1024
#define MT_RING 2048          // shared-memory ring of stream words (power of two >= 1078 + 623)
#define RNG_WINDOW 20560      // 19937 + 623 words determine any 624 later consecutive words
#define RNG_CHUNK 24920       // 40 * 623
#define RNG_WIN_PAD 20608     // window words staged in shared memory (32*623 + 4*155 + 36, rounded)
#define RNG_PAR_MIN (3 * RNG_CHUNK)  // deficits below this are produced serially
#define RNG_CHUNK_SMEM ((RNG_WIN_PAD + MT_RING + MT_N) * 4)

// exponents below 19937 of the characteristic polynomial of MT19937 (bithtm_b200/_mtjump.py: PHI)
#define RNG_PHI_LOW 134
__constant__ int c_phi_low[RNG_PHI_LOW] = {
    0, 1189, 1416, 1585, 1643, 1870, 2493, 2773, 3000, 3227, 3454, 3681, 3908, 4135, 4362, 4753, 5661, 6337, 6569,
    7129, 7477, 7525, 7583, 7752, 7979, 8206, 9505, 9901, 9969, 10128, 10693, 10761, 10920, 11089, 11147, 11157,
    11215, 11321, 11374, 11384, 11485, 11611, 11712, 11717, 11838, 11881, 11944, 11997, 12277, 12335, 12393, 12504,
    12509, 12620, 12673, 12731, 12736, 12789, 12905, 12958, 12963, 13137, 13185, 13190, 13243, 13301, 13412, 13528,
    13533, 13639, 13697, 13760, 13813, 13866, 14093, 14151, 14209, 14320, 14325, 14436, 14547, 14552, 14605, 14721,
    14774, 14779, 14953, 15001, 15006, 15059, 15117, 15228, 15344, 15349, 15455, 15513, 15576, 15629, 15682, 15909,
    15967, 16025, 16136, 16141, 16252, 16363, 16368, 16421, 16537, 16590, 16595, 16817, 16822, 16875, 16933, 17044,
    17160, 17271, 17329, 17445, 17498, 17725, 17783, 17841, 17952, 18068, 18179, 18237, 18406, 18633, 18691, 18860,
    19087, 19314};

// rng64 slots
enum {
  R_PRODUCED = 0,
  R_CURSOR,
  R_OFF1,         // absolute word index of draw #1 (rand(k, c)) of the current step
  R_OFF2,         // draw #2 (rand(L, W+1))
  R_OFF3,         // draw #3 (rand(M))
  R_N2,           // doubles actually drawn for #2 / #3 (after capacity clamping)
  R_N3,
  R_PLAN_BASE,    // parallel production plan: chunks [0, PLAN_CHUNKS) start at PLAN_BASE
  R_PLAN_CHUNKS,
  R_STEP_BASE,    // cursor at the first draw of the current step
  R_READY3,       // published with draw #2: doubles draw #3 can take from words already produced
  R_EST,          // stream words a step is expected to draw (previous step + margin): speculation target
  R_REGION_LO,    // words [REGION_LO, PRODUCED) are in the ring without a gap (a lazy step leaves one below)
  R_DENSE_LO,     // no word at or above DENSE_LO and below PRODUCED was skipped (sparse chunk starts need this)
  R_LAZY,         // draw #2 of the current step is LAZY: its rand(L, W+1) matrix is not materialised
  R_JUMP_BASE,    // lazy step: window base of all its jumps (= R_OFF2)
  R_TAIL,         // lazy step: first word of the region produced after the skipped matrix (= R_OFF3)
  R_TAIL_CHUNKS,  //            its chunks, produced by jumps
  R_TAIL_WORDS,   //            words per tail chunk (a multiple of 623)
  R_GROW_HIST,    // growing rows of the last two steps (low / high 32 bits): lazy <-> dense policy
  R_COUNT = 32
};

// Lazy draw #2 (large networks).  rand(L, W+1) (projections.py:120) is quadratic in the number of active
// columns, and in steady state almost no learning segment grows, i.e. almost none of its rows is ever read
// (projections.py:114-115: n_add == 0).  A lazy step therefore only ADVANCES the stream cursor over the
// matrix and produces (a) the rows of the segments that do grow and (b) the words after the matrix (rand(M),
// the next step's rand(k, c), the next window) by JUMPS: with D = a + b * skip_gran, the 624 words at
// base + D are  XOR_{i : g_b[i]} x[base + a + i + j],  g_b = t^(b * skip_gran) mod phi from the table
// ctx.mt_skip, applied to the window of 20560 + skip_gran produced words that follows the cursor.  Every
// production job (a row, a chunk of the matrix, a chunk of the tail) is one such jump, computed by the whole
// grid in units of a few polynomial words, followed by the serial recurrence on one CTA.
#define RNG_JOB_STRIDE 640     // words per job slot of ctx.rng_jump
#define RNG_UNIT_MAX 48        // most polynomial words per work unit (6 thread groups x 8)
#define RNG_UNIT_WIN(uw) ((uw) * 32 + MT_N + 36)  // window words a unit of uw polynomial words reads
#define RNG_TAIL_CHUNK 7476    // 12 * 623: smallest tail chunk
#define RNG_LAZY_CHUNK RNG_WINDOW  // chunk of the matrix when a lazy step must produce all of it (a chunk that starts
                                   // inside the jump window must also end inside it)
#define RNG_MAX_TAIL_CHUNKS 64
#define RNG_ROW_SLOT0 RNG_MAX_TAIL_CHUNKS  // job slots: tail chunks first, then the rows / chunks of the matrix
#define RNG_UNIT_WIN_PAD 2208     // RNG_UNIT_WIN(RNG_UNIT_MAX) rounded up to 16 bytes
#define RNG_LAZY_SMEM_WORDS (RNG_UNIT_WIN_PAD + MT_N + 16)  // dynamic shared memory of the lazy phases: unit window +
                                                            // partial output, or MT_RING

__device__ __forceinline__ uint32_t mt_twist(uint32_t u, uint32_t v) {
  uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
  return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

__device__ __forceinline__ uint32_t rng_word(const bh_ctx& c, long long a) {
  return c.rng_ring[(unsigned long long)a & (unsigned long long)(c.rng_ring_words - 1)];
}

// random_sample(): the double made of stream words a and a + 1
__device__ __forceinline__ double rng_uniform(const bh_ctx& c, long long a) {
  const uint32_t hi = mt_temper(rng_word(c, a)) >> 5, lo = mt_temper(rng_word(c, a + 1)) >> 6;
  return ((double)hi * 67108864.0 + (double)lo) * (1.0 / 9007199254740992.0);
}

// Extend the stream inside one CTA.  x = shared ring holding words [lo, G) at x[a & 2047]
// (G - lo >= 624); generates up to at least `target` in whole waves and returns the new
// G.  Words with absolute index in [store_lo, store_hi) are also written to the global
// ring.  All threads call; ends with a barrier.
__device__ __noinline__ long long mt_generate(const bh_ctx& c, uint32_t* x, long long lo, long long G, long long target,
                                              long long store_lo, long long store_hi) {
  const int t = threadIdx.x, NT = blockDim.x;
  const unsigned M = MT_RING - 1;
  const unsigned gm = (unsigned)(c.rng_ring_words - 1);  // ring_words <= 2^31 words (checked by the host layer)
  uint32_t* ring = c.rng_ring;
  // narrow waves (the plain recurrence, 227 words) until 1078 generated words of history exist
#pragma unroll 1
  while (G < target && !((G - lo >= 1078) && (G >= 1078 + MT_N))) {
#pragma unroll 1
    for (int i = t; i < MT_N - MT_M; i += NT) {
      const long long a = G + i;
      const unsigned n = (unsigned)a;
      const uint32_t v = x[(n - 227) & M] ^ mt_twist(x[(n - 624) & M], x[(n - 623) & M]);
      x[n & M] = v;
      if (a >= store_lo && a < store_hi) ring[n & gm] = v;
    }
    __syncthreads();
    G += MT_N - MT_M;
  }
  // wide waves: 623 independent words per barrier; all index arithmetic in 32 bits
  if (G < target) {
    const long long waves = (target - G + (MT_N - 2)) / (MT_N - 1);
    // store window relative to G, clamped to 32 bits
    const long long rel_lo = store_lo - G, rel_hi = store_hi - G;
    const unsigned s_lo = rel_lo < 0 ? 0u : (rel_lo > 0x7fffffffLL ? 0x7fffffffu : (unsigned)rel_lo);
    const unsigned s_hi = rel_hi < 0 ? 0u : (rel_hi > 0x7fffffffLL ? 0x7fffffffu : (unsigned)rel_hi);
    const unsigned g0 = (unsigned)G;
    unsigned off = (unsigned)t;  // offset of this thread's word from G
    if (NT >= MT_N - 1) {
#pragma unroll 1
      for (long long w = 0; w < waves; ++w) {
        if (t < MT_N - 1) {
          const unsigned n = g0 + off;
          const uint32_t v = x[(n - 681) & M] ^ mt_twist(x[(n - 1078) & M], x[(n - 1077) & M]) ^
                             mt_twist(x[(n - 851) & M], x[(n - 850) & M]) ^ mt_twist(x[(n - 624) & M], x[(n - 623) & M]);
          x[n & M] = v;
          if (off >= s_lo && off < s_hi) ring[n & gm] = v;
        }
        off += MT_N - 1;
        __syncthreads();
      }
    } else {
#pragma unroll 1
      for (long long w = 0; w < waves; ++w) {
#pragma unroll 1
        for (int i = t; i < MT_N - 1; i += NT) {
          const unsigned o2 = (unsigned)(w * (MT_N - 1)) + i, n = g0 + o2;
          const uint32_t v = x[(n - 681) & M] ^ mt_twist(x[(n - 1078) & M], x[(n - 1077) & M]) ^
                             mt_twist(x[(n - 851) & M], x[(n - 850) & M]) ^ mt_twist(x[(n - 624) & M], x[(n - 623) & M]);
          x[n & M] = v;
          if (o2 >= s_lo && o2 < s_hi) ring[n & gm] = v;
        }
        __syncthreads();
      }
    }
    G += waves * (MT_N - 1);
  }
  return G;
}

// Serial producer (one CTA): make sure words [.., target) are in the ring.
__device__ __noinline__ void rng_produce_serial(const bh_ctx& c, uint32_t* x, long long target) {
  const int t = threadIdx.x, NT = blockDim.x;
  __syncthreads();
  const long long G0 = c.rng64[R_PRODUCED];
  if (G0 >= target) return;
  const long long avail = G0 - c.rng64[R_REGION_LO];  // contiguous history (>= 624 words)
  const long long hist = avail < 1078 + MT_N ? avail : 1078 + MT_N;
  const long long lo = G0 - hist;
#pragma unroll 1
  for (long long a = lo + t; a < G0; a += NT) x[(unsigned)a & (MT_RING - 1)] = rng_word(c, a);
  __syncthreads();
  const long long G = mt_generate(c, x, lo, G0, target, G0, 0x7fffffffffffffffLL);
  if (t == 0) c.rng64[R_PRODUCED] = G;
  __syncthreads();
}

// Parallel producer: chunk p of the current plan (whole CTA, blockDim.x >= 960).
// smem: RNG_CHUNK_SMEM bytes.
__device__ __noinline__ void rng_chunk(const bh_ctx& c, uint32_t* smem, int p) {
  uint32_t* s_win = smem;                    // [RNG_WIN_PAD]
  uint32_t* x = smem + RNG_WIN_PAD;          // [MT_RING]
  uint32_t* s_out = x + MT_RING;             // [MT_N]
  const int t = threadIdx.x, NT = blockDim.x;
  const long long base = c.rng64[R_PLAN_BASE];
  const long long cb = base + (long long)p * RNG_CHUNK;
  // sparse start: smallest s whose nearest operand (623 << s words back) precedes PLAN_BASE;
  // usable when the farthest one (19937 << s back) is still in the ring and was generated
  int sh = 0;
  while ((623LL << sh) <= (long long)p * RNG_CHUNK + MT_N) ++sh;
  const long long depth = 19937LL << sh;
#ifdef BH_TOPK_STAMPS
  if (blockIdx.x == 0 && t == 0) {
    unsigned long long t_;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
    reinterpret_cast<unsigned long long*>(c.blk + 7 * BH_BLK_STRIDE)[59] = t_;
  }
#endif
  const long long plan_end = base + c.rng64[R_PLAN_CHUNKS] * RNG_CHUNK;  // slots below plan_end - ring are reused
  long long lowest = plan_end - c.rng_ring_words > 1 ? plan_end - c.rng_ring_words : 1;
  if (c.rng64[R_DENSE_LO] > lowest) lowest = c.rng64[R_DENSE_LO];  // words a lazy step skipped never existed
  const bool sparse = cb - depth >= lowest;
  if (sparse) {
#pragma unroll 1
    for (int j = t; j < MT_N; j += NT) {
      const long long a0 = cb + j - depth;
      uint32_t v = 0u;
#pragma unroll 1
      for (int e0 = 0; e0 < RNG_PHI_LOW; e0 += 32) {  // 32 independent loads in flight per thread
        uint32_t w[32];
#pragma unroll
        for (int u = 0; u < 32; ++u)
          w[u] = e0 + u < RNG_PHI_LOW ? rng_word(c, a0 + ((long long)c_phi_low[e0 + u] << sh)) : 0u;
#pragma unroll
        for (int u = 0; u < 32; ++u) v ^= w[u];
      }
      s_out[j] = v;
    }
  } else {
#pragma unroll 1
    for (int i = t; i < RNG_WIN_PAD; i += NT) s_win[i] = i < RNG_WINDOW ? rng_word(c, base - RNG_WINDOW + i) : 0u;
#pragma unroll 1
    for (int i = t; i < MT_N; i += NT) s_out[i] = 0u;
  }
  __syncthreads();
#ifdef BH_TOPK_STAMPS
  if (blockIdx.x == 0 && t == 0) {
    unsigned long long t_;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
    reinterpret_cast<unsigned long long*>(c.blk + 7 * BH_BLK_STRIDE)[60] = t_;
  }
#endif
  // out[j] = XOR_{i : g[i]} win[i + j].  6 groups of 156 threads split the 624 words of
  // g; thread q of a group keeps outputs 4q..4q+3 and slides a register window over win.
  if (!sparse && t < 936) {
    const int grp = t / 156, q = t - grp * 156;
    const uint32_t* gp = c.mt_jump + (long long)p * MT_N;
    uint32_t a0 = 0u, a1 = 0u, a2 = 0u, a3 = 0u;
#pragma unroll 1
    for (int wi = grp * 104; wi < grp * 104 + 104; ++wi) {
      const uint32_t gw = __ldg(gp + wi);
      if (gw == 0u) continue;
      const uint4* wp = reinterpret_cast<const uint4*>(s_win + 32 * wi + 4 * q);
      uint32_t r[36];
#pragma unroll
      for (int m = 0; m < 9; ++m) {
        const uint4 v = wp[m];
        r[4 * m] = v.x;
        r[4 * m + 1] = v.y;
        r[4 * m + 2] = v.z;
        r[4 * m + 3] = v.w;
      }
#pragma unroll
      for (int b = 0; b < 32; ++b) {
        if (gw & (1u << b)) {
          a0 ^= r[b];
          a1 ^= r[b + 1];
          a2 ^= r[b + 2];
          a3 ^= r[b + 3];
        }
      }
    }
    atomicXor(&s_out[4 * q], a0);
    atomicXor(&s_out[4 * q + 1], a1);
    atomicXor(&s_out[4 * q + 2], a2);
    atomicXor(&s_out[4 * q + 3], a3);
  }
  __syncthreads();
  const unsigned long long gm = (unsigned long long)(c.rng_ring_words - 1);
#pragma unroll 1
  for (int j = t; j < MT_N; j += NT) {
    const uint32_t v = s_out[j];
    x[(unsigned)(cb + j) & (MT_RING - 1)] = v;
    c.rng_ring[(unsigned long long)(cb + j) & gm] = v;
  }
  __syncthreads();
  mt_generate(c, x, cb, cb + MT_N, cb + RNG_CHUNK, cb + MT_N, cb + RNG_CHUNK);
#ifdef BH_TOPK_STAMPS
  if (blockIdx.x == 0 && t == 0) {
    unsigned long long t_;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
    reinterpret_cast<unsigned long long*>(c.blk + 7 * BH_BLK_STRIDE)[61] = t_;
  }
#endif
}

// All CTAs: run the chunks of the pending plan (b of nb).  The plan is committed
// (PRODUCED advanced) by the next rng_commit_plan on the drawing CTA.
__device__ __forceinline__ void ph_rng_chunks(const bh_ctx& c, uint32_t* smem, int b, int nb) {
  const int n = (int)c.rng64[R_PLAN_CHUNKS];
#pragma unroll 1
  for (int p = b; p < n; p += nb) {
    rng_chunk(c, smem, p);
    __syncthreads();
  }
}

// Single thread of the drawing CTA: fold a finished parallel plan into PRODUCED.
__device__ __forceinline__ void rng_commit_plan(const bh_ctx& c) {
  const long long n = c.rng64[R_PLAN_CHUNKS];
  if (n > 0) {
    c.rng64[R_PRODUCED] = c.rng64[R_PLAN_BASE] + n * RNG_CHUNK;
    c.rng64[R_PLAN_CHUNKS] = 0;
  }
}

// Largest number of doubles a draw starting at `cursor` may take: a step (from step_base
// on) may use rng_step_words words, which the caller sized the ring and the jump table for.
__device__ __forceinline__ long long rng_room(const bh_ctx& c, long long step_base, long long cursor) {
  const long long cap = c.rng_step_words - (cursor - step_base);
  return cap > 0 ? cap / 2 : 0;
}

// One CTA.  Take `count` doubles at the cursor (thread 0 publishes *off_slot / *n_slot) and
// make sure they -- plus `lookahead` more words -- are produced, serially when the
// deficit is small, else by planning chunks for ph_rng_chunks (which must run next).
// `x` = MT_RING words of shared memory.  Returns through shared state only.
// Thread 0 of the drawing CTA, draw #2 of `count` doubles = rows of row_doubles (= W + 1) at cursor `cur`:
// decide whether the step is lazy and, if so, lay out its tail.  Returns the number of tail chunks (0: dense).
__device__ __forceinline__ int rng_plan_lazy(const bh_ctx& c, long long cur, long long count, int row_doubles,
                                             bool allow_lazy) {
  long long* r = c.rng64;
  const bool was_lazy = r[R_LAZY] != 0;
  // growing rows of the step before this one (g_prev) and of the one before that (g_before)
  const long long hist = r[R_GROW_HIST];
  const long long g_prev = hist < 0 ? 0x7fffffffLL : (long long)c.sc[BH_SC_NGROW];
  const long long g_before = hist < 0 ? 0x7fffffffLL : (hist & 0xffffffffLL);
  r[R_GROW_HIST] = (g_before << 32) | g_prev;
  c.sc[BH_SC_NGROW] = 0;
  r[R_LAZY] = 0;
  if (c.skip_polys <= 0 || !allow_lazy) return 0;  // (only the cooperative step kernels have the lazy phases)
  const long long D = 2 * count;
  if (D < c.skip_min || cur < 1) return 0;
  const long long n_chunks = (D + RNG_LAZY_CHUNK - 1) / RNG_LAZY_CHUNK;
  // a lazy step that finds many growing rows pays a jump per chunk of the matrix; dense production pays 134
  // loads per chunk start but needs unbroken history: go lazy once two steps in a row grew few rows, return
  // to dense when two steps in a row grew most of them
  if (c.lazy_policy == 0 &&
      (was_lazy ? (g_prev > n_chunks && g_before > n_chunks) : (g_prev > n_chunks / 4 || g_before > n_chunks / 4)))
    return 0;
  // tail: rand(M) + the next step's rand(k, c) + the next step's jump window (+ slack for larger M / W)
  const long long kc2 = 2LL * c.active_columns * c.cell_dim;
  const long long need = 2 * (2LL * c.sc[BH_SC_M] + 4LL * c.active_columns) + kc2 + c.skip_gran + RNG_WINDOW +
                         4LL * row_doubles + 4 * MT_N;
  // few chunks: every chunk start costs a jump (~80 us of one SM), its words only the serial recurrence
  long long q = c.tail_chunks > 0 ? c.tail_chunks : 4;
  if (q > RNG_MAX_TAIL_CHUNKS) q = RNG_MAX_TAIL_CHUNKS;
  long long words = ((need + q - 1) / q + (MT_N - 2)) / (MT_N - 1) * (MT_N - 1);
  if (words < RNG_TAIL_CHUNK) words = RNG_TAIL_CHUNK;
  q = (need + words - 1) / words;
  if ((D + q * words) / c.skip_gran + 1 > c.skip_polys) return 0;  // beyond the jump table
  r[R_LAZY] = 1;
  r[R_JUMP_BASE] = cur;
  r[R_TAIL] = cur + D;
  r[R_TAIL_CHUNKS] = q;
  r[R_TAIL_WORDS] = words;
  return (int)q;
}

__device__ __noinline__ void rng_draw(const bh_ctx& c, uint32_t* x, long long count, int off_slot, int n_slot,
                                      bool first_of_step, long long lookahead, bool may_plan,
                                      bool publish_next = false, int row_doubles = 0, bool allow_lazy = false) {
  __shared__ long long s_serial_target;
  if (threadIdx.x == 0) {
    long long* r = c.rng64;
    rng_commit_plan(c);
    const long long cur = r[R_CURSOR];
    if (first_of_step) r[R_STEP_BASE] = cur;
    const long long room = rng_room(c, r[R_STEP_BASE], cur);
    if (count > room) {
      count = room;
      atomicOr(&c.sc[BH_SC_STATUS], BH_ST_RAND_OVERFLOW);
    }
    r[off_slot] = cur;
    if (n_slot >= 0) r[n_slot] = count;
    const long long end = cur + 2 * count;
    r[R_CURSOR] = end;
    long long la = lookahead;
    if (la > c.rng_step_words / 2) la = c.rng_step_words / 2;
    long long target = end + la + MT_N;  // keep 624 words past the cursor for state export
    const int lazy_chunks = row_doubles > 0 ? rng_plan_lazy(c, cur, count, row_doubles, allow_lazy) : 0;
    if (lazy_chunks > 0)  // only the jump window is needed now; matrix rows and tail come from jumps
      target = cur + c.skip_gran + RNG_WINDOW + 2LL * row_doubles + MT_N;
    long long produced = r[R_PRODUCED];
    long long serial_target = target;
    if (lazy_chunks == 0 && may_plan && c.jump_polys > 0 && target - produced > RNG_PAR_MIN) {
      // the window must consist of generated words (absolute index >= 1)
      const long long need = produced < 1 + RNG_WINDOW ? 1 + RNG_WINDOW : produced;
      long long chunks = (target - need + RNG_CHUNK - 1) / RNG_CHUNK;
      if (chunks > c.jump_polys) chunks = c.jump_polys;  // the rest is produced serially by later draws
      if (chunks > 0) {
        serial_target = need;
        r[R_PLAN_BASE] = need;  // == PRODUCED after the serial part below
        r[R_PLAN_CHUNKS] = chunks;
      }
    }
    s_serial_target = serial_target;
  }
  __syncthreads();
  rng_produce_serial(c, x, s_serial_target);
  if (threadIdx.x == 0 && row_doubles > 0 && c.rng64[R_LAZY]) {
    // lazy step: from here on the produced region is the tail (generated before anything reads it)
    long long* r = c.rng64;
    const long long tail = r[R_TAIL], len = r[R_TAIL_CHUNKS] * r[R_TAIL_WORDS];
    r[R_PRODUCED] = tail + len;
    r[R_REGION_LO] = tail;
    r[R_DENSE_LO] = tail;
    long long ready = (len - MT_N) / 2;
    const long long room = rng_room(c, r[R_STEP_BASE], tail);
    r[R_OFF3] = tail;
    r[R_READY3] = ready < room ? ready : room;
  } else if (threadIdx.x == 0) {
    long long* r = c.rng64;
    if (r[R_PLAN_CHUNKS] > 0) r[R_PLAN_BASE] = r[R_PRODUCED];
    if (publish_next) {
      // draw #3 follows directly in the stream: where it starts, and how many doubles of it are covered
      // by words that exist once this draw (and its chunks) are done -- lets the fused kernel skip the
      // draw-#3 phase and its barrier (fused.cuh)
      const long long end = r[R_CURSOR];
      const long long produced = r[R_PLAN_CHUNKS] > 0 ? r[R_PLAN_BASE] + r[R_PLAN_CHUNKS] * RNG_CHUNK : r[R_PRODUCED];
      long long ready = (produced - end - MT_N) / 2;
      const long long room = rng_room(c, r[R_STEP_BASE], end);
      if (ready > room) ready = room;
      r[R_OFF3] = end;
      r[R_READY3] = ready > 0 ? ready : 0;
    }
  }
  __syncthreads();
}

// Thread 0 of the drawing CTA, after the last draw of a step: what the next step is expected to need.
__device__ __forceinline__ void rng_finish_step(const bh_ctx& c) {
  long long* r = c.rng64;
  long long used = r[R_CURSOR] - r[R_STEP_BASE];
  if (r[R_LAZY]) used -= 2 * r[R_N2];  // a lazy step never produces its matrix
  long long est = used + used / 4 + 2 * (long long)c.active_columns * c.cell_dim + 2 * MT_N;
  if (est > c.rng_step_words / 2) est = c.rng_step_words / 2;
  r[R_EST] = est;
}

// Draw #3 when R_READY3 covers it: bookkeeping only (thread 0 of the drawing CTA).
__device__ __forceinline__ void rng_draw3_commit(const bh_ctx& c, long long count) {
  long long* r = c.rng64;
  r[R_N3] = count;
  r[R_CURSOR] = r[R_OFF3] + 2 * count;
  rng_finish_step(c);
}

// One CTA, while it has nothing else to do: produce the words the current step is expected to
// draw (serial producer; with many-CTA production the draws plan their own chunks).
// `ahead`: count from the cursor (the rest of this step and the next one) instead of the step's start.
__device__ __noinline__ void ph_rng_speculate(const bh_ctx& c, int divisor = 1, bool ahead = false) {
  __shared__ uint32_t x[MT_RING];
  if (c.jump_polys > 0) return;
  rng_produce_serial(c, x, c.rng64[ahead ? R_CURSOR : R_STEP_BASE] + c.rng64[R_EST] / divisor);
}

// ------------------------------------------------------------------------------------
// lazy steps: production jobs by jumps
// ------------------------------------------------------------------------------------
// kind 0: tail chunk i;  1: row of rand(L, W+1) named by grow-list entry i;  2: chunk i of the whole matrix
__device__ __forceinline__ void rng_job(const bh_ctx& c, int kind, int i, int row_words, long long& dst, long long& n) {
  const long long* r = c.rng64;
  if (kind == 0) {
    dst = r[R_TAIL] + (long long)i * r[R_TAIL_WORDS];
    n = r[R_TAIL_WORDS];
  } else if (kind == 1) {
    dst = r[R_OFF2] + (long long)c.grow_list[3 * i] * row_words;
    n = row_words;
  } else {
    const long long D = 2 * r[R_N2], off = (long long)i * RNG_LAZY_CHUNK;
    dst = r[R_OFF2] + off;
    n = D - off < RNG_LAZY_CHUNK ? D - off : RNG_LAZY_CHUNK;
  }
}

// All CTAs: the jumps of jobs [0, n_jobs) of one kind.  The n_jobs * 624 polynomial words are cut into units of
// uw words (a multiple of 6, chosen so that every CTA gets one unit when that is possible).  Unit (job, w0):
// out[job][j] ^= XOR_{i in [32 w0, 32 (w0 + uw)) : g[i]} x[src + i + j].  blockDim.x >= 960.
// smem: RNG_LAZY_SMEM_WORDS.  A barrier must follow before the slots are read.
__device__ __noinline__ void ph_rng_jumps(const bh_ctx& c, uint32_t* smem, int kind, int n_jobs, int row_words, int slot0,
                                          int b, int nb) {
  uint32_t* s_win = smem;                    // [RNG_UNIT_WIN(uw)], 16-byte aligned
  uint32_t* s_out = smem + RNG_UNIT_WIN_PAD;  // [MT_N]
  const int t = threadIdx.x, NT = blockDim.x;
  const long long base = c.rng64[R_JUMP_BASE];
  // unit size: the multiple of 6 words that minimises rounds x (words per unit + a fixed cost per unit) -- with
  // fewer CTAs than units of the largest size, a smaller unit avoids a last round that only a few CTAs work in
  // (8 jobs on 92 CTAs: 104 units of 48 words = 2 rounds of 48; 168 units of 30 words = 2 rounds of 30)
  int uw = 6;
  {
    long long best = -1;
#pragma unroll 1
    for (int cand = 6; cand <= RNG_UNIT_MAX; cand += 6) {
      const long long units = (long long)n_jobs * ((MT_N + cand - 1) / cand);
      const long long cost = (units + nb - 1) / nb * (cand + 12);
      if (best < 0 || cost < best) {
        best = cost;
        uw = cand;
      }
    }
  }
  const int per_poly = (MT_N + uw - 1) / uw, per_group = uw / 6;
  const long long n_units = (long long)n_jobs * per_poly;
#pragma unroll 1
  for (long long u = b; u < n_units; u += nb) {
    const int job = (int)(u / per_poly), w0 = (int)(u - (long long)job * per_poly) * uw;
    long long dst, n;
    rng_job(c, kind, job, row_words, dst, n);
    const long long D = dst - base;
    const long long poly = D / c.skip_gran - 1;  // row of mt_skip; -1: the job starts inside the produced window
    if (poly < 0) continue;                      // (uniform over the CTA)
    const long long src = base + D % c.skip_gran + 32LL * w0;
    const uint32_t* gp = c.mt_skip + poly * MT_N;
    const int win = RNG_UNIT_WIN(uw);
#pragma unroll 1
    for (int i = t; i < win; i += NT) s_win[i] = rng_word(c, src + i);
#pragma unroll 1
    for (int i = t; i < MT_N; i += NT) s_out[i] = 0u;
    __syncthreads();
    if (t < 936) {
      const int grp = t / 156, q = t - grp * 156;
      uint32_t a0 = 0u, a1 = 0u, a2 = 0u, a3 = 0u;
#pragma unroll 1
      for (int wl = grp * per_group; wl < (grp + 1) * per_group; ++wl) {
        const uint32_t gw = w0 + wl < MT_N ? __ldg(gp + w0 + wl) : 0u;
        if (gw == 0u) continue;
        const uint4* wp = reinterpret_cast<const uint4*>(s_win + 32 * wl + 4 * q);
        uint32_t rr[36];
#pragma unroll
        for (int m = 0; m < 9; ++m) {
          const uint4 v = wp[m];
          rr[4 * m] = v.x;
          rr[4 * m + 1] = v.y;
          rr[4 * m + 2] = v.z;
          rr[4 * m + 3] = v.w;
        }
#pragma unroll
        for (int bb = 0; bb < 32; ++bb) {
          if (gw & (1u << bb)) {
            a0 ^= rr[bb];
            a1 ^= rr[bb + 1];
            a2 ^= rr[bb + 2];
            a3 ^= rr[bb + 3];
          }
        }
      }
      atomicXor(&s_out[4 * q], a0);
      atomicXor(&s_out[4 * q + 1], a1);
      atomicXor(&s_out[4 * q + 2], a2);
      atomicXor(&s_out[4 * q + 3], a3);
    }
    __syncthreads();
    uint32_t* out = c.rng_jump + (long long)(slot0 + job) * RNG_JOB_STRIDE;
#pragma unroll 1
    for (int i = t; i < MT_N; i += NT) {
      const uint32_t v = s_out[i];
      if (v) atomicXor(&out[i], v);
    }
    __syncthreads();
  }
}

// One CTA: job slot -> ring (the 624 jumped words), then the serial recurrence up to the job's length; the slot
// is left zeroed for its next use.  smem: MT_RING words.  Jobs inside the produced window need nothing.
__device__ __noinline__ void rng_job_generate(const bh_ctx& c, uint32_t* x, int kind, int job, int row_words, int slot0) {
  const int t = threadIdx.x, NT = blockDim.x;
  long long dst, n;
  rng_job(c, kind, job, row_words, dst, n);
  if ((dst - c.rng64[R_JUMP_BASE]) / c.skip_gran < 1) return;
  uint32_t* slot = c.rng_jump + (long long)(slot0 + job) * RNG_JOB_STRIDE;
  const unsigned long long gm = (unsigned long long)(c.rng_ring_words - 1);
  __syncthreads();
#pragma unroll 1
  for (int j = t; j < MT_N; j += NT) {
    const uint32_t v = __ldcg(slot + j);
    slot[j] = 0u;
    x[(unsigned)(dst + j) & (MT_RING - 1)] = v;
    c.rng_ring[(unsigned long long)(dst + j) & gm] = v;
  }
  __syncthreads();
  if (n > MT_N) mt_generate(c, x, dst, dst + MT_N, dst + n, dst + MT_N, dst + n);
}

// The lazy part of a step after its stage-1 learning pass (all CTAs of a cooperative kernel; `sync` = its grid
// barrier).  Round 1 (tail jumps) was done alongside stage 1.  Here: the jumps of the growing rows, their
// words, and -- through `grow` -- the growth itself.  Leaves a barrier behind only when rows grew.
template <typename Sync, typename Grow>
__device__ __forceinline__ void ph_rng_lazy_rows(const bh_ctx& c, uint32_t* smem, int b, int nb, Sync sync, Grow grow) {
  const int G = c.sc[BH_SC_NGROW];  // uniform: read after the barrier that ended stage 1
  if (G <= 0) return;
  const int cur = c.sc[BH_SC_STEP] & 1;
  const int row_words = 2 * (c.sc[BH_SC_W0 + (cur ^ 1)] + 1);
  const long long D = 2 * c.rng64[R_N2];
  const int n_chunks = (int)((D + RNG_LAZY_CHUNK - 1) / RNG_LAZY_CHUNK);
  const bool by_rows = G <= n_chunks && G <= c.job_cap - RNG_ROW_SLOT0;
  if (by_rows) {
    ph_rng_jumps(c, smem, 1, G, row_words, RNG_ROW_SLOT0, b, nb);
    sync();
    grow(true);  // each CTA generates the words of its rows, then grows them
  } else {
    // most rows grow: produce the whole matrix in chunks (check_ctx sizes job_cap for the largest step)
    const int nj = n_chunks < c.job_cap - RNG_ROW_SLOT0 ? n_chunks : c.job_cap - RNG_ROW_SLOT0;
    ph_rng_jumps(c, smem, 2, nj, row_words, RNG_ROW_SLOT0, b, nb);
    sync();
#pragma unroll 1
    for (int j = b; j < nj; j += nb) {
      rng_job_generate(c, smem, 2, j, row_words, RNG_ROW_SLOT0);
      __syncthreads();
    }
    sync();
    grow(false);
  }
  sync();
}

// The tail of a lazy step is generated while the other CTAs scan the segments: by the LAST rng_tail_ctas()
// CTAs of the grid (which then take no part in the scan), chunk q by the q-th of them.
__device__ __forceinline__ int rng_tail_ctas(const bh_ctx& c, int nb) {
  const int Q = (int)c.rng64[R_TAIL_CHUNKS];
  return Q < nb - 1 ? Q : (nb - 1 > 0 ? nb - 1 : 0);  // a single CTA generates first and scans afterwards
}
__device__ __forceinline__ void ph_rng_lazy_tail(const bh_ctx& c, uint32_t* smem, int b, int nb) {
  const int Q = (int)c.rng64[R_TAIL_CHUNKS], ngen = rng_tail_ctas(c, nb);
  const int first = ngen > 0 ? nb - ngen : 0, stride = ngen > 0 ? ngen : 1;
  if (b < first) return;
#pragma unroll 1
  for (int q = b - first; q < Q; q += stride) {
    rng_job_generate(c, smem, 0, q, 0, 0);
    __syncthreads();
  }
}

// Host state -> ring (single CTA): key = words [0, 624), cursor = pos.
__device__ __forceinline__ void ph_rng_import(const bh_ctx& c) {
#pragma unroll 1
  for (int i = threadIdx.x; i < MT_N; i += blockDim.x) c.rng_ring[i] = c.mt_key[i];
  if (threadIdx.x == 0) {
    long long* r = c.rng64;
    r[R_PRODUCED] = MT_N;
    r[R_CURSOR] = c.sc[BH_SC_MT_POS];
    r[R_PLAN_CHUNKS] = 0;
    r[R_STEP_BASE] = c.sc[BH_SC_MT_POS];
    r[R_REGION_LO] = 0;
    r[R_DENSE_LO] = 1;
    r[R_LAZY] = 0;
    r[R_TAIL_CHUNKS] = 0;
    r[R_GROW_HIST] = -1;  // unknown: the first steps draw densely
  }
}

// State at the cursor as (624 key words, pos): out[i] = key word i, out[624] = pos.
// Threads gid of gsz cooperate; requires PRODUCED >= CURSOR and PRODUCED >= 624 (always
// true after an import) and no uncommitted plan covering the cursor.
__device__ __forceinline__ void rng_export(const bh_ctx& c, int* out, int gid, int gsz) {
  const long long n = c.rng64[R_PLAN_CHUNKS];
  const long long produced = n > 0 ? c.rng64[R_PLAN_BASE] + n * RNG_CHUNK : c.rng64[R_PRODUCED];
  const long long cur = c.rng64[R_CURSOR];
  const long long start = cur < produced - MT_N ? cur : produced - MT_N;
#pragma unroll 1
  for (int i = gid; i <= MT_N; i += gsz) out[i] = i < MT_N ? (int)rng_word(c, start + i) : (int)(cur - start);
}
