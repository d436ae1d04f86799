// Device-side MT19937 producing np.random's legacy float64 stream.
//
// The reference draws every random number from the global legacy np.random
// stream (networks.py:87, projections.py:120,235).  How many are drawn per step
// (L*(W+1), M) is decided on the device, so the generator lives on the device:
// the host uploads np.random.get_state() (624 words + position) and can read it
// back at any time, which keeps the caller's np.random in lock-step.
//
// One CTA regenerates the 624-word state in three dependency waves
// (227 + 227 + 170 words: x[n] = x[n-227] ^ twist(x[n-624], x[n-623])) and emits
// random_sample() doubles: (a >> 5) * 2^26 + (b >> 6)) / 2^53 from two tempered
// words.
#pragma once

#include "common.cuh"

#define MT_N 624
#define MT_M 397
#define MT_THREADS 1024

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

// x[0..624) holds the current state block; writes the next block to x[624..1248).
// Called by all threads of the CTA (blockDim.x >= 256).
__device__ __forceinline__ void mt_next_block(uint32_t* x) {
  const int t = threadIdx.x;
  if (t < MT_N - MT_M) x[MT_N + t] = x[t + MT_M] ^ mt_twist(x[t], x[t + 1]);
  __syncthreads();
  {
    int kk = (MT_N - MT_M) + t;  // 227 .. 453
    if (t < MT_N - MT_M) x[MT_N + kk] = x[MT_N + kk - (MT_N - MT_M)] ^ mt_twist(x[kk], x[kk + 1]);
  }
  __syncthreads();
  {
    int kk = 2 * (MT_N - MT_M) + t;  // 454 .. 623
    if (kk < MT_N - 1) x[MT_N + kk] = x[MT_N + kk - (MT_N - MT_M)] ^ mt_twist(x[kk], x[kk + 1]);
    if (kk == MT_N - 1) x[MT_N + kk] = x[MT_N + kk - (MT_N - MT_M)] ^ mt_twist(x[kk], x[MT_N]);
  }
  __syncthreads();
}

#define MT_RING 2048  // shared-memory ring of stream words (power of two >= 1078 + 623)

// Emit `count` doubles to out[0..count) continuing the stream at (key, *pos_io).
// Whole CTA (blockDim.x >= 256 threads); x is shared memory of MT_RING words.
//
// The stream is kept as a ring of raw words indexed by absolute position.  After one
// classic block (to have 1078 words of history) it is extended 623 words per barrier
// with the recurrence expanded three times,
//     x[n] = x[n-681] ^ f(n-1078) ^ f(n-851) ^ f(n-624),  f(m) = twist(x[m], x[m+1]),
// whose operands all precede a 623-word wave.  Any 624 consecutive stream words are a
// valid np.random key, so the final state is the last 624 words generated plus the
// offset of the first unconsumed one.
__device__ __noinline__ void mt_fill_block(uint32_t* x, uint32_t* key, int* pos_io, double* out, long long count) {
  const int t = threadIdx.x, NT = blockDim.x;
  const unsigned M = MT_RING - 1;
#pragma unroll 1
  for (int i = t; i < MT_N; i += NT) x[i] = key[i];
  const long long p = *pos_io;           // absolute index of the first unconsumed word
  const long long need_end = p + 2 * count;
  long long G = MT_N;                    // words available: [0, G)
  long long done = 0;                    // doubles emitted so far
  __syncthreads();
#define MT_EMIT()                                                                          \
  do {                                                                                     \
    long long upto_ = G > p ? (G - p) / 2 : 0;                                             \
    if (upto_ > count) upto_ = count;                                                      \
    _Pragma("unroll 1") for (long long q = done + t; q < upto_; q += NT) {                 \
      const unsigned w_ = (unsigned)(p + 2 * q);                                           \
      const uint32_t a_ = mt_temper(x[w_ & M]) >> 5, b_ = mt_temper(x[(w_ + 1) & M]) >> 6; \
      out[q] = ((double)a_ * 67108864.0 + (double)b_) * (1.0 / 9007199254740992.0);        \
    }                                                                                      \
    done = upto_ > done ? upto_ : done;                                                    \
  } while (0)
  MT_EMIT();
  if (need_end > G) {  // classic regeneration of one block: words [624, 1248)
    mt_next_block(x);
    G = 2 * MT_N;
    MT_EMIT();
  }
#pragma unroll 1
  while (need_end > G) {  // 623 independent words per barrier
#pragma unroll 1
    for (int i = t; i < MT_N - 1; i += NT) {
      const unsigned n = (unsigned)G + i;
      x[n & M] = x[(n - 681) & M] ^ mt_twist(x[(n - 1078) & M], x[(n - 1077) & M]) ^
                 mt_twist(x[(n - 851) & M], x[(n - 850) & M]) ^ mt_twist(x[(n - 624) & M], x[(n - 623) & M]);
    }
    __syncthreads();
    G += MT_N - 1;
    MT_EMIT();
  }
#undef MT_EMIT
  __syncthreads();
  const long long start = G - MT_N;
#pragma unroll 1
  for (int i = t; i < MT_N; i += NT) key[i] = x[(unsigned)(start + i) & M];
  if (t == 0) *pos_io = (int)(need_end - start);
}
