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
#define MT_THREADS 256

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

// Emit `count` doubles to out[0..count) continuing the stream at (key, *pos_io).
// Whole CTA (blockDim.x >= 256 threads); x is shared memory of 2*MT_N words.
__device__ __noinline__ void mt_fill_block(uint32_t* x, uint32_t* key, int* pos_io, double* out, long long count) {
  const int t = threadIdx.x, NT = blockDim.x;
  #pragma unroll 1
  for (int i = t; i < MT_N; i += NT) x[i] = key[i];
  int p = *pos_io;
  __syncthreads();
  long long done = 0;
  while (done < count) {
    long long need_words = 2 * (count - done);
    bool gen = (long long)p + need_words > MT_N;
    if (gen) mt_next_block(x);
    int avail = (gen ? 2 * MT_N : MT_N) - p;
    long long pairs = avail / 2;
    if (pairs > count - done) pairs = count - done;
    #pragma unroll 1
    for (int q = t; q < pairs; q += NT) {
      uint32_t a = mt_temper(x[p + 2 * q]) >> 5;
      uint32_t b = mt_temper(x[p + 2 * q + 1]) >> 6;
      out[done + q] = ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
    }
    p += 2 * (int)pairs;
    done += pairs;
    __syncthreads();
    if (gen) {  // then p > 624: the new block becomes the current one
      #pragma unroll 1
      for (int i = t; i < MT_N; i += NT) x[i] = x[MT_N + i];  // disjoint halves
      p -= MT_N;
      __syncthreads();
    }
  }
  #pragma unroll 1
  for (int i = t; i < MT_N; i += NT) key[i] = x[i];
  if (t == 0) *pos_io = p;
}
