// Micro-benchmark: how fast can ONE CTA run the MT19937 wide-wave recurrence (623 words per barrier)?
//   v0: as csrc/mt19937.cuh mt_generate (global store of every word inside the wave loop)
//   v1: no global stores at all (upper bound of the smem recurrence itself)
//   v2: words kept in registers, stored to global every 8 waves
//   v3: v2 with 640-thread CTA
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/mt_wave_bench tools/mt_wave_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define MT_N 624
#define MT_RING 2048
__device__ __forceinline__ uint32_t mt_twist(uint32_t u, uint32_t v) {
  uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
  return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
template <int V>
__global__ void k(uint32_t* ring, int waves, unsigned long long* ns) {
  __shared__ uint32_t x[MT_RING];
  const int t = threadIdx.x;
  const unsigned M = MT_RING - 1;
  for (int i = t; i < MT_RING; i += blockDim.x) x[i] = i * 2654435761u;
  __syncthreads();
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  unsigned off = t, g0 = 1702;
  uint32_t keep[8];
  for (int w = 0; w < waves; ++w) {
    if (t < MT_N - 1) {
      const unsigned n = g0 + off;
      const uint32_t v = x[(n - 681) & M] ^ mt_twist(x[(n - 1078) & M], x[(n - 1077) & M]) ^
                         mt_twist(x[(n - 851) & M], x[(n - 850) & M]) ^ mt_twist(x[(n - 624) & M], x[(n - 623) & M]);
      x[n & M] = v;
      if (V == 0) ring[n & 0xfffff] = v;
      if (V >= 2) {
        keep[w & 7] = v;
        if ((w & 7) == 7) {
#pragma unroll
          for (int j = 0; j < 8; ++j) ring[(n - (7 - j) * (MT_N - 1)) & 0xfffff] = keep[j];
        }
      }
    }
    off += MT_N - 1;
    __syncthreads();
  }
  unsigned long long t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  if (t == 0) ns[0] = t1 - t0;
  if (V == 1 && t < MT_N) ring[t] = x[(g0 + off) & M];
}
int main() {
  uint32_t* ring;
  unsigned long long* ns;
  cudaMalloc(&ring, 4 << 20);
  cudaMallocManaged(&ns, 8);
  const int waves = 400;
  for (int rep = 0; rep < 2; ++rep) {
    k<0><<<1, 1024>>>(ring, waves, ns); cudaDeviceSynchronize(); printf("v0 (store per wave, 1024 thr): %.1f ns/wave\n", (double)ns[0] / waves);
    k<1><<<1, 1024>>>(ring, waves, ns); cudaDeviceSynchronize(); printf("v1 (no global stores)        : %.1f ns/wave\n", (double)ns[0] / waves);
    k<2><<<1, 1024>>>(ring, waves, ns); cudaDeviceSynchronize(); printf("v2 (stores every 8 waves)    : %.1f ns/wave\n", (double)ns[0] / waves);
    k<2><<<1, 640>>>(ring, waves, ns); cudaDeviceSynchronize(); printf("v3 (v2, 640 threads)         : %.1f ns/wave\n", (double)ns[0] / waves);
    k<0><<<1, 640>>>(ring, waves, ns); cudaDeviceSynchronize(); printf("v4 (v0, 640 threads)         : %.1f ns/wave\n", (double)ns[0] / waves);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
