// Microbenchmark: cost of the inter-phase barriers and of a dependent global-load
// chain right after a barrier, on one thread-block cluster.  nvcc -arch=sm_100a.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__device__ __forceinline__ void bar_full() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void bar_light() {
  __syncthreads();
  if (threadIdx.x == 0) __threadfence();
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ unsigned long long gt() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int KIND>
__global__ void __launch_bounds__(1024, 1) k_bar(int iters, unsigned long long* out, int* chain, int chain_len) {
  unsigned long long t0 = gt();
  int acc = 0;
  for (int i = 0; i < iters; ++i) {
    if (KIND == 0) bar_full();
    else if (KIND == 1) bar_light();
    else __syncthreads();
    if (chain_len) {  // dependent chain of global loads by every thread of CTA 0 (same addresses)
      int p = (i * 7) & 1023;
      for (int j = 0; j < chain_len; ++j) p = chain[p];
      acc += p;
    }
  }
  unsigned long long t1 = gt();
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = acc; }
}

template <int KIND>
void run(const char* name, int nb, int threads, int chain_len, unsigned long long* d_out, int* d_chain) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nb); cfg.blockDim = dim3(threads);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = nb; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaFuncSetAttribute(k_bar<KIND>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_bar<KIND>, iters, d_out, d_chain, chain_len);
    if (e != cudaSuccess) { printf("%s launch failed: %s\n", name, cudaGetErrorString(e)); return; }
    cudaDeviceSynchronize();
  }
  unsigned long long h[2];
  cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
  printf("%-28s cluster=%2d threads=%4d chain=%d : %.3f us/iter\n", name, nb, threads, chain_len, h[0] / 1e3 / iters);
}

int main() {
  unsigned long long* d_out; int* d_chain;
  cudaMalloc(&d_out, 64); cudaMalloc(&d_chain, 4096);
  int h[1024];
  for (int i = 0; i < 1024; ++i) h[i] = (i * 37 + 11) & 1023;
  cudaMemcpy(d_chain, h, 4096, cudaMemcpyHostToDevice);
  for (int nb : {1, 8, 16}) {
    for (int th : {256, 1024}) {
      run<2>("syncthreads only", nb, th, 0, d_out, d_chain);
      run<0>("cluster full rel/acq", nb, th, 0, d_out, d_chain);
      run<1>("cluster light (t0 fence)", nb, th, 0, d_out, d_chain);
      run<0>("full + 4 dependent loads", nb, th, 4, d_out, d_chain);
      run<1>("light + 4 dependent loads", nb, th, 4, d_out, d_chain);
      run<2>("syncthreads + 4 dep loads", nb, th, 4, d_out, d_chain);
    }
  }
  return 0;
}
