// Micro-benchmark for the round-2 question "how much of a cfg2 phase is the L2 round trip?" (DESIGN.md
// section 8, next step 1).  Run once at the end of round 1 on B200: barriers only 1433 ns, via global
// memory 2271 ns, via DSMEM 1821 ns per iteration.
//
// One 16-CTA cluster of 1024-thread CTAs (the shape of k_step_fused<1>).  Per iteration every CTA
// produces 128 64-bit keys (2048 in all, cfg2's boosted keys), CTA 0 consumes all of them (2 per
// thread, block-reduced to one value that feeds the next iteration, so nothing can be hoisted):
//   mode 0  barrier only             2 x barrier.cluster per iteration, no data
//   mode 1  through global memory    st.global -> barrier.cluster -> ld.global (what the step kernel does)
//   mode 2  through DSMEM            st.shared::cluster into CTA 0 -> barrier.cluster -> ld.shared
// Prints ns per iteration for each mode; (mode 1 - mode 2) is what one hand-off can save, times ~8
// hand-offs per cfg2 step.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/dsmem_latency tools/experiments/dsmem_latency.cu
//   timeout 30 tools/_build/dsmem_latency
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

namespace cg = cooperative_groups;

constexpr int CTAS = 16, THREADS = 1024, KEYS = 2048, KEYS_PER_CTA = KEYS / CTAS, ITERS = 2000;

__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

__global__ void __launch_bounds__(THREADS, 1) k_handoff(int mode, unsigned long long* gkeys, unsigned long long* result,
                                                       unsigned long long* ns_out) {
  __shared__ unsigned long long s_keys[KEYS];  // used in CTA 0 (mode 2: written by every CTA of the cluster)
  __shared__ unsigned long long s_red[32];
  __shared__ unsigned long long s_seed;
  cg::cluster_group cluster = cg::this_cluster();
  const int b = (int)cluster.block_rank(), t = threadIdx.x;
  unsigned long long* remote = cluster.map_shared_rank(s_keys, 0);  // CTA 0's copy
  unsigned long long seed = 1;
  unsigned long long t0 = 0, t1 = 0;
  cluster.sync();
  if (b == 0 && t == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int it = 0; it < ITERS; ++it) {
    // produce: 128 keys per CTA, depending on the previous iteration's result
    if (t < KEYS_PER_CTA) {
      const int i = b * KEYS_PER_CTA + t;
      const unsigned long long key = seed * 0x9E3779B97F4A7C15ull + (unsigned long long)i;
      if (mode == 1) gkeys[i] = key;
      else if (mode == 2) remote[i] = key;
    }
    cluster_barrier();
    // consume on CTA 0: two keys per thread, block sum
    if (b == 0) {
      unsigned long long v = 0;
      if (mode == 1) v = __ldcg(gkeys + t) + __ldcg(gkeys + t + THREADS);
      else if (mode == 2) v = s_keys[t] + s_keys[t + THREADS];
      else v = seed + t;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((t & 31) == 0) s_red[t >> 5] = v;
      __syncthreads();
      if (t < 32) {
        v = s_red[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (t == 0) {
          s_seed = v;
          if (mode == 1) result[0] = v;  // the other CTAs read the result the way the step kernel publishes one
        }
      }
      __syncthreads();
    }
    // publish the result to every CTA: through global memory (mode 1) or by reading CTA 0's shared memory
    cluster_barrier();
    if (mode == 1) {
      seed = __ldcg(result);
    } else {
      const unsigned long long* rs = cluster.map_shared_rank(&s_seed, 0);
      seed = *rs;
    }
  }
  cluster.sync();
  if (b == 0 && t == 0) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    ns_out[mode] = t1 - t0;
    result[1 + mode] = seed;
  }
}

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      return 2;                                                                        \
    }                                                                                  \
  } while (0)

int main() {
  unsigned long long *gkeys, *result, *ns;
  CK(cudaMalloc(&gkeys, KEYS * 8));
  CK(cudaMalloc(&result, 8 * 8));
  CK(cudaMalloc(&ns, 3 * 8));
  CK(cudaMemset(result, 0, 8 * 8));
  CK(cudaFuncSetAttribute(k_handoff, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CTAS);
  cfg.blockDim = dim3(THREADS);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep)  // first pass warms up
    for (int mode = 0; mode < 3; ++mode) {
      CK(cudaLaunchKernelEx(&cfg, k_handoff, mode, gkeys, result, ns));
      CK(cudaDeviceSynchronize());
    }
  unsigned long long h_ns[3], h_res[8];
  CK(cudaMemcpy(h_ns, ns, sizeof(h_ns), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h_res, result, sizeof(h_res), cudaMemcpyDeviceToHost));
  printf("{\"iters\": %d, \"ns_per_iter\": {\"barriers_only\": %.1f, \"via_global\": %.1f, \"via_dsmem\": %.1f}, "
         "\"results_agree\": %s}\n",
         ITERS, (double)h_ns[0] / ITERS, (double)h_ns[1] / ITERS, (double)h_ns[2] / ITERS,
         h_res[2] == h_res[3] ? "true" : "false");
  return 0;
}
