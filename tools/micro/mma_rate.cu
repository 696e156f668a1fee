// mma_rate.cu -- how long does one tcgen05.mma.kind::tf32 (M = 128, K = 8) take
// as a function of N, operand source (A from shared memory / tensor memory) and
// accumulator dependence (one chain / alternating accumulators)?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a \
//        -I fb-bev_b200/csrc tools/micro/mma_rate.cu -o build/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tc5.cuh"
using namespace fbbev;

struct Res { long long clk; };

// mode: 0 SS one accumulator, 1 SS two accumulators alternating, 2 TS one acc,
// 3 TS two acc, 4 SS four accumulators
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) rate_kernel(Res* out, int n_mma) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t dummy[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 25000; i += 128)
    reinterpret_cast<float*>(smem)[i] = 0.f;
  if (warp == 0) {
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&dummy[i]), 1); fence_mbar_init(); }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 128 * 40 * 4 * 2;
    const uint32_t lbo_b = N * 16u;
    long long t0 = clock64(), t1 = 0;
    if (elect_one()) {
      for (int i = 0; i < n_mma; ++i) {
        const int k = i % 5;
        const uint32_t acc = (MODE == 1 || MODE == 3) ? (i & 1) * 256u : (MODE == 4) ? (i & 3) * 96u : 0u;
        if (false) {}
        const uint64_t db = smem_desc(b0 + 2u * k * lbo_b, lbo_b, 128);
        if (MODE == 2 || MODE == 3) {
          mma_tf32_ts(tmem + acc, tmem + 400u + 8u * k, db, idesc, 1u);
        } else {
          const uint64_t da = smem_desc(a0 + 2u * k * 2048u, 2048u, 128);
          mma_tf32(tmem + acc, da, db, idesc, 1u);
        }
        if (MODE >= 5 && i % 15 == 14) {
          tc_commit(smem_u32(&dummy[0]));
          if (MODE >= 6) tc_commit(smem_u32(&dummy[1]));
          if (MODE == 7) fence_proxy_async();
          if (MODE == 8) {   // wait for the stage's own commit (full drain)
            mbar_wait(smem_u32(&dummy[0]), (i / 15) & 1);
          }
        }
      }
      tc_commit(smem_u32(&bar));
    }
    __syncwarp();
    t1 = clock64();
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0].clk = t1 - t0; out[1].clk = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

template <int N, int MODE>
void run(const char* name) {
  Res* d;
  cudaMalloc(&d, 2 * sizeof(Res));
  const int smem = 25000 * 4;
  cudaFuncSetAttribute(rate_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int n_mma : {60, 600}) {
    Res h[2];
    for (int rep = 0; rep < 2; ++rep) {
      rate_kernel<N, MODE><<<148, 128, smem>>>(d, n_mma);
      cudaError_t e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-28s N=%3d n=%4d  issue %7.1f clk/mma   complete %7.1f clk/mma\n", name, N, n_mma,
           (double)h[0].clk / n_mma, (double)h[1].clk / n_mma);
  }
  cudaFree(d);
}

int main() {
  run<80, 0>("SS one accumulator");
  run<80, 1>("SS two accumulators");
  run<80, 4>("SS four accumulators");
  run<80, 2>("TS one accumulator");
  run<80, 3>("TS two accumulators");
  run<80, 5>("SS commit/15");
  run<80, 6>("SS 2 commits/15");
  run<80, 7>("SS 2 commits+proxy fence/15");
  run<80, 8>("SS commit+drain/15");
  run<160, 0>("SS one accumulator");
  run<160, 2>("TS one accumulator");
  run<240, 0>("SS one accumulator");
  run<96, 0>("SS one accumulator");
  run<128, 0>("SS one accumulator");
  run<64, 0>("SS one accumulator");
  return 0;
}
