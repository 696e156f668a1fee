// Micro-benchmark: how fast can B200 write a (C, ZYX) fp32 volume with
// different store patterns?  (development aid; results in profiles/)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t err_ = (x); if (err_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(err_), __LINE__); exit(1);} } while (0)

__global__ void linear_default(float4* o, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) o[i] = make_float4(0, 0, 0, 0);
}
__global__ void linear_cs(float4* o, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) __stcs(o + i, make_float4(0, 0, 0, 0));
}
// one CTA per tile of T voxels; each warp writes rows (channels) of T floats at stride zyx
template <int T, int POLICY>
__global__ void tile_rows(float* out, int c, int64_t zyx) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  constexpr int LPR = T / 4; constexpr int RPW = 32 / LPR > 0 ? 32 / LPR : 1;
  float* base = out + (int64_t)blockIdx.x * T;
  if (LPR <= 32) {
    const int g = lane % LPR;
    for (int row = warp * RPW + lane / LPR; row < c; row += nw * RPW) {
      float4* p = reinterpret_cast<float4*>(base + (int64_t)row * zyx + 4 * g);
      if (POLICY == 0) *p = make_float4(0, 0, 0, 0); else if (POLICY == 1) __stcs(p, make_float4(0, 0, 0, 0)); else __stcg(p, make_float4(0,0,0,0));
    }
  } else {
    for (int row = warp; row < c; row += nw)
      for (int g = lane; g < LPR; g += 32) {
        float4* p = reinterpret_cast<float4*>(base + (int64_t)row * zyx + 4 * g);
        if (POLICY == 0) *p = make_float4(0, 0, 0, 0); else if (POLICY == 1) __stcs(p, make_float4(0, 0, 0, 0)); else __stcg(p, make_float4(0,0,0,0));
      }
  }
}
// same tile pattern, rows written by TMA bulk copies from a zero row in smem
template <int T>
__global__ void tile_rows_bulk(float* out, int c, int64_t zyx) {
  __shared__ __align__(128) float zrow[T];
  for (int i = threadIdx.x; i < T; i += blockDim.x) zrow[i] = 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  float* base = out + (int64_t)blockIdx.x * T;
  if (threadIdx.x < c) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(zrow);
    for (int row = threadIdx.x; row < c; row += blockDim.x)
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + (int64_t)row * zyx), "r"(s), "r"(T * 4) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}
// persistent version of tile_rows: grid = k*SMs, each CTA loops over tiles
template <int T, int POLICY>
__global__ void tile_rows_persistent(float* out, int c, int64_t zyx, int n_tiles) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  constexpr int LPR = T / 4; constexpr int RPW = 32 / LPR;
  const int g = lane % LPR;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    float* base = out + (int64_t)tile * T;
    for (int row = warp * RPW + lane / LPR; row < c; row += nw * RPW) {
      float4* p = reinterpret_cast<float4*>(base + (int64_t)row * zyx + 4 * g);
      if (POLICY == 0) *p = make_float4(0, 0, 0, 0); else __stcs(p, make_float4(0, 0, 0, 0));
    }
  }
}

template <typename F> float timeit(F f, float* flush, size_t flush_bytes, int iters = 20) {
  cudaEvent_t s, e; CK(cudaEventCreate(&s)); CK(cudaEventCreate(&e));
  for (int i = 0; i < 3; ++i) f();
  float tot = 0;
  for (int i = 0; i < iters; ++i) {
    CK(cudaMemsetAsync(flush, 0, flush_bytes));
    CK(cudaEventRecord(s)); f(); CK(cudaEventRecord(e)); CK(cudaEventSynchronize(e));
    float ms; CK(cudaEventElapsedTime(&ms, s, e)); tot += ms;
  }
  CK(cudaGetLastError());
  return tot / iters * 1e3f;
}

int main() {
  const int c = 80; const int64_t zyx = 640000; const size_t n = (size_t)c * zyx;
  float* out; CK(cudaMalloc(&out, n * 4)); float* flush; const size_t fb = 256u << 20; CK(cudaMalloc(&flush, fb));
  printf("volume %.1f MB\n", n * 4 / 1e6);
  auto rep = [&](const char* name, float us) { printf("%-44s %8.1f us  %7.1f GB/s\n", name, us, n * 4 / us / 1e3); };
  rep("cudaMemset", timeit([&] { CK(cudaMemsetAsync(out, 0, n * 4)); }, flush, fb));
  rep("linear float4 default (148*8 CTAs x256)", timeit([&] { linear_default<<<148 * 8, 256>>>((float4*)out, n / 4); }, flush, fb));
  rep("linear float4 st.cs", timeit([&] { linear_cs<<<148 * 8, 256>>>((float4*)out, n / 4); }, flush, fb));
  rep("tile T=128 rows default, 256 thr", timeit([&] { tile_rows<128, 0><<<zyx / 128, 256>>>(out, c, zyx); }, flush, fb));
  rep("tile T=128 rows st.cs, 256 thr", timeit([&] { tile_rows<128, 1><<<zyx / 128, 256>>>(out, c, zyx); }, flush, fb));
  rep("tile T=128 rows st.cg, 256 thr", timeit([&] { tile_rows<128, 2><<<zyx / 128, 256>>>(out, c, zyx); }, flush, fb));
  rep("tile T=128 rows st.cs, 128 thr", timeit([&] { tile_rows<128, 1><<<zyx / 128, 128>>>(out, c, zyx); }, flush, fb));
  rep("tile T=64 rows st.cs, 128 thr", timeit([&] { tile_rows<64, 1><<<zyx / 64, 128>>>(out, c, zyx); }, flush, fb));
  rep("tile T=256 rows st.cs, 256 thr", timeit([&] { tile_rows<256, 1><<<zyx / 256, 256>>>(out, c, zyx); }, flush, fb));
  rep("tile T=512 rows st.cs, 256 thr", timeit([&] { tile_rows<512, 1><<<zyx / 512, 256>>>(out, c, zyx); }, flush, fb));
  rep("tile T=1280 rows default, 256 thr", timeit([&] { tile_rows<1280, 0><<<zyx / 1280, 256>>>(out, c, zyx); }, flush, fb));
  rep("tile T=128 bulk(TMA) zero row, 128 thr", timeit([&] { tile_rows_bulk<128><<<zyx / 128, 128>>>(out, c, zyx); }, flush, fb));
  rep("tile T=256 bulk(TMA) zero row, 128 thr", timeit([&] { tile_rows_bulk<256><<<zyx / 256, 128>>>(out, c, zyx); }, flush, fb));
  rep("tile T=512 bulk(TMA) zero row, 128 thr", timeit([&] { tile_rows_bulk<512><<<zyx / 512, 128>>>(out, c, zyx); }, flush, fb));
  rep("tile T=128 persistent 148*4 x256 st.cs", timeit([&] { tile_rows_persistent<128, 1><<<148 * 4, 256>>>(out, c, zyx, zyx / 128); }, flush, fb));
  rep("tile T=128 persistent 148*8 x256 default", timeit([&] { tile_rows_persistent<128, 0><<<148 * 8, 256>>>(out, c, zyx, zyx / 128); }, flush, fb));
  return 0;
}
