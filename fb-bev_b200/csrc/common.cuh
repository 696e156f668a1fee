// common.cuh -- shared helpers for the sm_100a kernels behind libfbbev_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fbbev_b200.h"

#define FBBEV_API extern "C" __attribute__((visibility("default")))

namespace fbbev {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

static inline cudaStream_t as_stream(fbbev_stream_t s) {
  return reinterpret_cast<cudaStream_t>(s);
}

// Returns the pending launch error (positive cudaError_t) or 0.
static inline int launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? FBBEV_OK : static_cast<int>(e);
}

// Diagnostics only: cumulative number of kernel launches issued by this library
// (bench.py reports it as `gpu_launches`).  Defined in capi.cu.
void count_launch(int n = 1);

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Streaming (evict-first) 128-bit / 32-bit stores for write-once outputs.
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
  __stcs(p, v);
}
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }

}  // namespace fbbev
