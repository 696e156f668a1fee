// oracle/ref_shim.cu -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the reference's own bev_pool_v2 CUDA kernels
// (mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu) from the place they live in
// the reference checkout -- the path is injected by oracle/Makefile through
// -DREF_BEV_POOL_CU -- and exposes its two host launchers
// (`bev_pool_v2`, bev_pool_cuda.cu:120-128; `bev_pool_v2_grad`, :130-137) under
// C names so tests/ and bench.py can call the UNMODIFIED reference kernels on
// the GPU box through ctypes.  No reference source is copied into this repo;
// the built library lands in oracle/_ref/ (git-ignored).
#ifndef REF_BEV_POOL_CU
#error "build through oracle/Makefile (needs -DREF_BEV_POOL_CU=\"...\")"
#endif
#include REF_BEV_POOL_CU

#include <cuda_runtime.h>

extern "C" {

// Same argument order as the reference launcher (bev_pool_cuda.cu:120-121).
// The reference launches on the legacy default stream and returns nothing.
__attribute__((visibility("default"))) int ref_bev_pool_v2(
    int c, int n_intervals, const float* depth, const float* feat,
    const int* ranks_depth, const int* ranks_feat, const int* ranks_bev,
    const int* interval_starts, const int* interval_lengths, float* out) {
  bev_pool_v2(c, n_intervals, depth, feat, ranks_depth, ranks_feat, ranks_bev,
              interval_starts, interval_lengths, out);
  return (int)cudaGetLastError();
}

__attribute__((visibility("default"))) int ref_bev_pool_v2_grad(
    int c, int n_intervals, const float* out_grad, const float* depth,
    const float* feat, const int* ranks_depth, const int* ranks_feat,
    const int* ranks_bev, const int* interval_starts,
    const int* interval_lengths, float* depth_grad, float* feat_grad) {
  bev_pool_v2_grad(c, n_intervals, out_grad, depth, feat, ranks_depth,
                   ranks_feat, ranks_bev, interval_starts, interval_lengths,
                   depth_grad, feat_grad);
  return (int)cudaGetLastError();
}

}  // extern "C"
