// bev_pool_bwd.cu -- backward of the lift-splat voxel pooling for sm_100a.
//
// Replaces  bev_pool_grad_kernel + launcher  mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu:64-118, 130-137
//           bev_pool_v2_backward              mmdet3d/ops/bev_pool_v2/src/bev_pool.cpp:72-102
//
// The reference runs ONE THREAD per feature-pixel interval and walks all C
// channels serially inside it (two nested loops, :88-117).  Here one warp owns
// an interval and its lanes own channels: every out_grad row is read once,
// coalesced, and feeds both gradients --
//   depth_grad[rd_k]   = sum_c out_grad[rb_k, c] * feat[rf, c]   (warp reduction)
//   feat_grad[rf, c]   = sum_k out_grad[rb_k, c] * depth[rd_k]   (per-lane FMA chain,
//                        same point order as the reference)
// LAYOUT = 0: out_grad is (B,Z,Y,X,C) as the reference op receives it after
//             `out_grad.contiguous()` (bev_pool.py:67);
// LAYOUT = 1: out_grad is (B,C,Z,Y,X), the layout the gradient actually arrives
//             in (the forward returns the permuted tensor, bev_pool.py:89), so the
//             fused path skips the full-volume transpose copy.
#include "common.cuh"

namespace fbbev {

constexpr int kBwdThreads = 256;
constexpr int kBwdWarps = kBwdThreads / kWarp;

template <int NCH, int LAYOUT>
__global__ void __launch_bounds__(kBwdThreads) bev_pool_grad_kernel(
    const float* __restrict__ out_grad, const float* __restrict__ depth,
    const float* __restrict__ feat, const int* __restrict__ ranks_depth,
    const int* __restrict__ ranks_feat, const int* __restrict__ ranks_bev,
    const int* __restrict__ interval_starts,
    const int* __restrict__ interval_lengths, int n_intervals, int c,
    int64_t zyx, float* __restrict__ depth_grad,
    float* __restrict__ feat_grad) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * kBwdWarps + (threadIdx.x >> 5);
  if (i >= n_intervals) return;
  const int start = __ldg(interval_starts + i);
  const int len = __ldg(interval_lengths + i);
  const int rf0 = __ldg(ranks_feat + start);

  float fg[NCH];
#pragma unroll
  for (int r = 0; r < NCH; ++r) fg[r] = 0.f;
  for (int base = 0; base < len; base += kWarp) {
    const int n = min(kWarp, len - base);
    int rb = 0, rd = 0, rf = 0;
    float d = 0.f;
    if (lane < n) {
      rb = __ldg(ranks_bev + start + base + lane);
      rd = __ldg(ranks_depth + start + base + lane);
      rf = __ldg(ranks_feat + start + base + lane);
      d = __ldg(depth + rd);
    }
    for (int k = 0; k < n; ++k) {
      const int rbk = __shfl_sync(kFull, rb, k);
      const float dk = __shfl_sync(kFull, d, k);
      // the reference reads the point's own feat row (:90); inside a
      // ranks_feat run this is the same row every time (L1 hit)
      const float* f = feat + (int64_t)__shfl_sync(kFull, rf, k) * c;
      const float* og;
      int64_t cs;
      if (LAYOUT == 0) {
        og = out_grad + (int64_t)rbk * c;
        cs = 1;
      } else {
        const int64_t b = rbk / zyx;
        og = out_grad + b * c * zyx + (rbk - b * zyx);
        cs = zyx;
      }
      float dot = 0.f;
#pragma unroll
      for (int r = 0; r < NCH; ++r) {
        const int ch = lane + kWarp * r;
        if (ch < c) {
          const float g = __ldg(og + ch * cs);
          dot = fmaf(g, __ldg(f + ch), dot);
          fg[r] = fmaf(g, dk, fg[r]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(kFull, dot, o);
      // the reference indexes depth_grad with the point's own ranks_depth (:100)
      const int rdk = __shfl_sync(kFull, rd, k);
      if (lane == 0) depth_grad[rdk] = dot;
    }
  }
  // the reference writes feat_grad at ranks_feat[interval_start] (:115)
#pragma unroll
  for (int r = 0; r < NCH; ++r) {
    const int ch = lane + kWarp * r;
    if (ch < c) feat_grad[(int64_t)rf0 * c + ch] = fg[r];
  }
}

template <int LAYOUT>
static int launch_grad(const float* out_grad, const float* depth,
                       const float* feat, const int* ranks_depth,
                       const int* ranks_feat, const int* ranks_bev,
                       const int* interval_starts, const int* interval_lengths,
                       int n_intervals, int c, int64_t zyx, float* depth_grad,
                       float* feat_grad, cudaStream_t st) {
  const unsigned grid = (unsigned)ceil_div64(n_intervals, kBwdWarps);
  const int nch = (c + kWarp - 1) / kWarp;
  count_launch();
#define FBBEV_BWD_CASE(N)                                                     \
  bev_pool_grad_kernel<N, LAYOUT><<<grid, kBwdThreads, 0, st>>>(              \
      out_grad, depth, feat, ranks_depth, ranks_feat, ranks_bev,              \
      interval_starts, interval_lengths, n_intervals, c, zyx, depth_grad,     \
      feat_grad)
  if (nch <= 1) FBBEV_BWD_CASE(1);
  else if (nch <= 2) FBBEV_BWD_CASE(2);
  else if (nch <= 3) FBBEV_BWD_CASE(3);
  else if (nch <= 4) FBBEV_BWD_CASE(4);
  else if (nch <= 8) FBBEV_BWD_CASE(8);
  else return FBBEV_ERR_UNSUPPORTED;
#undef FBBEV_BWD_CASE
  return launch_status();
}

}  // namespace fbbev

using namespace fbbev;

FBBEV_API int fbbev_bev_pool_v2_bwd(
    const float* out_grad, const float* depth, const float* feat,
    const int32_t* ranks_depth, const int32_t* ranks_feat,
    const int32_t* ranks_bev, const int32_t* interval_starts,
    const int32_t* interval_lengths, int32_t n_intervals, int32_t c,
    float* depth_grad, float* feat_grad, fbbev_stream_t stream) {
  if (n_intervals < 0 || c <= 0) return FBBEV_ERR_INVALID_ARGUMENT;
  if (n_intervals == 0) return FBBEV_OK;
  if (!out_grad || !depth || !feat || !ranks_depth || !ranks_feat ||
      !ranks_bev || !interval_starts || !interval_lengths || !depth_grad ||
      !feat_grad)
    return FBBEV_ERR_INVALID_ARGUMENT;
  return launch_grad<0>(out_grad, depth, feat, ranks_depth, ranks_feat,
                        ranks_bev, interval_starts, interval_lengths,
                        n_intervals, c, 1, depth_grad, feat_grad,
                        as_stream(stream));
}

// Same contract, but out_grad is (B,C,Z,Y,X) with Z*Y*X = n_voxels_per_sample:
// the gradient of the tensor bev_pool_v2() returns (bev_pool.py:89), consumed
// in place without the transpose copy of bev_pool.py:67.
FBBEV_API int fbbev_bev_pool_v2_bwd_bczyx(
    const float* out_grad, const float* depth, const float* feat,
    const int32_t* ranks_depth, const int32_t* ranks_feat,
    const int32_t* ranks_bev, const int32_t* interval_starts,
    const int32_t* interval_lengths, int32_t n_intervals, int32_t c,
    int64_t n_voxels_per_sample, float* depth_grad, float* feat_grad,
    fbbev_stream_t stream) {
  if (n_intervals < 0 || c <= 0 || n_voxels_per_sample <= 0)
    return FBBEV_ERR_INVALID_ARGUMENT;
  if (n_intervals == 0) return FBBEV_OK;
  if (!out_grad || !depth || !feat || !ranks_depth || !ranks_feat ||
      !ranks_bev || !interval_starts || !interval_lengths || !depth_grad ||
      !feat_grad)
    return FBBEV_ERR_INVALID_ARGUMENT;
  return launch_grad<1>(out_grad, depth, feat, ranks_depth, ranks_feat,
                        ranks_bev, interval_starts, interval_lengths,
                        n_intervals, c, n_voxels_per_sample, depth_grad,
                        feat_grad, as_stream(stream));
}
