// msda_common.cuh -- sampling helpers shared by the deformable-attention kernels
#pragma once
#include "common.cuh"

namespace fbbev {

// scalar bilinear (depth look-up): value laid out [pixel][stride] floats
__device__ __forceinline__ float sample_scalar(const float* __restrict__ val,
                                               int H, int W, int stride,
                                               float h_im, float w_im) {
  if (!(h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W))
    return 0.f;
  const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
  const int h_high = h_low + 1, w_high = w_low + 1;
  const float lh = h_im - (float)h_low, lw = w_im - (float)w_low;
  const float hh = 1.f - lh, hw = 1.f - lw;
  const float* p1 = val + ((int64_t)h_low * W + w_low) * stride;
  float v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f;
  if (h_low >= 0 && w_low >= 0) v1 = __ldg(p1);
  if (h_low >= 0 && w_high <= W - 1) v2 = __ldg(p1 + stride);
  if (h_high <= H - 1 && w_low >= 0) v3 = __ldg(p1 + (int64_t)W * stride);
  if (h_high <= H - 1 && w_high <= W - 1)
    v4 = __ldg(p1 + (int64_t)W * stride + stride);
  const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
  return w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4;
}

// pixel coordinate of a normalised location: loc * size - 0.5, product rounded
// to fp32 before the subtraction exactly as in mmcv's kernel (no FMA contraction)
__device__ __forceinline__ float pix(float loc, int size) {
  return __fsub_rn(__fmul_rn(loc, (float)size), 0.5f);
}

}  // namespace fbbev
