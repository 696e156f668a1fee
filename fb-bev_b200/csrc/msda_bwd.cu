// msda_bwd.cu -- multi-scale deformable attention, backward, for sm_100a.
//
// Drop-in for mmcv `_ext.ms_deform_attn_backward`
//   (call site .../bevformer_utils/multi_scale_deformable_attn_function.py:159-169;
//    arithmetic: mmcv-full 1.5.2 ms_deformable_col2im_gpu_kernel_*, the published
//    Deformable-DETR col2im).
//
// One thread per (batch, query, head): it owns every (level, point) of that head,
// so grad_sampling_loc and grad_attn_weight are plain stores (the reference
// reduces them across CH threads through shared memory); only grad_value, which
// many queries hit, uses red.global.add.f32.
#include "common.cuh"

namespace fbbev {

constexpr int kMsdaBwdThreads = 128;

template <int CH>
__global__ void __launch_bounds__(kMsdaBwdThreads) msda_bwd_kernel(
    const float* __restrict__ value, const int64_t* __restrict__ shapes,
    const int64_t* __restrict__ lstart, const float* __restrict__ loc,
    const float* __restrict__ attw, const float* __restrict__ grad_out,
    int64_t n_items, int n_value, int heads, int levels, int nq, int points,
    float* __restrict__ grad_value, float* __restrict__ grad_loc,
    float* __restrict__ grad_attw) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_items) return;
  const int m = (int)(idx % heads);
  const int64_t bq = idx / heads;
  const int b = (int)(bq / nq);
  const int E = heads * CH;
  float go[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) go[c] = __ldg(grad_out + bq * E + m * CH + c);
  const int64_t wbase = idx * levels * points;
  for (int l = 0; l < levels; ++l) {
    const int H = (int)__ldg(shapes + 2 * l), W = (int)__ldg(shapes + 2 * l + 1);
    const int64_t voff = ((int64_t)b * n_value + __ldg(lstart + l)) * E + m * CH;
    const float* val = value + voff;
    float* gval = grad_value + voff;
    for (int p = 0; p < points; ++p) {
      const int64_t wi = wbase + (int64_t)l * points + p;
      const float2 xy = __ldg(reinterpret_cast<const float2*>(loc) + wi);
      const float wgt = __ldg(attw + wi);
      const float h = __fsub_rn(__fmul_rn(xy.y, (float)H), 0.5f);
      const float w = __fsub_rn(__fmul_rn(xy.x, (float)W), 0.5f);
      if (!(h > -1.f && w > -1.f && h < (float)H && w < (float)W)) continue;
      const int h_low = (int)floorf(h), w_low = (int)floorf(w);
      const int h_high = h_low + 1, w_high = w_low + 1;
      const float lh = h - (float)h_low, lw = w - (float)w_low;
      const float hh = 1.f - lh, hw = 1.f - lw;
      const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
      const bool ok1 = h_low >= 0 && w_low >= 0;
      const bool ok2 = h_low >= 0 && w_high <= W - 1;
      const bool ok3 = h_high <= H - 1 && w_low >= 0;
      const bool ok4 = h_high <= H - 1 && w_high <= W - 1;
      const int64_t o1 = ((int64_t)h_low * W + w_low) * E;
      const int64_t o2 = o1 + E, o3 = o1 + (int64_t)W * E, o4 = o3 + E;
      float g_h = 0.f, g_w = 0.f, g_a = 0.f;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const float top = go[c] * wgt;
        float v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f;
        if (ok1) { v1 = __ldg(val + o1 + c); atomicAdd(gval + o1 + c, w1 * top); }
        if (ok2) { v2 = __ldg(val + o2 + c); atomicAdd(gval + o2 + c, w2 * top); }
        if (ok3) { v3 = __ldg(val + o3 + c); atomicAdd(gval + o3 + c, w3 * top); }
        if (ok4) { v4 = __ldg(val + o4 + c); atomicAdd(gval + o4 + c, w4 * top); }
        g_h += (-hw * v1 - lw * v2 + hw * v3 + lw * v4) * top;
        g_w += (-hh * v1 + hh * v2 - lh * v3 + lh * v4) * top;
        g_a += (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4) * go[c];
      }
      grad_attw[wi] = g_a;
      reinterpret_cast<float2*>(grad_loc)[wi] =
          make_float2((float)W * g_w, (float)H * g_h);
    }
  }
}

// Any head width: one thread per (batch, query, head), channels walked serially.
__global__ void __launch_bounds__(kMsdaBwdThreads) msda_bwd_generic_kernel(
    const float* __restrict__ value, const int64_t* __restrict__ shapes,
    const int64_t* __restrict__ lstart, const float* __restrict__ loc,
    const float* __restrict__ attw, const float* __restrict__ grad_out,
    int64_t n_items, int n_value, int heads, int ch, int levels, int nq,
    int points, float* __restrict__ grad_value, float* __restrict__ grad_loc,
    float* __restrict__ grad_attw) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_items) return;
  const int m = (int)(idx % heads);
  const int64_t bq = idx / heads;
  const int b = (int)(bq / nq);
  const int E = heads * ch;
  const float* go = grad_out + bq * E + m * ch;
  const int64_t wbase = idx * levels * points;
  for (int l = 0; l < levels; ++l) {
    const int H = (int)__ldg(shapes + 2 * l), W = (int)__ldg(shapes + 2 * l + 1);
    const int64_t voff = ((int64_t)b * n_value + __ldg(lstart + l)) * E + m * ch;
    const float* val = value + voff;
    float* gval = grad_value + voff;
    for (int p = 0; p < points; ++p) {
      const int64_t wi = wbase + (int64_t)l * points + p;
      const float2 xy = __ldg(reinterpret_cast<const float2*>(loc) + wi);
      const float wgt = __ldg(attw + wi);
      const float h = __fsub_rn(__fmul_rn(xy.y, (float)H), 0.5f);
      const float w = __fsub_rn(__fmul_rn(xy.x, (float)W), 0.5f);
      if (!(h > -1.f && w > -1.f && h < (float)H && w < (float)W)) continue;
      const int h_low = (int)floorf(h), w_low = (int)floorf(w);
      const int h_high = h_low + 1, w_high = w_low + 1;
      const float lh = h - (float)h_low, lw = w - (float)w_low;
      const float hh = 1.f - lh, hw = 1.f - lw;
      const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
      const bool ok1 = h_low >= 0 && w_low >= 0;
      const bool ok2 = h_low >= 0 && w_high <= W - 1;
      const bool ok3 = h_high <= H - 1 && w_low >= 0;
      const bool ok4 = h_high <= H - 1 && w_high <= W - 1;
      const int64_t o1 = ((int64_t)h_low * W + w_low) * E;
      const int64_t o2 = o1 + E, o3 = o1 + (int64_t)W * E, o4 = o3 + E;
      float g_h = 0.f, g_w = 0.f, g_a = 0.f;
      for (int c = 0; c < ch; ++c) {
        const float g = __ldg(go + c);
        const float top = g * wgt;
        float v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f;
        if (ok1) { v1 = __ldg(val + o1 + c); atomicAdd(gval + o1 + c, w1 * top); }
        if (ok2) { v2 = __ldg(val + o2 + c); atomicAdd(gval + o2 + c, w2 * top); }
        if (ok3) { v3 = __ldg(val + o3 + c); atomicAdd(gval + o3 + c, w3 * top); }
        if (ok4) { v4 = __ldg(val + o4 + c); atomicAdd(gval + o4 + c, w4 * top); }
        g_h += (-hw * v1 - lw * v2 + hw * v3 + lw * v4) * top;
        g_w += (-hh * v1 + hh * v2 - lh * v3 + lh * v4) * top;
        g_a += (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4) * g;
      }
      grad_attw[wi] = g_a;
      reinterpret_cast<float2*>(grad_loc)[wi] =
          make_float2((float)W * g_w, (float)H * g_h);
    }
  }
}

}  // namespace fbbev

using namespace fbbev;

FBBEV_API int fbbev_msda_bwd(const float* value, const int64_t* spatial_shapes,
                             const int64_t* level_start, const float* loc,
                             const float* attw, const float* grad_out,
                             int32_t bs, int32_t n_value, int32_t heads,
                             int32_t ch, int32_t levels, int32_t nq,
                             int32_t points, float* grad_value, float* grad_loc,
                             float* grad_attw, fbbev_stream_t stream) {
  if (bs < 0 || nq < 0 || n_value <= 0 || heads <= 0 || ch <= 0 ||
      levels <= 0 || points <= 0)
    return FBBEV_ERR_INVALID_ARGUMENT;
  const int64_t n_items = (int64_t)bs * nq * heads;
  if (n_items == 0) return FBBEV_OK;
  if (!value || !spatial_shapes || !level_start || !loc || !attw ||
      !grad_out || !grad_value || !grad_loc || !grad_attw)
    return FBBEV_ERR_INVALID_ARGUMENT;
  const unsigned grid = (unsigned)ceil_div64(n_items, kMsdaBwdThreads);
  cudaStream_t st = as_stream(stream);
  count_launch();
#define FBBEV_LAUNCH(CHV)                                                     \
  msda_bwd_kernel<CHV><<<grid, kMsdaBwdThreads, 0, st>>>(                     \
      value, spatial_shapes, level_start, loc, attw, grad_out, n_items,       \
      n_value, heads, levels, nq, points, grad_value, grad_loc, grad_attw);   \
  break
  switch (ch) {
    case 4: FBBEV_LAUNCH(4);
    case 8: FBBEV_LAUNCH(8);
    case 10: FBBEV_LAUNCH(10);
    case 16: FBBEV_LAUNCH(16);
    case 20: FBBEV_LAUNCH(20);
    case 32: FBBEV_LAUNCH(32);
    case 64: FBBEV_LAUNCH(64);
    default:
      msda_bwd_generic_kernel<<<grid, kMsdaBwdThreads, 0, st>>>(
          value, spatial_shapes, level_start, loc, attw, grad_out, n_items,
          n_value, heads, ch, levels, nq, points, grad_value, grad_loc,
          grad_attw);
      break;
  }
#undef FBBEV_LAUNCH
  return launch_status();
}
