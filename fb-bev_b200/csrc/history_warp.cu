// history_warp.cu -- temporal alignment of the BEV / voxel history (sm_100a).
//
// Replaces, in FBOCC.fuse_history
//   mmdet3d/models/fbbev/detectors/fbocc.py:207-319 (+ generate_grid :170-205)
// the chain
//   grid = rt_flow @ (x, y, z, 1)                 (n,h,w,z,4,4 batched matmul)
//   grid = grid[..., :3] / (w-1, h-1, z-1) * 2 - 1
//   sampled = F.grid_sample(history (n, T*C, Z, H, W), grid, align_corners=True)
//   feats_cat = torch.cat([curr_bev, sampled], 1)
// i.e. a (n, H, W, Z, 4, 4) matmul, a 5-D grid tensor, a 3-D grid_sample over
// the whole history (410 MB per sample for 16 x 80 channels of 8 x 100 x 100)
// and a concatenation copy of it, with ONE pass: every output voxel evaluates
// its source location from the 4x4 flow, and the trilinear sample of each
// history channel is written straight into its slot of the concatenated buffer.
//
// grid_sample semantics reproduced (ATen GridSampler, 3-D, bilinear, zeros
// padding, align_corners=True): source index = ((g + 1) / 2) * (size - 1) with
// g the normalised coordinate; the eight corners weigh by the opposite volumes;
// corners outside the volume contribute zero.
//
// Work decomposition: thread = one output voxel x kChPerThread channels (the
// corner offsets and weights are computed once and reused); consecutive threads
// are consecutive x, so stores are coalesced and the eight gathers are
// coalesced whenever the flow is close to a rigid shift (it is an ego-motion).
// HBM-bound: one read + one write of the history per step.
#include "common.cuh"

namespace fbbev {

constexpr int kWarpThreads = 256;
constexpr int kChPerThread = 16;

struct WarpParams {
  const float* hist;   // (n, MC, Z, H, W)
  const float* flow;   // (n, 4, 4) row-major, voxel index -> voxel index
  float* out;          // (n, C_total, Z, H, W); channels [ch_off, ch_off + MC)
  int n, MC, Z, H, W, C_total, ch_off;
  int64_t hist_bstride;  // floats between samples of `hist`
};

__global__ void __launch_bounds__(kWarpThreads) history_warp_kernel(WarpParams P) {
  const int64_t zhw = (int64_t)P.Z * P.H * P.W;
  const int64_t v = (int64_t)blockIdx.x * kWarpThreads + threadIdx.x;
  if (v >= zhw) return;
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * kChPerThread;
  const int x = (int)(v % P.W), y = (int)((v / P.W) % P.H),
            z = (int)(v / ((int64_t)P.W * P.H));
  const float* f = P.flow + b * 16;
  const float fx = (float)x, fy = (float)y, fz = (float)z;
  // rt_flow @ (x, y, z, 1)   (fbocc.py:199)
  const float px = fmaf(__ldg(f + 2), fz, fmaf(__ldg(f + 1), fy, __ldg(f + 0) * fx)) + __ldg(f + 3);
  const float py = fmaf(__ldg(f + 6), fz, fmaf(__ldg(f + 5), fy, __ldg(f + 4) * fx)) + __ldg(f + 7);
  const float pz = fmaf(__ldg(f + 10), fz, fmaf(__ldg(f + 9), fy, __ldg(f + 8) * fx)) + __ldg(f + 11);
  // normalise (:203) and un-normalise (grid_sampler_unnormalize, align_corners)
  const float sx = (float)(P.W - 1), sy = (float)(P.H - 1), sz = (float)(P.Z - 1);
  const float gx = __fsub_rn(__fmul_rn(__fdiv_rn(px, sx), 2.f), 1.f);
  const float gy = __fsub_rn(__fmul_rn(__fdiv_rn(py, sy), 2.f), 1.f);
  const float gz = __fsub_rn(__fmul_rn(__fdiv_rn(pz, sz), 2.f), 1.f);
  const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), sx);
  const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.f), 2.f), sy);
  const float iz = __fmul_rn(__fdiv_rn(__fadd_rn(gz, 1.f), 2.f), sz);
  const float x0f = floorf(ix), y0f = floorf(iy), z0f = floorf(iz);
  const float tx = ix - x0f, ty = iy - y0f, tz = iz - z0f;
  // NaN / huge coordinates: every corner fails the bounds test below
  const bool finite = fabsf(ix) < 1e9f && fabsf(iy) < 1e9f && fabsf(iz) < 1e9f;
  const int x0 = finite ? (int)x0f : -2, y0 = finite ? (int)y0f : -2,
            z0 = finite ? (int)z0f : -2;
  float wgt[8];
  int64_t off[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
    const int xx = x0 + dx, yy = y0 + dy, zz = z0 + dz;
    const bool in = xx >= 0 && xx < P.W && yy >= 0 && yy < P.H && zz >= 0 &&
                    zz < P.Z;
    const float w = (dx ? tx : 1.f - tx) * (dy ? ty : 1.f - ty) *
                    (dz ? tz : 1.f - tz);
    wgt[k] = in ? w : 0.f;
    off[k] = in ? ((int64_t)zz * P.H + yy) * P.W + xx : 0;
  }
  const float* src = P.hist + (int64_t)b * P.hist_bstride + (int64_t)c0 * zhw;
  float* dst = P.out + ((int64_t)b * P.C_total + P.ch_off + c0) * zhw + v;
  const int nc = min(kChPerThread, P.MC - c0);
  for (int c = 0; c < nc; ++c) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc = fmaf(__ldg(src + off[k]), wgt[k], acc);
    __stcs(dst, acc);
    src += zhw;
    dst += zhw;
  }
}

}  // namespace fbbev

using namespace fbbev;

FBBEV_API int fbbev_history_warp(const float* history,
                                 int64_t history_batch_stride, const float* flow,
                                 int32_t n, int32_t mc, int32_t Z, int32_t H,
                                 int32_t W, float* out, int32_t c_total,
                                 int32_t ch_offset, fbbev_stream_t stream) {
  if (n < 0 || mc < 0 || Z <= 0 || H <= 0 || W <= 0 || c_total < mc ||
      ch_offset < 0 || ch_offset + mc > c_total ||
      history_batch_stride < (int64_t)mc * Z * H * W)
    return FBBEV_ERR_INVALID_ARGUMENT;
  if (n == 0 || mc == 0) return FBBEV_OK;
  if (!history || !flow || !out) return FBBEV_ERR_INVALID_ARGUMENT;
  if (n > 65535 || (mc + kChPerThread - 1) / kChPerThread > 65535)
    return FBBEV_ERR_UNSUPPORTED;
  WarpParams P;
  P.hist = history; P.flow = flow; P.out = out;
  P.n = n; P.MC = mc; P.Z = Z; P.H = H; P.W = W;
  P.C_total = c_total; P.ch_off = ch_offset;
  P.hist_bstride = history_batch_stride;
  const int64_t zhw = (int64_t)Z * H * W;
  const dim3 grid((unsigned)ceil_div64(zhw, kWarpThreads),
                  (unsigned)((mc + kChPerThread - 1) / kChPerThread),
                  (unsigned)n);
  count_launch();
  history_warp_kernel<<<grid, kWarpThreads, 0, as_stream(stream)>>>(P);
  return launch_status();
}
