// lift_tail.cu -- the producer side of the lift-splat op: the tail of the depth
// net, emitting `depth` and `feat` in the layouts the pooling kernels read.
//
// What it replaces (the step right before the path):
//   * CM_DepthNet.forward, mmdet3d/models/fbbev/modules/depth_net.py:359-363
//       depth = depth.softmax(dim=1); context.view(B, N, C, H, W)
//   * LSSViewTransformer(2 / BEVDepth).forward, mmdet3d/models/necks/
//     view_transformer.py:313-321, 710-718, 1094-1096
//       depth_digit = x[:, :D]; tran_feat = x[:, D:D + C]; depth = softmax(dim=1)
//   followed, inside the pooling op, by `feat.permute(0,1,3,4,2)` +
//   `feat.contiguous()` (view_transformer.py:530, bev_pool.py:19): a strided
//   softmax kernel plus a transposing copy in the reference.
//
// One launch:
//   softmax blocks   32 pixels x 4 depth lanes; lanes of a warp are consecutive
//                    pixels (the depth bins of a pixel are H*W floats apart in
//                    NCHW), the four warps split the bins; max / sum are
//                    combined through shared memory.  Three passes over the
//                    logits as torch's softmax: max, sum of expf(x - max),
//                    expf(x - max) / sum -- the second and third hit L1.
//   transpose blocks 32 pixels x 32 channels through a padded shared-memory
//                    tile: (C, H*W) -> (H*W, C), 128-byte rows on both sides.
// The two inputs may be channel slices of ONE tensor (the BEVDet-lineage
// depth_net output): each has its own per-image stride.
#include "common.cuh"

namespace fbbev {

struct LiftTailParams {
  const float* logits;   // image n, bin d, pixel p at logits[n*logits_ns + d*HW + p]
  const float* ctx;      // image n, channel c, pixel p at ctx[n*ctx_ns + c*HW + p]
  int64_t logits_ns, ctx_ns;
  float* depth;          // (BN, D, HW) dense
  float* feat;           // (BN, HW, C) dense
  int BN, D, C, HW;
  int px_tiles, c_tiles; // ceil(HW / 32), ceil(C / 32)
  int n_softmax_blocks;  // BN * px_tiles (0 when logits == nullptr)
};

__global__ void __launch_bounds__(128) lift_tail_kernel(LiftTailParams P) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ float s_red[4][32];
  __shared__ float s_tile[32][33];
  if ((int)blockIdx.x < P.n_softmax_blocks) {
    const int n = blockIdx.x / P.px_tiles, pt = blockIdx.x % P.px_tiles;
    const int p = pt * 32 + lane;
    const bool ok = p < P.HW;
    const float* src = P.logits + (int64_t)n * P.logits_ns + p;
    float* dst = P.depth + (int64_t)n * P.D * P.HW + p;
    // pass 1: max over this warp's bins d = warp, warp + 4, ...
    float mx = -INFINITY;
    if (ok)
      for (int d = warp; d < P.D; d += 4)
        mx = fmaxf(mx, __ldg(src + (int64_t)d * P.HW));
    s_red[warp][lane] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(s_red[0][lane], s_red[1][lane]),
               fmaxf(s_red[2][lane], s_red[3][lane]));
    __syncthreads();
    // pass 2: sum of exp(x - max)
    float sum = 0.f;
    if (ok)
      for (int d = warp; d < P.D; d += 4)
        sum += expf(__ldg(src + (int64_t)d * P.HW) - mx);
    s_red[warp][lane] = sum;
    __syncthreads();
    sum = (s_red[0][lane] + s_red[1][lane]) + (s_red[2][lane] + s_red[3][lane]);
    // pass 3: normalised probabilities
    if (ok)
      for (int d = warp; d < P.D; d += 4)
        dst[(int64_t)d * P.HW] =
            __fdiv_rn(expf(__ldg(src + (int64_t)d * P.HW) - mx), sum);
    return;
  }
  // ---- context: (C, HW) -> (HW, C) ----
  int t = blockIdx.x - P.n_softmax_blocks;
  const int ct = t % P.c_tiles;
  t /= P.c_tiles;
  const int pt = t % P.px_tiles, n = t / P.px_tiles;
  const int p0 = pt * 32, c0 = ct * 32;
  const float* src = P.ctx + (int64_t)n * P.ctx_ns;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + warp * 8 + i, p = p0 + lane;
    s_tile[warp * 8 + i][lane] =
        (c < P.C && p < P.HW) ? __ldg(src + (int64_t)c * P.HW + p) : 0.f;
  }
  __syncthreads();
  float* dst = P.feat + (int64_t)n * P.HW * P.C;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int p = p0 + warp * 8 + i, c = c0 + lane;
    if (p < P.HW && c < P.C) dst[(int64_t)p * P.C + c] = s_tile[lane][warp * 8 + i];
  }
}

}  // namespace fbbev

using namespace fbbev;

FBBEV_API int fbbev_lift_tail_fwd(const float* depth_logits,
                                  int64_t logits_image_stride,
                                  const float* context,
                                  int64_t context_image_stride, int32_t bn,
                                  int32_t d, int32_t c, int32_t hw,
                                  float* depth_out, float* feat_out,
                                  fbbev_stream_t stream) {
  if (bn < 0 || hw <= 0 || (!depth_logits && !context))
    return FBBEV_ERR_INVALID_ARGUMENT;
  if (depth_logits && (d <= 0 || !depth_out || logits_image_stride < (int64_t)d * hw))
    return FBBEV_ERR_INVALID_ARGUMENT;
  if (context && (c <= 0 || !feat_out || context_image_stride < (int64_t)c * hw))
    return FBBEV_ERR_INVALID_ARGUMENT;
  if (bn == 0) return FBBEV_OK;
  LiftTailParams P;
  P.logits = depth_logits; P.ctx = context;
  P.logits_ns = logits_image_stride; P.ctx_ns = context_image_stride;
  P.depth = depth_out; P.feat = feat_out;
  P.BN = bn; P.D = d; P.C = c; P.HW = hw;
  P.px_tiles = (hw + 31) / 32;
  P.c_tiles = context ? (c + 31) / 32 : 0;
  P.n_softmax_blocks = depth_logits ? bn * P.px_tiles : 0;
  const int64_t blocks =
      (int64_t)P.n_softmax_blocks + (int64_t)bn * P.px_tiles * P.c_tiles;
  if (blocks > INT32_MAX) return FBBEV_ERR_UNSUPPORTED;
  count_launch();
  lift_tail_kernel<<<(unsigned)blocks, 128, 0, as_stream(stream)>>>(P);
  return launch_status();
}
