// geometry.cu -- BEV reference points projected into every camera, one kernel.
//
// Replaces  bevformer_encoder.point_sampling
//           .../backward_projection/bevformer_utils/bevformer_encoder.py:92-120
// which the reference evaluates as ~20 eager PyTorch ops over a
// (B, N, Y, X, Z, 3) tensor (three broadcast batched matmuls through cuBLAS gemv
// -- 8 ms for 200x200x4 points x 6 cameras on B200 -- plus repeat / cat / permute
// copies).  The chain is the reference's:
//   p = inv(bda) ref;  p -= trans;  c = inv(rots inv(K)) p;
//   (u, v) = c.xy / max(c.z, eps);  (u, v, d) = post_rots (u, v, c.z) + post_trans;
//   u /= W_in;  v /= H_in;  mask = d > eps and eps < u, v < 1 - eps
// in fp32, every 3x3 product in the rounding order torch's broadcast matmul has
// on this device (mat3_apply_ref, common.cuh) and the two normalisations as the
// multiplications by a reciprocal that torch's CUDA `tensor /= python_scalar`
// performs (BinaryDivTrueKernel.cu), so ref_cam / depth / mask are bit-identical
// to the eager chain's (tests/test_backward_gpu.py::test_fused_point_sampling_bit_exact).
#include "common.cuh"

namespace fbbev {

struct SamplingParams {
  const float *X, *Y, *Z;          // voxel-centre coordinates per axis
  const float *inv_bda, *trans, *ego2cam, *post_rots, *post_trans;
  int nX, nY, nZ, B, N;
  int order;                       // FBBEV_ORDER_SEQ_* bits
  float inv_w, inv_h, eps, one_minus_eps;
};

__global__ void __launch_bounds__(256) point_sampling_kernel(
    SamplingParams P, float* __restrict__ ref_cam, float* __restrict__ depth,
    uint8_t* __restrict__ mask) {
  // one thread per (n, b, q, z) in the OUTPUT order (N, B, Y*X, Z)
  const int64_t nq = (int64_t)P.nY * P.nX;
  const int64_t total = (int64_t)P.N * P.B * nq * P.nZ;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int z = (int)(i % P.nZ);
  int64_t t = i / P.nZ;
  const int64_t q = t % nq;
  t /= nq;
  const int b = (int)(t % P.B);
  const int n = (int)(t / P.B);
  const int ix = (int)(q % P.nX), iy = (int)(q / P.nX);
  const int bn = b * P.N + n;

  float x, y, zz;
  mat3_apply_ref(P.inv_bda + b * 9, __ldg(P.X + ix), __ldg(P.Y + iy),
                 __ldg(P.Z + z), (P.order & 1) != 0, x, y, zz);
  const float* tr = P.trans + bn * 3;
  x = __fsub_rn(x, __ldg(tr + 0));
  y = __fsub_rn(y, __ldg(tr + 1));
  zz = __fsub_rn(zz, __ldg(tr + 2));
  float cx, cy, cz;
  mat3_apply_ref(P.ego2cam + bn * 9, x, y, zz, (P.order & 2) != 0, cx, cy, cz);
  const float den = fmaxf(cz, P.eps);
  const float u = __fdiv_rn(cx, den), v = __fdiv_rn(cy, den);
  float pu, pv, pd;
  mat3_apply_ref(P.post_rots + bn * 9, u, v, cz, (P.order & 4) != 0, pu, pv,
                 pd);
  const float* pt = P.post_trans + bn * 3;
  pu = __fadd_rn(pu, __ldg(pt + 0));
  pv = __fadd_rn(pv, __ldg(pt + 1));
  pd = __fadd_rn(pd, __ldg(pt + 2));
  pu = __fmul_rn(pu, P.inv_w);  // cam[..., 0] /= ogfW  on CUDA: a * (1 / b)
  pv = __fmul_rn(pv, P.inv_h);
  const bool m = pd > P.eps && pu > P.eps && pu < P.one_minus_eps &&
                 pv > P.eps && pv < P.one_minus_eps;
  reinterpret_cast<float2*>(ref_cam)[i] = make_float2(pu, pv);
  depth[i] = pd;
  mask[i] = m ? 1 : 0;
}

// bev_queries[b, q, :] = embedding[q, :] + lss_bev[b, :, q]
//   backward_projection.py:93-97: bev_embedding.weight.unsqueeze(1).repeat(1, bs, 1)
//   + lss_bev.flatten(2).permute(2, 0, 1), as ONE pass producing the
//   (bs, nq, E)-contiguous tensor the encoder works on (the reference-shaped
//   (nq, bs, E) tensor is its permuted view).  The (E, nq) -> (nq, E) transpose
//   of lss_bev goes through a padded shared-memory tile so that both the read
//   and the write are coalesced.
__global__ void __launch_bounds__(256) bev_query_init_kernel(
    const float* __restrict__ emb, const float* __restrict__ lss, int bs, int nq,
    int E, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int q0 = blockIdx.x * 32, e0 = blockIdx.y * 32, b = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  if (lss) {
    const float* src = lss + (int64_t)b * E * nq;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = e0 + ty + 8 * i, q = q0 + tx;
      tile[ty + 8 * i][tx] = (e < E && q < nq) ? __ldg(src + (int64_t)e * nq + q)
                                               : 0.f;
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + ty + 8 * i, e = e0 + tx;
    if (q < nq && e < E) {
      float v = __ldg(emb + (int64_t)q * E + e);
      if (lss) v = __fadd_rn(v, tile[tx][ty + 8 * i]);
      out[((int64_t)b * nq + q) * E + e] = v;
    }
  }
}

// Same, 128 queries x 32 channels per block with 128-bit accesses on both
// sides (nq % 4 == 0, E % 4 == 0, 16-byte aligned pointers): four independent
// 2 KB warp loads in flight per warp instead of four 128-byte ones.  The first
// kernel of a step reads its 25 MB of inputs from DRAM: memory-level
// parallelism per thread is what it is bound by.
__global__ void __launch_bounds__(256) bev_query_init_v4_kernel(
    const float* __restrict__ emb, const float* __restrict__ lss, int bs, int nq,
    int E, float* __restrict__ out) {
  __shared__ __align__(16) float tile[32][132];
  const int q0 = blockIdx.x * 128, e0 = blockIdx.y * 32, b = blockIdx.z;
  const int lane = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* src = lss + (int64_t)b * E * nq;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int e = e0 + ty + 8 * i, q = q0 + 4 * lane;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e < E && q < nq)
      v = __ldg(reinterpret_cast<const float4*>(src + (int64_t)e * nq + q));
    *reinterpret_cast<float4*>(&tile[ty + 8 * i][4 * lane]) = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = threadIdx.x + 256 * i;
    const int ql = idx >> 3, e4 = idx & 7;
    const int q = q0 + ql, e = e0 + 4 * e4;
    if (q < nq && e < E) {
      const float4 a =
          __ldg(reinterpret_cast<const float4*>(emb + (int64_t)q * E + e));
      float4 r;
      r.x = __fadd_rn(a.x, tile[4 * e4 + 0][ql]);
      r.y = __fadd_rn(a.y, tile[4 * e4 + 1][ql]);
      r.z = __fadd_rn(a.z, tile[4 * e4 + 2][ql]);
      r.w = __fadd_rn(a.w, tile[4 * e4 + 3][ql]);
      *reinterpret_cast<float4*>(out + ((int64_t)b * nq + q) * E + e) = r;
    }
  }
}

// out[b, e, q] = in[b, q, e]: the (bs, nq, E) token tensor the encoder ends with
// -> the (bs, E, bev_h, bev_w) map BackwardProjection returns
// (backward_projection.py:131-133: permute(0, 2, 1).view(...).contiguous()).
// 128 queries x 32 channels per block through shared memory, 128-bit accesses on
// both sides.
__global__ void __launch_bounds__(256) tokens_to_map_kernel(
    const float* __restrict__ in, int bs, int nq, int E, float* __restrict__ out) {
  __shared__ __align__(16) float tile[32][132];
  const int q0 = blockIdx.x * 128, e0 = blockIdx.y * 32, b = blockIdx.z;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = threadIdx.x + 256 * i;
    const int ql = idx >> 3, e4 = idx & 7;
    const int q = q0 + ql, e = e0 + 4 * e4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < nq && e < E)
      v = __ldg(reinterpret_cast<const float4*>(in + ((int64_t)b * nq + q) * E + e));
    tile[4 * e4 + 0][ql] = v.x;
    tile[4 * e4 + 1][ql] = v.y;
    tile[4 * e4 + 2][ql] = v.z;
    tile[4 * e4 + 3][ql] = v.w;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int e = e0 + ty + 8 * i, q = q0 + 4 * lane;
    if (e < E && q < nq)
      st_stream(reinterpret_cast<float4*>(out + ((int64_t)b * E + e) * nq + q),
                *reinterpret_cast<const float4*>(&tile[ty + 8 * i][4 * lane]));
  }
}

}  // namespace fbbev

using namespace fbbev;

FBBEV_API int fbbev_tokens_to_map(const float* tokens, int32_t bs, int32_t nq,
                                  int32_t E, float* out, fbbev_stream_t stream) {
  if (bs <= 0 || nq <= 0 || E <= 0 || !tokens || !out)
    return FBBEV_ERR_INVALID_ARGUMENT;
  if (nq % 4 || E % 4 || bs > 65535 || (E + 31) / 32 > 65535 ||
      ((reinterpret_cast<uintptr_t>(tokens) | reinterpret_cast<uintptr_t>(out)) & 15))
    return FBBEV_ERR_UNSUPPORTED;
  const dim3 grid((unsigned)((nq + 127) / 128), (unsigned)((E + 31) / 32),
                  (unsigned)bs);
  count_launch();
  tokens_to_map_kernel<<<grid, 256, 0, as_stream(stream)>>>(tokens, bs, nq, E, out);
  return launch_status();
}

FBBEV_API int fbbev_bev_query_init(const float* embedding, const float* lss_bev,
                                   int32_t bs, int32_t nq, int32_t E, float* out,
                                   fbbev_stream_t stream) {
  if (bs <= 0 || nq <= 0 || E <= 0 || !embedding || !out)
    return FBBEV_ERR_INVALID_ARGUMENT;
  if (bs > 65535 || (E + 31) / 32 > 65535) return FBBEV_ERR_UNSUPPORTED;
  count_launch();
  const bool aligned =
      ((reinterpret_cast<uintptr_t>(embedding) | reinterpret_cast<uintptr_t>(out) |
        reinterpret_cast<uintptr_t>(lss_bev)) & 15) == 0;
  if (lss_bev && aligned && nq % 4 == 0 && E % 4 == 0) {
    const dim3 grid((unsigned)((nq + 127) / 128), (unsigned)((E + 31) / 32),
                    (unsigned)bs);
    bev_query_init_v4_kernel<<<grid, 256, 0, as_stream(stream)>>>(
        embedding, lss_bev, bs, nq, E, out);
    return launch_status();
  }
  const dim3 grid((unsigned)((nq + 31) / 32), (unsigned)((E + 31) / 32),
                  (unsigned)bs);
  bev_query_init_kernel<<<grid, 256, 0, as_stream(stream)>>>(embedding, lss_bev,
                                                              bs, nq, E, out);
  return launch_status();
}

FBBEV_API int fbbev_point_sampling(
    const float* X, const float* Y, const float* Z, int32_t nX, int32_t nY,
    int32_t nZ, const float* inv_bda, const float* trans, const float* ego2cam,
    const float* post_rots, const float* post_trans, int32_t order_flags,
    int32_t B, int32_t N, float w_in, float h_in, float eps,
    float one_minus_eps, float* ref_cam, float* depth, uint8_t* mask,
    fbbev_stream_t stream) {
  if (nX <= 0 || nY <= 0 || nZ <= 0 || B <= 0 || N <= 0)
    return FBBEV_ERR_INVALID_ARGUMENT;
  if (!X || !Y || !Z || !inv_bda || !trans || !ego2cam || !post_rots ||
      !post_trans || !ref_cam || !depth || !mask)
    return FBBEV_ERR_INVALID_ARGUMENT;
  SamplingParams P;
  P.X = X; P.Y = Y; P.Z = Z;
  P.inv_bda = inv_bda; P.trans = trans; P.ego2cam = ego2cam;
  P.post_rots = post_rots; P.post_trans = post_trans;
  P.nX = nX; P.nY = nY; P.nZ = nZ; P.B = B; P.N = N;
  P.order = order_flags;
  P.inv_w = 1.0f / w_in;  // accscalar_t(1.0) / b, rounded once to fp32
  P.inv_h = 1.0f / h_in;
  P.eps = eps;
  P.one_minus_eps = one_minus_eps;
  const int64_t total = (int64_t)N * B * nY * nX * nZ;
  count_launch();
  point_sampling_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0,
                          as_stream(stream)>>>(P, ref_cam, depth, mask);
  return launch_status();
}

// ---------------------------------------------------------------------------
// per_cam_mask & bev_mask for the fused cross-attention
//   DA_SpatialCrossAttention.forward, spatial_cross_attention_depth.py:156-169:
//     per_cam_mask_list_ = per_cam_mask_list & bev_mask[None, :, :, None]
//     index = per_cam_mask_[j].sum(-1).nonzero()
//     if len(index) == 0: index = per_cam_mask_list[i][j].sum(-1).nonzero()[0:1]
// and :213-214  count = per_cam_mask_list_.sum(-1) > 0.
// Output bytes: 1 = anchor visible under the masked list (the query is
// processed for that camera AND counted), 2 = the empty-camera rule: the first
// query the camera sees at all is processed although bev_mask excludes it, and
// is not counted.  The attention kernels test `byte != 0` for visibility and
// `byte & 1` for the count.
// ---------------------------------------------------------------------------
namespace fbbev {

__global__ void bev_mask_fold_init_kernel(int* any_first, int n_pairs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pairs) {
    any_first[i] = 0;
    any_first[n_pairs + i] = INT32_MAX;
  }
}

__global__ void __launch_bounds__(256) bev_mask_fold_kernel(
    const uint8_t* __restrict__ mask, const uint8_t* __restrict__ bev_mask,
    int bs, int nq, int Z, int n_pairs, uint8_t* __restrict__ out,
    int* __restrict__ any_first) {
  const int pair = blockIdx.y;          // n * bs + b
  const int b = pair % bs;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  bool vis = false, kept = false;
  if (q < nq) {
    const int64_t base = ((int64_t)pair * nq + q) * Z;
    const bool bev = __ldg(bev_mask + (int64_t)b * nq + q) != 0;
    for (int z = 0; z < Z; ++z) {
      const bool m = __ldg(mask + base + z) != 0;
      vis |= m;
      out[base + z] = (m && bev) ? 1 : 0;
    }
    kept = vis && bev;
  }
  const unsigned any_kept = __ballot_sync(kFull, kept);
  const int first = __reduce_min_sync(kFull, vis ? q : INT32_MAX);
  if ((threadIdx.x & 31) == 0) {
    if (any_kept) any_first[pair] = 1;                   // benign race: all write 1
    if (first != INT32_MAX) atomicMin(any_first + n_pairs + pair, first);
  }
}

__global__ void bev_mask_fold_fix_kernel(const uint8_t* __restrict__ mask, int nq,
                                         int Z, int n_pairs,
                                         const int* __restrict__ any_first,
                                         uint8_t* __restrict__ out) {
  const int pair = blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= n_pairs) return;
  const int q0 = any_first[n_pairs + pair];
  if (any_first[pair] != 0 || q0 == INT32_MAX) return;
  const int64_t base = ((int64_t)pair * nq + q0) * Z;
  for (int z = 0; z < Z; ++z) out[base + z] = mask[base + z] != 0 ? 2 : 0;
}

}  // namespace fbbev

FBBEV_API size_t fbbev_bev_mask_fold_workspace_bytes(int32_t bs, int32_t n_cams) {
  if (bs <= 0 || n_cams <= 0) return 0;
  return (size_t)bs * n_cams * 2 * sizeof(int);
}

FBBEV_API int fbbev_bev_mask_fold(const uint8_t* mask, const uint8_t* bev_mask,
                                  int32_t bs, int32_t n_cams, int32_t nq,
                                  int32_t Z, uint8_t* mask_out, void* workspace,
                                  size_t workspace_bytes, fbbev_stream_t stream) {
  if (!mask || !bev_mask || !mask_out || !workspace || bs <= 0 || n_cams <= 0 ||
      nq <= 0 || Z <= 0)
    return FBBEV_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < fbbev_bev_mask_fold_workspace_bytes(bs, n_cams))
    return FBBEV_ERR_WORKSPACE_TOO_SMALL;
  const int n_pairs = bs * n_cams;
  if (n_pairs > 65535) return FBBEV_ERR_UNSUPPORTED;
  int* af = static_cast<int*>(workspace);
  cudaStream_t st = as_stream(stream);
  count_launch(3);
  bev_mask_fold_init_kernel<<<(n_pairs + 255) / 256, 256, 0, st>>>(af, n_pairs);
  bev_mask_fold_kernel<<<dim3((nq + 255) / 256, n_pairs), 256, 0, st>>>(
      mask, bev_mask, bs, nq, Z, n_pairs, mask_out, af);
  bev_mask_fold_fix_kernel<<<(n_pairs + 255) / 256, 256, 0, st>>>(
      mask, nq, Z, n_pairs, af, mask_out);
  return launch_status();
}
