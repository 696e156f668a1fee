/*
 * fbbev_oracle.c -- CPU restatement of the FB-BEV / FB-OCC view-transformation
 * hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it; the product
 * path (fb-bev_b200/) never links, imports or calls anything in oracle/.
 *
 * Every function cites the reference file:line (paths relative to the
 * reference checkout, NVlabs/FB-BEV @ 6e25469) whose algorithm it restates.
 *
 * Pinning status:
 *   - voxel_prepare / bev_pool fwd+bwd : pinned against the reference's own
 *     known-answer test (mmdet3d/ops/bev_pool_v2/bev_pool.py:145-176), against
 *     golden vectors produced by importing the reference's Python
 *     (tests/golden/gen_golden.py) and, on the GPU box, against the reference's
 *     own bev_pool_cuda.cu compiled into oracle/_ref/.
 *   - msda_forward / msda_backward : the arithmetic lives in mmcv-full 1.5.2
 *     (`ms_deform_attn_forward`), which is NOT vendored in the reference and is
 *     not installable here -> "parity unpinned" by the reference itself.  The
 *     restatement follows the published Deformable-DETR im2col algorithm and is
 *     cross-checked against torch grid_sample (the formula mmcv documents as
 *     its CPU fallback `multi_scale_deformable_attn_pytorch`) and against the
 *     independent HF transformers implementation.
 *
 * Build: see oracle/Makefile (gcc -O3 -fopenmp, function multi-versioning so
 * the same .so is safe on any x86-64 host and still uses AVX2/AVX-512 when
 * present).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#if defined(__x86_64__) && defined(__GNUC__) && !defined(FBBEV_NO_CLONES)
#define HOT __attribute__((target_clones("default", "avx2,fma", "avx512f")))
#else
#define HOT
#endif

#define API __attribute__((visibility("default")))

API int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

API void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* torch `.long()` on the device the reference runs on (CUDA cvt.rzi.s64.f32):
 * truncation toward zero, NaN -> 0, saturating at the int64 limits. */
static inline int64_t trunc_to_i64(float f) {
  if (isnan(f)) return 0;
  if (f >= 9223372036854775807.0f) return INT64_MAX;
  if (f <= -9223372036854775808.0f) return INT64_MIN;
  return (int64_t)f;
}

/* ------------------------------------------------------------------------
 * voxel_pooling_prepare_v2
 *   reference: mmdet3d/models/fbbev/view_transformation/forward_projection/
 *              view_transformer.py:547-605
 *
 * coor   : (B, N, D, H, W, 3) fp32 ego-frame frustum points (get_lidar_coor)
 * lo, iv : grid lower bound / interval, float32 (create_grid_infos :384-386)
 * gs     : grid size as FLOAT32 (reference keeps it as a float tensor, :387,
 *          and compares integer coordinates against it, :578-580)
 * rank_mode 0: rank in exact int64 arithmetic.
 * rank_mode 1: rank in float32 exactly as the reference does (:586-589,
 *              `long * 0-dim float32 -> float32`), cast to int32 at :603.
 *              Identical to mode 0 while B*Z*Y*X < 2^24.
 * Order inside one voxel: ascending point index (what a stable sort gives;
 * the reference's `argsort()` at :590 leaves it unspecified).
 *
 * Outputs must hold n_points entries each; returns n_kept, n_intervals.
 * ---------------------------------------------------------------------- */
API int oracle_voxel_prepare(const float *coor, int B, int N, int D, int H,
                             int W, const float *lo, const float *iv,
                             const float *gs, int rank_mode,
                             int32_t *ranks_bev, int32_t *ranks_depth,
                             int32_t *ranks_feat, int32_t *interval_starts,
                             int32_t *interval_lengths, int64_t *n_kept_out,
                             int64_t *n_int_out) {
  const int64_t n_pts = (int64_t)B * N * D * H * W;
  const int64_t per_b = (int64_t)N * D * H * W;
  const int64_t hw = (int64_t)H * W;
  const int64_t gx = (int64_t)gs[0], gy = (int64_t)gs[1], gz = (int64_t)gs[2];
  const int64_t n_vox = (int64_t)B * gx * gy * gz;
  *n_kept_out = 0;
  *n_int_out = 0;
  if (n_pts == 0) return 0;

  int64_t *rank = (int64_t *)malloc(sizeof(int64_t) * (size_t)n_pts);
  if (!rank) return -1;
  /* float32 products the reference forms once as 0-dim tensors (:586-588) */
  const float gs210 = (gs[2] * gs[1]) * gs[0];
  const float gs10 = gs[1] * gs[0];
  int64_t max_rank = n_vox;

  for (int64_t p = 0; p < n_pts; ++p) {
    /* :570-572  ((coor - lower) / interval).long() -- fp32 sub, fp32 div */
    float fx = (coor[3 * p + 0] - lo[0]) / iv[0];
    float fy = (coor[3 * p + 1] - lo[1]) / iv[1];
    float fz = (coor[3 * p + 2] - lo[2]) / iv[2];
    int64_t cx = trunc_to_i64(fx), cy = trunc_to_i64(fy), cz = trunc_to_i64(fz);
    int64_t b = p / per_b; /* :573-575 */
    /* :578-580 bounds test; integer coordinate promoted to float32 against
     * the float32 grid_size */
    int keep = (cx >= 0) && ((float)cx < gs[0]) && (cy >= 0) &&
               ((float)cy < gs[1]) && (cz >= 0) && ((float)cz < gs[2]);
    if (!keep) {
      rank[p] = -1;
      continue;
    }
    if (rank_mode == 1) {
      float r = (float)b * gs210;          /* :586-587 */
      r = r + (float)cz * gs10;            /* :588 */
      r = r + ((float)cy * gs[0] + (float)cx); /* :589 */
      rank[p] = (int64_t)(int32_t)r;       /* :603 .int() */
      if (rank[p] + 1 > max_rank) max_rank = rank[p] + 1;
    } else {
      rank[p] = ((b * gz + cz) * gy + cy) * gx + cx;
    }
  }

  /* stable counting sort by rank (== argsort + gather, :590-592) */
  int64_t *count = (int64_t *)calloc((size_t)max_rank + 1, sizeof(int64_t));
  if (!count) {
    free(rank);
    return -1;
  }
  int64_t n_kept = 0;
  for (int64_t p = 0; p < n_pts; ++p)
    if (rank[p] >= 0) {
      count[rank[p] + 1]++;
      n_kept++;
    }
  /* interval list = runs of equal rank (:594-602) */
  int64_t n_int = 0;
  for (int64_t v = 0; v < max_rank; ++v) {
    if (count[v + 1] > 0) {
      interval_starts[n_int] = (int32_t)count[v]; /* count[v] is prefix so far */
      interval_lengths[n_int] = (int32_t)count[v + 1];
      n_int++;
    }
    count[v + 1] += count[v];
  }
  for (int64_t p = 0; p < n_pts; ++p) {
    if (rank[p] < 0) continue;
    int64_t dst = count[rank[p]]++;
    ranks_bev[dst] = (int32_t)rank[p];
    ranks_depth[dst] = (int32_t)p; /* :561-562 arange(num_points) */
    /* :563-568 arange(num_points // D).reshape(B,N,1,H,W).expand(.., D, ..) */
    int64_t bn = p / ((int64_t)D * hw);
    ranks_feat[dst] = (int32_t)(bn * hw + p % hw);
  }
  free(count);
  free(rank);
  *n_kept_out = n_kept;
  *n_int_out = n_int;
  return 0;
}

/* ------------------------------------------------------------------------
 * bev_pool_v2 forward kernel
 *   reference: mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu:18-45
 * out is (B,Z,Y,X,C) and must be zero-filled by the caller (bev_pool.py:25);
 * one plain store per (interval, channel), no accumulation into `out`.
 * The reference compiles `psum += feat * depth` with nvcc's default
 * -fmad=true, i.e. one fused multiply-add per point, in point order.
 * ---------------------------------------------------------------------- */
HOT API void oracle_bev_pool_v2_fwd(int c, int n_intervals, const float *depth,
                                    const float *feat,
                                    const int32_t *ranks_depth,
                                    const int32_t *ranks_feat,
                                    const int32_t *ranks_bev,
                                    const int32_t *interval_starts,
                                    const int32_t *interval_lengths,
                                    float *out) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < n_intervals; ++i) {
    const int s = interval_starts[i];
    const int len = interval_lengths[i];
    float *o = out + (int64_t)ranks_bev[s] * c; /* :42-44 */
    for (int ch = 0; ch < c; ++ch) o[ch] = 0.f;
    for (int k = 0; k < len; ++k) { /* :36-40 */
      const float d = depth[ranks_depth[s + k]];
      const float *f = feat + (int64_t)ranks_feat[s + k] * c;
      for (int ch = 0; ch < c; ++ch) o[ch] = fmaf(f[ch], d, o[ch]);
    }
  }
}

/* (B,Z,Y,X,C) -> (B,C,Z,Y,X): `x.permute(0,4,1,2,3).contiguous()`
 *   reference: mmdet3d/ops/bev_pool_v2/bev_pool.py:89 */
HOT API void oracle_permute_bzyxc_to_bczyx(const float *in, float *out, int B,
                                           int64_t zyx, int c) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int ch = 0; ch < c; ++ch) {
      const float *src = in + (int64_t)b * zyx * c + ch;
      float *dst = out + ((int64_t)b * c + ch) * zyx;
      for (int64_t v = 0; v < zyx; ++v) dst[v] = src[v * c];
    }
}

/* The op exactly as the reference ships it (bev_pool.py:15-39, 84-90):
 * new_zeros (B,Z,Y,X,C) + kernel + permute/contiguous to (B,C,Z,Y,X).
 * `scratch` holds B*zyx*c floats.  This is what bench.py times as the CPU
 * baseline ("port"). */
HOT API void oracle_bev_pool_v2_op(int c, int n_intervals, const float *depth,
                                   const float *feat,
                                   const int32_t *ranks_depth,
                                   const int32_t *ranks_feat,
                                   const int32_t *ranks_bev,
                                   const int32_t *interval_starts,
                                   const int32_t *interval_lengths, int B,
                                   int64_t zyx, float *scratch,
                                   float *out_bczyx) {
  const int64_t n = (int64_t)B * zyx * c;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) scratch[i] = 0.f; /* bev_pool.py:25 */
  oracle_bev_pool_v2_fwd(c, n_intervals, depth, feat, ranks_depth, ranks_feat,
                         ranks_bev, interval_starts, interval_lengths, scratch);
  oracle_permute_bzyxc_to_bczyx(scratch, out_bczyx, B, zyx, c);
}

/* ------------------------------------------------------------------------
 * bev_pool_v2 backward kernel
 *   reference: mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu:64-118
 * Intervals here are runs of equal ranks_feat (bev_pool.py:45-55).
 * depth_grad / feat_grad must be zero-filled by the caller (bev_pool.py:65-66).
 * out_grad is (B,Z,Y,X,C).
 * ---------------------------------------------------------------------- */
HOT API void oracle_bev_pool_v2_bwd(int c, int n_intervals,
                                    const float *out_grad, const float *depth,
                                    const float *feat,
                                    const int32_t *ranks_depth,
                                    const int32_t *ranks_feat,
                                    const int32_t *ranks_bev,
                                    const int32_t *interval_starts,
                                    const int32_t *interval_lengths,
                                    float *depth_grad, float *feat_grad) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < n_intervals; ++i) {
    const int s = interval_starts[i];
    const int len = interval_lengths[i];
    for (int k = 0; k < len; ++k) { /* :88-102 */
      const float *og = out_grad + (int64_t)ranks_bev[s + k] * c;
      const float *f = feat + (int64_t)ranks_feat[s + k] * c;
      float g = 0.f;
      for (int ch = 0; ch < c; ++ch) g = fmaf(og[ch], f[ch], g);
      depth_grad[ranks_depth[s + k]] = g;
    }
    float *fg = feat_grad + (int64_t)ranks_feat[s] * c; /* :104-117 */
    for (int ch = 0; ch < c; ++ch) {
      float g = 0.f;
      for (int k = 0; k < len; ++k)
        g = fmaf(out_grad[(int64_t)ranks_bev[s + k] * c + ch],
                 depth[ranks_depth[s + k]], g);
      fg[ch] = g;
    }
  }
}

/* ------------------------------------------------------------------------
 * Multi-scale deformable attention, forward.
 *   boundary: ext_module.ms_deform_attn_forward as called at
 *     .../bevformer_utils/multi_scale_deformable_attn_function.py:127-133
 *   arithmetic: mmcv-full 1.5.2 (un-vendored) ms_deformable_im2col_gpu_kernel
 *     / ms_deform_attn_im2col_bilinear -- the published Deformable-DETR
 *     algorithm: pixel = loc*size - 0.5, sample iff -1 < h < H and -1 < w < W,
 *     bilinear with zero padding, weighted sum over levels x points.
 * value (bs, n_value, heads, ch); shapes/level_start int64 (L,2)/(L,);
 * loc (bs, nq, heads, L, P, 2) as (x, y); attw (bs, nq, heads, L, P);
 * out (bs, nq, heads*ch).
 * ---------------------------------------------------------------------- */
static inline float msda_bilinear(const float *val, int H, int W, int heads,
                                  int ch, float h, float w, int m, int c) {
  const int h_low = (int)floorf(h), w_low = (int)floorf(w);
  const int h_high = h_low + 1, w_high = w_low + 1;
  const float lh = h - h_low, lw = w - w_low;
  const float hh = 1.f - lh, hw = 1.f - lw;
  const int64_t w_stride = (int64_t)heads * ch;
  const int64_t h_stride = (int64_t)W * w_stride;
  const int64_t base = (int64_t)m * ch + c;
  float v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f;
  if (h_low >= 0 && w_low >= 0) v1 = val[h_low * h_stride + w_low * w_stride + base];
  if (h_low >= 0 && w_high <= W - 1) v2 = val[h_low * h_stride + w_high * w_stride + base];
  if (h_high <= H - 1 && w_low >= 0) v3 = val[h_high * h_stride + w_low * w_stride + base];
  if (h_high <= H - 1 && w_high <= W - 1) v4 = val[h_high * h_stride + w_high * w_stride + base];
  const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
  return w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4;
}

API void oracle_msda_fwd(const float *value, const int64_t *spatial_shapes,
                         const int64_t *level_start, const float *loc,
                         const float *attw, int bs, int n_value, int heads,
                         int ch, int levels, int nq, int points, float *out) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < bs; ++b)
    for (int q = 0; q < nq; ++q)
      for (int m = 0; m < heads; ++m)
        for (int c = 0; c < ch; ++c) {
          float col = 0.f;
          const int64_t wbase = (((int64_t)b * nq + q) * heads + m) * levels * points;
          for (int l = 0; l < levels; ++l) {
            const int H = (int)spatial_shapes[2 * l], W = (int)spatial_shapes[2 * l + 1];
            const float *val = value + ((int64_t)b * n_value + level_start[l]) * heads * ch;
            for (int p = 0; p < points; ++p) {
              const int64_t wi = wbase + (int64_t)l * points + p;
              const float lx = loc[2 * wi], ly = loc[2 * wi + 1];
              const float wgt = attw[wi];
              const float h_im = ly * H - 0.5f, w_im = lx * W - 0.5f;
              if (h_im > -1 && w_im > -1 && h_im < H && w_im < W)
                col += msda_bilinear(val, H, W, heads, ch, h_im, w_im, m, c) * wgt;
            }
          }
          out[((int64_t)b * nq + q) * heads * ch + (int64_t)m * ch + c] = col;
        }
}

/* Multi-scale deformable attention, backward.
 *   boundary: ext_module.ms_deform_attn_backward as called at
 *     .../multi_scale_deformable_attn_function.py:159-169
 * grad_value / grad_loc / grad_attw must be zero-filled by the caller
 * (:155-157).  Serial (accumulates into grad_value). */
API void oracle_msda_bwd(const float *value, const int64_t *spatial_shapes,
                         const int64_t *level_start, const float *loc,
                         const float *attw, const float *grad_out, int bs,
                         int n_value, int heads, int ch, int levels, int nq,
                         int points, float *grad_value, float *grad_loc,
                         float *grad_attw) {
  for (int b = 0; b < bs; ++b)
    for (int q = 0; q < nq; ++q)
      for (int m = 0; m < heads; ++m) {
        const int64_t wbase = (((int64_t)b * nq + q) * heads + m) * levels * points;
        const float *go = grad_out + ((int64_t)b * nq + q) * heads * ch + (int64_t)m * ch;
        for (int l = 0; l < levels; ++l) {
          const int H = (int)spatial_shapes[2 * l], W = (int)spatial_shapes[2 * l + 1];
          const int64_t voff = ((int64_t)b * n_value + level_start[l]) * heads * ch;
          const float *val = value + voff;
          float *gval = grad_value + voff;
          const int64_t w_stride = (int64_t)heads * ch;
          const int64_t h_stride = (int64_t)W * w_stride;
          for (int p = 0; p < points; ++p) {
            const int64_t wi = wbase + (int64_t)l * points + p;
            const float lx = loc[2 * wi], ly = loc[2 * wi + 1];
            const float wgt = attw[wi];
            const float h = ly * H - 0.5f, w = lx * W - 0.5f;
            if (!(h > -1 && w > -1 && h < H && w < W)) continue;
            const int h_low = (int)floorf(h), w_low = (int)floorf(w);
            const int h_high = h_low + 1, w_high = w_low + 1;
            const float lh = h - h_low, lw = w - w_low;
            const float hh = 1.f - lh, hw = 1.f - lw;
            const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
            float g_w = 0.f, g_h = 0.f, g_a = 0.f;
            for (int c = 0; c < ch; ++c) {
              const int64_t base = (int64_t)m * ch + c;
              const float top = go[c] * wgt;
              float v1 = 0, v2 = 0, v3 = 0, v4 = 0;
              if (h_low >= 0 && w_low >= 0) {
                const int64_t o = h_low * h_stride + w_low * w_stride + base;
                v1 = val[o]; gval[o] += w1 * top;
              }
              if (h_low >= 0 && w_high <= W - 1) {
                const int64_t o = h_low * h_stride + w_high * w_stride + base;
                v2 = val[o]; gval[o] += w2 * top;
              }
              if (h_high <= H - 1 && w_low >= 0) {
                const int64_t o = h_high * h_stride + w_low * w_stride + base;
                v3 = val[o]; gval[o] += w3 * top;
              }
              if (h_high <= H - 1 && w_high <= W - 1) {
                const int64_t o = h_high * h_stride + w_high * w_stride + base;
                v4 = val[o]; gval[o] += w4 * top;
              }
              g_h += (-hw * v1 - lw * v2 + hw * v3 + lw * v4) * top;
              g_w += (-hh * v1 + hh * v2 - lh * v3 + lh * v4) * top;
              g_a += (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4) * go[c];
            }
            grad_attw[wi] = g_a;
            grad_loc[2 * wi] = W * g_w;
            grad_loc[2 * wi + 1] = H * g_h;
          }
        }
      }
}
