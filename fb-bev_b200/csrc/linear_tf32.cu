// linear_tf32.cu -- the row-wise Linears of the backward projection on the
// 5th-generation tensor cores (tcgen05, accumulators in TMEM), fp32 in / fp32 out.
//
// What it replaces: every nn.Linear of the BEVFormer encoder layer the reference
// builds for the backward projection -- sampling_offsets / attention_weights /
// value_proj / output_proj of the self- and cross-attention
// (spatial_cross_attention_depth.py:420-427, mmcv MultiScaleDeformableAttention)
// and the FFN (bevformer_encoder.py:251-377 op order) -- together with what
// follows each of them row by row: bias, ReLU, the residual add and the
// LayerNorm(embed_dims).  The reference runs them as cuBLAS fp32 GEMMs plus
// separate elementwise / LayerNorm kernels.
//
// Numerics: the 1e-4 parity bar excludes plain TF32 (10-bit mantissa), so every
// product is formed as 3xTF32:  x = x_hi + x_lo, w = w_hi + w_lo with *_hi the
// upper 19 bits and *_lo the exact fp32 remainder;
//     x.w ~= x_hi.w_hi + x_lo.w_hi + x_hi.w_lo        (error ~2^-21 relative)
// three tcgen05.mma.kind::tf32 per k-step into the same fp32 TMEM accumulator.
//
// Structure (one persistent CTA per SM, 128-row tiles, 9 warps):
//   warps 4-7  loaders: X rows -> registers -> hi / lo split -> shared memory in
//              the UMMA canonical K-major layout (8x16-byte core matrices, no
//              swizzle); one lane also starts the 1-D bulk copy (TMA) of the
//              pre-packed weight block of the stage;
//   warp  8    one lane issues the MMAs (M=128, N<=160, K=8 per instruction),
//              tcgen05.commit releases the stage / publishes the accumulator;
//   warps 0-3  epilogue: tcgen05.ld (lane == row), bias / ReLU / residual /
//              LayerNorm in registers, row stores.
//   Ring of K-blocks of 40 floats (5 k-steps) between loaders and MMA, two
//   accumulators in TMEM between MMA and epilogue.
#include "common.cuh"

namespace fbbev {

constexpr int kKB = 40;                          // floats of K per stage
constexpr int kChunks = kKB / 4;                 // 16-byte chunks per row
constexpr int kTileM = 128;
constexpr int kAPart = kTileM * kKB * 4;         // bytes of A_hi (== A_lo)
constexpr int kAChunkStride = (kTileM / 8) * 128;  // bytes between K chunks
constexpr int kLinThreads = 288;
constexpr int kMaxN = 160;
constexpr int kTmemCols = 512;
constexpr int kSmemLimit = 232448 - 1024;

// ------------------------------- PTX wrappers --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t tx) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
               "r"(tx)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spins > (1u << 26)) __trap();  // a protocol bug must not hang the GPU
  }
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src,
                                         uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 "
      "[%0];" ::"r"(bar)
      : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, tf32 operands, fp32 accumulate
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major operand without swizzle: 8-row x 16-byte core matrices; `lbo` bytes
// between the two K chunks of one MMA, `sbo` bytes between 8-row groups
// (cute::UMMA::SmemDescriptor, version 1, LayoutType::SWIZZLE_NONE).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo,
                                              uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) |
         ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]),
        "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
        "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------ weight packing -------------------------------
// W [N][K] row-major -> [k-block][hi, lo][chunk 0..9][Npad / 8][8 rows][4]:
// exactly the shared-memory image of one stage, so a stage's weights are one
// contiguous bulk copy.  Rows >= N and columns >= K are zero.
__global__ void linear_pack_kernel(const float* __restrict__ w, int n, int k,
                                   int npad, int n_kb, float* __restrict__ out) {
  const int64_t total = (int64_t)n_kb * kChunks * npad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(i % npad);
    const int ch = (int)((i / npad) % kChunks);
    const int kb = (int)(i / ((int64_t)npad * kChunks));
    float hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = kb * kKB + ch * 4 + j;
      const float v = (row < n && col < k) ? w[(int64_t)row * k + col] : 0.f;
      hi[j] = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
      lo[j] = v - hi[j];
    }
    const int64_t part = (int64_t)kChunks * npad * 4;  // floats of one hi / lo part
    const int64_t off =
        (int64_t)kb * 2 * part + (int64_t)ch * npad * 4 + (row / 8) * 32 + (row % 8) * 4;
    *reinterpret_cast<float4*>(out + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<float4*>(out + off + part) =
        make_float4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// -------------------------------- the kernel ---------------------------------
struct LinearParams {
  const float* x;         // [M][K]
  const float* wp;        // packed weights
  const float* bias;      // [N] or null
  const float* residual;  // [M][N] or null
  const float* gamma;     // LayerNorm weight [N] or null
  const float* beta;      // LayerNorm bias [N] or null
  float* y;               // [M][N]
  int M, K, N, npad, n_kb, relu, n_tiles, stages;
  int64_t ldx, ldr, ldy;  // row strides in floats
  float eps;
};

template <int NC16, bool LN>
__global__ void __launch_bounds__(kLinThreads, 1)
    linear_tf32_kernel(const LinearParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int npad = p.npad, S = p.stages;
  const uint32_t w_part = (uint32_t)npad * kKB * 4;        // bytes of W_hi
  const uint32_t stage_bytes = 2u * kAPart + 2u * w_part;
  unsigned char* ctrl = smem + (size_t)S * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ctrl);
  const uint32_t bar_full = smem_u32(bars);                // [S]
  const uint32_t bar_empty = bar_full + 8u * S;            // [S]
  const uint32_t bar_tfull = bar_empty + 8u * S;           // [2]
  const uint32_t bar_tempty = bar_tfull + 16u;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctrl + 16 * S + 32);
  const uint32_t smem_base = smem_u32(smem);

  if (warp == 8) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        mbar_init(bar_full + 8u * s, 128 + 1);  // loader threads + expect_tx
        mbar_init(bar_empty + 8u * s, 1);       // tcgen05.commit
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(bar_tfull + 8u * a, 1);       // tcgen05.commit
        mbar_init(bar_tempty + 8u * a, 128);    // epilogue threads
      }
      fence_mbar_init();
    }
    __syncwarp();
    asm volatile(
        "tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
            smem_u32(tmem_slot)),
        "r"(kTmemCols)
        : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::
                     : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 4 && warp < 8) {
    // ================================ loaders ================================
    const int lw = warp - 4;
    const int lt = threadIdx.x - 128;
    const int r_lo = lane & 15, c_lo = lane >> 4;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int row0 = tile * kTileM;
      for (int kb = 0; kb < p.n_kb; ++kb, ++it) {
        const uint32_t s = it % S, ph = (it / S) & 1u;
        mbar_wait(bar_empty + 8u * s, ph ^ 1u);
        const uint32_t st = smem_base + s * stage_bytes;
        if (lt == 0) {
          mbar_arrive_expect_tx(bar_full + 8u * s, 2u * w_part);
          bulk_g2s(st + 2u * kAPart, p.wp + (size_t)kb * (2u * w_part / 4u),
                   2u * w_part, bar_full + 8u * s);
        }
        float4 v[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) {
          const int idx = lw * 10 + i;
          const int row = (idx / 5) * 16 + r_lo;
          const int ch = (idx % 5) * 2 + c_lo;
          const int grow = min(row0 + row, p.M - 1);
          const int col = kb * kKB + ch * 4;
          v[i] = col < p.K ? __ldg(reinterpret_cast<const float4*>(
                                 p.x + (size_t)grow * p.ldx + col))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 10; ++i) {
          const int idx = lw * 10 + i;
          const int row = (idx / 5) * 16 + r_lo;
          const int ch = (idx % 5) * 2 + c_lo;
          float4 hi, lo;
          hi.x = __uint_as_float(__float_as_uint(v[i].x) & 0xFFFFE000u);
          hi.y = __uint_as_float(__float_as_uint(v[i].y) & 0xFFFFE000u);
          hi.z = __uint_as_float(__float_as_uint(v[i].z) & 0xFFFFE000u);
          hi.w = __uint_as_float(__float_as_uint(v[i].w) & 0xFFFFE000u);
          lo.x = v[i].x - hi.x; lo.y = v[i].y - hi.y;
          lo.z = v[i].z - hi.z; lo.w = v[i].w - hi.w;
          const uint32_t off =
              (uint32_t)(ch * (kTileM / 8) + (row >> 3)) * 128u + (row & 7) * 16u;
          unsigned char* a = smem + (size_t)s * stage_bytes + off;
          *reinterpret_cast<float4*>(a) = hi;
          *reinterpret_cast<float4*>(a + kAPart) = lo;
        }
        fence_proxy_async();
        mbar_arrive(bar_full + 8u * s);
      }
    }
  } else if (warp == 8) {
    // ================================ MMA issue ===============================
    if (lane == 0) {
      // kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = npad
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) |
                             ((uint32_t)(npad >> 3) << 17) | (8u << 24);
      const uint32_t lbo_b = (uint32_t)npad * 16u;
      uint32_t it = 0, tc = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tc) {
        const uint32_t acc = tc & 1u, aph = (tc >> 1) & 1u;
        mbar_wait(bar_tempty + 8u * acc, aph ^ 1u);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * (uint32_t)npad;
        for (int kb = 0; kb < p.n_kb; ++kb, ++it) {
          const uint32_t s = it % S, ph = (it / S) & 1u;
          mbar_wait(bar_full + 8u * s, ph);
          tc_fence_after();
          const uint32_t a_hi = smem_base + s * stage_bytes;
          const uint32_t a_lo = a_hi + kAPart;
          const uint32_t w_hi = a_lo + kAPart;
          const uint32_t w_lo = w_hi + w_part;
#pragma unroll
          for (int k = 0; k < kChunks / 2; ++k) {
            const uint32_t ao = 2u * k * kAChunkStride, bo = 2u * k * lbo_b;
            const uint64_t dah = smem_desc(a_hi + ao, kAChunkStride, 128);
            const uint64_t dal = smem_desc(a_lo + ao, kAChunkStride, 128);
            const uint64_t dbh = smem_desc(w_hi + bo, lbo_b, 128);
            const uint64_t dbl = smem_desc(w_lo + bo, lbo_b, 128);
            mma_tf32(d, dah, dbh, idesc, (kb | k) != 0);
            mma_tf32(d, dal, dbh, idesc, 1u);
            mma_tf32(d, dah, dbl, idesc, 1u);
          }
          tc_commit(bar_empty + 8u * s);
        }
        tc_commit(bar_tfull + 8u * acc);
      }
    }
    __syncwarp();
  } else {
    // ================================ epilogue ================================
    const int N = p.N;
    uint32_t tc = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tc) {
      const uint32_t acc = tc & 1u, aph = (tc >> 1) & 1u;
      mbar_wait(bar_tfull + 8u * acc, aph);
      tc_fence_after();
      const uint32_t taddr =
          tmem_base + ((uint32_t)(warp * 32) << 16) + acc * (uint32_t)npad;
      float v[NC16 * 16];
#pragma unroll
      for (int c = 0; c < NC16; ++c) tmem_ld16(taddr + 16u * c, v + 16 * c);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar_tempty + 8u * acc);

      const int grow = tile * kTileM + warp * 32 + lane;
      if (grow < p.M) {
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < NC16 * 16; ++j)
            if (j < N) v[j] += __ldg(p.bias + j);
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < NC16 * 16; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (p.residual) {
          const float4* r =
              reinterpret_cast<const float4*>(p.residual + (size_t)grow * p.ldr);
#pragma unroll
          for (int j = 0; j < NC16 * 4; ++j)
            if (4 * j < N) {
              const float4 t = __ldg(r + j);
              v[4 * j] += t.x; v[4 * j + 1] += t.y;
              v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
            }
        }
        if (LN) {
          float mean = 0.f;
#pragma unroll
          for (int j = 0; j < NC16 * 16; ++j)
            if (j < N) mean += v[j];
          mean /= (float)N;
          float var = 0.f;
#pragma unroll
          for (int j = 0; j < NC16 * 16; ++j)
            if (j < N) {
              const float dlt = v[j] - mean;
              var = fmaf(dlt, dlt, var);
            }
          const float rstd = rsqrtf(var / (float)N + p.eps);
#pragma unroll
          for (int j = 0; j < NC16 * 16; ++j)
            if (j < N)
              v[j] = fmaf((v[j] - mean) * rstd, __ldg(p.gamma + j),
                          __ldg(p.beta + j));
        }
        float4* o = reinterpret_cast<float4*>(p.y + (size_t)grow * p.ldy);
#pragma unroll
        for (int j = 0; j < NC16 * 4; ++j)
          if (4 * j < N)
            o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(
                     tmem_base),
                 "r"(kTmemCols)
                 : "memory");
  }
}

static inline int pad16(int n) { return (n + 15) / 16 * 16; }
static inline int n_kblocks(int k) { return (k + kKB - 1) / kKB; }

template <int NC16>
static int launch_linear(const LinearParams& p, bool ln, size_t smem, int grid,
                         cudaStream_t st) {
  auto kern = ln ? linear_tf32_kernel<NC16, true> : linear_tf32_kernel<NC16, false>;
  cudaError_t e = cudaFuncSetAttribute(
      kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<grid, kLinThreads, smem, st>>>(p);
  return launch_status();
}

}  // namespace fbbev

using namespace fbbev;

FBBEV_API size_t fbbev_linear_packed_bytes(int32_t n, int32_t k) {
  if (n <= 0 || k <= 0) return 0;
  return (size_t)n_kblocks(k) * 2 * kKB * pad16(n) * sizeof(float);
}

FBBEV_API int fbbev_linear_pack(const float* weight, int32_t n, int32_t k,
                                float* packed, fbbev_stream_t stream) {
  if (!weight || !packed || n <= 0 || k <= 0) return FBBEV_ERR_INVALID_ARGUMENT;
  if (reinterpret_cast<uintptr_t>(packed) & 15) return FBBEV_ERR_INVALID_ARGUMENT;
  const int npad = pad16(n), nkb = n_kblocks(k);
  const int64_t total = (int64_t)nkb * kChunks * npad;
  count_launch();
  linear_pack_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0,
                       as_stream(stream)>>>(weight, n, k, npad, nkb, packed);
  return launch_status();
}

FBBEV_API int fbbev_linear_fwd(const float* x, int64_t ldx, const float* packed,
                               const float* bias, const float* residual,
                               int64_t ldr, const float* ln_weight,
                               const float* ln_bias, int64_t m, int32_t k,
                               int32_t n, int32_t relu, float ln_eps, float* y,
                               int64_t ldy, fbbev_stream_t stream) {
  if (!x || !packed || !y || m < 0 || k <= 0 || n <= 0)
    return FBBEV_ERR_INVALID_ARGUMENT;
  if ((ln_weight == nullptr) != (ln_bias == nullptr))
    return FBBEV_ERR_INVALID_ARGUMENT;
  if (k % 4 != 0 || n % 4 != 0 || n > kMaxN || m > (int64_t)1 << 30)
    return FBBEV_ERR_UNSUPPORTED;
  if (ldx < k || ldy < n || (residual && ldr < n) || ldx % 4 || ldy % 4 ||
      (residual && ldr % 4))
    return FBBEV_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
       reinterpret_cast<uintptr_t>(packed) |
       reinterpret_cast<uintptr_t>(residual)) & 15)
    return FBBEV_ERR_INVALID_ARGUMENT;
  if (m == 0) return FBBEV_OK;
  LinearParams p;
  p.x = x; p.wp = packed; p.bias = bias; p.residual = residual;
  p.gamma = ln_weight; p.beta = ln_bias; p.y = y;
  p.M = (int)m; p.K = k; p.N = n; p.npad = pad16(n); p.n_kb = n_kblocks(k);
  p.relu = relu; p.eps = ln_eps;
  p.ldx = ldx; p.ldr = ldr; p.ldy = ldy;
  p.n_tiles = (int)ceil_div64(m, kTileM);
  const size_t stage = 2 * (size_t)kAPart + 2 * (size_t)p.npad * kKB * 4;
  int stages = (int)((kSmemLimit - 256) / stage);
  stages = stages > 4 ? 4 : stages;
  if (stages < 2) return FBBEV_ERR_UNSUPPORTED;
  p.stages = stages;
  const size_t smem = stages * stage + 256;
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
  }
  const int grid = p.n_tiles < n_sm ? p.n_tiles : n_sm;
  const bool ln = ln_weight != nullptr;
  cudaStream_t st = as_stream(stream);
  count_launch();
  switch (p.npad / 16) {
    case 1: return launch_linear<1>(p, ln, smem, grid, st);
    case 2: return launch_linear<2>(p, ln, smem, grid, st);
    case 3: return launch_linear<3>(p, ln, smem, grid, st);
    case 4: return launch_linear<4>(p, ln, smem, grid, st);
    case 5: return launch_linear<5>(p, ln, smem, grid, st);
    case 6: return launch_linear<6>(p, ln, smem, grid, st);
    case 7: return launch_linear<7>(p, ln, smem, grid, st);
    case 8: return launch_linear<8>(p, ln, smem, grid, st);
    case 9: return launch_linear<9>(p, ln, smem, grid, st);
    case 10: return launch_linear<10>(p, ln, smem, grid, st);
    default: return FBBEV_ERR_UNSUPPORTED;
  }
}
