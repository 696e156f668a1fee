// bev_pool_split.cu -- dense lift-splat pooling as two lean kernels.
//
// Evidence behind the structure (profiles/):
//  * tools/micro/store_pattern.cu: writing the (B,C,Z,Y,X) volume as per-tile
//    channel rows (128-bit stores, 512 B per row and tile) runs at memset speed
//    on B200 (35 us for 204.8 MB) -- the output pattern is not the limit;
//  * every fused single-kernel variant measured (one tile per CTA, persistent,
//    cp.async-pipelined, TMA-store) stayed at 60-100 us with the SM issue slots
//    ~55 % busy (37-51 M warp instructions) and long-scoreboard / barrier stalls:
//    index chasing (tile table -> index words -> depth / feat rows) inside the
//    CTA that owns the 40 KB output tile starves the store stream.
// So the index chasing is moved out of the CTA that owns an output tile, and
// both kinds of work run in ONE launch as producer / consumer CTAs:
//
//  producers (blockIdx < n_sum_ctas) -- "interval sums": warp w folds the 32
//      kept points [32w, 32w+32) -- perfectly balanced, so the dense voxels next
//      to a camera (up to 63 points on the 200x200x16 grid, thousands on the
//      1-camera 128x128 grid: the reference kernel's and every tile-owning
//      kernel's tail) are spread over many warps.  Coalesced index loads; every
//      run of equal voxel rank is folded by a 4-lane group (8 runs per warp
//      instruction) with 128-bit feat loads and one FMA per point and channel
//      in point order (bev_pool_cuda.cu:36-40).  Interval sums go to compact
//      rows V[interval][C] (43 MB for the 200x200x16 grid; they never leave
//      L2), the part of an interval that spills into later slices to carry rows
//      X[slice][C].  A producer CTA publishes a flag when its rows are written.
//  consumers (the other CTAs) -- "dense write": a CTA owns T consecutive voxel
//      ranks x all C channels.  Empty tiles stream zeros at once; the others
//      wait for the flags of the (one or two) producers that cover their
//      points, copy the rows of their contiguous intervals V[i0:i1] into shared
//      memory with cp.async, add carry rows in slice order, and stream the tile
//      out channel row by channel row with 128-bit evict-first stores.
//  Producers have the lowest block indices, so they are resident before any
//  consumer can wait on them; consumers never block producers.
//  Every output element and every V / X row is written exactly once; no
//  floating-point atomics; results are deterministic.
//
// Requires C % 4 == 0 and (Z*Y*X) % 4 == 0, 16-byte aligned out.
#include <algorithm>
#include <cstdlib>

#include "bev_pool_split.h"

namespace fbbev {

constexpr int kPoolThreads = 256;  // both roles
constexpr int kSumThreads = kPoolThreads;
constexpr int kWrThreads = kPoolThreads;
constexpr int kWrWarps = kWrThreads / kWarp;

__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(sdst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ------------------------------- plan --------------------------------------
// tile_first[t] = first interval of tile t (tile_first[n_tiles] = n);
// seg_rank[i]   = voxel rank of interval i;
// warp_first[w] = first interval starting at or after point 32*w: K1's warp w
//                 owns intervals [warp_first[w], warp_first[w+1]) -- whole
//                 intervals only, about kPtsPerWarp points each, so a dense
//                 region near a camera is spread over many warps instead of
//                 serialising one (the reference kernel's and every
//                 tile-owning kernel's tail, see profiles/);
// meta = {n_intervals, n_kept_points, n_warps}.
constexpr int kPtsPerWarp = 32;

__global__ void split_plan_kernel(
    const int* __restrict__ ranks_bev, const int* __restrict__ interval_starts,
    const int* __restrict__ interval_lengths, int n_intervals_max,
    const int* __restrict__ n_intervals_dev, int n_warps_max, int64_t zyx,
    int tiles_per_b, int T, int64_t n_tiles, int* __restrict__ tile_first,
    int* __restrict__ seg_rank, int* __restrict__ warp_first,
    int* __restrict__ meta) {
  const int n = n_intervals_dev ? min(*n_intervals_dev, n_intervals_max)
                                : n_intervals_max;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t gsz = (int64_t)gridDim.x * blockDim.x;
  if (n <= 0) {
    if (gid == 0) meta[0] = meta[1] = meta[2] = 0;
    for (int64_t t = gid; t <= n_tiles; t += gsz) tile_first[t] = 0;
    return;
  }
  const int n_kept = interval_starts[n - 1] + interval_lengths[n - 1];
  const int n_warps = min((n_kept + kPtsPerWarp - 1) / kPtsPerWarp, n_warps_max);
  if (gid == 0) {
    meta[0] = n;
    meta[1] = n_kept;
    meta[2] = n_warps;
  }
  auto tile_of = [&](int64_t rank) {
    const int64_t b = rank / zyx;
    const int64_t t = b * tiles_per_b + (rank - b * zyx) / T;
    return max((int64_t)0, min(t, n_tiles - 1));
  };
  for (int64_t i = gid; i < n; i += gsz) {
    const int rank = ranks_bev[interval_starts[i]];
    seg_rank[i] = rank;
    const int64_t t = tile_of(rank);
    const int64_t tp =
        i > 0 ? tile_of(ranks_bev[interval_starts[i - 1]]) : (int64_t)-1;
    for (int64_t u = tp + 1; u <= t; ++u) tile_first[u] = (int)i;
    if (i == n - 1)
      for (int64_t u = t + 1; u <= n_tiles; ++u) tile_first[u] = n;
  }
  for (int64_t w = gid; w <= n_warps; w += gsz) {
    int lo = 0, hi = n;  // lower_bound(interval_starts, 32*w)
    const int key = (int)w * kPtsPerWarp;
    if (w == n_warps) lo = n;
    else
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (interval_starts[mid] < key) lo = mid + 1; else hi = mid;
      }
    warp_first[w] = lo;
  }
}

// --------------------------- K1: interval sums -----------------------------
// Warp w folds exactly the 32 kept points [32w, 32w+32): perfectly balanced, so
// the dense voxels next to a camera (up to 63 points each on the 200x200x16
// grid, thousands on the 1-camera 128x128 grid) are spread over many warps.
// Two dependent loads reach the data (plan table -> index words, coalesced).
// Inside the slice every run of equal voxel rank (a segment == an interval or
// a piece of one) is folded by an LG-lane group -- 8 segments per warp
// instruction for LG = 4 -- with 128-bit feat loads and one FMA per point and
// channel in point order (the reference's order, bev_pool_cuda.cu:36-40).
// An interval that STARTS in the slice is stored to its row V[interval]; the
// leading part of an interval that started in an earlier slice goes to the
// slice's carry row X[w] and is added, in slice order, by K2.  Every row is
// written exactly once; no atomics; deterministic.
template <int LG, int VPL>
__device__ __forceinline__ void interval_sums_role(
    int cta, const float* __restrict__ depth, const float* __restrict__ feat,
    const int* __restrict__ ranks_depth, const int* __restrict__ ranks_feat,
    const int* __restrict__ ranks_bev, const int* __restrict__ interval_starts,
    const int* __restrict__ warp_first, const int* __restrict__ meta, int c,
    float* __restrict__ V, float* __restrict__ X) {
  constexpr int WPC = kSumThreads / kWarp;
  constexpr int GPW = kWarp / LG;  // segments folded per warp instruction
  __shared__ int s_k0[WPC][kPtsPerWarp + 1];
  __shared__ int s_rf[WPC][kPtsPerWarp];
  __shared__ float s_d[WPC][kPtsPerWarp];
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const int64_t w = (int64_t)cta * WPC + wi;
  if (w >= meta[2]) return;
  const int n = meta[0];
  const int base = (int)w * kPtsPerWarp;
  const int cnt = min(kPtsPerWarp, meta[1] - base);
  const int lb = __ldg(warp_first + w);
  int rbL = -1, rfL = 0;
  float dL = 0.f;
  if (lane < cnt) {
    rbL = __ldg(ranks_bev + base + lane);
    rfL = __ldg(ranks_feat + base + lane);
    dL = __ldg(depth + __ldg(ranks_depth + base + lane));
  }
  // does the slice begin inside an interval that started earlier?
  const int carry = !(lb < n && __ldg(interval_starts + lb) == base);
  const int prev = __shfl_up_sync(kFull, rbL, 1);
  const bool is_start = lane < cnt && (lane == 0 || rbL != prev);
  const unsigned starts = __ballot_sync(kFull, is_start);
  const int nseg = __popc(starts);
  s_rf[wi][lane] = rfL;
  s_d[wi][lane] = dL;
  if (is_start) s_k0[wi][__popc(starts & ((1u << lane) - 1u))] = lane;
  if (lane == 0) s_k0[wi][nseg] = cnt;
  __syncwarp();

  const int gl = lane % LG, g = lane / LG;
  const int c4 = c >> 2;
  const float4* feat4 = reinterpret_cast<const float4*>(feat);
  for (int j = g; j < nseg; j += GPW) {
    const int k0 = s_k0[wi][j], k1 = s_k0[wi][j + 1];
    float4 acc[VPL];
#pragma unroll
    for (int q = 0; q < VPL; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = k0; k < k1; ++k) {
      const float d = s_d[wi][k];
      const float4* f = feat4 + (int64_t)s_rf[wi][k] * c4;
#pragma unroll
      for (int q = 0; q < VPL; ++q) {
        const int vi = gl + LG * q;
        if (vi < c4) {
          const float4 x = __ldg(f + vi);
          acc[q].x = fmaf(x.x, d, acc[q].x);
          acc[q].y = fmaf(x.y, d, acc[q].y);
          acc[q].z = fmaf(x.z, d, acc[q].z);
          acc[q].w = fmaf(x.w, d, acc[q].w);
        }
      }
    }
    float4* dst = (j == 0 && carry)
                      ? reinterpret_cast<float4*>(X) + w * c4
                      : reinterpret_cast<float4*>(V) +
                            (int64_t)(lb + j - carry) * c4;
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
      const int vi = gl + LG * q;
      if (vi < c4) dst[vi] = acc[q];
    }
  }
}

// ---------------------------- K2: dense write ------------------------------
// smem: rows[<=T][c + 4] | slot[T] (0 = empty voxel, else row + 1)
//       | carry_lo[T], carry_n[T] (carry rows X[lo+1 .. lo+n] of each interval)
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int T>
__device__ __forceinline__ void dense_write_role(
    int tile, const float* __restrict__ V, const float* __restrict__ X,
    const int* __restrict__ tile_first, const int* __restrict__ seg_rank,
    const int* __restrict__ interval_starts,
    const int* __restrict__ interval_lengths, const int* __restrict__ meta,
    const int* __restrict__ flags, int c, int64_t zyx, int tiles_per_b,
    float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int pitch = c + 4;  // 16-byte aligned rows
  float* rows = reinterpret_cast<float*>(smem_raw);              // [T][pitch]
  int* slot = reinterpret_cast<int*>(rows + (size_t)T * pitch);  // [T]
  int* carry_lo = slot + T;
  int* carry_n = carry_lo + T;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = tile / tiles_per_b;
  const int64_t v0 = (int64_t)(tile - b * tiles_per_b) * T;
  const int nv = (int)min((int64_t)T, zyx - v0);
  const int64_t rank0 = (int64_t)b * zyx + v0;
  const int i0 = __ldg(tile_first + tile);
  const int i1 = __ldg(tile_first + tile + 1);
  const int nrows = min(i1 - i0, T);
  const int c4 = c >> 2;
  float* obase = out + (int64_t)b * c * zyx + v0;
  constexpr int LPR = T / 4;        // lanes per channel row
  constexpr int RPW = kWarp / LPR;  // rows per warp instruction
  const int g = lane % LPR;
  float* o = obase + 4 * g + (int64_t)(warp * RPW + lane / LPR) * zyx;
  const int64_t step = (int64_t)kWrWarps * RPW * zyx;

  if (nrows <= 0) {  // empty tile: pure zero stream
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (4 * g < nv)
      for (int row = warp * RPW + lane / LPR; row < c;
           row += kWrWarps * RPW, o += step)
        st_stream(reinterpret_cast<float4*>(o), z);
    return;
  }
  // wait for the producers that fold this tile's points
  if (tid == 0) {
    const int pa = __ldg(interval_starts + i0);
    const int pb = i1 < meta[0] ? __ldg(interval_starts + i1) : meta[1];
    constexpr int kPtsPerCta = kPtsPerWarp * (kSumThreads / kWarp);
    for (int f = pa / kPtsPerCta; f <= (pb - 1) / kPtsPerCta; ++f)
      while (ld_acquire(flags + f) == 0) __nanosleep(64);
  }
  __syncthreads();
  // rows of this tile's intervals: contiguous in V, copied asynchronously
  {
    const float4* src = reinterpret_cast<const float4*>(V) + (int64_t)i0 * c4;
    const int total = nrows * c4;
    for (int q = tid; q < total; q += kWrThreads) {
      const int r = q / c4, v = q - r * c4;
      cp_async16(rows + (size_t)r * pitch + 4 * v, src + q);
    }
  }
  for (int q = tid; q < T; q += kWrThreads) slot[q] = 0;
  __syncthreads();
  int any_carry = 0;
  for (int r = tid; r < nrows; r += kWrThreads) {
    const int64_t vl = (int64_t)__ldg(seg_rank + i0 + r) - rank0;
    if (vl >= 0 && vl < nv) slot[vl] = r + 1;
    const int st = __ldg(interval_starts + i0 + r);
    const int ln = __ldg(interval_lengths + i0 + r);
    const int lo = st / kPtsPerWarp;
    const int nx = (st + ln - 1) / kPtsPerWarp - lo;  // carry rows to add
    carry_lo[r] = lo;
    carry_n[r] = nx;
    any_carry |= nx > 0;
  }
  cp_async_commit_wait_all();
  if (__syncthreads_or(any_carry)) {
    // add the carry rows of intervals that span several K1 slices, in slice
    // order (deterministic); 8 loads in flight per thread
    const float4* X4 = reinterpret_cast<const float4*>(X);
    const int total = nrows * c4;
    for (int q = tid; q < total; q += kWrThreads) {
      const int r = q / c4, v = q - r * c4;
      const int nx = carry_n[r];
      if (nx == 0) continue;
      float4* dst = reinterpret_cast<float4*>(rows + (size_t)r * pitch + 4 * v);
      float4 a = *dst;
      const float4* src = X4 + (int64_t)(carry_lo[r] + 1) * c4 + v;
      for (int k = 0; k < nx; k += 8) {
        float4 t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          t[j] = (k + j < nx) ? __ldcg(src + (int64_t)(k + j) * c4)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (k + j < nx) {
            a.x += t[j].x; a.y += t[j].y; a.z += t[j].z; a.w += t[j].w;
          }
        }
      }
      *dst = a;
    }
    __syncthreads();
  }

  if (4 * g < nv) {
    const int4 s4 = *reinterpret_cast<const int4*>(slot + 4 * g);
    const bool gany = (s4.x | s4.y | s4.z | s4.w) != 0;
    const float* rx = rows + (size_t)(s4.x - 1) * pitch;
    const float* ry = rows + (size_t)(s4.y - 1) * pitch;
    const float* rz = rows + (size_t)(s4.z - 1) * pitch;
    const float* rw = rows + (size_t)(s4.w - 1) * pitch;
    for (int row = warp * RPW + lane / LPR; row < c;
         row += kWrWarps * RPW, o += step) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gany) {
        if (s4.x) v.x = rx[row];
        if (s4.y) v.y = ry[row];
        if (s4.z) v.z = rz[row];
        if (s4.w) v.w = rw[row];
      }
      st_stream(reinterpret_cast<float4*>(o), v);
    }
  }
}

// ------------------------- one launch, two roles ---------------------------
template <int T, int LG, int VPL>
__global__ void __launch_bounds__(kPoolThreads) pool_split_kernel(
    const float* __restrict__ depth, const float* __restrict__ feat,
    const int* __restrict__ ranks_depth, const int* __restrict__ ranks_feat,
    const int* __restrict__ ranks_bev, const int* __restrict__ interval_starts,
    const int* __restrict__ interval_lengths,
    const int* __restrict__ tile_first, const int* __restrict__ seg_rank,
    const int* __restrict__ warp_first, const int* __restrict__ meta,
    int* __restrict__ flags, int n_sum_ctas, int c, int64_t zyx,
    int tiles_per_b, float* __restrict__ V, float* __restrict__ X,
    float* __restrict__ out) {
  if ((int)blockIdx.x < n_sum_ctas) {
    interval_sums_role<LG, VPL>(blockIdx.x, depth, feat, ranks_depth,
                                ranks_feat, ranks_bev, interval_starts,
                                warp_first, meta, c, V, X);
    // publish: every thread's row stores, then the flag
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flags + blockIdx.x),
                   "r"(1)
                   : "memory");
    }
    return;
  }
  dense_write_role<T>(blockIdx.x - n_sum_ctas, V, X, tile_first, seg_rank,
                      interval_starts, interval_lengths, meta, flags, c, zyx,
                      tiles_per_b, out);
}

// ------------------------------ host side ---------------------------------
static inline size_t write_smem_bytes(int T, int c) {
  return (size_t)T * (c + 4) * 4 + (size_t)3 * T * 4;
}

// FBBEV_POOL_TILE overrides the tile size (32 / 64 / 128 voxels) for tuning.
static int split_pick_tile(int c) {
  static const char* env = getenv("FBBEV_POOL_TILE");
  if (env) {
    const int t = atoi(env);
    if (t == 32 || t == 64 || t == 128) return t;
  }
  if (write_smem_bytes(128, c) <= 45 * 1024) return 128;
  if (write_smem_bytes(64, c) <= 45 * 1024) return 64;
  return 32;
}

bool split_supported(int c, int64_t zyx) {
  return c % 4 == 0 && zyx % 4 == 0 && c >= 4 && c <= 1024 &&
         write_smem_bytes(32, c) <= 200 * 1024;
}

static inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

struct SplitWs {
  int* tile_first;
  int* seg_rank;
  int* warp_first;
  int* meta;
  int* flags;  // [n_sum_ctas] producer-done flags (cleared per launch)
  float* V;    // [n_intervals_max][c]  interval sums
  float* X;    // [n_warps_max][c]      carry rows of producer slices
  int n_warps_max, n_sum_ctas;
  size_t bytes;
};

static SplitWs split_layout(void* ws, int batch, int64_t zyx,
                            int n_intervals_max, int n_points_max, int c) {
  const int64_t tiles = (int64_t)batch * ceil_div64(zyx, 32);
  char* p = static_cast<char*>(ws);
  SplitWs w;
  size_t off = 0;
  w.tile_first = reinterpret_cast<int*>(p + off);
  off += up256((size_t)(tiles + 1) * 4);
  w.seg_rank = reinterpret_cast<int*>(p + off);
  off += up256((size_t)std::max(n_intervals_max, 1) * 4);
  w.n_warps_max = (int)ceil_div64(std::max(n_points_max, 1), kPtsPerWarp);
  w.n_sum_ctas = (int)ceil_div64(w.n_warps_max, kSumThreads / kWarp);
  w.warp_first = reinterpret_cast<int*>(p + off);
  off += up256((size_t)(w.n_warps_max + 1) * 4);
  w.meta = reinterpret_cast<int*>(p + off);
  off += 256;
  w.flags = reinterpret_cast<int*>(p + off);
  off += up256((size_t)w.n_sum_ctas * 4);
  w.V = reinterpret_cast<float*>(p + off);
  off += up256((size_t)std::max(n_intervals_max, 1) * c * 4);
  w.X = reinterpret_cast<float*>(p + off);
  off += up256((size_t)w.n_warps_max * c * 4);
  w.bytes = off;
  return w;
}

size_t split_workspace_bytes(int batch, int64_t zyx, int n_intervals_max,
                             int n_points_max, int c) {
  return split_layout(nullptr, batch, zyx, n_intervals_max, n_points_max, c)
      .bytes;
}

int split_plan(const int* ranks_bev, const int* interval_starts,
               const int* interval_lengths, int n_intervals_max,
               const int* n_intervals_dev, int n_points_max, int c, int batch,
               int64_t zyx, void* workspace, cudaStream_t st) {
  const SplitWs w =
      split_layout(workspace, batch, zyx, n_intervals_max, n_points_max, c);
  const int T = split_pick_tile(c);
  const int tiles_per_b = (int)ceil_div64(zyx, T);
  const int64_t n_tiles = (int64_t)batch * tiles_per_b;
  const int threads = 256;
  const int64_t work = n_intervals_max > 0
                           ? std::max<int64_t>(n_intervals_max, w.n_warps_max + 1)
                           : n_tiles + 1;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div64(work, threads),
                                                    (int64_t)1 << 20);
  count_launch();
  split_plan_kernel<<<grid, threads, 0, st>>>(
      ranks_bev, interval_starts, interval_lengths, n_intervals_max,
      n_intervals_dev, w.n_warps_max, zyx, tiles_per_b, T, n_tiles,
      w.tile_first, w.seg_rank, w.warp_first, w.meta);
  return launch_status();
}

template <int T, int LG, int VPL>
static int launch_pool(const SplitWs& w, const float* depth, const float* feat,
                       const int* ranks_depth, const int* ranks_feat,
                       const int* ranks_bev, const int* interval_starts,
                       const int* interval_lengths, int n_sum_ctas, int c,
                       int64_t zyx, int tiles_per_b, int64_t n_tiles,
                       float* out, cudaStream_t st) {
  const size_t smem = write_smem_bytes(T, c);
  auto k = pool_split_kernel<T, LG, VPL>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(
        k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  k<<<(unsigned)(n_sum_ctas + n_tiles), kPoolThreads, smem, st>>>(
      depth, feat, ranks_depth, ranks_feat, ranks_bev, interval_starts,
      interval_lengths, w.tile_first, w.seg_rank, w.warp_first, w.meta,
      w.flags, n_sum_ctas, c, zyx, tiles_per_b, w.V, w.X, out);
  return launch_status();
}

template <int T>
static int launch_pool_t(const SplitWs& w, const float* depth,
                         const float* feat, const int* ranks_depth,
                         const int* ranks_feat, const int* ranks_bev,
                         const int* interval_starts,
                         const int* interval_lengths, int n_sum_ctas, int c,
                         int64_t zyx, int tiles_per_b, int64_t n_tiles,
                         float* out, cudaStream_t st) {
  const int c4 = c / 4;
#define FBBEV_POOL_CASE(LGV, VPLV)                                            \
  return launch_pool<T, LGV, VPLV>(w, depth, feat, ranks_depth, ranks_feat,   \
                                   ranks_bev, interval_starts,                \
                                   interval_lengths, n_sum_ctas, c, zyx,      \
                                   tiles_per_b, n_tiles, out, st)
  // group width LG and float4 columns per lane VPL with LG * VPL >= C / 4
  if (c4 <= 4) FBBEV_POOL_CASE(4, 1);
  if (c4 <= 8) FBBEV_POOL_CASE(4, 2);
  if (c4 <= 16) FBBEV_POOL_CASE(4, 4);
  if (c4 <= 20) FBBEV_POOL_CASE(4, 5);
  if (c4 <= 32) FBBEV_POOL_CASE(4, 8);
  if (c4 <= 64) FBBEV_POOL_CASE(8, 8);
  if (c4 <= 128) FBBEV_POOL_CASE(16, 8);
  FBBEV_POOL_CASE(32, 8);
#undef FBBEV_POOL_CASE
}

int split_launch(const float* depth, const float* feat, const int* ranks_depth,
                 const int* ranks_feat, const int* ranks_bev,
                 const int* interval_starts, const int* interval_lengths,
                 int n_intervals_max, int n_points_max, int c, int batch,
                 int64_t zyx, float* out, void* workspace, cudaStream_t st) {
  if (reinterpret_cast<uintptr_t>(out) & 15) return FBBEV_ERR_INVALID_ARGUMENT;
  const SplitWs w =
      split_layout(workspace, batch, zyx, n_intervals_max, n_points_max, c);
  const int T = split_pick_tile(c);
  const int tiles_per_b = (int)ceil_div64(zyx, T);
  const int64_t n_tiles = (int64_t)batch * tiles_per_b;
  // no producers at all when the index is empty (the plan left meta = 0 and
  // tile_first = 0, so every tile is an empty tile)
  const int n_sum_ctas = n_intervals_max > 0 ? w.n_sum_ctas : 0;
  if (n_sum_ctas > 0) {
    cudaError_t e = cudaMemsetAsync(w.flags, 0, (size_t)n_sum_ctas * 4, st);
    if (e != cudaSuccess) return (int)e;
  }
  count_launch();
  switch (T) {
    case 128:
      return launch_pool_t<128>(w, depth, feat, ranks_depth, ranks_feat,
                                ranks_bev, interval_starts, interval_lengths,
                                n_sum_ctas, c, zyx, tiles_per_b, n_tiles, out,
                                st);
    case 64:
      return launch_pool_t<64>(w, depth, feat, ranks_depth, ranks_feat,
                               ranks_bev, interval_starts, interval_lengths,
                               n_sum_ctas, c, zyx, tiles_per_b, n_tiles, out,
                               st);
    default:
      return launch_pool_t<32>(w, depth, feat, ranks_depth, ranks_feat,
                               ranks_bev, interval_starts, interval_lengths,
                               n_sum_ctas, c, zyx, tiles_per_b, n_tiles, out,
                               st);
  }
}

}  // namespace fbbev
