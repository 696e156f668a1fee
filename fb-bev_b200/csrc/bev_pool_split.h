// bev_pool_split.h -- internal interface of the two-kernel dense pooling path
// (bev_pool_split.cu), used by the dispatcher in bev_pool_fwd.cu.
#pragma once
#include "common.cuh"

namespace fbbev {

bool split_supported(int c, int64_t zyx);
size_t split_workspace_bytes(int batch, int64_t zyx, int n_intervals_max,
                             int n_points_max, int c);
int split_plan(const int* ranks_bev, const int* interval_starts,
               const int* interval_lengths, int n_intervals_max,
               const int* n_intervals_dev, int n_points_max, int c, int batch,
               int64_t zyx, void* workspace, cudaStream_t st);
// Where the plan tables of a dense-pooling workspace live, so that the index
// builder (voxel_prepare.cu) can fill them while it scans the histogram
// instead of a separate plan launch.
struct SplitPlanPtrs {
  int *tile_first, *seg_rank, *warp_first, *meta;
  int T, tiles_per_b, n_warps_max;
  int64_t n_tiles;
};
SplitPlanPtrs split_plan_ptrs(void* workspace, int batch, int64_t zyx,
                              int n_intervals_max, int n_points_max, int c);
// true when the dense op runs the two-kernel path for this shape
bool dense_uses_split(int c, int64_t zyx);
// stages: K1 (interval sums) and / or K2 (dense write, optionally + `add`, a
// (B, C, yx_n) map broadcast over Z)
constexpr int kSplitSums = 1, kSplitWrite = 2;
int split_launch(const float* depth, const float* feat, const int* ranks_depth,
                 const int* ranks_feat, const int* ranks_bev,
                 const int* interval_starts,
                 const int* interval_lengths, int n_intervals_max,
                 int n_points_max, int c, int batch, int64_t zyx, float* out,
                 void* workspace, cudaStream_t st,
                 int stages = kSplitSums | kSplitWrite,
                 const float* add = nullptr, int yx_n = 0);
// token-major Z-mean (B, yx_n, C) of the planned + summed volume
int split_zmean(const int* interval_starts, const int* interval_lengths,
                int n_intervals_max, int n_points_max, int c, int batch,
                int64_t zyx, int yx_n, float* lss, void* workspace,
                cudaStream_t st);

}  // namespace fbbev
