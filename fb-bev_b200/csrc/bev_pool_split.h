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
int split_launch(const float* depth, const float* feat, const int* ranks_depth,
                 const int* ranks_feat, const int* ranks_bev,
                 const int* interval_starts,
                 const int* interval_lengths, int n_intervals_max,
                 int n_points_max, int c, int batch, int64_t zyx, float* out,
                 void* workspace, cudaStream_t st);

}  // namespace fbbev
