// da_sca_smem.h -- shared-memory-resident depth-aware spatial cross-attention
// (da_sca_smem.cu), selected by fbbev_da_sca_fwd when the shape allows.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace fbbev {

size_t da_sca_smem_workspace_bytes(int bs, int n_cams);
bool da_sca_smem_eligible(int n_cams, int n_value, int heads, int ch, int levels,
                          int points, int Z);
// stages: the prologue (per-camera counts + zero-fill of `out`; needs only the
// mask) and / or the main kernel
constexpr int kScaPrologue = 1, kScaMain = 2;
int da_sca_smem_launch(const float* value, const float* depth_prob,
                       const float* ref_cam, const float* ref_depth,
                       const uint8_t* mask, const float* offsets,
                       const float* logits, const int64_t* shapes, float d_min,
                       float d_step, int bs, int n_cams, int nq, int n_value,
                       int DC, float* out, void* workspace, cudaStream_t st,
                       int stages = kScaPrologue | kScaMain);

}  // namespace fbbev
