"""CPU restatement of the depth-aware spatial cross-attention core and a CPU
runner for the backward projection.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Used by tests/ (as the
checker of the fused CUDA kernel on seeded inputs larger than the golden
fixtures) and by bench.py's CPU-baseline / ``--impl reference`` legs.

``da_sca_core_cpu`` follows the reference's algorithm step by step
(spatial_cross_attention_depth.py, line numbers in comments): per-camera query
selection, zero-padded re-batching, one-hot depth bins, the depth look-up MSDA
launch, depth re-weighting, the main MSDA launch, scatter-add over cameras and
the division by the per-query camera count.  MSDA itself is the C oracle
(``oracle_msda_fwd``).  It is pinned against the golden vectors recorded from
the reference's own Python in tests/test_oracle.py.
"""
import contextlib

import numpy as np
import torch
import torch.nn.functional as F

from . import cpu


def _msda(value, shapes, lsi, loc, attw):
    out = cpu.msda_fwd(value.contiguous().numpy(), shapes.numpy(),
                       lsi.numpy(), loc.contiguous().numpy(),
                       attw.contiguous().numpy())
    return torch.from_numpy(out)


def msda_fused_cpu(value, spatial_shapes, level_start_index, reference_points,
                   sampling_offsets, attention_logits, map_width=0):
    """mmcv MultiScaleDeformableAttention.forward core: softmax, sampling
    locations = ref + offsets / (W, H), ms_deform_attn_forward."""
    shapes = spatial_shapes.cpu().long()
    lsi = level_start_index.cpu().long()
    w = attention_logits.flatten(3).softmax(-1).view_as(attention_logits)
    wh = torch.stack([shapes[..., 1], shapes[..., 0]], -1).float()
    loc = reference_points[:, :, None, :, None, :] + \
        sampling_offsets / wh[None, None, None, :, None, :]
    return _msda(value.float(), shapes, lsi, loc, w)


def da_sca_core_cpu(value, depth_prob, reference_points_cam, bev_query_depth,
                    per_cam_mask, sampling_offsets, attention_logits,
                    spatial_shapes, level_start_index, dbound, num_Z_anchors,
                    prepared=None):
    """Same contract as ``fbbev_da_sca_fwd`` / ops.da_spatial_cross_attention_core.

    value (bs*N, n_value, heads, ch); depth_prob (bs*N, H0*W0, DC);
    reference_points_cam (N, bs, nq, Z, 2); bev_query_depth (N, bs, nq, Z);
    per_cam_mask (N, bs, nq, Z) bool; sampling_offsets (bs, nq, heads, L, P, 2);
    attention_logits (bs, nq, heads, L, P).  Returns (bs, nq, heads*ch)."""
    shapes = spatial_shapes.cpu().long()
    lsi = level_start_index.cpu().long()
    N, bs, nq, Z, _ = reference_points_cam.shape
    _, n_value, heads, ch = value.shape
    L, P = sampling_offsets.shape[3], sampling_offsets.shape[4]
    DC = depth_prob.shape[-1]
    E = heads * ch
    if bev_query_depth.dim() == 4:
        bev_query_depth = bev_query_depth[..., None]
    # :163-169 per (sample, camera) list of visible queries
    seen = per_cam_mask.sum(-1) > 0
    idxs = [[seen[i, j].nonzero().squeeze(-1) for i in range(N)]
            for j in range(bs)]
    max_len = max(1, max(len(i) for r in idxs for i in r))
    # :173-186 zero-padded re-batch (offsets/logits are row-wise functions of
    # the query, so re-batching them == applying the Linears to re-batched
    # queries)
    off_re = sampling_offsets.new_zeros(bs, N, max_len, heads, L, P, 2)
    log_re = attention_logits.new_zeros(bs, N, max_len, heads, L, P)
    ref_re = reference_points_cam.new_zeros(bs, N, max_len, Z, 2)
    dep_re = reference_points_cam.new_zeros(bs, N, max_len, Z, 1)
    for j in range(bs):
        for i in range(N):
            k = idxs[j][i]
            off_re[j, i, :len(k)] = sampling_offsets[j, k]
            log_re[j, i, :len(k)] = attention_logits[j, k]
            ref_re[j, i, :len(k)] = reference_points_cam[i, j, k]
            dep_re[j, i, :len(k)] = bev_query_depth[i, j, k]
    # :196-199 depth bin one-hot
    bins = torch.floor((dep_re - dbound[0]) / dbound[2])
    bins = torch.clip(bins, 0, DC - 1).to(torch.long)
    onehot = F.one_hot(bins.squeeze(-1), num_classes=DC)
    B2 = bs * N
    off = off_re.view(B2, max_len, heads, L, P, 2)
    w = log_re.view(B2, max_len, heads, L * P).softmax(-1).view(
        B2, max_len, heads, L, P)                                   # :540
    ref = ref_re.view(B2, max_len, Z, 2)
    wh = torch.stack([shapes[..., 1], shapes[..., 0]], -1).float()
    off = off / wh[None, None, None, :, None, :]                    # :558
    loc = ref[:, :, None, None, None, :, :] + off.view(
        B2, max_len, heads, L, P // Z, Z, 2)                        # :563
    loc = loc.view(B2, max_len, heads, L, P, 2)                     # :567
    # :584-591 depth look-up
    dref = ref.reshape(B2, max_len * Z, 1, 1, 1, 2)
    dw = _msda(depth_prob.unsqueeze(2).float(), shapes[0:1], lsi[0:1], dref,
               torch.ones_like(dref[..., 0]))
    dw = (dw.reshape(B2, max_len, Z, -1) *
          onehot.view(B2, max_len, Z, DC)).sum(-1)
    dw = dw.unsqueeze(2).repeat(1, 1, P // Z, 1).reshape(B2, max_len, P)
    w = w * dw[:, :, None, None, :]                                 # :592
    out = _msda(value.float(), shapes, lsi, loc, w).view(
        bs, N, max_len, E)                                          # :593-595
    # :208-216 scatter back and average over the cameras that see a query
    slots = out.new_zeros(bs, nq, E)
    for j in range(bs):
        for i in range(N):
            k = idxs[j][i]
            slots[j, k] += out[j, i, :len(k)]
    count = seen.permute(1, 2, 0).sum(-1)
    count = torch.clamp(count, min=1.0)
    return slots / count[..., None]


@contextlib.contextmanager
def _cpu_kernels():
    """Run the plugin's host-side Python with the CUDA ops replaced by the
    oracle (the plugin itself has no CPU path)."""
    from fbbev_b200.view_transformation import backward_projection as bp_mod
    saved = (bp_mod.ms_deform_attn_fused,
             bp_mod.da_spatial_cross_attention_core)
    bp_mod.ms_deform_attn_fused = msda_fused_cpu
    bp_mod.da_spatial_cross_attention_core = da_sca_core_cpu
    try:
        yield
    finally:
        (bp_mod.ms_deform_attn_fused,
         bp_mod.da_spatial_cross_attention_core) = saved


@torch.no_grad()
def backward_projection_cpu(bp, mlvl_feats, lss_bev, cam_params,
                            pred_img_depth):
    """BackwardProjection.forward on CPU tensors: the module's own dense
    layers (torch CPU) around the oracle's attention cores."""
    with _cpu_kernels():
        return bp(mlvl_feats, None, lss_bev=lss_bev, cam_params=cam_params,
                  pred_img_depth=pred_img_depth)
