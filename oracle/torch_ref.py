"""PyTorch-CPU float32 restatements of the third-party (mmcv-full 1.5.2) pieces
the reference's backward projection calls but does not vendor.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

mmcv is not installable in this environment, so these follow mmcv's published
behaviour and are anchored on the reference's own call sites:

* ``multi_scale_deformable_attn_pytorch`` -- imported by the reference at
  .../bevformer_utils/spatial_cross_attention_depth.py:7 and used as its CPU
  branch (:597-598); mmcv documents it as the pure-PyTorch equivalent of
  ``ms_deform_attn_forward``.
* ``ms_deform_attn_forward`` -- the ``_ext`` op bound at
  .../multi_scale_deformable_attn_function.py:18-19, called at :127-133.
* ``MultiScaleDeformableAttention`` -- the module the FB-OCC config names as the
  encoder layer's ``self_attn`` (fbocc-r50 config :176-180).
* ``FFN`` -- ``mmcv.cnn.bricks.transformer.FFN`` (config :194-201).

They are cross-checked in tests/ against the C oracle (im2col formulation) and
against HF transformers' independent implementation.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def multi_scale_deformable_attn_pytorch(value, value_spatial_shapes,
                                        sampling_locations, attention_weights):
    """grid_sample formulation: pixel = loc*size - 0.5 (align_corners=False),
    zero padding, bilinear; weighted sum over levels x points."""
    bs, _, num_heads, embed_dims = value.shape
    _, num_queries, num_heads, num_levels, num_points, _ = \
        sampling_locations.shape
    shapes = [(int(h), int(w)) for h, w in value_spatial_shapes]
    value_list = value.split([h * w for h, w in shapes], dim=1)
    sampling_grids = 2 * sampling_locations - 1
    sampling_value_list = []
    for level, (H_, W_) in enumerate(shapes):
        value_l_ = value_list[level].flatten(2).transpose(1, 2).reshape(
            bs * num_heads, embed_dims, H_, W_)
        sampling_grid_l_ = sampling_grids[:, :, :, level].transpose(
            1, 2).flatten(0, 1)
        sampling_value_l_ = F.grid_sample(
            value_l_, sampling_grid_l_, mode='bilinear',
            padding_mode='zeros', align_corners=False)
        sampling_value_list.append(sampling_value_l_)
    attention_weights = attention_weights.transpose(1, 2).reshape(
        bs * num_heads, 1, num_queries, num_levels * num_points)
    output = (torch.stack(sampling_value_list, dim=-2).flatten(-2) *
              attention_weights).sum(-1).view(bs, num_heads * embed_dims,
                                              num_queries)
    return output.transpose(1, 2).contiguous()


def ms_deform_attn_forward(value, spatial_shapes, level_start_index,
                           sampling_locations, attention_weights,
                           im2col_step=64):
    """``ext_module.ms_deform_attn_forward`` through the C oracle (im2col
    formulation, oracle/fbbev_oracle.c:oracle_msda_fwd)."""
    from . import cpu
    out = cpu.msda_fwd(value.detach().cpu().numpy(),
                       spatial_shapes.cpu().numpy(),
                       level_start_index.cpu().numpy(),
                       sampling_locations.detach().cpu().numpy(),
                       attention_weights.detach().cpu().numpy())
    return torch.from_numpy(out)


def ms_deform_attn_backward(value, spatial_shapes, level_start_index,
                            sampling_locations, attention_weights, grad_output,
                            grad_value, grad_sampling_loc, grad_attn_weight,
                            im2col_step=64):
    from . import cpu
    gv, gl, ga = cpu.msda_bwd(value.detach().numpy(), spatial_shapes.numpy(),
                              level_start_index.numpy(),
                              sampling_locations.detach().numpy(),
                              attention_weights.detach().numpy(),
                              grad_output.detach().numpy())
    grad_value.copy_(torch.from_numpy(gv))
    grad_sampling_loc.copy_(torch.from_numpy(gl))
    grad_attn_weight.copy_(torch.from_numpy(ga))


def xavier_init(module, gain=1, bias=0, distribution='normal'):
    if hasattr(module, 'weight') and module.weight is not None:
        if distribution == 'uniform':
            nn.init.xavier_uniform_(module.weight, gain=gain)
        else:
            nn.init.xavier_normal_(module.weight, gain=gain)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def constant_init(module, val, bias=0):
    if hasattr(module, 'weight') and module.weight is not None:
        nn.init.constant_(module.weight, val)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


class MultiScaleDeformableAttention(nn.Module):
    """mmcv.ops.MultiScaleDeformableAttention (1.5.2 behaviour)."""

    def __init__(self, embed_dims=256, num_heads=8, num_levels=4,
                 num_points=4, im2col_step=64, dropout=0.1, batch_first=False,
                 norm_cfg=None, init_cfg=None):
        super().__init__()
        assert embed_dims % num_heads == 0
        self.norm_cfg = norm_cfg
        self.dropout = nn.Dropout(dropout)
        self.batch_first = batch_first
        self.im2col_step = im2col_step
        self.embed_dims = embed_dims
        self.num_levels = num_levels
        self.num_heads = num_heads
        self.num_points = num_points
        self.sampling_offsets = nn.Linear(
            embed_dims, num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(
            embed_dims, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.init_weights()

    def init_weights(self):
        constant_init(self.sampling_offsets, 0.)
        thetas = torch.arange(self.num_heads, dtype=torch.float32) * (
            2.0 * math.pi / self.num_heads)
        grid_init = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid_init = (grid_init / grid_init.abs().max(-1, keepdim=True)[0]
                     ).view(self.num_heads, 1, 1, 2).repeat(
                         1, self.num_levels, self.num_points, 1)
        for i in range(self.num_points):
            grid_init[:, :, i, :] *= i + 1
        self.sampling_offsets.bias.data = grid_init.view(-1)
        constant_init(self.attention_weights, val=0., bias=0.)
        xavier_init(self.value_proj, distribution='uniform', bias=0.)
        xavier_init(self.output_proj, distribution='uniform', bias=0.)

    def forward(self, query, key=None, value=None, identity=None,
                query_pos=None, key_padding_mask=None, reference_points=None,
                spatial_shapes=None, level_start_index=None, **kwargs):
        if value is None:
            value = query
        if identity is None:
            identity = query
        if query_pos is not None:
            query = query + query_pos
        if not self.batch_first:
            query = query.permute(1, 0, 2)
            value = value.permute(1, 0, 2)
        bs, num_query, _ = query.shape
        bs, num_value, _ = value.shape
        assert (spatial_shapes[:, 0] * spatial_shapes[:, 1]).sum() == num_value
        value = self.value_proj(value)
        if key_padding_mask is not None:
            value = value.masked_fill(key_padding_mask[..., None], 0.0)
        value = value.view(bs, num_value, self.num_heads, -1)
        sampling_offsets = self.sampling_offsets(query).view(
            bs, num_query, self.num_heads, self.num_levels, self.num_points, 2)
        attention_weights = self.attention_weights(query).view(
            bs, num_query, self.num_heads, self.num_levels * self.num_points)
        attention_weights = attention_weights.softmax(-1)
        attention_weights = attention_weights.view(
            bs, num_query, self.num_heads, self.num_levels, self.num_points)
        assert reference_points.shape[-1] == 2
        offset_normalizer = torch.stack(
            [spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
        sampling_locations = reference_points[:, :, None, :, None, :] \
            + sampling_offsets \
            / offset_normalizer[None, None, None, :, None, :]
        output = ms_deform_attn_forward(
            value, spatial_shapes, level_start_index, sampling_locations,
            attention_weights, self.im2col_step)
        output = self.output_proj(output)
        if not self.batch_first:
            output = output.permute(1, 0, 2)
        return self.dropout(output) + identity


class FFN(nn.Module):
    """mmcv.cnn.bricks.transformer.FFN (1.5.2 behaviour, ReLU only)."""

    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2,
                 act_cfg=dict(type='ReLU', inplace=True), ffn_drop=0.,
                 dropout_layer=None, add_identity=True, init_cfg=None,
                 **kwargs):
        super().__init__()
        assert num_fcs >= 2
        assert act_cfg.get('type', 'ReLU') == 'ReLU'
        self.embed_dims = embed_dims
        self.feedforward_channels = feedforward_channels
        self.num_fcs = num_fcs
        layers = []
        in_channels = embed_dims
        for _ in range(num_fcs - 1):
            layers.append(nn.Sequential(
                nn.Linear(in_channels, feedforward_channels),
                nn.ReLU(inplace=True), nn.Dropout(ffn_drop)))
            in_channels = feedforward_channels
        layers.append(nn.Linear(feedforward_channels, embed_dims))
        layers.append(nn.Dropout(ffn_drop))
        self.layers = nn.Sequential(*layers)
        self.dropout_layer = nn.Identity()
        self.add_identity = add_identity

    def forward(self, x, identity=None):
        out = self.layers(x)
        if not self.add_identity:
            return self.dropout_layer(out)
        if identity is None:
            identity = x
        return identity + self.dropout_layer(out)


def to_numpy_tree(obj):
    if isinstance(obj, torch.Tensor):
        return obj.detach().cpu().numpy()
    if isinstance(obj, (list, tuple)):
        return type(obj)(to_numpy_tree(o) for o in obj)
    if isinstance(obj, dict):
        return {k: to_numpy_tree(v) for k, v in obj.items()}
    return np.asarray(obj)
