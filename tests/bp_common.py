"""Helpers shared by the backward-projection tests: rebuild the module from the
config recorded in a golden fixture and load the reference's weights."""
import numpy as np
import torch

from conftest import load_golden


def bp_cfg_from_golden(g):
    E = int(g["E"])
    bev_h, bev_w = int(g["bev_h"]), int(g["bev_w"])
    n_levels = len(g["level_shapes"])
    pc_range = [float(v) for v in g["pc_range"]]
    grid_bev = dict(x=[float(v) for v in g["grid_x"]],
                    y=[float(v) for v in g["grid_y"]],
                    z=[float(v) for v in g["grid_z"]])
    input_size = tuple(int(v) for v in g["input_size"])
    dbound = [float(v) for v in g["dbound"]]
    return dict(
        type='BackwardProjection', bev_h=bev_h, bev_w=bev_w, in_channels=E,
        out_channels=E, pc_range=pc_range,
        transformer=dict(
            type='BEVFormer', use_cams_embeds=False, embed_dims=E,
            encoder=dict(
                type='bevformer_encoder', num_layers=1, pc_range=pc_range,
                grid_config=grid_bev, data_config=dict(input_size=input_size),
                return_intermediate=False,
                transformerlayers=dict(
                    type='BEVFormerEncoderLayer',
                    attn_cfgs=[
                        dict(type='MultiScaleDeformableAttention',
                             embed_dims=E, dropout=0.0, num_levels=1),
                        dict(type='DA_SpatialCrossAttention',
                             pc_range=pc_range, dbound=dbound, dropout=0.0,
                             deformable_attention=dict(
                                 type='DA_MSDeformableAttention',
                                 embed_dims=E, num_points=8,
                                 num_levels=n_levels),
                             embed_dims=E)],
                    ffn_cfgs=dict(type='FFN', embed_dims=E,
                                  feedforward_channels=E * 4, ffn_drop=0.0,
                                  act_cfg=dict(type='ReLU', inplace=True)),
                    feedforward_channels=E * 4, ffn_dropout=0.0,
                    operation_order=('self_attn', 'norm', 'cross_attn', 'norm',
                                     'ffn', 'norm')))),
        positional_encoding=dict(type='CustormLearnedPositionalEncoding',
                                 num_feats=E // 2, row_num_embed=bev_h,
                                 col_num_embed=bev_w))


def build_bp(name, device="cpu"):
    from fbbev_b200.registry import build_head
    g = load_golden(name)
    bp = build_head(bp_cfg_from_golden(g))
    sd = {k[4:]: torch.from_numpy(v) for k, v in g.items()
          if k.startswith("sd::")}
    missing, unexpected = bp.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return g, bp.to(device).eval()


def cam_params(g, device="cpu"):
    return [torch.from_numpy(g[k]).to(device) for k in
            ("rots", "trans", "intrins", "post_rots", "post_trans", "bda")]
