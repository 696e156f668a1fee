"""Golden fixtures for the backward projection, produced by running the
reference's own classes (see gen_golden.py / ref_import.py).

b_sca_*   one DA_SpatialCrossAttention.forward call: inputs, module weights
          and output (spatial_cross_attention_depth.py:86-223, 465-601).
b_bp_*    a whole BackwardProjection.forward (backward_projection.py:85-133)
          including BEVFormer, the encoder's point_sampling and the layer's
          self-attention / LayerNorm / FFN: cam params, inputs, state_dict and
          output, plus the encoder's intermediate geometry.
"""
import numpy as np
import torch

from fbbev_b200 import synthetic


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


def _perturb(module, seed, scale=0.02):
    """Module's own init + small seeded noise so offsets / weights are
    non-trivial."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in module.parameters():
            p.add_(torch.randn(p.shape, generator=g) * scale)


def bp_config(bev_h, bev_w, E, n_levels, pc_range, grid_bev, input_size, dbound,
              num_points=8):
    return dict(
        type='BackwardProjection', bev_h=bev_h, bev_w=bev_w, in_channels=E,
        out_channels=E, pc_range=pc_range,
        transformer=dict(
            type='BEVFormer', use_cams_embeds=False, embed_dims=E,
            encoder=dict(
                type='bevformer_encoder', num_layers=1, pc_range=pc_range,
                grid_config=grid_bev,
                data_config=dict(input_size=input_size),
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
                                 embed_dims=E, num_points=num_points,
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


def gen_backward(ref, save):
    from ref_import import REGS

    def run_bp(name, B, bev_hw, E, level_shapes, seed):
        bev_h, bev_w = bev_hw
        pc_range = [-40, -40, -1.0, 40, 40, 5.4]
        step = 80.0 / bev_w
        grid_bev = dict(x=[-40, 40, step], y=[-40, 40, 80.0 / bev_h],
                        z=[-1, 5.4, 1.6])
        input_size = (128, 352)
        dbound = [2.0, 42.0, 1.0]
        DC = 40
        cfg = bp_config(bev_h, bev_w, E, len(level_shapes), pc_range, grid_bev,
                        input_size, dbound)
        torch.manual_seed(seed)
        bp = REGS['HEADS'].build(cfg)
        bp.init_weights()
        _perturb(bp, seed + 100)
        bp.eval()
        g = torch.Generator().manual_seed(seed)
        cam = synthetic.make_cam_params(B, 6, input_size, jitter=1.0, seed=seed)
        mlvl = [torch.randn(B, 6, E, h, w, generator=g) for h, w in level_shapes]
        H0, W0 = level_shapes[0]
        depth = torch.randn(B, 6, DC, H0, W0, generator=g).softmax(2)
        lss_bev = torch.randn(B, E, bev_h, bev_w, generator=g) * 0.1

        # capture the cross-attention call and the encoder geometry
        cap = {}
        sca = bp.transformer.encoder.layers[0].attentions[1]
        orig = sca.forward

        def spy(*a, **k):
            out = orig(*a, **k)
            cap['q'], cap['key'] = a[0], a[1]
            cap.update({kk: k[kk] for kk in (
                'query_pos', 'reference_points_cam', 'bev_query_depth',
                'per_cam_mask_list', 'spatial_shapes', 'level_start_index',
                'pred_img_depth')})
            cap['out'] = out
            return out
        sca.forward = spy
        with torch.no_grad():
            out = bp(mlvl, None, lss_bev=lss_bev, cam_params=cam,
                     pred_img_depth=depth)
        sca.forward = orig

        # bev_mask (spatial_cross_attention_depth.py:156-169): (a) a random
        # half of the BEV cells; (b) every cell camera 0 sees is excluded, so
        # camera 0's masked list is empty and the reference's empty-camera
        # rule (:166-167) adds its first visible query, uncounted
        per_cam = cap['per_cam_mask_list']                 # (N, B, nq, Z)
        gm = torch.Generator().manual_seed(seed + 7)
        mask_a = torch.rand(B, bev_h * bev_w, generator=gm) > 0.5
        mask_b = ~per_cam[0].any(-1)
        assert per_cam[0].any() and (per_cam & mask_b[None, :, :, None]).any()
        with torch.no_grad():
            out_a = bp(mlvl, None, lss_bev=lss_bev, cam_params=cam,
                       pred_img_depth=depth,
                       bev_mask=mask_a.view(B, bev_h, bev_w))
            out_b = bp(mlvl, None, lss_bev=lss_bev, cam_params=cam,
                       pred_img_depth=depth,
                       bev_mask=mask_b.view(B, bev_h, bev_w))

        sd = {'sd::' + k: _np(v) for k, v in bp.state_dict().items()}
        arrays = dict(
            bev_h=np.array(bev_h), bev_w=np.array(bev_w), E=np.array(E),
            level_shapes=np.array(level_shapes), pc_range=np.array(pc_range),
            grid_x=np.array(grid_bev['x']), grid_y=np.array(grid_bev['y']),
            grid_z=np.array(grid_bev['z']), input_size=np.array(input_size),
            dbound=np.array(dbound),
            rots=_np(cam[0]), trans=_np(cam[1]), intrins=_np(cam[2]),
            post_rots=_np(cam[3]), post_trans=_np(cam[4]), bda=_np(cam[5]),
            depth=_np(depth), lss_bev=_np(lss_bev), out=_np(out),
            sca_query=_np(cap['q']), sca_key=_np(cap['key']),
            sca_query_pos=_np(cap['query_pos']),
            reference_points_cam=_np(cap['reference_points_cam']),
            bev_query_depth=_np(cap['bev_query_depth']),
            per_cam_mask=_np(cap['per_cam_mask_list']),
            spatial_shapes=_np(cap['spatial_shapes']),
            level_start_index=_np(cap['level_start_index']),
            sca_out=_np(cap['out']),
            bev_mask_a=_np(mask_a), out_bev_mask_a=_np(out_a),
            bev_mask_b=_np(mask_b), out_bev_mask_b=_np(out_b), **sd)
        for i, f in enumerate(mlvl):
            arrays[f'feat{i}'] = _np(f)
        save(name, **arrays)

    # shipped-like: E=80 (head_dim 10), one level, 12x12 BEV
    run_bp('b_bp_e80_1lvl', B=1, bev_hw=(12, 12), E=80,
           level_shapes=[(8, 22)], seed=21)
    # FB-BEV-like: E=64 (head_dim 8), three levels, batch 2
    run_bp('b_bp_e64_3lvl', B=2, bev_hw=(10, 14), E=64,
           level_shapes=[(8, 22), (4, 11), (2, 6)], seed=22)
