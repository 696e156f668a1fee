"""Tiny on-device self-check of the backward projection used by
``__graft_entry__.smoke()``: the fused depth-aware cross-attention kernel must
agree with the same module's reference-shaped (re-batched, unfused) path."""
import torch


def smoke_backward_projection(dev="cuda:0"):
    from .registry import build_head
    from . import synthetic
    E, bev = 80, 10
    pc_range = [-40, -40, -1.0, 40, 40, 5.4]
    cfg = dict(
        type='BackwardProjection', bev_h=bev, bev_w=bev, in_channels=E,
        out_channels=E, pc_range=pc_range,
        transformer=dict(
            type='BEVFormer', use_cams_embeds=False, embed_dims=E,
            encoder=dict(
                type='bevformer_encoder', num_layers=1, pc_range=pc_range,
                grid_config=dict(x=[-40, 40, 8.0], y=[-40, 40, 8.0],
                                 z=[-1, 5.4, 1.6]),
                data_config=dict(input_size=(128, 352)),
                transformerlayers=dict(
                    type='BEVFormerEncoderLayer',
                    attn_cfgs=[
                        dict(type='MultiScaleDeformableAttention',
                             embed_dims=E, dropout=0.0, num_levels=1),
                        dict(type='DA_SpatialCrossAttention',
                             pc_range=pc_range, dbound=[2.0, 42.0, 1.0],
                             dropout=0.0,
                             deformable_attention=dict(
                                 type='DA_MSDeformableAttention',
                                 embed_dims=E, num_points=8, num_levels=1),
                             embed_dims=E)],
                    ffn_cfgs=dict(type='FFN', embed_dims=E,
                                  feedforward_channels=4 * E, ffn_drop=0.0),
                    operation_order=('self_attn', 'norm', 'cross_attn', 'norm',
                                     'ffn', 'norm')))),
        positional_encoding=dict(type='CustormLearnedPositionalEncoding',
                                 num_feats=E // 2, row_num_embed=bev,
                                 col_num_embed=bev))
    torch.manual_seed(0)
    bp = build_head(cfg)
    bp.init_weights()
    with torch.no_grad():
        for p in bp.parameters():
            p.add_(torch.randn_like(p) * 0.02)
    bp = bp.to(dev).eval()
    cam = synthetic.make_cam_params(1, 6, (128, 352), device=dev)
    g = torch.Generator().manual_seed(1)
    feat = torch.randn(1, 6, E, 8, 22, generator=g).to(dev)
    depth = torch.randn(1, 6, 40, 8, 22, generator=g).softmax(2).to(dev)
    lss = (torch.randn(1, E, bev, bev, generator=g) * 0.1).to(dev)
    # capture the cross-attention call, then replay it through the
    # reference-shaped path (an all-true bev_mask selects it)
    sca = bp.transformer.encoder.layers[0].attentions[1]
    cap = {}
    orig = sca.forward

    def spy(*a, **k):
        cap['a'], cap['k'] = a, k
        cap['out'] = orig(*a, **k)
        return cap['out']
    sca.forward = spy
    with torch.no_grad():
        out = bp([feat], None, lss_bev=lss, cam_params=cam,
                 pred_img_depth=depth)
        sca.forward = orig
        kw = dict(cap['k'])
        kw['bev_mask'] = torch.ones(1, bev * bev, dtype=torch.bool, device=dev)
        unfused = sca(*cap['a'], **kw)
    fused = cap['out']
    err = (fused - unfused).abs().max().item()
    assert torch.isfinite(out).all()
    assert err <= 1e-4, f"fused vs re-batched cross-attention: {err}"
    return err
