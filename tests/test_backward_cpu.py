"""CPU tests of the host-side mirror of the backward projection: state-dict
keys, registry/config contract and the plain-PyTorch geometry must equal the
reference's own Python (golden fixtures from tests/golden/gen_golden.py)."""
import numpy as np
import pytest
import torch

from bp_common import bp_cfg_from_golden, build_bp, cam_params
from conftest import load_golden

CASES = ["b_bp_e80_1lvl", "b_bp_e64_3lvl"]


@pytest.mark.parametrize("case", CASES)
def test_state_dict_keys_and_shapes_match_reference(case):
    from fbbev_b200.registry import build_head
    g = load_golden(case)
    bp = build_head(bp_cfg_from_golden(g))
    ref = {k[4:]: v.shape for k, v in g.items() if k.startswith("sd::")}
    mine = {k: tuple(v.shape) for k, v in bp.state_dict().items()}
    assert set(mine) == set(ref)
    for k in ref:
        assert mine[k] == tuple(ref[k]), k


@pytest.mark.parametrize("case", CASES)
def test_point_sampling_bit_exact(case):
    """bevformer_encoder.get_reference_points / point_sampling
    (bevformer_encoder.py:52-120) on CPU == the reference's tensors."""
    g, bp = build_bp(case)
    enc = bp.transformer.encoder
    bev_h, bev_w = int(g["bev_h"]), int(g["bev_w"])
    ref_3d = enc.get_reference_points(bev_h, bev_w, dim='3d', device='cpu')
    _, ref_cam, mask, depth = enc.point_sampling(ref_3d, enc.pc_range, None,
                                                 cam_params=cam_params(g))
    np.testing.assert_array_equal(ref_cam.numpy(), g["reference_points_cam"])
    np.testing.assert_array_equal(mask.numpy(), g["per_cam_mask"])
    np.testing.assert_array_equal(depth.numpy(), g["bev_query_depth"])
    ref_2d = enc.get_reference_points(bev_h, bev_w, dim='2d', bs=2,
                                      device='cpu')
    assert ref_2d.shape == (2, bev_h * bev_w, 1, 2)
    assert torch.allclose(ref_2d[0, 0, 0], torch.tensor([0.5 / bev_w,
                                                         0.5 / bev_h]))


def test_positional_encoding_and_init():
    from fbbev_b200.view_transformation.backward_projection import (
        CustormLearnedPositionalEncoding, DA_MSDeformableAttention,
        MultiScaleDeformableAttention)
    pe = CustormLearnedPositionalEncoding(40, 100, 100)
    out = pe(2, 100, 100, 'cpu')
    assert out.shape == (2, 80, 100, 100)
    assert torch.equal(out[0, :40, 3, 7], pe.col_embed.weight[7])
    assert torch.equal(out[1, 40:, 3, 7], pe.row_embed.weight[3])
    # sampling-offset bias: ring direction per head, radius p+1, anchors share
    da = DA_MSDeformableAttention(embed_dims=80, num_points=8, num_levels=1)
    b = da.sampling_offsets.bias.view(8, 1, 2, 4, 2)
    assert torch.allclose(b[0, 0, 0, :, :], torch.tensor([1.0, 0.0]).expand(4, 2))
    assert torch.allclose(b[0, 0, 1, :, :], torch.tensor([2.0, 0.0]).expand(4, 2))
    assert torch.all(da.sampling_offsets.weight == 0)
    assert torch.all(da.attention_weights.weight == 0)
    assert da.output_proj is None
    sa = MultiScaleDeformableAttention(embed_dims=80, num_levels=1)
    b = sa.sampling_offsets.bias.view(8, 1, 4, 2)
    assert torch.allclose(b[2, 0, 3], torch.tensor([0.0, 4.0]), atol=1e-6)


def test_layer_injects_batch_first_and_builds_from_config():
    """custom_base_transformer_layer.py:132-136 injects batch_first into every
    attention config; FFN embed_dims defaults to the layer's."""
    g = load_golden(CASES[0])
    from fbbev_b200.registry import build_head
    bp = build_head(bp_cfg_from_golden(g))
    layer = bp.transformer.encoder.layers[0]
    assert layer.attentions[0].batch_first is True
    assert layer.attentions[1].batch_first is True
    assert layer.attentions[1].deformable_attention.num_Z_anchors == 4
    assert layer.operation_order[0] == 'self_attn' and not layer.pre_norm
    assert len(layer.norms) == 3 and len(layer.ffns) == 1
    assert layer.ffns[0].layers[0][0].out_features == 4 * int(g["E"])


def test_linear_host_decisions():
    """Host-side rules of the tensor-core Linear path (no GPU needed): which
    LayerNorm widths fit, and that CPU tensors / autograd keep the torch path."""
    import torch.nn as nn
    from fbbev_b200.ops import linear as lin
    from fbbev_b200.view_transformation import backward_projection as bp
    assert lin.MAX_N == 192
    assert [n for n in range(16, 193, 16) if lin.ln_supported(n)] == \
        [16, 32, 48, 64, 80]
    x = torch.randn(7, 80)
    assert not bp._fused_linear_on(x)                      # CPU tensor
    with torch.no_grad():
        assert not bp._fused_linear_on(x)                  # still CPU
    fc, norm = nn.Linear(80, 80), nn.LayerNorm(80)
    res = torch.randn(7, 80)
    # folding the norm into the producer is the same arithmetic as the
    # reference's op order (here through the torch path)
    ffn = bp.FFN(embed_dims=80, feedforward_channels=160, ffn_drop=0.0)
    want = norm(ffn(x, res))
    got = ffn(x, res, post_norm=norm)
    assert torch.equal(got, want)
    with pytest.raises(Exception):
        lin.linear_fused(x, fc.weight, fc.bias)            # no CPU fallback


def test_token_major_positional_encoding_cache():
    from fbbev_b200.view_transformation.backward_projection import \
        bevformer_encoder
    enc = bevformer_encoder.__new__(bevformer_encoder)
    enc.__dict__.update({'_parameters': {}, '_modules': {}, '_buffers': {}})
    root = torch.randn(1, 8, 4, 5)
    view = root.flatten(2).permute(2, 0, 1).permute(1, 0, 2)   # (bs, nq, E)
    assert not view.is_contiguous()
    assert enc._token_major(view) is view                      # autograd on
    with torch.no_grad():
        a = enc._token_major(view)
        b = enc._token_major(root.flatten(2).permute(2, 0, 1).permute(1, 0, 2))
        assert a.is_contiguous() and torch.equal(a, view) and a is b
        root.mul_(2.0)                                         # version bump
        c = enc._token_major(root.flatten(2).permute(2, 0, 1).permute(1, 0, 2))
        assert c is not a and torch.equal(c, view)


def test_linear_train_backward_formulas(monkeypatch):
    """ops.linear.LinearTF32Function: the three gradient products of the
    training route (dx = g W, dW = g^T x, db = sum g, dresidual = g, ReLU mask
    from the saved output) against autograd, with the tcgen05 forward replaced
    by its definition so that the check runs on the CPU in float64."""
    import torch.nn.functional as F
    from fbbev_b200.ops import linear as lin

    def forward_definition(x, weight, bias=None, relu=False, residual=None,
                           **kw):
        y = F.linear(x, weight, bias)
        y = y.relu() if relu else y
        return y if residual is None else y + residual
    monkeypatch.setattr(lin, "linear_fused", forward_definition)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5, 7, 8, generator=g, dtype=torch.float64,
                    requires_grad=True)
    w = torch.randn(12, 8, generator=g, dtype=torch.float64, requires_grad=True)
    b = torch.randn(12, generator=g, dtype=torch.float64, requires_grad=True)
    r = torch.randn(5, 7, 12, generator=g, dtype=torch.float64,
                    requires_grad=True)
    assert torch.autograd.gradcheck(
        lambda x, w, b: lin.linear_train(x, w, b, relu=True), (x, w, b))
    assert torch.autograd.gradcheck(
        lambda x, w, b, r: lin.linear_train(x, w, b, residual=r), (x, w, b, r))
    assert torch.autograd.gradcheck(
        lambda x, w: lin.linear_train(x, w, None), (x, w))
    # relu + residual composes (the kernel is never asked for both)
    assert torch.autograd.gradcheck(
        lambda x, w, b, r: lin.linear_train(x, w, b, relu=True, residual=r),
        (x, w, b, r))
