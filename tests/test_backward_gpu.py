"""GPU parity tests of the backward projection (depth-aware spatial
cross-attention).  CUDA results come through the C ABI and are compared with

* the golden fixtures recorded from the reference's own Python classes
  (mmcv's MSDA kernel served by the C oracle) -- tests/golden/b_*.npz;
* the C oracle / torch grid_sample restatement on seeded random inputs.

Tolerance 1e-4 absolute (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from bp_common import build_bp, cam_params
from conftest import load_golden
from test_oracle import _rand_msda

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ATOL = 1e-4
CASES = ["b_bp_e80_1lvl", "b_bp_e64_3lvl"]


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("ch", [3, 4, 8, 10, 16, 20, 32, 40, 64])
def test_msda_forward_vs_oracle(oracle_cpu, ch):
    from fbbev_b200.ops.ms_deform_attn import ms_deform_attn_forward
    value, shapes, lsi, loc, attw = _rand_msda(ch, ch=ch, nq=37)
    want = oracle_cpu.msda_fwd(value.numpy(), shapes.numpy(), lsi.numpy(),
                               loc.numpy(), attw.numpy())
    got = ms_deform_attn_forward(value.to(DEV), shapes.to(DEV), lsi.to(DEV),
                                 loc.to(DEV), attw.to(DEV), 64)
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=0, atol=2e-5)


def test_msda_forward_backward_only_config_full_size(oracle_cpu):
    """BASELINE.json configs[3]: 200x200 BEV queries, 4-scale 8-head
    MSDeformAttn (embed 256 -> 32 channels per head), full size."""
    from fbbev_b200.ops.ms_deform_attn import ms_deform_attn_forward
    value, shapes, lsi, loc, attw = _rand_msda(
        11, bs=1, nq=200 * 200, heads=8, ch=32,
        shapes=((32, 88), (16, 44), (8, 22), (4, 11)), points=4)
    attw = attw / attw.flatten(3).sum(-1)[..., None, None]
    want = oracle_cpu.msda_fwd(value.numpy(), shapes.numpy(), lsi.numpy(),
                               loc.numpy(), attw.numpy())
    got = ms_deform_attn_forward(value.to(DEV), shapes.to(DEV), lsi.to(DEV),
                                 loc.to(DEV), attw.to(DEV), 64)
    assert got.shape == (1, 200 * 200, 256)
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=0, atol=ATOL)


def test_msda_edge_locations(oracle_cpu):
    """Locations on / outside the border, NaN and huge values: zero padding and
    the (-1, H) x (-1, W) validity window of the im2col algorithm."""
    from fbbev_b200.ops.ms_deform_attn import ms_deform_attn_forward
    value, shapes, lsi, loc, attw = _rand_msda(3, bs=1, nq=16, points=4)
    special = torch.tensor([0.0, 1.0, -1e-3, 1.0 + 1e-3, -0.2, 1.2, 1e9, -1e9,
                            0.5 / 7, 1 - 0.5 / 7])
    flat = loc.view(-1, 2)
    flat[:len(special), 0] = special
    flat[len(special):2 * len(special), 1] = special
    flat[40, 0] = float('nan')
    flat[41, 1] = float('inf')
    want = oracle_cpu.msda_fwd(value.numpy(), shapes.numpy(), lsi.numpy(),
                               loc.numpy(), attw.numpy())
    got = ms_deform_attn_forward(value.to(DEV), shapes.to(DEV), lsi.to(DEV),
                                 loc.to(DEV), attw.to(DEV))
    assert np.isfinite(want).all()
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=0, atol=2e-5)


@pytest.mark.parametrize("ch", [10, 12])
def test_msda_backward_vs_oracle(oracle_cpu, ch):
    from fbbev_b200.ops.ms_deform_attn import \
        MultiScaleDeformableAttnFunction_fp32
    value, shapes, lsi, loc, attw = _rand_msda(5, ch=ch, nq=29)
    v = value.to(DEV).requires_grad_()
    l = loc.to(DEV).requires_grad_()
    a = attw.to(DEV).requires_grad_()
    out = MultiScaleDeformableAttnFunction_fp32.apply(v, shapes.to(DEV),
                                                      lsi.to(DEV), l, a, 64)
    go = torch.randn(out.shape, generator=torch.Generator().manual_seed(6))
    out.backward(go.to(DEV))
    gv, gl, ga = oracle_cpu.msda_bwd(value.numpy(), shapes.numpy(),
                                     lsi.numpy(), loc.numpy(), attw.numpy(),
                                     go.numpy())
    np.testing.assert_allclose(v.grad.cpu().numpy(), gv, rtol=0, atol=5e-5)
    np.testing.assert_allclose(a.grad.cpu().numpy(), ga, rtol=0, atol=5e-5)
    np.testing.assert_allclose(l.grad.cpu().numpy(), gl, rtol=0, atol=5e-4)


def test_fused_self_attention_core_vs_unfused(oracle_cpu):
    """fbbev_msda_fused_fwd == softmax + loc arithmetic + ms_deform_attn."""
    from fbbev_b200.ops.ms_deform_attn import ms_deform_attn_fused
    value, shapes, lsi, loc, attw = _rand_msda(7, ch=10, nq=50, points=4)
    bs, nq, heads, levels, points, _ = loc.shape
    g = torch.Generator().manual_seed(8)
    ref = torch.rand(bs, nq, levels, 2, generator=g)
    off = torch.randn(bs, nq, heads, levels, points, 2, generator=g) * 2
    logits = torch.randn(bs, nq, heads, levels, points, generator=g)
    wh = torch.stack([shapes[:, 1], shapes[:, 0]], -1).float()
    loc = ref[:, :, None, :, None, :] + off / wh[None, None, None, :, None, :]
    w = logits.flatten(3).softmax(-1).view_as(logits)
    want = oracle_cpu.msda_fwd(value.numpy(), shapes.numpy(), lsi.numpy(),
                               loc.numpy(), w.numpy())
    got = ms_deform_attn_fused(value.to(DEV), shapes.to(DEV), lsi.to(DEV),
                               ref.to(DEV), off.to(DEV), logits.to(DEV))
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=0, atol=2e-5)


@pytest.mark.parametrize("case", CASES)
def test_da_sca_vs_reference_golden(case):
    """One DA_SpatialCrossAttention.forward on the reference's recorded inputs
    (its own query / masks / reference points) == the reference's output."""
    g, bp = build_bp(case, DEV)
    sca = bp.transformer.encoder.layers[0].attentions[1]
    with torch.no_grad():
        out = sca(t(g["sca_query"]), t(g["sca_key"]), t(g["sca_key"]), None,
                  query_pos=t(g["sca_query_pos"]),
                  reference_points_cam=t(g["reference_points_cam"]),
                  spatial_shapes=t(g["spatial_shapes"]),
                  level_start_index=t(g["level_start_index"]),
                  bev_query_depth=t(g["bev_query_depth"]),
                  pred_img_depth=t(g["depth"]),
                  per_cam_mask_list=t(g["per_cam_mask"]))
    np.testing.assert_allclose(out.cpu().numpy(), g["sca_out"], rtol=0,
                               atol=ATOL)


@pytest.mark.parametrize("case", CASES)
def test_da_sca_rebatch_path_equals_fused(case):
    """The reference-shaped re-batching path (used when a bev_mask is given)
    and the fused kernel agree; an all-true bev_mask must change nothing."""
    g, bp = build_bp(case, DEV)
    sca = bp.transformer.encoder.layers[0].attentions[1]
    kw = dict(query_pos=t(g["sca_query_pos"]),
              reference_points_cam=t(g["reference_points_cam"]),
              spatial_shapes=t(g["spatial_shapes"]),
              level_start_index=t(g["level_start_index"]),
              bev_query_depth=t(g["bev_query_depth"]),
              pred_img_depth=t(g["depth"]),
              per_cam_mask_list=t(g["per_cam_mask"]))
    q, k = t(g["sca_query"]), t(g["sca_key"])
    with torch.no_grad():
        fused = sca(q, k, k, None, **kw)
        mask = torch.ones(q.shape[:2], dtype=torch.bool, device=DEV)
        sca.rebatch_bev_mask = True       # the literal per-camera loops
        rebatch = sca(q, k, k, None, bev_mask=mask, **kw)
        sca.rebatch_bev_mask = False      # device-side mask fold + fused kernel
        folded = sca(q, k, k, None, bev_mask=mask, **kw)
    # (the camera-resident kernel adds the cameras' contributions in no fixed
    # order: last-bit differences between two runs are expected)
    assert (folded - fused).abs().max().item() <= 2e-6
    assert (fused - rebatch).abs().max().item() <= ATOL
    np.testing.assert_allclose(rebatch.cpu().numpy(), g["sca_out"], rtol=0,
                               atol=ATOL)


@pytest.mark.parametrize("case", CASES)
def test_backward_projection_vs_reference_golden(case):
    """Whole BackwardProjection.forward (embedding + pos-enc + BEVFormer +
    encoder geometry + self-attn + LN + DA-SCA + LN + FFN + LN) with the
    reference's weights == the reference's output."""
    g, bp = build_bp(case, DEV)
    n_lvl = len(g["level_shapes"])
    mlvl = [t(g[f"feat{i}"]) for i in range(n_lvl)]
    with torch.no_grad():
        out = bp(mlvl, None, lss_bev=t(g["lss_bev"]),
                 cam_params=cam_params(g, DEV), pred_img_depth=t(g["depth"]))
    assert tuple(out.shape) == g["out"].shape
    # measured 1.4e-5 / 1.1e-5 (tools/micro/bp_stage_err.py); bar = north_star
    np.testing.assert_allclose(out.cpu().numpy(), g["out"], rtol=0, atol=ATOL)
    # geometry on the device agrees with the reference's CPU tensors
    enc = bp.transformer.encoder
    ref_3d = enc.get_reference_points(int(g["bev_h"]), int(g["bev_w"]),
                                      dim='3d', device=DEV)
    _, ref_cam, mask, depth = enc.point_sampling(
        ref_3d, enc.pc_range, None, cam_params=cam_params(g, DEV))
    agree = (mask.cpu().numpy() == g["per_cam_mask"]).mean()
    assert agree >= 0.999
    vis = g["per_cam_mask"]
    np.testing.assert_allclose(ref_cam.cpu().numpy()[vis],
                               g["reference_points_cam"][vis], atol=1e-4)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("which", ["a", "b"])
@pytest.mark.parametrize("rebatch", [False, True])
def test_backward_projection_bev_mask_vs_reference_golden(case, which, rebatch):
    """``bev_mask`` (spatial_cross_attention_depth.py:156-169): outputs of the
    reference's own BackwardProjection with (a) a random half of the BEV cells
    masked and (b) a mask that empties camera 0's list, which triggers the
    reference's empty-camera rule (:166-167: the camera's first visible query
    is processed although masked, and not counted, :213-214).  The default
    route folds the mask on the device (``fbbev_bev_mask_fold``) and runs the
    fused kernels without any host synchronisation; ``rebatch`` runs the
    reference-shaped loops."""
    g, bp = build_bp(case, DEV)
    sca = bp.transformer.encoder.layers[0].attentions[1]
    sca.rebatch_bev_mask = rebatch
    n_lvl = len(g["level_shapes"])
    mlvl = [t(g[f"feat{i}"]) for i in range(n_lvl)]
    bev_h, bev_w = int(g["bev_h"]), int(g["bev_w"])
    mask = t(g[f"bev_mask_{which}"]).view(-1, bev_h, bev_w)
    with torch.no_grad():
        out = bp(mlvl, None, lss_bev=t(g["lss_bev"]),
                 cam_params=cam_params(g, DEV), pred_img_depth=t(g["depth"]),
                 bev_mask=mask)
    np.testing.assert_allclose(out.cpu().numpy(), g[f"out_bev_mask_{which}"],
                               rtol=0, atol=ATOL)


def test_bev_mask_fold_encoding():
    """fbbev_bev_mask_fold against its definition: 1 where mask & bev_mask;
    a (camera, sample) pair left without any query gets 2 on the anchors of the
    first query it sees at all; pairs that see nothing stay empty."""
    from fbbev_b200.ops.ms_deform_attn import bev_mask_fold
    gen = torch.Generator().manual_seed(5)
    n_cams, bs, nq, Z = 5, 3, 1000, 4
    mask = torch.rand(n_cams, bs, nq, Z, generator=gen) > 0.7
    mask[3, 1] = False                          # camera 3 of sample 1: blind
    bev = torch.rand(bs, nq, generator=gen) > 0.5
    bev[2] &= ~mask[1, 2].any(-1)               # camera 1 of sample 2: emptied
    bev[0] = False                              # every camera of sample 0
    got = bev_mask_fold(mask.to(DEV), bev.to(DEV)).cpu()
    want = (mask & bev[None, :, :, None]).to(torch.uint8)
    for n in range(n_cams):
        for b in range(bs):
            if not want[n, b].any() and mask[n, b].any():
                q0 = int(mask[n, b].any(-1).nonzero()[0])
                want[n, b, q0] = mask[n, b, q0].to(torch.uint8) * 2
    assert torch.equal(got, want)
    assert (got == 2).any() and not got[3, 1].any()


def _bp_module(bev, E, levels, input_size, B, points=8, dbound=(2.0, 42.0, 0.5),
               z_step=1.6, seed=3):
    """A BackwardProjection of the given size with perturbed init weights."""
    from fbbev_b200.registry import build_head
    from bp_common import bp_cfg_from_golden
    g = dict(E=E, bev_h=bev[0], bev_w=bev[1], level_shapes=levels,
             pc_range=[-40, -40, -1.0, 40, 40, 5.4],
             grid_x=[-40, 40, 80.0 / bev[1]], grid_y=[-40, 40, 80.0 / bev[0]],
             grid_z=[-1, 5.4, z_step], input_size=input_size, dbound=dbound)
    cfg = bp_cfg_from_golden(g)
    cfg['transformer']['encoder']['transformerlayers']['attn_cfgs'][1][
        'deformable_attention']['num_points'] = points
    torch.manual_seed(seed)
    bp = build_head(cfg)
    bp.init_weights()
    gen = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for p in bp.parameters():
            p.add_(torch.randn(p.shape, generator=gen) * 0.02)
    return bp.to(DEV).eval()


@pytest.mark.parametrize("case", CASES + ["bench", "b16", "cfg3", "onecam"])
def test_fused_point_sampling_bit_exact(case):
    """fbbev_point_sampling (the encoder's default) == the reference's eager
    chain ``point_sampling`` (bevformer_encoder.py:92-120) on this device, bit
    for bit: reference_points_cam, depth and the visibility mask."""
    from fbbev_b200 import synthetic
    from test_forward_gpu import _augment
    if case in CASES:
        g, bp = build_bp(case, DEV)
        cams = cam_params(g, DEV)
        bev = (int(g["bev_h"]), int(g["bev_w"]))
    else:
        B, N, bev, inp = {"bench": (1, 6, (200, 200), (256, 704)),
                          "b16": (16, 6, (100, 100), (256, 704)),
                          "cfg3": (1, 6, (200, 200), (512, 1408)),
                          "onecam": (1, 1, (64, 64), (256, 704))}[case]
        bp = _bp_module(bev, 64, [(4, 11)], inp, B)
        cams = synthetic.make_cam_params(B, N, inp, device=DEV, jitter=1.0)
        if case != "bench":
            cams = _augment(cams)
    enc = bp.transformer.encoder
    ref_cam, mask, depth = enc.point_sampling_fused(cams)
    ref_3d = enc.get_reference_points(bev[0], bev[1], dim='3d', device=DEV)
    _, ref_cam_t, mask_t, depth_t = enc.point_sampling(ref_3d, enc.pc_range,
                                                       None, cam_params=cams)
    assert mask.dtype == torch.bool and mask.any()
    assert torch.equal(mask, mask_t)
    assert torch.equal(ref_cam, ref_cam_t)
    assert torch.equal(depth, depth_t)


@pytest.mark.parametrize("case", CASES)
def test_fused_point_sampling_vs_torch_and_golden(case):
    """fbbev_point_sampling (the encoder's default) against the eager-PyTorch
    chain and the reference's recorded tensors."""
    g, bp = build_bp(case, DEV)
    enc = bp.transformer.encoder
    cams = cam_params(g, DEV)
    ref_cam, mask, depth = enc.point_sampling_fused(cams)
    ref_3d = enc.get_reference_points(int(g["bev_h"]), int(g["bev_w"]),
                                      dim='3d', device=DEV)
    _, ref_cam_t, mask_t, depth_t = enc.point_sampling(ref_3d, enc.pc_range,
                                                       None, cam_params=cams)
    assert ref_cam.shape == ref_cam_t.shape and depth.shape == depth_t.shape
    assert mask.dtype == torch.bool and mask.shape == mask_t.shape
    assert (mask == mask_t).float().mean().item() >= 0.9999
    both = (mask & mask_t).cpu().numpy()
    np.testing.assert_allclose(ref_cam.cpu().numpy()[both],
                               ref_cam_t.cpu().numpy()[both], atol=2e-5)
    np.testing.assert_allclose(depth.cpu().numpy()[both],
                               depth_t.cpu().numpy()[both], rtol=1e-5,
                               atol=1e-4)
    assert (mask.cpu().numpy() == g["per_cam_mask"]).mean() >= 0.999
    vis = g["per_cam_mask"] & mask.cpu().numpy()
    np.testing.assert_allclose(ref_cam.cpu().numpy()[vis],
                               g["reference_points_cam"][vis], atol=1e-4)


class _GridSampleMSDA:
    """Stand-in for MultiScaleDeformableAttnFunction_fp32 on CPU: mmcv's
    documented pure-PyTorch equivalent, differentiable by autograd."""

    @staticmethod
    def apply(value, shapes, lstart, loc, attw, im2col_step):
        from oracle.torch_ref import multi_scale_deformable_attn_pytorch
        return multi_scale_deformable_attn_pytorch(value, shapes, loc, attw)


@pytest.mark.parametrize("case", CASES)
def test_training_gradients_reach_every_input(case, monkeypatch):
    """With autograd recording, BackwardProjection must take the differentiable
    route (the fused sampling kernels are forward-only): same forward result,
    and gradients w.r.t. every parameter, the image features, the depth
    distribution and lss_bev equal those of a pure-PyTorch (grid_sample)
    evaluation of the same module on the CPU -- what the reference gets through
    MultiScaleDeformableAttnFunction (multi_scale_deformable_attn_function.py:
    142-172)."""
    from fbbev_b200.ops import ms_deform_attn as ops

    def run(dev):
        g, bp = build_bp(case, dev)
        n_lvl = len(g["level_shapes"])
        tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        mlvl = [tt(g[f"feat{i}"]).requires_grad_() for i in range(n_lvl)]
        depth = tt(g["depth"]).requires_grad_()
        lss = tt(g["lss_bev"]).requires_grad_()
        out = bp(mlvl, None, lss_bev=lss, cam_params=cam_params(g, dev),
                 pred_img_depth=depth)
        wgt = torch.randn(out.shape, generator=torch.Generator().manual_seed(9))
        (out * wgt.to(dev)).sum().backward()
        grads = {n: p.grad for n, p in bp.named_parameters()}
        grads.update({f"feat{i}": f.grad for i, f in enumerate(mlvl)})
        grads.update(depth=depth.grad, lss=lss.grad)
        return g, out.detach(), grads

    g, out_gpu, grads_gpu = run(DEV)
    np.testing.assert_allclose(out_gpu.cpu().numpy(), g["out"], rtol=0,
                               atol=ATOL)
    monkeypatch.setattr(ops, "MultiScaleDeformableAttnFunction_fp32",
                        _GridSampleMSDA)
    _, out_cpu, grads_cpu = run("cpu")
    np.testing.assert_allclose(out_cpu.numpy(), g["out"], rtol=0, atol=ATOL)
    used = 0
    for name, gc in grads_cpu.items():
        gg = grads_gpu[name]
        if gc is None:      # e.g. cams_embeds * 0 keeps a zero gradient
            continue
        assert gg is not None, f"no gradient reached {name}"
        scale = max(1.0, float(gc.abs().max()))
        err = float((gg.cpu() - gc).abs().max())
        assert err <= 5e-4 * scale, (name, err, scale)
        used += int(float(gc.abs().max()) > 0)
    # value_proj / sampling_offsets / attention_weights of BOTH attentions, the
    # image features and the depth distribution all carry signal
    for must in ("depth", "feat0", "lss"):
        assert float(grads_gpu[must].abs().max()) > 0
    for frag in ("attentions.0.value_proj.weight",
                 "attentions.0.sampling_offsets.weight",
                 "attentions.1.deformable_attention.value_proj.weight",
                 "attentions.1.deformable_attention.sampling_offsets.weight",
                 "attentions.1.deformable_attention.attention_weights.weight"):
        hit = [n for n in grads_gpu if n.endswith(frag)]
        assert hit and float(grads_gpu[hit[0]].abs().max()) > 0, frag
    assert used > 10


# ------------------------------------------------ full-size configurations --
def _bp_inputs(B, E, levels, DC, bev, inp, seed=17):
    from fbbev_b200 import synthetic
    g = torch.Generator().manual_seed(seed)
    cams = synthetic.make_cam_params(B, 6, inp, jitter=1.0, seed=seed)
    mlvl = [torch.randn(B, 6, E, h, w, generator=g) for h, w in levels]
    H0, W0 = levels[0]
    depth = torch.randn(B, 6, DC, H0, W0, generator=g).softmax(2)
    lss = torch.randn(B, E, *bev, generator=g) * 0.1
    return cams, mlvl, depth, lss


def _bp_vs_cpu_oracle(bp, cams, mlvl, depth, lss):
    """GPU module (fused kernels through the C ABI) against the same module's
    host logic around the oracle's attention cores (oracle/backward_ref.py:
    the reference's re-batching algorithm, MSDA by the C oracle)."""
    import copy
    from oracle.backward_ref import backward_projection_cpu
    with torch.no_grad():
        got = bp(
            [f.to(DEV) for f in mlvl], None, lss_bev=lss.to(DEV),
            cam_params=[c.to(DEV) for c in cams], pred_img_depth=depth.to(DEV))
    bp_cpu = copy.deepcopy(bp).cpu()
    want = backward_projection_cpu(bp_cpu, mlvl, lss, list(cams), depth)
    return got.cpu().numpy(), want.numpy()


def test_backward_only_config_full_size():
    """BASELINE.json configs[3] THROUGH the plugin: 200x200 BEV queries,
    embed 256 (8 heads x 32 channels), 4 feature levels (32x88 ... 4x11),
    8 points / 4 Z anchors, 80 depth bins: fbbev_msda_fused_fwd and
    fbbev_da_sca_fwd at <32, 4> plus the tcgen05 Linears at K = N = 256."""
    levels = [(32, 88), (16, 44), (8, 22), (4, 11)]
    bev, inp = (200, 200), (512, 1408)
    bp = _bp_module(bev, 256, levels, inp, 1)
    cams, mlvl, depth, lss = _bp_inputs(1, 256, levels, 80, bev, inp)
    got, want = _bp_vs_cpu_oracle(bp, cams, mlvl, depth, lss)
    assert got.shape == (1, 256, 200, 200)
    np.testing.assert_allclose(got, want, rtol=0, atol=ATOL)


def test_sixteen_frame_config_full_size():
    """BASELINE.json configs[2]: 16 frames (B = 16) of the FB-OCC R50 geometry
    through BackwardProjection in ONE call (200x200 queries, E = 80), against
    the CPU oracle; plus frame independence: frame b of
    the batched call == the same frame run alone."""
    levels = [(16, 44)]
    bev, inp = (200, 200), (256, 704)
    bp = _bp_module(bev, 80, levels, inp, 16)
    cams, mlvl, depth, lss = _bp_inputs(16, 80, levels, 80, bev, inp)
    got, want = _bp_vs_cpu_oracle(bp, cams, mlvl, depth, lss)
    assert got.shape == (16, 80, 200, 200)
    np.testing.assert_allclose(got, want, rtol=0, atol=ATOL)
    for b in (0, 9, 15):
        with torch.no_grad():
            one = bp([f[b:b + 1].to(DEV) for f in mlvl], None,
                     lss_bev=lss[b:b + 1].to(DEV),
                     cam_params=[c[b:b + 1].to(DEV) for c in cams],
                     pred_img_depth=depth[b:b + 1].to(DEV))
        np.testing.assert_allclose(one.cpu().numpy()[0], got[b], rtol=0,
                                   atol=1e-5)


@pytest.mark.parametrize("bs,nq,hw", [(1, 997, (16, 44)), (2, 4100, (8, 22)),
                                      (3, 64, (4, 6))])
def test_da_sca_smem_kernel_vs_global_kernel_and_oracle(bs, nq, hw):
    """The camera-resident (shared-memory) cross-attention kernel against the
    global-memory kernel of the same entry point and against the CPU oracle, on
    random inputs with the awkward cases: a camera that sees nothing, queries
    seen by 0 / 1 / 3+ cameras, reference points outside the image, ragged
    query counts (not a multiple of 32)."""
    from fbbev_b200 import _lib
    from fbbev_b200.ops.ms_deform_attn import da_spatial_cross_attention_core
    from oracle.backward_ref import da_sca_core_cpu
    g = torch.Generator().manual_seed(31 + nq)
    N, heads, ch, L, P, Z, DC = 6, 8, 10, 1, 8, 4, 40
    H, W = hw
    value = torch.randn(bs * N, H * W, heads, ch, generator=g)
    depth_prob = torch.randn(bs * N, H * W, DC, generator=g).softmax(-1)
    ref = torch.rand(N, bs, nq, Z, 2, generator=g) * 1.4 - 0.2
    qdepth = torch.rand(N, bs, nq, Z, generator=g) * 50 - 2
    mask = torch.rand(N, bs, nq, Z, generator=g) < 0.12
    mask[3] = False                       # camera 3 sees nothing
    mask[:, :, 5] = True                  # query 5: every camera but 3
    mask[3, :, 5] = False
    off = torch.randn(bs, nq, heads, L, P, 2, generator=g) * 2
    lg = torch.randn(bs, nq, heads, L, P, generator=g)
    ss = torch.tensor([[H, W]])
    ls = torch.tensor([0])
    dbound = [2.0, 42.0, 1.0]
    want = da_sca_core_cpu(value, depth_prob, ref, qdepth, mask, off, lg, ss,
                           ls, dbound, Z).numpy()
    args = [a.to(DEV) for a in (value, depth_prob, ref, qdepth, mask, off, lg,
                                ss, ls)]
    launches0 = _lib.lib().fbbev_debug_launch_count()
    got = da_spatial_cross_attention_core(*args, dbound, Z)
    assert _lib.lib().fbbev_debug_launch_count() - launches0 == 2  # smem path
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=0, atol=2e-5)
    # same entry point without a workspace -> global-memory kernel
    L_ = _lib.lib()
    out = torch.empty_like(got)
    v, dp, r, qd, mk, o, l, s, lsi = args
    rc = L_.fbbev_da_sca_fwd(
        _lib.ptr(v), _lib.ptr(dp), _lib.ptr(r), _lib.ptr(qd),
        _lib.ptr(mk.to(torch.uint8)), _lib.ptr(o), _lib.ptr(l), _lib.ptr(s),
        _lib.ptr(lsi), _lib.c_floats(dbound), bs, N, nq, H * W, heads, ch, L, P,
        Z, DC, _lib.ptr(out), None, 0, 0, _lib.stream_ptr(torch.device(DEV)))
    assert rc == 0
    torch.cuda.synchronize()
    np.testing.assert_allclose(out.cpu().numpy(), want, rtol=0, atol=2e-5)
    assert (out - got).abs().max().item() <= 2e-6


@pytest.mark.parametrize("bs,hw,E", [(1, (200, 200), 80), (3, (13, 7), 64),
                                     (2, (5, 33), 20), (2, (6, 22), 20),
                                     (3, (10, 14), 64)])
def test_bev_query_init_kernel(bs, hw, E):
    """fbbev_bev_query_init == embedding.unsqueeze(1).repeat(1, bs, 1) +
    lss_bev.flatten(2).permute(2, 0, 1) (backward_projection.py:93-97), bit for
    bit (one fp32 add), ragged tile edges included."""
    from fbbev_b200.ops.ms_deform_attn import bev_query_init
    g = torch.Generator().manual_seed(bs * 7 + E)
    nq = hw[0] * hw[1]
    emb = torch.randn(nq, E, generator=g).to(DEV)
    lss = torch.randn(bs, E, *hw, generator=g).to(DEV)
    got = bev_query_init(emb, lss)
    want = emb.unsqueeze(1).repeat(1, bs, 1) + lss.flatten(2).permute(2, 0, 1)
    assert got.shape == want.shape == (nq, bs, E)
    assert torch.equal(got, want)
    assert got.permute(1, 0, 2).is_contiguous()
    # and the way back: (bs, nq, E) tokens -> (bs, E, h, w) map
    from fbbev_b200.ops.ms_deform_attn import tokens_to_map
    tok = got.permute(1, 0, 2).contiguous()
    back = tokens_to_map(tok, *hw)
    ref = tok.permute(0, 2, 1).reshape(bs, E, *hw)
    if nq % 4 == 0 and E % 4 == 0:
        assert back is not None and torch.equal(back, ref)
        slot = torch.empty(ref.shape, device=DEV)
        assert tokens_to_map(tok, *hw, out=slot) is slot and torch.equal(slot, ref)
    else:
        assert back is None


@pytest.mark.parametrize("B,bev,grid,inp", [
    (1, (200, 200), "fbocc_200", (256, 704)),     # the bench workload
    (2, (100, 100), "fbocc_shipped", (256, 704)),  # shipped FB-OCC grid, 2 frames
])
def test_forward_backward_readd_glue(B, bev, grid, inp):
    """FBOCC.extract_img_bev_feat's three lines around the projections
    (fbocc.py:339, 357-366) with the volume written once: mean(-1) from the
    interval sums, the re-add fused into the dense write == the literal
    sequence forward -> mean(-1) -> backward -> refined[..., None] + bev."""
    from fbbev_b200 import synthetic
    from fbbev_b200.view_transformation.forward_projection import (
        LSSViewTransformerFunction3D, forward_backward_readd)
    vt = LSSViewTransformerFunction3D(synthetic.GRID_CONFIGS[grid], inp, 16)
    H, W = inp[0] // 16, inp[1] // 16
    bp = _bp_module(bev, 80, [(H, W)], inp, B,
                    z_step=1.6, dbound=(2.0, 42.0, 0.5))
    cams = synthetic.make_cam_params(B, 6, inp, device=DEV, jitter=1.0)
    depth, feat = synthetic.make_depth_feat(B, 6, vt.D, H, W, 80, device=DEV)
    with torch.no_grad():
        bev_feat = vt(cams, feat, depth)                       # (B,C,Y,X,Z)
        lss = bev_feat.mean(-1)
        refined = bp([feat], None, lss_bev=lss, cam_params=cams,
                     pred_img_depth=depth)
        want = refined[..., None] + bev_feat
        dv = vt.forward_deferred(cams, feat, depth)
        assert dv is not None
        np.testing.assert_allclose(dv.mean_z().cpu().numpy(), lss.cpu().numpy(),
                                   rtol=0, atol=1e-5)
        assert torch.equal(dv.materialize().permute(0, 1, 3, 4, 2), bev_feat)
        got, refined2 = forward_backward_readd(vt, bp, cams, feat, depth)
        got_noadd, _ = forward_backward_readd(vt, bp, cams, feat, depth,
                                              readd=False)
    assert got.shape == want.shape
    np.testing.assert_allclose(refined2.cpu().numpy(), refined.cpu().numpy(),
                               rtol=0, atol=ATOL)
    assert float((got - want).abs().max()) <= ATOL
    assert torch.equal(got_noadd, refined2)
