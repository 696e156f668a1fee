"""GPU parity tests of the callers either side of the pooling op (SURVEY.md
section 8 f3 / f4): the BEVDet-lineage transformers with the depth-threshold
sparsification, and the tail of the depth net as the producer of ``depth`` /
``feat``.  CUDA results come through the C ABI and are compared with

* the golden fixtures recorded from the reference's own classes
  (tests/golden/l_*.npz, generator tests/golden/gen_golden_lineage.py);
* the CPU oracle (oracle/lift_ref.py, oracle/cpu.py) on seeded random inputs.

Integer index path bit-exact; float outputs <= 1e-4 absolute (north_star),
the softmax itself <= 2e-6."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ATOL = 1e-4


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.fixture(autouse=True)
def _fp32_convolutions():
    """The goldens were produced in fp32 on CPU.  The `depth_net` / `context_conv`
    convolutions in front of the path are library calls (cuDNN), which default
    to TF32 on this GPU: switch that off so the comparison sees the path's own
    arithmetic."""
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


# ------------------------------------------------------------- lift tail ---
@pytest.mark.parametrize("bn,d,c,h,w", [(6, 80, 80, 16, 44), (1, 59, 64, 64, 176),
                                        (12, 118, 80, 32, 88), (3, 7, 5, 3, 11),
                                        (2, 10, 33, 4, 9)])
def test_lift_tail_vs_oracle(bn, d, c, h, w):
    """fbbev_lift_tail_fwd == softmax(dim=1) + permute(0,2,3,1).contiguous()
    on the shapes of BASELINE.json's configs and on ragged ones (H*W and C not
    multiples of the 32-wide tiles); the two inputs given as channel slices of
    ONE tensor, as the BEVDet-lineage forward passes them."""
    from fbbev_b200.ops.lift_tail import lift_tail
    from oracle import lift_ref
    gen = torch.Generator().manual_seed(bn * 1000 + d)
    x = torch.randn(bn, d + c + 3, h, w, generator=gen) * 4.0
    xd = x.to(DEV)
    depth, feat = lift_tail(xd[:, :d], xd[:, d:d + c])
    want_d, want_f = lift_ref.lift_tail(x[:, :d].numpy(), x[:, d:d + c].numpy())
    assert depth.shape == (bn, d, h, w) and feat.shape == (bn, h, w, c)
    np.testing.assert_allclose(depth.cpu().numpy(), want_d, rtol=0, atol=2e-6)
    assert np.array_equal(feat.cpu().numpy(), want_f)          # a pure copy
    # against torch on the same device
    ref = xd[:, :d].softmax(dim=1)
    assert (depth - ref).abs().max().item() <= 2e-6
    assert abs(float(depth.sum(1).mean()) - 1.0) < 1e-5
    # separate contiguous tensors take the same route
    d2, f2 = lift_tail(xd[:, :d].contiguous(), xd[:, d:d + c].contiguous())
    assert torch.equal(d2, depth) and torch.equal(f2, feat)


def test_lift_tail_extreme_logits():
    """+-inf-free extremes: a huge logit owns the whole probability, equal
    logits give exactly 1 / D (the `uniform` mode, depth_digit * 0)."""
    from fbbev_b200.ops.lift_tail import lift_tail
    x = torch.zeros(2, 8, 3, 5, device=DEV)
    c = torch.zeros(2, 4, 3, 5, device=DEV)
    depth, _ = lift_tail(x, c)
    assert torch.equal(depth, torch.full_like(depth, 0.125))
    x[:, 3] = 1e4
    x[:, 5] = -1e4
    depth, _ = lift_tail(x, c)
    assert torch.equal(depth[:, 3], torch.ones_like(depth[:, 3]))
    assert float(depth.sum()) == 2 * 3 * 5


def test_lift_tail_gradients():
    from fbbev_b200.ops.lift_tail import lift_tail
    gen = torch.Generator().manual_seed(4)
    lg = torch.randn(3, 9, 4, 6, generator=gen).to(DEV).requires_grad_()
    cx = torch.randn(3, 5, 4, 6, generator=gen).to(DEV).requires_grad_()
    wd = torch.randn(3, 9, 4, 6, generator=gen).to(DEV)
    wf = torch.randn(3, 4, 6, 5, generator=gen).to(DEV)
    depth, feat = lift_tail(lg, cx)
    ((depth * wd).sum() + (feat * wf).sum()).backward()
    g1, g2 = lg.grad.clone(), cx.grad.clone()
    lg.grad = cx.grad = None
    ((lg.softmax(1) * wd).sum() + (cx.permute(0, 2, 3, 1) * wf).sum()).backward()
    assert (g1 - lg.grad).abs().max().item() <= 1e-6
    assert torch.equal(g2, cx.grad)


# ------------------------------------------- depth-threshold index builder --
@pytest.mark.parametrize("fused", [True, False])
def test_prepare_sparse_vs_oracle_exact(oracle_cpu, fused):
    """Index with the 0.01 depth threshold (necks/view_transformer.py:556-557)
    on BASELINE configs[1]'s rig: bit-exact against the oracle, through the
    coordinate route and the fused-geometry route."""
    from fbbev_b200 import synthetic
    from fbbev_b200.view_transformation.bevdet_lineage import \
        LSSViewTransformer2
    from oracle import lift_ref
    grid = synthetic.GRID_CONFIGS["fbocc_200"]
    vt = LSSViewTransformer2(grid, (256, 704), 16, in_channels=8,
                             out_channels=8)
    vt.fused_geometry = fused
    B = 2
    cam = synthetic.make_cam_params(B, 6, (256, 704), device=DEV, jitter=1.0,
                                    seed=3)
    gen = torch.Generator().manual_seed(3)
    depth = (torch.randn(B, 6, vt.D, 16, 44, generator=gen) * 3).softmax(2)
    idx = vt._build_index(cam, depth.to(DEV))
    n_kept, n_int = idx.counts.tolist()
    coor = vt.get_lidar_coor(*cam).cpu().numpy()
    rb, rd, rf, st, ln = lift_ref.prepare_sparse(
        coor, vt.grid_lower_bound.numpy(), vt.grid_interval.numpy(),
        vt.grid_size.numpy(), depth.numpy(), 0.01)
    full = oracle_cpu.voxel_prepare(coor, vt.grid_lower_bound.numpy(),
                                    vt.grid_interval.numpy(),
                                    vt.grid_size.numpy())
    assert 0.05 * len(full[0]) < len(rb) < 0.9 * len(full[0])
    assert n_kept == len(rb) and n_int == len(st)
    assert np.array_equal(idx.ranks_bev[:n_kept].cpu().numpy(), rb)
    assert np.array_equal(idx.ranks_depth[:n_kept].cpu().numpy(), rd)
    assert np.array_equal(idx.ranks_feat[:n_kept].cpu().numpy(), rf)
    assert np.array_equal(idx.interval_starts[:n_int].cpu().numpy(), st)
    assert np.array_equal(idx.interval_lengths[:n_int].cpu().numpy(), ln)


# ------------------------------------------------ BEVDet-lineage modules ---
class _Recorded(torch.nn.Module):
    """Stands for the depth-net BODY of LSSViewTransformerBEVDepth (out of
    scope): replays the output the reference's net produced."""

    def __init__(self, outs):
        super().__init__()
        self.outs = list(outs)

    def forward(self, x, mlp_input):
        return self.outs.pop(0)


def _build(g, cls_name, accelerate, recorded=None):
    from fbbev_b200.view_transformation import bevdet_lineage as bl
    grid = dict(x=list(g["grid_x"]), y=list(g["grid_y"]), z=list(g["grid_z"]),
                depth=list(g["grid_depth"]))
    kw = dict(grid_config=grid,
              input_size=tuple(int(v) for v in g["input_size"]),
              downsample=int(g["downsample"]),
              in_channels=int(g["in_channels"]),
              out_channels=int(g["out_channels"]), accelerate=accelerate)
    if recorded is not None:
        m = getattr(bl, cls_name)(depthnet_cfg=dict(module=recorded), **kw)
    else:
        m = getattr(bl, cls_name)(**kw)
        m.load_state_dict({"depth_net.weight": torch.from_numpy(
            g["depth_net.weight"]), "depth_net.bias": torch.from_numpy(
                g["depth_net.bias"])}, strict=True)
    return m.to(DEV).eval()


def _z_split(bev_collapsed, z):
    """(B, Z*C, Y, X) (torch.cat(unbind(2), 1)) -> (B, C, Z, Y, X)."""
    B, zc, Y, X = bev_collapsed.shape
    return bev_collapsed.reshape(B, z, zc // z, Y, X).transpose(0, 2, 1, 3, 4)


@pytest.mark.parametrize("case,cls_name", [
    ("l_lss_v1", "LSSViewTransformer"), ("l_lss_v2", "LSSViewTransformer2"),
    ("l_lss_bevdepth", "LSSViewTransformerBEVDepth")])
@pytest.mark.parametrize("fused", [True, False])
def test_lss_lineage_vs_reference_golden(case, cls_name, fused):
    """forward() of the three BEVDet-lineage transformers == the outputs of
    the reference's own classes (necks/view_transformer.py), incl. the depth
    threshold.  accelerate=True: LSSViewTransformer matches the reference's
    cached path; for the two thresholded classes the reference's cached path
    applies its depth flags to the wrong list entries (see oracle/lift_ref.py
    ::prepare_sparse_cached, which reproduces it), so the product is held to
    the aligned filter = the reference's UNcached result in the squeezed
    layout."""
    g = load_golden(case)
    cam = [t(g[k]) for k in ("rots", "trans", "intrins", "post_rots",
                             "post_trans", "bda")]
    bevdepth = "net_out" in g

    def inputs(which):
        inp = [t(g["x" if which == 1 else "x2"])] + cam
        if bevdepth:
            inp.append(t(g["mlp_input"]))
        return inp

    def rec(*keys):
        return _Recorded([t(g[k]) for k in keys]) if bevdepth else None

    with torch.no_grad():
        m = _build(g, cls_name, False, rec("net_out"))
        m.fused_geometry = fused
        bev, depth, digit = m(inputs(1), return_depth_digit=True)
        np.testing.assert_allclose(depth.cpu().numpy(), g["depth_plain"],
                                   rtol=0, atol=2e-6)
        np.testing.assert_allclose(digit.cpu().numpy(), g["digit_plain"],
                                   rtol=0, atol=ATOL)
        assert tuple(bev.shape) == g["bev_plain"].shape
        np.testing.assert_allclose(bev.cpu().numpy(), g["bev_plain"], rtol=0,
                                   atol=ATOL)
        if bevdepth:
            mlp = m.get_mlp_input(*cam)
            np.testing.assert_array_equal(mlp.cpu().numpy(), g["mlp_input"])

        acc = _build(g, cls_name, True, rec("net_out", "net_out2"))
        acc.fused_geometry = fused
        bev_a, _ = acc(inputs(1))
        z = int(acc.grid_size[2])
        if cls_name == "LSSViewTransformer":
            want_a, want_b = g["bev_acc"], g["bev_acc_second"]
        else:
            want_a = _z_split(g["bev_plain"], z)
            want_b = None
            assert int(acc.kept.sum()) == int(g["n_kept_geom"])
            assert np.array_equal(acc.kept.cpu().numpy(), g["kept_mask"])
        assert tuple(bev_a.shape) == want_a.shape
        np.testing.assert_allclose(bev_a.cpu().numpy(), want_a, rtol=0,
                                   atol=ATOL)
        assert acc.initial_flag is False
        assert len(acc.ranks_bev) == int(g["n_kept_geom"])
        # second call: cached cameras (moving them must have no effect)
        moved = inputs(2)
        moved[2] = moved[2] + 5.0
        bev_b, _ = acc(moved)
        if want_b is None:
            plain2 = _build(g, cls_name, False, rec("net_out2"))
            plain2.fused_geometry = fused
            want_b = _z_split(plain2(inputs(2))[0].cpu().numpy(), z)
        np.testing.assert_allclose(bev_b.cpu().numpy(), want_b, rtol=0,
                                   atol=ATOL)
    assert sorted(k for k in _build(g, "LSSViewTransformer2", False)
                  .state_dict()) == ["depth_net.bias", "depth_net.weight"] \
        if not bevdepth else True


# ------------------------------------------------------ CM_DepthNet tail ----
@pytest.mark.parametrize("channels_last", [False, True])
def test_cm_depth_net_tail_vs_reference_golden(channels_last):
    """CM_DepthNetTail on the captured inputs of the reference's tail
    (depth_net.py:346-363): ``feat`` is the reference's context in the
    (B,N,H,W,C) layout the pooling op reads, ``depth`` its softmax.  With a
    channels-last input the 1x1 convolution runs as the tcgen05 row-wise
    Linear straight into that layout."""
    from fbbev_b200.view_transformation.depth_net_tail import CM_DepthNetTail
    g = load_golden("l_cm_tail")
    B, N = int(g["B"]), int(g["N"])
    mid, C = g["ctx_in"].shape[1], g["context"].shape[2]
    tail = CM_DepthNetTail(mid, C)
    tail.load_state_dict({
        "context_conv.weight": torch.from_numpy(g["context_conv_weight"]),
        "context_conv.bias": torch.from_numpy(g["context_conv_bias"])})
    tail = tail.to(DEV).eval()
    ctx = t(g["ctx_in"])
    if channels_last:
        ctx = ctx.contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        feat, depth = tail(ctx, t(g["logits"]), B, N)
    assert feat.is_contiguous() and depth.is_contiguous()
    np.testing.assert_allclose(depth.cpu().numpy(), g["depth"], rtol=0,
                               atol=2e-6)
    np.testing.assert_allclose(feat.cpu().numpy(),
                               g["context"].transpose(0, 1, 3, 4, 2), rtol=0,
                               atol=ATOL)


def test_nhwc_context_is_the_same_pooling():
    """LSSViewTransformerFunction3D.forward(context_layout='nhwc') on the
    producer's (B,N,H,W,C) tensor == the NCHW call, bit for bit: only the
    permute + copy in front of the kernel disappears."""
    from test_forward_gpu import make_case
    vt, cam, depth, feat = make_case("shipped", 2)
    want = vt(cam, feat, depth)
    nhwc = feat.permute(0, 1, 3, 4, 2).contiguous()
    got = vt(cam, nhwc, depth, context_layout='nhwc')
    assert torch.equal(got, want)
