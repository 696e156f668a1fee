"""CPU tests: pin the oracle against the reference's own known-answer test and
the golden vectors produced by the reference's Python (tests/golden/)."""
import numpy as np
import pytest
import torch

from conftest import canon_index, load_golden

F_CASES = ["f_small_6cam", "f_unit_1cam", "f_negative_trunc"]


def kat_inputs():
    """The reference's only golden vector for the hot path:
    mmdet3d/ops/bev_pool_v2/bev_pool.py:145-176 (test_bev_pool_v2)."""
    depth = np.array([0.3, 0.4, 0.2, 0.1, 0.7, 0.6, 0.8, 0.9],
                     np.float32).reshape(1, 1, 2, 2, 2)
    feat = np.ones((1, 1, 2, 2, 2), np.float32)
    ranks_depth = np.array([0, 4, 1, 6], np.int32)
    ranks_feat = np.array([0, 0, 1, 2], np.int32)
    ranks_bev = np.array([0, 0, 1, 1], np.int32)
    interval_starts = np.array([0, 2], np.int32)
    interval_lengths = np.array([2, 2], np.int32)
    return (depth, feat, ranks_depth, ranks_feat, ranks_bev, (1, 1, 2, 2, 2),
            interval_starts, interval_lengths)


KAT_GRAD_DEPTH = np.array([2., 2., 0., 0., 2., 0., 2., 0.], np.float32)
KAT_GRAD_FEAT = np.array([1.0, 1.0, 0.4, 0.4, 0.8, 0.8, 0., 0.], np.float32)


def test_kat_forward(oracle_cpu):
    args = kat_inputs()
    out = oracle_cpu.bev_pool_v2(*args)
    assert out.shape == (1, 2, 1, 2, 2)  # (B, C, Z, Y, X)
    assert np.isclose(out.sum(), 4.4, rtol=0, atol=1e-6)  # bev_pool.py:169
    # hand-derivable: voxel0 = .3+.7, voxel1 = .4+.8 for both channels
    zyxc = oracle_cpu.bev_pool_v2_fwd(*args)
    np.testing.assert_allclose(zyxc.reshape(4, 2)[0], [1.0, 1.0], atol=1e-7)
    np.testing.assert_allclose(zyxc.reshape(4, 2)[1], [1.2, 1.2], atol=1e-6)
    assert np.all(zyxc.reshape(4, 2)[2:] == 0)


def test_kat_backward(oracle_cpu):
    depth, feat, rd, rf, rb, shape, st, ln = kat_inputs()
    og = np.ones(shape, np.float32)  # d(sum)/d(out)
    dg, fg = oracle_cpu.bev_pool_v2_bwd(og, depth, feat, rd, rf, rb)
    np.testing.assert_allclose(dg.ravel(), KAT_GRAD_DEPTH)  # :170-173
    np.testing.assert_allclose(fg.ravel(), KAT_GRAD_FEAT, atol=1e-6)  # :174-176


@pytest.mark.parametrize("case", F_CASES)
def test_prepare_matches_reference_python(oracle_cpu, case):
    g = load_golden(case)
    rb, rd, rf, st, ln = oracle_cpu.voxel_prepare(
        g["coor"], g["grid_lower_bound"], g["grid_interval"], g["grid_size"])
    # bit-exact integer path
    np.testing.assert_array_equal(rb, g["ranks_bev"])
    np.testing.assert_array_equal(st, g["interval_starts"])
    np.testing.assert_array_equal(ln, g["interval_lengths"])
    a = canon_index(rb, rd, rf)
    b = canon_index(g["ranks_bev"], g["ranks_depth"], g["ranks_feat"])
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)
    # the oracle's own order is the canonical (stable) one
    for x, y in zip(a, (rb, rd, rf)):
        np.testing.assert_array_equal(x, y)
    # float32 rank arithmetic of the reference == exact integers below 2^24
    rb1 = oracle_cpu.voxel_prepare(g["coor"], g["grid_lower_bound"],
                                   g["grid_interval"], g["grid_size"],
                                   rank_mode=1)[0]
    np.testing.assert_array_equal(rb1, rb)


def test_prepare_negative_truncation_keeps_cell0(oracle_cpu):
    """`.long()` truncates toward zero: coordinates in (-1, 0) land in cell 0
    and are KEPT (view_transformer.py:570-580)."""
    g = load_golden("f_negative_trunc")
    coor = g["coor"].reshape(-1, 3)
    rel = (coor - g["grid_lower_bound"]) / g["grid_interval"]
    inside_by_trunc = np.all((np.trunc(rel) >= 0) &
                             (np.trunc(rel) < g["grid_size"]), axis=1)
    inside_by_floor = np.all((np.floor(rel) >= 0) &
                             (np.floor(rel) < g["grid_size"]), axis=1)
    assert inside_by_trunc.sum() > inside_by_floor.sum()
    assert len(g["ranks_bev"]) == inside_by_trunc.sum()


def test_prepare_empty(oracle_cpu):
    g = load_golden("f_empty")
    out = oracle_cpu.voxel_prepare(g["coor"], g["grid_lower_bound"],
                                   g["grid_interval"], g["grid_size"])
    assert all(o is None for o in out)
    assert "ranks_bev" not in g  # the reference returned None as well
    assert np.all(g["bev_feat"] == 0)


@pytest.mark.parametrize("case", F_CASES)
def test_pool_matches_reference_glue(oracle_cpu, case):
    """voxel_pooling_v2 (view_transformer.py:521-545) end to end."""
    g = load_golden(case)
    B, N, C, H, W = g["feat"].shape
    gs = g["grid_size"].astype(int)
    feat_nhwc = np.ascontiguousarray(g["feat"].transpose(0, 1, 3, 4, 2))
    out = oracle_cpu.bev_pool_v2(
        g["depth"], feat_nhwc, g["ranks_depth"], g["ranks_feat"],
        g["ranks_bev"], (B, gs[2], gs[1], gs[0], C), g["interval_starts"],
        g["interval_lengths"])
    bev = out.transpose(0, 1, 3, 4, 2)  # (B,C,Z,Y,X) -> (B,C,Y,X,Z)
    assert tuple(g["bev_feat_shape"]) == bev.shape
    np.testing.assert_array_equal(bev, g["bev_feat"])


def test_float32_rank_overflow_documented(oracle_cpu):
    """Above 2^24 voxels the reference's float32 rank arithmetic
    (view_transformer.py:586-589) collides; the exact mode does not."""
    lo = np.array([0, 0, 0], np.float32)
    iv = np.array([1, 1, 1], np.float32)
    gs = np.array([4096, 4096, 2], np.float32)  # 2^25 voxels
    coor = np.array([[[[[[4092.5, 4095.5, 1.5], [4093.5, 4095.5, 1.5]]]]]],
                    np.float32)  # (B=1,N=1,D=1,H=1,W=2,3)
    exact = oracle_cpu.voxel_prepare(coor, lo, iv, gs, rank_mode=0)
    f32 = oracle_cpu.voxel_prepare(coor, lo, iv, gs, rank_mode=1)
    assert len(exact[3]) == 2          # two distinct voxels
    assert len(f32[3]) == 1            # merged by float32 rounding


def _rand_msda(seed, bs=2, nq=13, heads=4, ch=10, shapes=((5, 7), (3, 4)),
               points=6):
    g = torch.Generator().manual_seed(seed)
    shapes_t = torch.tensor(shapes, dtype=torch.long)
    lsi = torch.cat((shapes_t.new_zeros(1), shapes_t.prod(1).cumsum(0)[:-1]))
    n_value = int(shapes_t.prod(1).sum())
    value = torch.randn(bs, n_value, heads, ch, generator=g)
    # locations deliberately spill outside [0,1] to exercise zero padding
    loc = torch.rand(bs, nq, heads, len(shapes), points, 2, generator=g) * 1.4 - 0.2
    attw = torch.rand(bs, nq, heads, len(shapes), points, generator=g)
    return value, shapes_t, lsi, loc, attw


@pytest.mark.parametrize("ch", [10, 32])
def test_msda_oracle_vs_grid_sample_and_hf(oracle_cpu, ch):
    """The C im2col restatement == mmcv's documented PyTorch formulation
    (grid_sample) == HF transformers' independent implementation."""
    from oracle import torch_ref
    value, shapes, lsi, loc, attw = _rand_msda(0, ch=ch)
    c = oracle_cpu.msda_fwd(value.numpy(), shapes.numpy(), lsi.numpy(),
                            loc.numpy(), attw.numpy())
    t = torch_ref.multi_scale_deformable_attn_pytorch(value, shapes, loc, attw)
    np.testing.assert_allclose(c, t.numpy(), rtol=0, atol=2e-5)
    try:
        from transformers.models.deformable_detr.modeling_deformable_detr \
            import MultiScaleDeformableAttention as HF
    except Exception:  # pragma: no cover
        pytest.skip("transformers implementation not importable")
    hf = HF()
    shapes_list = [(int(h), int(w)) for h, w in shapes]
    h = hf(value, shapes, shapes_list, lsi, loc, attw, 64)
    np.testing.assert_allclose(c, h.detach().numpy(), rtol=0, atol=2e-5)


def test_msda_backward_oracle_vs_autograd(oracle_cpu):
    from oracle import torch_ref
    value, shapes, lsi, loc, attw = _rand_msda(1)
    value.requires_grad_(), loc.requires_grad_(), attw.requires_grad_()
    out = torch_ref.multi_scale_deformable_attn_pytorch(value, shapes, loc,
                                                        attw)
    go = torch.randn(out.shape, generator=torch.Generator().manual_seed(2))
    out.backward(go)
    gv, gl, ga = oracle_cpu.msda_bwd(
        value.detach().numpy(), shapes.numpy(), lsi.numpy(),
        loc.detach().numpy(), attw.detach().numpy(), go.numpy())
    np.testing.assert_allclose(gv, value.grad.numpy(), atol=5e-5)
    np.testing.assert_allclose(ga, attw.grad.numpy(), atol=5e-5)
    np.testing.assert_allclose(gl, loc.grad.numpy(), atol=5e-4)


@pytest.mark.parametrize("case", ["b_bp_e80_1lvl", "b_bp_e64_3lvl"])
def test_backward_projection_cpu_oracle_vs_reference_golden(case):
    """oracle/backward_ref.py (the CPU restatement used as the checker of the
    fused CUDA kernel and as bench.py's CPU baseline) reproduces the output of
    the reference's own BackwardProjection / DA_SpatialCrossAttention."""
    from bp_common import build_bp, cam_params
    from oracle import backward_ref
    g, bp = build_bp(case)
    n_lvl = len(g["level_shapes"])
    mlvl = [torch.from_numpy(g[f"feat{i}"]) for i in range(n_lvl)]
    out = backward_ref.backward_projection_cpu(
        bp, mlvl, torch.from_numpy(g["lss_bev"]), cam_params(g),
        torch.from_numpy(g["depth"]))
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=0, atol=2e-5)
    # the cross-attention alone, on the reference's recorded inputs
    sca = bp.transformer.encoder.layers[0].attentions[1]
    with backward_ref._cpu_kernels(), torch.no_grad():
        o = sca(torch.from_numpy(g["sca_query"]),
                torch.from_numpy(g["sca_key"]), torch.from_numpy(g["sca_key"]),
                None, query_pos=torch.from_numpy(g["sca_query_pos"]),
                reference_points_cam=torch.from_numpy(
                    g["reference_points_cam"]),
                spatial_shapes=torch.from_numpy(g["spatial_shapes"]),
                level_start_index=torch.from_numpy(g["level_start_index"]),
                bev_query_depth=torch.from_numpy(g["bev_query_depth"]),
                pred_img_depth=torch.from_numpy(g["depth"]),
                per_cam_mask_list=torch.from_numpy(g["per_cam_mask"]))
    np.testing.assert_allclose(o.numpy(), g["sca_out"], rtol=0, atol=2e-5)


def test_inference_caches_follow_parameter_updates():
    """With gradients off the plugin caches constants of the weights (the
    positional encoding and its token-major copy).  An in-place parameter update
    -- an optimizer step, load_state_dict -- must invalidate them."""
    import copy
    from bp_common import build_bp, cam_params
    from oracle import backward_ref
    g, bp = build_bp("b_bp_e80_1lvl")
    feats = [torch.from_numpy(g[f"feat{i}"])
             for i in range(len(g["level_shapes"]))]
    args = (feats, torch.from_numpy(g["lss_bev"]), cam_params(g),
            torch.from_numpy(g["depth"]))
    first = backward_ref.backward_projection_cpu(bp, *args)
    again = backward_ref.backward_projection_cpu(bp, *args)   # served by caches
    assert torch.equal(first, again)
    with torch.no_grad():
        bp.positional_encoding.row_embed.weight.add_(0.25)
        bp.positional_encoding.col_embed.weight.mul_(0.5)
    fresh = copy.deepcopy(bp)
    for m in fresh.modules():                                  # drop every cache
        for k in [k for k in m.__dict__ if k.startswith('_pos')]:
            del m.__dict__[k]
    want = backward_ref.backward_projection_cpu(fresh, *args)
    got = backward_ref.backward_projection_cpu(bp, *args)
    assert not torch.equal(first, got)
    assert torch.equal(got, want)


@pytest.mark.parametrize("seed", range(6))
def test_prepare_oracle_vs_numpy_restatement_ragged(oracle_cpu, seed):
    """The C restatement of voxel_pooling_prepare_v2 against an independent
    numpy restatement of the same reference lines (view_transformer.py:563-605)
    on ragged random frustums: whole cameras outside the grid, points on cell
    faces, coordinates in (-1, 0), a batch element with nothing kept."""
    rng = np.random.default_rng(seed)
    B, N, D, H, W = 2 + seed % 2, 1 + seed % 3, 3 + seed, 2 + seed % 4, 5
    lo = np.array([-4.0, -3.0, -1.0], np.float32)
    iv = np.array([0.5, 0.75, 1.0], np.float32)
    gs = np.array([16, 8, 3], np.float32)
    coor = (rng.random((B, N, D, H, W, 3), dtype=np.float32) * 14 - 6).astype(
        np.float32)
    coor[..., ::2, :, 0] = np.round(coor[..., ::2, :, 0] * 2) / 2   # on faces
    coor[0, 0] += 100.0                                             # camera out
    if B > 2:
        coor[2] -= 100.0                                            # empty frame
    got = oracle_cpu.voxel_prepare(coor, lo, iv, gs)
    # numpy restatement: fp32 subtract, fp32 divide, truncate toward zero
    n = B * N * D * H * W
    rel = ((coor.reshape(-1, 3) - lo) / iv).astype(np.float32)
    cell = np.trunc(rel).astype(np.int64)
    batch = np.repeat(np.arange(B), n // B)
    kept = np.all((cell >= 0) & (cell < gs.astype(np.int64)), axis=1)
    ranks_depth = np.arange(n, dtype=np.int64)
    ranks_feat = (np.arange(n // D).reshape(B, N, 1, H, W) +
                  np.zeros((1, 1, D, 1, 1), np.int64)).reshape(-1)
    X, Y, Z = (int(v) for v in gs)
    rank = (batch * (Z * Y * X) + cell[:, 2] * (Y * X) + cell[:, 1] * X +
            cell[:, 0])
    rank, rd, rf = rank[kept], ranks_depth[kept], ranks_feat[kept]
    order = np.argsort(rank, kind="stable")
    rank, rd, rf = rank[order], rd[order], rf[order]
    if len(rank) == 0:
        assert all(o is None for o in got)
        return
    starts = np.flatnonzero(np.r_[True, rank[1:] != rank[:-1]])
    lengths = np.diff(np.r_[starts, len(rank)])
    rb, grd, grf, st, ln = got
    np.testing.assert_array_equal(rb, rank.astype(np.int32))
    np.testing.assert_array_equal(st, starts.astype(np.int32))
    np.testing.assert_array_equal(ln, lengths.astype(np.int32))
    np.testing.assert_array_equal(grd, rd.astype(np.int32))   # stable order
    np.testing.assert_array_equal(grf, rf.astype(np.int32))


# ------------------------------------------- eager "reference on GPU" ------
@pytest.mark.parametrize("case", ["b_bp_e80_1lvl", "b_bp_e64_3lvl"])
def test_eager_reference_restatement_vs_golden(case):
    """oracle/gpu_ref.py (the comparator bench.py times as "the reference on
    the same B200") reproduces the reference's own recorded outputs: the
    backward projection with the re-batching cross-attention and the
    grid_sample MSDA, and voxel_pooling_prepare_v2 op for op.  Device-agnostic
    torch ops, so it is pinned here on the CPU."""
    import torch
    from bp_common import build_bp, cam_params
    from oracle import gpu_ref
    g, bp = build_bp(case, "cpu")
    n_lvl = len(g["level_shapes"])
    tt = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    mlvl = [tt(g[f"feat{i}"]) for i in range(n_lvl)]
    with torch.no_grad(), gpu_ref.eager_reference_mode(bp.transformer.encoder):
        out = bp(mlvl, None, lss_bev=tt(g["lss_bev"]),
                 cam_params=cam_params(g, "cpu"), pred_img_depth=tt(g["depth"]))
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=0, atol=1e-5)


@pytest.mark.parametrize("case", ["f_small_6cam", "f_unit_1cam",
                                  "f_negative_trunc"])
def test_eager_prepare_restatement_vs_golden(case):
    import torch
    from oracle import gpu_ref
    g = load_golden(case)
    rb, rd, rf, st, ln = gpu_ref.voxel_pooling_prepare_v2(
        torch.from_numpy(g["coor"]), torch.from_numpy(g["grid_lower_bound"]),
        torch.from_numpy(g["grid_interval"]), torch.from_numpy(g["grid_size"]))
    assert np.array_equal(rb.numpy(), g["ranks_bev"])
    assert np.array_equal(st.numpy(), g["interval_starts"])
    assert np.array_equal(ln.numpy(), g["interval_lengths"])
    a = canon_index(rb.numpy(), rd.numpy(), rf.numpy())
    b = canon_index(g["ranks_bev"], g["ranks_depth"], g["ranks_feat"])
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    feat = torch.from_numpy(g["feat"]).permute(0, 1, 3, 4, 2).contiguous()
    B, C, Y, X, Z = g["bev_feat"].shape
    out = gpu_ref.bev_pool_v2_index_add(torch.from_numpy(g["depth"]), feat, rd,
                                        rf, rb, (B, Z, Y, X, C))
    np.testing.assert_allclose(out.permute(0, 1, 3, 4, 2).numpy(),
                               g["bev_feat"], rtol=0, atol=1e-5)


# ---------------------------------------------------------------------------
# BEVDet-lineage callers and the depth-net tail
# ---------------------------------------------------------------------------
def _lineage_inputs(g):
    """depth_net output and frustum coordinates of a l_lss_* fixture, formed
    on CPU with the golden's weights (the 1x1 conv is the step before the
    path) and the reference's geometry chain (pinned by the f_* goldens)."""
    import torch
    from fbbev_b200.view_transformation.bevdet_lineage import \
        LSSViewTransformer
    grid = dict(x=list(g["grid_x"]), y=list(g["grid_y"]), z=list(g["grid_z"]),
                depth=list(g["grid_depth"]))
    vt = LSSViewTransformer(grid, tuple(int(v) for v in g["input_size"]),
                            int(g["downsample"]),
                            in_channels=int(g["in_channels"]),
                            out_channels=int(g["out_channels"]))
    cam = [torch.from_numpy(g[k]) for k in
           ("rots", "trans", "intrins", "post_rots", "post_trans", "bda")]
    coor = vt.get_lidar_coor(*cam).numpy()
    outs = []
    for key, rec in (("x", "net_out"), ("x2", "net_out2")):
        if rec in g:
            outs.append(g[rec])
            continue
        x = torch.from_numpy(g[key])
        B, N, Cin, H, W = x.shape
        w = torch.from_numpy(g["depth_net.weight"])
        b = torch.from_numpy(g["depth_net.bias"])
        outs.append(torch.nn.functional.conv2d(
            x.view(B * N, Cin, H, W), w, b).numpy())
    B, N = g["x"].shape[:2]
    H, W = g["x"].shape[-2:]
    outs = [o.reshape(B, N, -1, H, W) for o in outs]
    return vt, coor, outs


@pytest.mark.parametrize("case,thresh", [("l_lss_v1", None),
                                         ("l_lss_v2", 0.01),
                                         ("l_lss_bevdepth", 0.01)])
def test_lineage_oracle_vs_reference_golden(oracle_cpu, case, thresh):
    """oracle.lift_ref (softmax + NHWC + thresholded index + pooling) against
    the outputs of the reference's own LSSViewTransformer / LSSViewTransformer2
    / LSSViewTransformerBEVDepth (necks/view_transformer.py), accelerate off
    and on, first and second call."""
    from oracle import lift_ref
    g = load_golden(case)
    vt, coor, (o1, o2) = _lineage_inputs(g)
    lo, iv, gs = (vt.grid_lower_bound.numpy(), vt.grid_interval.numpy(),
                  vt.grid_size.numpy())
    D, C = vt.D, int(g["out_channels"])
    bev, depth = lift_ref.lss_forward(o1, coor, lo, iv, gs, D, C, thresh)
    np.testing.assert_allclose(depth, g["depth_plain"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(bev, g["bev_plain"], rtol=0, atol=1e-5)
    bev_a, _ = lift_ref.lss_forward(o1, coor, lo, iv, gs, D, C, thresh, True)
    np.testing.assert_allclose(bev_a, g["bev_acc"], rtol=0, atol=1e-5)
    bev_b, _ = lift_ref.lss_forward(o2, coor, lo, iv, gs, D, C, thresh, True)
    np.testing.assert_allclose(bev_b, g["bev_acc_second"], rtol=0, atol=1e-5)
    if thresh is not None:     # the threshold must actually bite in the fixture
        assert float(g["frac_below_thresh"]) > 0.02
        full, _ = lift_ref.lss_forward(o1, coor, lo, iv, gs, D, C, None)
        assert np.abs(full - bev).max() > 1e-4


def test_cm_depth_net_tail_oracle_vs_reference_golden():
    """oracle.lift_ref.lift_tail on the captured inputs of CM_DepthNet's tail
    (depth_net.py:346-363) == the reference's returned (context, depth)."""
    import torch
    from oracle import lift_ref
    g = load_golden("l_cm_tail")
    ctx = torch.nn.functional.conv2d(
        torch.from_numpy(g["ctx_in"]), torch.from_numpy(g["context_conv_weight"]),
        torch.from_numpy(g["context_conv_bias"])).numpy()
    depth, feat = lift_ref.lift_tail(g["logits"], ctx)
    B, N = int(g["B"]), int(g["N"])
    np.testing.assert_allclose(depth.reshape(g["depth"].shape), g["depth"],
                               rtol=0, atol=1e-6)
    np.testing.assert_allclose(
        feat.reshape(B, N, *feat.shape[1:]),
        g["context"].transpose(0, 1, 3, 4, 2), rtol=0, atol=1e-6)
