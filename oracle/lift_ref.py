"""CPU restatement of the callers either side of the pooling op (SURVEY.md
section 8 f3 / f4).  TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

* :func:`lift_tail`        depth_net.py:359-363 / necks/view_transformer.py:313-321
                           (+ the NHWC copy of view_transformer.py:530, bev_pool.py:19)
* :func:`prepare_sparse`   necks/view_transformer.py:520-578 (depth-threshold
                           sparsification, :556-557) on top of ``oracle.cpu.voxel_prepare``
* :func:`lss_forward`      LSSViewTransformer(2).forward, necks/view_transformer.py
                           :296-324 / :693-721, accelerate off (:290-293) and on (:266-289,
                           :645-687)

Pinned by ``tests/golden/l_*.npz`` -- outputs of the reference's own classes
(``tests/test_oracle.py::test_lineage_oracle_vs_reference_golden``).
"""
import numpy as np

from . import cpu


def lift_tail(depth_logits, context):
    """depth_logits (BN, D, H, W), context (BN, C, H, W) ->
    softmax over D (float32, max / exp / sum / divide) and context as
    (BN, H, W, C)."""
    x = np.asarray(depth_logits, np.float32)
    m = x.max(1, keepdims=True)
    e = np.exp(x - m, dtype=np.float32)
    depth = e / e.sum(1, keepdims=True, dtype=np.float32)
    feat = np.ascontiguousarray(
        np.asarray(context, np.float32).transpose(0, 2, 3, 1))
    return depth.astype(np.float32), feat


def _intervals(ranks_bev):
    kept = np.ones(len(ranks_bev), bool)           # :488-491
    kept[1:] = ranks_bev[1:] != ranks_bev[:-1]
    st = np.nonzero(kept)[0].astype(np.int32)
    ln = np.zeros_like(st)
    ln[:-1] = st[1:] - st[:-1]
    ln[-1] = len(ranks_bev) - st[-1]
    return st, ln


def prepare_sparse(coor, lo, iv, gs, depth, thresh=0.01):
    """Index of the points that are inside the grid AND have depth probability
    > thresh, sorted by voxel rank (stable).  Filtering the sorted geometric
    list (what the reference's cached-index path does, :657-661) and filtering
    before the sort (:556-557) give the same list."""
    rb, rd, rf, _, _ = cpu.voxel_prepare(coor, lo, iv, gs)
    if rb is None:
        return None, None, None, None, None
    keep = np.asarray(depth, np.float32).reshape(-1)[rd] > np.float32(thresh)
    rb, rd, rf = rb[keep], rd[keep], rf[keep]
    if len(rb) == 0:
        return None, None, None, None, None
    st, ln = _intervals(rb)
    return rb, rd, rf, st, ln


def reference_sorted_list(coor, lo, iv, gs):
    """The in-grid point list exactly as ``voxel_pooling_prepare_v2_inf``
    (:580-637) builds it with THIS torch build on CPU: float32 rank arithmetic
    and ``argsort()`` without ``stable=True`` -- the order inside a voxel is
    whatever torch's sort leaves (it is not the stable order)."""
    import torch
    c = torch.from_numpy(np.asarray(coor, np.float32))
    B, N, D, H, W, _ = c.shape
    n = B * N * D * H * W
    rd = torch.arange(0, n, dtype=torch.int)
    rf = torch.arange(0, n // D, dtype=torch.int).reshape(B, N, 1, H, W)
    rf = rf.expand(B, N, D, H, W).flatten()
    lo_t, iv_t, gs_t = (torch.from_numpy(np.asarray(a, np.float32))
                        for a in (lo, iv, gs))
    cc = ((c - lo_t) / iv_t).long().view(n, 3)
    bidx = torch.arange(0, B).reshape(B, 1).expand(B, n // B).reshape(n, 1)
    cc = torch.cat((cc, bidx.to(cc)), 1)
    kept = (cc[:, 0] >= 0) & (cc[:, 0] < gs_t[0]) & (cc[:, 1] >= 0) & \
        (cc[:, 1] < gs_t[1]) & (cc[:, 2] >= 0) & (cc[:, 2] < gs_t[2])
    cc, rd, rf = cc[kept], rd[kept], rf[kept]
    rb = cc[:, 3] * (gs_t[2] * gs_t[1] * gs_t[0])
    rb += cc[:, 2] * (gs_t[1] * gs_t[0])
    rb += cc[:, 1] * gs_t[0] + cc[:, 0]
    order = rb.argsort()
    return (kept.numpy(), rb[order].int().numpy(), rd[order].numpy(),
            rf[order].numpy())


def prepare_sparse_cached(coor, lo, iv, gs, depth, thresh=0.01):
    """The cached-index form, AS THE REFERENCE COMPUTES IT (:657-661):

        depth_kept = (depth.view(-1) > 0.01)[self.kept]
        new_ranks_* = self.ranks_*[depth_kept]

    ``self.kept`` is the in-grid mask over the points in their ORIGINAL order,
    while ``self.ranks_*`` are sorted by voxel: entry j of the sorted list is
    kept when the j-th in-grid point in original order passes the threshold.
    Flags and list are misaligned, and since the order inside a voxel comes
    from an unstable argsort the result is not even a function of the inputs
    alone (it follows the sort implementation).  This restatement reproduces
    the reference's CPU result (same torch build) to show that reading is
    right; the product implements the aligned filter -- what the uncached
    path computes (:556-557) -- and DESIGN.md lists the difference."""
    kept, rb, rd, rf = reference_sorted_list(coor, lo, iv, gs)
    keep = (np.asarray(depth, np.float32).reshape(-1) >
            np.float32(thresh))[kept]
    rb, rd, rf = rb[keep], rd[keep], rf[keep]
    if len(rb) == 0:
        return None, None, None, None, None
    st, ln = _intervals(rb)
    return rb, rd, rf, st, ln


def lss_forward(net_out, coor, lo, iv, gs, D, C, thresh=None, accelerate=False):
    """net_out (B, N, D + C, H, W): output of ``depth_net``; coor
    (B, N, D, H, W, 3) from get_lidar_coor.  Returns (bev, depth): bev with Z
    collapsed into channels (:191), or squeezed when ``accelerate`` (:283)."""
    B, N, _, H, W = net_out.shape
    x = net_out.reshape(B * N, -1, H, W)
    depth, feat = lift_tail(x[:, :D], x[:, D:D + C])
    depth5 = depth.reshape(B, N, D, H, W)
    if thresh is None:
        idx = cpu.voxel_prepare(coor, lo, iv, gs)
    elif accelerate:
        idx = prepare_sparse_cached(coor, lo, iv, gs, depth5, thresh)
    else:
        idx = prepare_sparse(coor, lo, iv, gs, depth5, thresh)
    gz, gy, gx = int(gs[2]), int(gs[1]), int(gs[0])
    if idx[0] is None:
        vol = np.zeros((B, C, gz, gy, gx), np.float32)
    else:
        rb, rd, rf, st, ln = idx
        vol = cpu.bev_pool_v2(depth5, feat.reshape(B, N, H, W, C), rd, rf, rb,
                              (B, gz, gy, gx, C), st, ln)
    if accelerate:
        bev = vol.squeeze(2) if vol.shape[2] == 1 else vol   # torch squeeze
    else:
        bev = np.concatenate([vol[:, :, z] for z in range(gz)], 1)
    return bev, depth
