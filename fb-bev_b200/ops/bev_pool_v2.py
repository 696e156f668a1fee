"""``bev_pool_v2`` -- host-side mirror of the reference operator.

Mirrors ``mmdet3d/ops/bev_pool_v2/bev_pool.py`` (names, argument order and
meaning, output layout):

* ``QuickCumsumCuda``   -- bev_pool.py:12-81: autograd function over the
  reference-layout kernels; ``forward`` returns ``(B, Z, Y, X, C)``.
* ``bev_pool_v2()``     -- bev_pool.py:84-90: returns the contiguous
  ``(B, C, Z, Y, X)`` tensor.  Here it is ONE fused kernel
  (``fbbev_bev_pool_v2_fwd_dense``) instead of memset + kernel + transpose copy.
* ``voxel_pooling_prepare_v2`` -- device implementation of
  view_transformer.py:547-605 (``fbbev_voxel_prepare``).

All compute is in ``libfbbev_b200.so``; tensors must be CUDA tensors.
"""
import torch

from .. import _lib

__all__ = ['DeferredVolume', 'bev_pool_v2_deferred',
           'bev_pool_v2', 'bev_pool_v2_dense', 'QuickCumsumCuda',
           'voxel_pooling_prepare_v2', 'voxel_pooling_prepare_from_cams',
           'VoxelIndex']


def _feat_intervals(ranks_feat, ranks_depth, ranks_bev):
    """Re-sort the point list by feature pixel and rebuild intervals --
    QuickCumsumCuda.backward, bev_pool.py:45-55 (torch ops; plumbing)."""
    order = torch.sort(ranks_feat, stable=True)[1]
    ranks_feat, ranks_depth, ranks_bev = \
        ranks_feat[order], ranks_depth[order], ranks_bev[order]
    kept = torch.ones(ranks_bev.shape[0], device=ranks_bev.device,
                      dtype=torch.bool)
    kept[1:] = ranks_feat[1:] != ranks_feat[:-1]
    interval_starts_bp = torch.where(kept)[0].int()
    interval_lengths_bp = torch.zeros_like(interval_starts_bp)
    interval_lengths_bp[:-1] = interval_starts_bp[1:] - interval_starts_bp[:-1]
    interval_lengths_bp[-1] = ranks_bev.shape[0] - interval_starts_bp[-1]
    return (ranks_feat.contiguous(), ranks_depth.contiguous(),
            ranks_bev.contiguous(), interval_starts_bp.contiguous(),
            interval_lengths_bp.contiguous())


class QuickCumsumCuda(torch.autograd.Function):
    """Same contract as the reference class (bev_pool.py:12-81)."""

    @staticmethod
    def forward(ctx, depth, feat, ranks_depth, ranks_feat, ranks_bev,
                bev_feat_shape, interval_starts, interval_lengths):
        dev = _lib.require_cuda(depth, feat, ranks_depth, ranks_feat,
                                ranks_bev, interval_starts, interval_lengths)
        ranks_bev = ranks_bev.int()
        depth = depth.contiguous().float()
        feat = feat.contiguous().float()
        ranks_depth = ranks_depth.contiguous().int()
        ranks_feat = ranks_feat.contiguous().int()
        interval_lengths = interval_lengths.contiguous().int()
        interval_starts = interval_starts.contiguous().int()

        out = feat.new_zeros(bev_feat_shape)
        with torch.cuda.device(dev):
            rc = _lib.lib().fbbev_bev_pool_v2_fwd(
                _lib.ptr(depth), _lib.ptr(feat), _lib.ptr(ranks_depth),
                _lib.ptr(ranks_feat), _lib.ptr(ranks_bev),
                _lib.ptr(interval_starts), _lib.ptr(interval_lengths),
                interval_lengths.shape[0], feat.shape[-1], _lib.ptr(out),
                _lib.stream_ptr(dev))
        _lib.check(rc, 'fbbev_bev_pool_v2_fwd')
        ctx.save_for_backward(ranks_bev, depth, feat, ranks_feat, ranks_depth)
        return out

    @staticmethod
    def backward(ctx, out_grad):
        ranks_bev, depth, feat, ranks_feat, ranks_depth = ctx.saved_tensors
        dev = depth.device
        ranks_feat, ranks_depth, ranks_bev, starts_bp, lengths_bp = \
            _feat_intervals(ranks_feat, ranks_depth, ranks_bev)
        depth_grad = depth.new_zeros(depth.shape)
        feat_grad = feat.new_zeros(feat.shape)
        out_grad = out_grad.contiguous().float()
        with torch.cuda.device(dev):
            rc = _lib.lib().fbbev_bev_pool_v2_bwd(
                _lib.ptr(out_grad), _lib.ptr(depth), _lib.ptr(feat),
                _lib.ptr(ranks_depth), _lib.ptr(ranks_feat),
                _lib.ptr(ranks_bev), _lib.ptr(starts_bp), _lib.ptr(lengths_bp),
                lengths_bp.shape[0], feat.shape[-1], _lib.ptr(depth_grad),
                _lib.ptr(feat_grad), _lib.stream_ptr(dev))
        _lib.check(rc, 'fbbev_bev_pool_v2_bwd')
        return depth_grad, feat_grad, None, None, None, None, None, None


# Optional instrumentation: when set to an object with ``before(stream)`` /
# ``after(stream)`` methods, they are called around the launch of the dense
# pooling kernel (bench.py uses it to bracket that one kernel with CUDA events).
KERNEL_HOOK = None


def _usable_plan(plan, shape, n_points):
    """A plan filled by the index builder fits this call when it was laid out
    for the same volume / channel count (its workspace layout depends on them)."""
    if plan is None:
        return None
    ws, nbytes, pshape, cap = plan
    if tuple(pshape) != tuple(int(v) for v in shape):
        return None
    return ws, nbytes, cap


def _dense_forward(depth, feat, ranks_depth, ranks_feat, ranks_bev,
                   bev_feat_shape, interval_starts, interval_lengths,
                   n_intervals_dev, plan=None):
    dev = _lib.require_cuda(depth, feat, ranks_depth, ranks_feat, ranks_bev,
                            interval_starts, interval_lengths, n_intervals_dev)
    B, Z, Y, X, C = (int(s) for s in bev_feat_shape)
    assert feat.shape[-1] == C, (feat.shape, bev_feat_shape)
    L = _lib.lib()
    n_points = ranks_bev.shape[0]
    if C > 512 or B * Z * Y * X > 2 ** 31 - 1:
        # shapes the dense kernels do not cover (FBBEV_ERR_UNSUPPORTED): the
        # reference's own op sequence on the drop-in interval kernel --
        # zero-filled (B,Z,Y,X,C) + kernel + permute (bev_pool.py:25-36, 89)
        n_int = interval_lengths.shape[0] if n_intervals_dev is None \
            else int(n_intervals_dev.item())
        out = feat.new_zeros((B, Z, Y, X, C))
        with torch.cuda.device(dev):
            rc = L.fbbev_bev_pool_v2_fwd(
                _lib.ptr(depth), _lib.ptr(feat), _lib.ptr(ranks_depth),
                _lib.ptr(ranks_feat), _lib.ptr(ranks_bev),
                _lib.ptr(interval_starts), _lib.ptr(interval_lengths), n_int,
                C, _lib.ptr(out), _lib.stream_ptr(dev))
        _lib.check(rc, 'fbbev_bev_pool_v2_fwd')
        return out.permute(0, 4, 1, 2, 3).contiguous()
    out = torch.empty((B, C, Z, Y, X), dtype=torch.float32, device=dev)
    # an index can never hold more intervals than there are voxels: bounds the
    # V[n_intervals][C] workspace of the sync-free path (whose index buffers
    # are n_points long) -- 1.7 GB -> 0.8 GB for 16 frames of 200x200x16
    n_int_cap = min(interval_lengths.shape[0], B * Z * Y * X)
    ready = _usable_plan(plan, (B, Z, Y, X, C), n_points)
    if ready is not None and ready[2] == n_int_cap:
        ws, ws_bytes, _ = ready            # planned by the index builder
    else:
        ready = None
        ws_bytes = L.fbbev_bev_pool_v2_dense_workspace_bytes(
            B, Z * Y * X, n_int_cap, n_points, C)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    hook = KERNEL_HOOK
    with torch.cuda.device(dev):
        sp = _lib.stream_ptr(dev)
        if ready is None:
            rc = L.fbbev_bev_pool_v2_plan(
                _lib.ptr(ranks_bev), _lib.ptr(interval_starts),
                _lib.ptr(interval_lengths), n_int_cap,
                _lib.ptr(n_intervals_dev), n_points, C, B, Z * Y * X,
                _lib.ptr(ws), ws_bytes, sp)
            _lib.check(rc, 'fbbev_bev_pool_v2_plan')
        if hook is not None:
            hook.before()
        rc = L.fbbev_bev_pool_v2_fwd_dense_planned(
            _lib.ptr(depth), _lib.ptr(feat), _lib.ptr(ranks_depth),
            _lib.ptr(ranks_feat), _lib.ptr(ranks_bev),
            _lib.ptr(interval_starts), _lib.ptr(interval_lengths),
            n_int_cap, n_points, C, B, Z * Y * X,
            _lib.ptr(out), _lib.ptr(ws), ws_bytes, sp)
        if hook is not None:
            hook.after()
    _lib.check(rc, 'fbbev_bev_pool_v2_fwd_dense_planned')
    return out


class _BevPoolV2Dense(torch.autograd.Function):
    """Fused op: (B,C,Z,Y,X) out, gradient consumed in that layout."""

    @staticmethod
    def forward(ctx, depth, feat, ranks_depth, ranks_feat, ranks_bev,
                bev_feat_shape, interval_starts, interval_lengths,
                n_intervals_dev, n_kept_dev, plan=None):
        ranks_bev = ranks_bev.contiguous().int()
        depth = depth.contiguous().float()
        feat = feat.contiguous().float()
        ranks_depth = ranks_depth.contiguous().int()
        ranks_feat = ranks_feat.contiguous().int()
        interval_lengths = interval_lengths.contiguous().int()
        interval_starts = interval_starts.contiguous().int()
        out = _dense_forward(depth, feat, ranks_depth, ranks_feat, ranks_bev,
                             bev_feat_shape, interval_starts, interval_lengths,
                             n_intervals_dev, plan)
        ctx.save_for_backward(ranks_bev, depth, feat, ranks_feat, ranks_depth,
                              n_kept_dev)
        ctx.has_count = n_kept_dev is not None
        return out

    @staticmethod
    def backward(ctx, out_grad):
        ranks_bev, depth, feat, ranks_feat, ranks_depth, n_kept_dev = \
            ctx.saved_tensors
        dev = depth.device
        if n_kept_dev is not None:
            # padded index buffers from the sync-free path: trim (one sync,
            # backward only)
            n = int(n_kept_dev.item())
            ranks_bev, ranks_feat, ranks_depth = \
                ranks_bev[:n], ranks_feat[:n], ranks_depth[:n]
        depth_grad = depth.new_zeros(depth.shape)
        feat_grad = feat.new_zeros(feat.shape)
        if ranks_bev.shape[0] > 0:
            ranks_feat, ranks_depth, ranks_bev, starts_bp, lengths_bp = \
                _feat_intervals(ranks_feat, ranks_depth, ranks_bev)
            out_grad = out_grad.contiguous().float()  # (B,C,Z,Y,X)
            B, C, Z, Y, X = out_grad.shape
            with torch.cuda.device(dev):
                rc = _lib.lib().fbbev_bev_pool_v2_bwd_bczyx(
                    _lib.ptr(out_grad), _lib.ptr(depth), _lib.ptr(feat),
                    _lib.ptr(ranks_depth), _lib.ptr(ranks_feat),
                    _lib.ptr(ranks_bev), _lib.ptr(starts_bp),
                    _lib.ptr(lengths_bp), lengths_bp.shape[0], C, Z * Y * X,
                    _lib.ptr(depth_grad), _lib.ptr(feat_grad),
                    _lib.stream_ptr(dev))
            _lib.check(rc, 'fbbev_bev_pool_v2_bwd_bczyx')
        return (depth_grad, feat_grad) + (None,) * 9


def bev_pool_v2_dense(depth, feat, ranks_depth, ranks_feat, ranks_bev,
                      bev_feat_shape, interval_starts, interval_lengths,
                      n_intervals_dev=None, n_kept_dev=None, plan=None):
    """Fused ``bev_pool_v2``: returns contiguous ``(B, C, Z, Y, X)``.

    ``n_intervals_dev`` / ``n_kept_dev`` (0-dim int32 CUDA tensors) mark the
    live prefix of over-allocated index buffers, as produced by
    :func:`voxel_pooling_prepare_v2` with ``sync=False``.
    """
    return _BevPoolV2Dense.apply(depth, feat, ranks_depth, ranks_feat,
                                 ranks_bev, bev_feat_shape, interval_starts,
                                 interval_lengths, n_intervals_dev, n_kept_dev,
                                 plan)


class DeferredVolume:
    """The pooled volume held as interval sums, materialised on demand.

    FBOCC's glue around the two projections (fbocc.py:339, 357-366)

        bev_feat = forward_projection(...)
        refined  = backward_projection(..., lss_bev=bev_feat.mean(-1), ...)
        bev_feat = refined[..., None] + bev_feat

    reads the dense volume twice and writes it twice (820 MB for 200x200x16 x
    80).  Here the forward projection stops after the interval-sum stage;
    :meth:`mean_z` produces ``bev_feat.mean(-1)`` from the ~n_int interval sums
    and :meth:`materialize` writes the volume ONCE, adding the refined BEV to
    every Z slice on the way out.  Inference only (no autograd)."""

    def __init__(self, shape, idx, ws, ws_bytes, n_int_cap, n_points):
        self.shape = shape            # (B, Z, Y, X, C)
        self._idx, self._ws, self._ws_bytes = idx, ws, ws_bytes
        self._cap, self._n_points = n_int_cap, n_points

    def mean_z(self):
        """``volume.mean(-1)`` of the (B, C, Y, X, Z) view: (B, C, Y, X), a
        channels-last-strided view of a token-major (B, Y*X, C) buffer (the
        layout BackwardProjection's BEV queries use)."""
        B, Z, Y, X, C = self.shape
        st, ln = self._idx
        dev = self._ws.device
        lss = torch.empty((B, Y * X, C), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.lib().fbbev_bev_pool_v2_zmean_planned(
                _lib.ptr(st), _lib.ptr(ln), self._cap, self._n_points, C, B,
                Z * Y * X, Y * X, _lib.ptr(lss), _lib.ptr(self._ws),
                self._ws_bytes, _lib.stream_ptr(dev))
        _lib.check(rc, 'fbbev_bev_pool_v2_zmean_planned')
        return lss.view(B, Y, X, C).permute(0, 3, 1, 2)

    def materialize(self, add=None):
        """The contiguous (B, C, Z, Y, X) volume, plus ``add[..., None]``
        ((B, C, Y, X), e.g. the refined BEV) broadcast over Z when given."""
        B, Z, Y, X, C = self.shape
        st, ln = self._idx
        dev = self._ws.device
        out = torch.empty((B, C, Z, Y, X), dtype=torch.float32, device=dev)
        if add is not None:
            _lib.require_cuda(add)
            assert tuple(add.shape) == (B, C, Y, X), add.shape
            add = add.contiguous().float()
        with torch.cuda.device(dev):
            rc = _lib.lib().fbbev_bev_pool_v2_write_planned(
                _lib.ptr(st), _lib.ptr(ln), self._cap, self._n_points, C, B,
                Z * Y * X, Y * X, _lib.ptr(add), _lib.ptr(out),
                _lib.ptr(self._ws), self._ws_bytes, _lib.stream_ptr(dev))
        _lib.check(rc, 'fbbev_bev_pool_v2_write_planned')
        return out


def deferred_supported(bev_feat_shape):
    B, Z, Y, X, C = (int(s) for s in bev_feat_shape)
    return (C % 4 == 0 and C <= 512 and (Z * Y * X) % 4 == 0 and
            (Y * X) % 4 == 0 and B * Z * Y * X <= 2 ** 31 - 1)


def bev_pool_v2_deferred(depth, feat, ranks_depth, ranks_feat, ranks_bev,
                         bev_feat_shape, interval_starts, interval_lengths,
                         n_intervals_dev=None, plan=None):
    """Plan + interval sums of the dense op; returns a :class:`DeferredVolume`
    (``fbbev_bev_pool_v2_plan`` + ``fbbev_bev_pool_v2_sums_planned``)."""
    dev = _lib.require_cuda(depth, feat, ranks_depth, ranks_feat, ranks_bev,
                            interval_starts, interval_lengths, n_intervals_dev)
    B, Z, Y, X, C = (int(s) for s in bev_feat_shape)
    assert feat.shape[-1] == C and deferred_supported(bev_feat_shape)
    depth = depth.contiguous().float()
    feat = feat.contiguous().float()
    rb, rd, rf = (t.contiguous().int() for t in
                  (ranks_bev, ranks_depth, ranks_feat))
    st, ln = interval_starts.contiguous().int(), interval_lengths.contiguous().int()
    L = _lib.lib()
    n_points = rb.shape[0]
    cap = min(ln.shape[0], B * Z * Y * X)
    ready = _usable_plan(plan, (B, Z, Y, X, C), n_points)
    if ready is not None and ready[2] == cap:
        ws, ws_bytes, _ = ready
    else:
        ready = None
        ws_bytes = L.fbbev_bev_pool_v2_dense_workspace_bytes(B, Z * Y * X, cap,
                                                             n_points, C)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        sp = _lib.stream_ptr(dev)
        if ready is None:
            _lib.check(L.fbbev_bev_pool_v2_plan(
                _lib.ptr(rb), _lib.ptr(st), _lib.ptr(ln), cap,
                _lib.ptr(n_intervals_dev), n_points, C, B, Z * Y * X,
                _lib.ptr(ws), ws_bytes, sp), 'fbbev_bev_pool_v2_plan')
        _lib.check(L.fbbev_bev_pool_v2_sums_planned(
            _lib.ptr(depth), _lib.ptr(feat), _lib.ptr(rd), _lib.ptr(rf),
            _lib.ptr(rb), _lib.ptr(st), _lib.ptr(ln), cap, n_points, C, B,
            Z * Y * X, _lib.ptr(ws), ws_bytes, sp),
            'fbbev_bev_pool_v2_sums_planned')
    return DeferredVolume((B, Z, Y, X, C), (st, ln), ws, ws_bytes, cap, n_points)


def bev_pool_v2(depth, feat, ranks_depth, ranks_feat, ranks_bev,
                bev_feat_shape, interval_starts, interval_lengths):
    """Drop-in for the reference ``bev_pool_v2`` (bev_pool.py:84-90).

    depth ``(B,N,D,H,W)``, feat ``(B,N,H,W,C)`` (any strides), int index
    tensors as returned by ``voxel_pooling_prepare_v2``; ``bev_feat_shape`` =
    ``(B, Z, Y, X, C)``.  Returns contiguous ``(B, C, Z, Y, X)``.
    Requires the interval list to be ordered by voxel rank (it always is when
    it comes from ``voxel_pooling_prepare_v2``); use ``QuickCumsumCuda`` for
    arbitrary interval lists.
    """
    return bev_pool_v2_dense(depth, feat, ranks_depth, ranks_feat, ranks_bev,
                             bev_feat_shape, interval_starts, interval_lengths)


class VoxelIndex:
    """Index tensors of one ``voxel_pooling_prepare_v2`` call, device-resident.

    The five index tensors are over-allocated (``n_points`` entries); the live
    prefix lengths are ``counts[0]`` (points kept) and ``counts[1]``
    (intervals), int32 on the device.  Nothing here forces a host sync.
    """

    def __init__(self, ranks_bev, ranks_depth, ranks_feat, interval_starts,
                 interval_lengths, counts, plan=None):
        self.ranks_bev = ranks_bev
        self.ranks_depth = ranks_depth
        self.ranks_feat = ranks_feat
        self.interval_starts = interval_starts
        self.interval_lengths = interval_lengths
        self.counts = counts
        # (workspace, bytes, shape (B, Z, Y, X, C), n_int_cap) of a dense-pooling
        # plan the index builder filled while scanning (no plan launch needed)
        self.plan = plan

    @property
    def n_kept_dev(self):
        return self.counts[0]

    @property
    def n_intervals_dev(self):
        return self.counts[1]

    def trimmed(self):
        """Exact-length tensors as the reference returns them (host sync);
        five ``None`` when nothing is kept (view_transformer.py:598-599)."""
        n_kept, n_int = (int(v) for v in self.counts.tolist())
        if n_int == 0:
            return None, None, None, None, None
        return (self.ranks_bev[:n_kept], self.ranks_depth[:n_kept],
                self.ranks_feat[:n_kept], self.interval_starts[:n_int],
                self.interval_lengths[:n_int])


def _plan_workspace(L, dev, B, gs, n_pts, pool_channels):
    """Dense-pooling workspace for the index builder to fill the plan into, or
    ``(None, 0, None)`` when the shape is not covered."""
    if not pool_channels:
        return None, 0, None
    X, Y, Z = int(gs[0]), int(gs[1]), int(gs[2])
    zyx = Z * Y * X
    C = int(pool_channels)
    if C > 512 or B * zyx > 2 ** 31 - 1 or \
            not L.fbbev_voxel_prepare_can_plan(C, zyx):
        return None, 0, None
    cap = min(n_pts, B * zyx)
    nbytes = L.fbbev_bev_pool_v2_dense_workspace_bytes(B, zyx, cap, n_pts, C)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    return ws, nbytes, (ws, nbytes, (B, Z, Y, X, C), cap)


def _depth_arg(depth, n_pts, dev):
    """Contiguous fp32 depth tensor for the depth-threshold sparsification
    (necks/view_transformer.py:556-557), or None."""
    if depth is None:
        return None
    d = depth.detach().contiguous().float()
    assert d.numel() == n_pts and d.device == dev
    return d


def voxel_pooling_prepare_v2(coor, grid_lower_bound, grid_interval, grid_size,
                             pool_channels=None, depth=None,
                             depth_thresh=0.01):
    """Device ``voxel_pooling_prepare_v2`` (view_transformer.py:547-605).

    coor ``(B,N,D,H,W,3)`` fp32 CUDA; the three grid descriptors are the
    3-element float32 tensors of ``create_grid_infos`` (:384-387) -- host
    tensors or python sequences.  Returns a :class:`VoxelIndex`;
    ``pool_channels`` = C of the pooling call that will consume the index lets
    the builder fill that call's plan on the way (``VoxelIndex.plan``).
    ``depth`` (B,N,D,H,W): additionally drop the points whose depth
    probability is <= ``depth_thresh`` -- LSSViewTransformer2's sparsification
    (necks/view_transformer.py:556-557).
    """
    dev = _lib.require_cuda(coor)
    coor = coor.contiguous().float()
    B, N, D, H, W, three = coor.shape
    assert three == 3
    lo = [float(v) for v in torch.as_tensor(grid_lower_bound).float().cpu()]
    iv = [float(v) for v in torch.as_tensor(grid_interval).float().cpu()]
    gs = [float(v) for v in torch.as_tensor(grid_size).float().cpu()]
    n_pts = B * N * D * H * W
    n_vox = B * int(gs[0]) * int(gs[1]) * int(gs[2])
    L = _lib.lib()
    idx = torch.empty((5, n_pts), dtype=torch.int32, device=dev)
    counts = torch.empty(2, dtype=torch.int32, device=dev)
    ws_bytes = L.fbbev_voxel_prepare_workspace_bytes(n_pts, n_vox)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    pws, pbytes, plan = _plan_workspace(L, dev, B, gs, n_pts, pool_channels)
    dprob = _depth_arg(depth, n_pts, dev)
    with torch.cuda.device(dev):
        rc = L.fbbev_voxel_prepare_sparse(
            _lib.ptr(coor), _lib.ptr(dprob), float(depth_thresh), B, N, D, H,
            W, _lib.c_floats(lo), _lib.c_floats(iv),
            _lib.c_floats(gs), _lib.ptr(idx[0]), _lib.ptr(idx[1]),
            _lib.ptr(idx[2]), _lib.ptr(idx[3]), _lib.ptr(idx[4]),
            _lib.ptr(counts), _lib.ptr(ws), ws_bytes, int(pool_channels or 0),
            _lib.ptr(pws), pbytes, _lib.stream_ptr(dev))
    _lib.check(rc, 'fbbev_voxel_prepare')
    return VoxelIndex(idx[0], idx[1], idx[2], idx[3], idx[4], counts, plan)


def voxel_pooling_prepare_from_cams(frustum_axes, inv_post_rots, post_trans,
                                    cam2ego, trans, bda, depth_bins,
                                    grid_lower_bound, grid_interval,
                                    grid_size, pool_channels=None, depth=None,
                                    depth_thresh=0.01):
    """``get_lidar_coor`` + ``voxel_pooling_prepare_v2`` in one pass
    (view_transformer.py:458-498, 547-605): the (B,N,D,H,W,3) coordinate
    tensor is never materialised (``fbbev_voxel_prepare_cams``).

    frustum_axes = (u [W], v [H], d [D]) CUDA float tensors;
    inv_post_rots / cam2ego (B,N,3,3); post_trans / trans (B,N,3); bda (B,3,3).
    Returns a :class:`VoxelIndex`, bit-identical to the two-step route: the
    kernel rounds every 3x3 product as torch's broadcast matmul does on this
    device (see include/fbbev_b200.h, ``_lib.matmul_order_flags``)."""
    fu, fv, fd = (t.contiguous().float() for t in frustum_axes)
    dev = _lib.require_cuda(fu, fv, fd, inv_post_rots, post_trans, cam2ego,
                            trans, bda)
    order = _lib.matmul_order_flags(inv_post_rots, cam2ego, bda)
    mats = [t.contiguous().float() for t in
            (inv_post_rots, post_trans, cam2ego, trans, bda)]
    B, N = mats[0].shape[:2]
    D, H, W = fd.shape[0], fv.shape[0], fu.shape[0]
    assert D == depth_bins
    lo = [float(v) for v in torch.as_tensor(grid_lower_bound).float().cpu()]
    iv = [float(v) for v in torch.as_tensor(grid_interval).float().cpu()]
    gs = [float(v) for v in torch.as_tensor(grid_size).float().cpu()]
    n_pts = B * N * D * H * W
    n_vox = B * int(gs[0]) * int(gs[1]) * int(gs[2])
    L = _lib.lib()
    idx = torch.empty((5, n_pts), dtype=torch.int32, device=dev)
    counts = torch.empty(2, dtype=torch.int32, device=dev)
    ws_bytes = L.fbbev_voxel_prepare_workspace_bytes(n_pts, n_vox)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    pws, pbytes, plan = _plan_workspace(L, dev, B, gs, n_pts, pool_channels)
    dprob = _depth_arg(depth, n_pts, dev)
    with torch.cuda.device(dev):
        rc = L.fbbev_voxel_prepare_cams_sparse(
            _lib.ptr(fu), _lib.ptr(fv), _lib.ptr(fd), _lib.ptr(mats[0]),
            _lib.ptr(mats[1]), _lib.ptr(mats[2]), _lib.ptr(mats[3]),
            _lib.ptr(mats[4]), order, _lib.ptr(dprob), float(depth_thresh), B,
            N, D, H, W, _lib.c_floats(lo),
            _lib.c_floats(iv), _lib.c_floats(gs), _lib.ptr(idx[0]),
            _lib.ptr(idx[1]), _lib.ptr(idx[2]), _lib.ptr(idx[3]),
            _lib.ptr(idx[4]), _lib.ptr(counts), _lib.ptr(ws), ws_bytes,
            int(pool_channels or 0), _lib.ptr(pws), pbytes,
            _lib.stream_ptr(dev))
    _lib.check(rc, 'fbbev_voxel_prepare_cams')
    return VoxelIndex(idx[0], idx[1], idx[2], idx[3], idx[4], counts, plan)
