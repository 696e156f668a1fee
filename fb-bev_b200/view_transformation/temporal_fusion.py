"""Temporal fusion of the voxel history -- the stage right after the path.

Mirror of the history branch of ``FBOCC`` in
``mmdet3d/models/fbbev/detectors/fbocc.py``:

==============================  =============================================
here                            reference
==============================  =============================================
generate_forward_transformation_matrix   fbocc.py:37-42
TemporalFusion.generate_flow    FBOCC.generate_grid, the 4x4 product (:181-197)
TemporalFusion.fuse_history     FBOCC.fuse_history (:207-319)
==============================  =============================================

``TemporalFusion`` owns exactly the state and the two conv stacks the detector
keeps for this (attribute names as in the reference, so the ``history_*``
entries of an FB-OCC checkpoint load under the same keys):
``history_bev / history_seq_ids / history_forward_augs / history_sweep_time``,
``history_keyframe_time_conv`` and ``history_keyframe_cat_conv``.

What changed underneath: the reference builds a (n, H, W, Z, 4, 4) batched
matmul and a 5-D sampling grid, runs a 3-D ``F.grid_sample`` over the whole
history (T x C = 1280 channels of 8 x 100 x 100 -> 410 MB per sample), copies it
again in ``torch.cat([curr_bev, sampled])`` and a third time in
``feats_cat[:, :-C].detach().clone()``.  Here ``fbbev_history_warp`` evaluates
the flow per voxel and writes the trilinear samples straight into the
concatenated buffer (one read + one write of the history), and the buffer is
double-buffered so that the next step's history is a VIEW of this step's
concatenation: no ``cat``, no ``clone``.  The two 1x1x1 convolutions stay
PyTorch (cuDNN GEMMs, outside the memory-bound part).
"""
import torch
import torch.nn as nn

from .. import _lib
from ..registry import BaseModule

__all__ = ['generate_forward_transformation_matrix', 'history_warp',
           'TemporalFusion']


def generate_forward_transformation_matrix(bda, img_meta_dict=None):
    """(b, 3, 3) BEV augmentation -> homogeneous (b, 4, 4) (fbocc.py:37-42)."""
    b = bda.size(0)
    hom = torch.eye(4, device=bda.device, dtype=bda.dtype)[None].repeat(b, 1, 1)
    hom[:, :3, :3] = bda
    return hom


def history_warp(history, flow, out, ch_offset):
    """``out[:, ch_offset:ch_offset + MC] = grid_sample(history, grid(flow))``
    (fbocc.py:199-204, 275): ``fbbev_history_warp``.

    history (n, MC, Z, H, W) fp32 CUDA, dense inside a sample (it may be a
    channel slice of a larger buffer); flow (n, 4, 4) voxel-index flow; out
    (n, C_total, Z, H, W) contiguous."""
    dev = _lib.require_cuda(history, flow, out)
    n, mc, Z, H, W = history.shape
    if history[0].is_contiguous() is False:
        history = history.contiguous()
    assert out.is_contiguous()
    assert history.dtype == out.dtype == torch.float32
    flow = flow.contiguous().float()
    assert out.shape[0] == n and tuple(out.shape[2:]) == (Z, H, W)
    bstride = history.stride(0) if n > 1 else mc * Z * H * W
    with torch.cuda.device(dev):
        rc = _lib.lib().fbbev_history_warp(
            _lib.ptr(history), bstride, _lib.ptr(flow), n, mc, Z, H, W,
            _lib.ptr(out), out.shape[1], int(ch_offset), _lib.stream_ptr(dev))
    _lib.check(rc, 'fbbev_history_warp')
    return out


class TemporalFusion(BaseModule):
    """The history branch of FBOCC as a module (fbocc.py:100-130, 207-319).

    Args mirror the detector's: ``single_bev_num_channels`` (80),
    ``history_cat_num`` (16), ``history_cat_conv_out_channels`` (None ->
    ``single_bev_num_channels``), ``interpolation_mode`` ('bilinear' only),
    ``do_history``; ``dx`` / ``bx`` are the forward projection's voxel size and
    first-voxel centre (``forward_projection.dx / .bx``, fbocc.py:183-188).
    """

    def __init__(self, dx, bx, single_bev_num_channels=80, history_cat_num=16,
                 history_cat_conv_out_channels=None, do_history=True,
                 interpolation_mode='bilinear', norm='BN'):
        super().__init__()
        if interpolation_mode != 'bilinear':
            raise NotImplementedError('history warp: bilinear only')
        self.single_bev_num_channels = C = single_bev_num_channels
        self.do_history = do_history
        self.interpolation_mode = interpolation_mode
        self.history_cat_num = history_cat_num
        self.history_cam_sweep_freq = 0.5  # seconds between frames (:105)
        out_ch = history_cat_conv_out_channels or C
        bn = nn.SyncBatchNorm if norm == 'SyncBN' else nn.BatchNorm3d
        # embed each frame with its temporal offset, then mix the frames
        self.history_keyframe_time_conv = nn.Sequential(
            nn.Conv3d(C + 1, C, kernel_size=1, padding=0, stride=1), bn(C),
            nn.ReLU(inplace=True))
        self.history_keyframe_cat_conv = nn.Sequential(
            nn.Conv3d(C * (history_cat_num + 1), out_ch, kernel_size=1,
                      padding=0, stride=1), bn(out_ch), nn.ReLU(inplace=True))
        self.register_buffer('dx', torch.as_tensor(dx).float().clone(),
                             persistent=False)
        self.register_buffer('bx', torch.as_tensor(bx).float().clone(),
                             persistent=False)
        self.history_sweep_time = None
        self.history_bev = None
        self.history_seq_ids = None
        self.history_forward_augs = None
        self._bufs = None   # two concatenation buffers, used alternately
        self._turn = 0

    def reset(self):
        self.history_bev = self.history_seq_ids = None
        self.history_forward_augs = self.history_sweep_time = None

    # -- fbocc.py:181-197 ---------------------------------------------------
    def generate_flow(self, history_forward_augs, forward_augs,
                      curr_to_prev_ego_rt):
        """Voxel-index flow (n, 4, 4): current grid -> metres -> undo the
        current augmentation -> previous ego frame -> previous augmentation ->
        previous grid."""
        dx, bx = self.dx.to(forward_augs), self.bx.to(forward_augs)
        feat2bev = torch.zeros((4, 4), dtype=forward_augs.dtype,
                               device=forward_augs.device)
        feat2bev[0, 0], feat2bev[1, 1], feat2bev[2, 2] = dx[0], dx[1], dx[2]
        feat2bev[0, 3] = bx[0] - dx[0] / 2.
        feat2bev[1, 3] = bx[1] - dx[1] / 2.
        feat2bev[2, 3] = bx[2] - dx[2] / 2.
        feat2bev[3, 3] = 1
        feat2bev = feat2bev.view(1, 4, 4)
        return (torch.inverse(feat2bev) @ history_forward_augs @
                curr_to_prev_ego_rt @ torch.inverse(forward_augs) @ feat2bev)

    def _concat_buffer(self, n, c_total, zhw, like):
        shape = (n, c_total) + tuple(zhw)
        if (self._bufs is None or self._bufs[0].shape != shape or
                self._bufs[0].device != like.device):
            self._bufs = [torch.empty(shape, dtype=torch.float32,
                                      device=like.device) for _ in range(2)]
            self._turn = 0
        self._turn ^= 1
        return self._bufs[self._turn]

    # -- fbocc.py:207-319 ---------------------------------------------------
    @torch.no_grad()
    def _align(self, curr_bev, seq_ids, start_of_sequence, forward_augs,
               curr_to_prev_ego_rt):
        """State update + warp.  Returns feats_cat (n, (T+1)*C, Z, H, W) and
        the (n, T+1) sweep times."""
        T, C = self.history_cat_num, self.single_bev_num_channels
        n, c_, Z, H, W = curr_bev.shape
        assert c_ == C
        if self.history_bev is None:                              # :228-238
            self.history_bev = curr_bev.repeat(1, T, 1, 1, 1)
            self.history_seq_ids = seq_ids.clone()
            self.history_forward_augs = forward_augs.clone()
            self.history_sweep_time = curr_bev.new_zeros(n, T)
        assert self.history_bev.dtype == torch.float32
        assert (self.history_seq_ids != seq_ids)[~start_of_sequence].sum() == 0, \
            "{}, {}, {}".format(self.history_seq_ids, seq_ids, start_of_sequence)
        self.history_sweep_time += 1                              # :252
        if start_of_sequence.sum() > 0:                           # :253-261
            self.history_bev[start_of_sequence] = \
                curr_bev[start_of_sequence].repeat(1, T, 1, 1, 1)
            self.history_sweep_time[start_of_sequence] = 0
            self.history_seq_ids[start_of_sequence] = seq_ids[start_of_sequence]
            self.history_forward_augs[start_of_sequence] = \
                forward_augs[start_of_sequence]
        flow = self.generate_flow(self.history_forward_augs, forward_augs,
                                  curr_to_prev_ego_rt)
        feats_cat = self._concat_buffer(n, (T + 1) * C, (Z, H, W), curr_bev)
        feats_cat[:, :C].copy_(curr_bev)
        history_warp(self.history_bev, flow, feats_cat, C)        # :263-275
        sweep = torch.cat([self.history_sweep_time.new_zeros(n, 1),
                           self.history_sweep_time], dim=1)       # :279-281
        # the new history = current frame + the T-1 most recent warped frames:
        # the leading channels of feats_cat (the reference clones them, :311)
        self.history_bev = feats_cat[:, :T * C]
        self.history_sweep_time = sweep[:, :-1]
        self.history_forward_augs = forward_augs.clone()
        return feats_cat, sweep

    def fuse_history(self, curr_bev, img_metas, bda):
        """curr_bev (n, C, H, W, Z) -- the (B, C, Y, X, Z) view the forward
        projection returns; img_metas: per-sample dicts with
        ``sequence_group_idx``, ``start_of_sequence``, ``curr_to_prev_ego_rt``
        (4x4); bda (n, 3, 3).  Returns (n, C_out, H, W, Z)."""
        assert curr_bev.dim() == 5, "voxel features (n, C, H, W, Z) expected"
        curr = curr_bev.permute(0, 1, 4, 2, 3).float()   # n, c, z, h, w (:211)
        if not curr.is_contiguous():
            curr = curr.contiguous()
        dev = curr.device
        seq_ids = torch.LongTensor([m['sequence_group_idx']
                                    for m in img_metas]).to(dev)
        start = torch.BoolTensor([m['start_of_sequence']
                                  for m in img_metas]).to(dev)
        forward_augs = generate_forward_transformation_matrix(bda.float())
        c2p = torch.stack([torch.as_tensor(m['curr_to_prev_ego_rt'])
                           for m in img_metas]).to(curr)
        feats_cat, sweep = self._align(curr, seq_ids, start, forward_augs, c2p)
        T, C = self.history_cat_num, self.single_bev_num_channels
        n, _, Z, H, W = feats_cat.shape
        if torch.is_grad_enabled() and curr.requires_grad:
            # training: the current frame keeps its graph (the history is
            # detached in the reference as well, :241)
            feats_cat = torch.cat([curr, feats_cat[:, C:]], dim=1)
        feats = feats_cat.reshape(n, T + 1, C, Z, H, W)           # :289-290
        feats = torch.cat(
            [feats, sweep[:, :, None, None, None, None].repeat(
                1, 1, 1, Z, H, W) * self.history_cam_sweep_freq], dim=2)
        feats = self.history_keyframe_time_conv(
            feats.reshape(-1, C + 1, Z, H, W)).reshape(n, T + 1, -1, Z, H, W)
        feats = self.history_keyframe_cat_conv(
            feats.reshape(n, -1, Z, H, W))                        # :307-309
        if not self.do_history:
            self.history_bev = None
        return feats.permute(0, 1, 3, 4, 2).clone()               # :313-317

    forward = fuse_history
