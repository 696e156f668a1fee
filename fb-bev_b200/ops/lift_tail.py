"""The tail of the depth net as the producer of the pooling op's inputs
(``fbbev_lift_tail_fwd``, csrc/lift_tail.cu).

Replaces, in one launch,

* ``depth = depth_digit.softmax(dim=1)``  (CM_DepthNet.forward,
  mmdet3d/models/fbbev/modules/depth_net.py:359; LSSViewTransformer.forward,
  mmdet3d/models/necks/view_transformer.py:313-321, 710-718, 1094-1096) and
* ``feat.permute(0, 1, 3, 4, 2)`` + ``feat.contiguous()`` -- the transposing
  copy the pooling op starts with (view_transformer.py:530, bev_pool.py:19),

so that ``feat`` reaches ``bev_pool_v2`` already in its (B, N, H, W, C) layout.
"""
import torch

from .. import _lib

__all__ = ['lift_tail', 'LiftTail']


def _image_stride(t, hw):
    """``t`` (BN, K, H, W): per-image stride when every image is a dense
    (K, H*W) block (a channel slice of a wider NCHW tensor qualifies)."""
    if t.stride(3) == 1 and t.stride(2) == t.shape[3] and t.stride(1) == hw:
        return t.stride(0)
    return None


def _lift_tail_fwd(depth_logits, context):
    """depth_logits (BN, D, H, W), context (BN, C, H, W) fp32 CUDA ->
    depth (BN, D, H, W) = softmax over D; feat (BN, H, W, C)."""
    dev = _lib.require_cuda(depth_logits, context)
    BN, D, H, W = depth_logits.shape
    C = context.shape[1]
    assert tuple(context.shape) == (BN, C, H, W)
    hw = H * W
    lg = depth_logits.float()
    cx = context.float()
    ls, cs = _image_stride(lg, hw), _image_stride(cx, hw)
    if ls is None:
        lg = lg.contiguous()
        ls = D * hw
    if cs is None:
        cx = cx.contiguous()
        cs = C * hw
    depth = torch.empty((BN, D, H, W), dtype=torch.float32, device=dev)
    feat = torch.empty((BN, H, W, C), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().fbbev_lift_tail_fwd(
            _lib.ptr(lg), ls, _lib.ptr(cx), cs, BN, D, C, hw, _lib.ptr(depth),
            _lib.ptr(feat), _lib.stream_ptr(dev))
    _lib.check(rc, 'fbbev_lift_tail_fwd')
    return depth, feat


class LiftTail(torch.autograd.Function):
    """Differentiable wrapper: the backward is the softmax backward and the
    inverse permutation (torch element-wise ops; training is not the timed
    path)."""

    @staticmethod
    def forward(ctx, depth_logits, context):
        depth, feat = _lift_tail_fwd(depth_logits.detach(), context.detach())
        ctx.save_for_backward(depth)
        return depth, feat

    @staticmethod
    def backward(ctx, g_depth, g_feat):
        depth, = ctx.saved_tensors
        g_logits = g_ctx = None
        if ctx.needs_input_grad[0] and g_depth is not None:
            g_logits = depth * (g_depth - (g_depth * depth).sum(1, keepdim=True))
        if ctx.needs_input_grad[1] and g_feat is not None:
            g_ctx = g_feat.permute(0, 3, 1, 2)
        return g_logits, g_ctx


def lift_tail(depth_logits, context):
    """``(softmax(depth_logits, dim=1), context.permute(0, 2, 3, 1).contiguous())``
    in one launch.  depth_logits (BN, D, H, W), context (BN, C, H, W); both may
    be channel slices of one tensor (``x[:, :D]``, ``x[:, D:D + C]``)."""
    if torch.is_grad_enabled() and (depth_logits.requires_grad or
                                    context.requires_grad):
        return LiftTail.apply(depth_logits, context)
    return _lift_tail_fwd(depth_logits, context)
