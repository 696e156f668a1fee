"""Frame sharding of the view transformation across ranks.

Both halves of the path are independent per sample/frame: the batch index is
only the top term of the voxel rank (view_transformer.py:573-575, 586-587) and
the cross-attention processes ``bs * num_cams`` independent camera batches
(spatial_cross_attention_depth.py:190-206).  So frames are split across ranks
(one process per GPU) and NO collective sits on the data path; cameras of one
frame stay together because the cross-attention averages over the cameras that
see a query (:213-216).

The reference has no collective on this path either.
``gather_bev`` is the optional exchange for a consumer that needs every frame's
BEV on every rank: one ``all_gather_into_tensor`` of the refined 2-D BEV
(B, C, bev_h, bev_w) -- 16x smaller than the voxel volume, which stays sharded.
"""
import torch
import torch.distributed as dist

__all__ = ['frame_slice', 'shard_frames', 'gather_bev', 'GatherBuffer']


def frame_slice(n_frames, rank, world_size):
    """Contiguous, balanced split: rank r owns frames [lo, hi)."""
    base, extra = divmod(n_frames, world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def shard_frames(tensors, rank, world_size, dim=0):
    """Slice every tensor of a (nested) tuple/list along the frame dim."""
    if isinstance(tensors, torch.Tensor):
        lo, hi = frame_slice(tensors.shape[dim], rank, world_size)
        return tensors.narrow(dim, lo, hi - lo)
    return type(tensors)(shard_frames(t, rank, world_size, dim)
                         for t in tensors)


def gather_bev(local_bev, group=None):
    """All-gather per-rank BEV tensors (b_local, C, H, W) along the frame dim.

    Every rank must hold the same number of local frames (pad the last shard
    if the frame count is not divisible).  Works on NCCL (CUDA tensors) and
    gloo (CPU tensors, used by the CPU tests)."""
    if not dist.is_available() or not dist.is_initialized():
        return local_bev
    world = dist.get_world_size(group)
    if world == 1:
        return local_bev
    local_bev = local_bev.contiguous()
    out = local_bev.new_empty((world * local_bev.shape[0],) +
                              tuple(local_bev.shape[1:]))
    dist.all_gather_into_tensor(out, local_bev, group=group)
    return out


class GatherBuffer:
    """Registered all-gather buffer for the refined BEV of every frame.

    ``slot`` is this rank's (b_local, C, H, W) view inside the gathered
    (b_total, C, H, W) tensor: pass it as ``BackwardProjection.forward(...,
    out=buf.slot)`` so the module's final layout pass writes straight into the
    exchange buffer, then ``buf.gather()`` runs the in-place
    ``all_gather_into_tensor`` (NCCL in-place semantics: input = the rank's own
    chunk of the output; SURVEY.md section 8e "fusion with the collective" --
    a zero-copy hand-off, no staging copy on either side)."""

    def __init__(self, b_local, C, H, W, device, dtype=torch.float32,
                 group=None):
        self.group = group
        init = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if init else 1
        self.rank = dist.get_rank(group) if init else 0
        self.full = torch.empty((self.world * b_local, C, H, W), dtype=dtype,
                                device=device)
        self.slot = self.full[self.rank * b_local:(self.rank + 1) * b_local]

    def gather(self):
        if self.world > 1:
            dist.all_gather_into_tensor(self.full, self.slot, group=self.group)
        return self.full
