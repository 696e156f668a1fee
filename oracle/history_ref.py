"""CPU restatement of the sampling half of ``FBOCC.fuse_history``
(mmdet3d/models/fbbev/detectors/fbocc.py:170-205, 263-275, 286).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Same contract as
``fbbev_b200.view_transformation.temporal_fusion.history_warp``; pinned against
the golden vectors recorded from the reference's own ``fuse_history``
(tests/golden/t_fuse_history.npz, tests/test_temporal_cpu.py)."""
import torch
import torch.nn.functional as F


def history_warp_cpu(history, flow, out, ch_offset):
    n, mc, z, h, w = history.shape
    dt, dev = history.dtype, history.device
    xs = torch.linspace(0, w - 1, w, dtype=dt, device=dev).view(1, w, 1).expand(h, w, z)
    ys = torch.linspace(0, h - 1, h, dtype=dt, device=dev).view(h, 1, 1).expand(h, w, z)
    zs = torch.linspace(0, z - 1, z, dtype=dt, device=dev).view(1, 1, z).expand(h, w, z)
    grid = torch.stack((xs, ys, zs, torch.ones_like(xs)), -1).view(
        1, h, w, z, 4).expand(n, h, w, z, 4).reshape(n, h, w, z, 4, 1)   # :174-177
    grid = flow.view(n, 1, 1, 1, 4, 4) @ grid                              # :199
    norm = torch.tensor([w - 1.0, h - 1.0, z - 1.0], dtype=dt, device=dev)
    grid = grid[:, :, :, :, :3, 0] / norm.view(1, 1, 1, 1, 3) * 2.0 - 1.0  # :202-203
    sampled = F.grid_sample(history, grid.permute(0, 3, 1, 2, 4),
                            align_corners=True, mode='bilinear')           # :275
    out[:, ch_offset:ch_offset + mc] = sampled
    return out
