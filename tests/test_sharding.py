"""world_size-2 gloo tests of the frame sharding (CPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_frame_slice_partitions():
    from fbbev_b200.sharding import frame_slice
    for n in (1, 2, 7, 16):
        for w in (1, 2, 3, 8):
            spans = [frame_slice(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fbbev_b200.sharding import gather_bev, shard_frames
    g = torch.Generator().manual_seed(0)
    frames = torch.randn(4, 3, 5, 6, generator=g)       # 4 frames total
    cams = (torch.randn(4, 6, 3, 3, generator=g), torch.randn(4, 6, 3, generator=g))
    local = shard_frames(frames, rank, world)
    local_cams = shard_frames(cams, rank, world)
    assert local.shape[0] == 2 and local_cams[0].shape[0] == 2
    # per-frame "processing" (stand-in for the view transformation)
    out = local * 2 + 1
    full = gather_bev(out)
    ok = torch.equal(full, frames * 2 + 1)
    # zero-copy hand-off: the producer writes into its slot of the gather
    # buffer, the collective runs in place
    from fbbev_b200.sharding import GatherBuffer
    buf = GatherBuffer(2, 3, 5, 6, "cpu")
    assert buf.slot.data_ptr() == buf.full[rank * 2:].data_ptr()
    buf.slot.copy_(local * 2 + 1)           # stands for forward(..., out=slot)
    ok = ok and torch.equal(buf.gather(), frames * 2 + 1)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, bool(ok), tuple(full.shape)))


def test_gather_bev_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert all(shape == (4, 3, 5, 6) for _, _, shape in res)
