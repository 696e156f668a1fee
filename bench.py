#!/usr/bin/env python
"""bench.py -- headline benchmark of the forward-backward view transformation.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A *step* is one pass of the hot path over one frame per GPU of synthetic input
(BASELINE.json configs[1]: FB-OCC R50 single frame, 6 cameras 256x704, feature
map 16x44, D = 80 depth bins, C = 80, 200x200x16 voxel grid, 200x200 BEV
queries, 8 heads, 1 level, 8 points, 4 Z anchors):

    F  LSSViewTransformerFunction3D.forward(cam_params, context, depth)
         = get_lidar_coor + voxel_pooling_prepare_v2 + bev_pool_v2
    B  BackwardProjection.forward([context], lss_bev, cam_params, depth)
         = embedding/pos-enc + point_sampling + self-attn + depth-aware
           spatial cross-attention + LayerNorms + FFN

i.e. every row of SURVEY.md section 8(a).  Metric: BEV voxels/s =
frames * Z*Y*X / step time (whole job, all ranks).

Launch contract: `python bench.py --gpus 1 ...` or, for N > 1,
`python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...`
(one rank per GPU, NCCL).  Frames are independent, so ranks shard frames and no
collective sits on the data path (weak scaling: one frame per GPU).

`--impl reference` times the reference algorithm's CPU implementation (the
oracle port: the reference ships no CPU kernel for this path) on the host
cores, on the same config.  Rank 0 only.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "bev_voxels_per_sec"
UNIT = "voxels/s"
WORKLOAD = "fbocc_r50_single_frame_6cam_256x704_voxel200x200x16"

# ----------------------------------------------------------------- workload --
CFG = dict(
    grid="fbocc_200", input_size=(256, 704), downsample=16, n_cams=6, C=80,
    bev=(200, 200), dbound=[2.0, 42.0, 0.5],
    pc_range=[-40, -40, -1.0, 40, 40, 5.4],
    grid_bev=dict(x=[-40, 40, 0.4], y=[-40, 40, 0.4], z=[-1, 5.4, 1.6]),
)


def bp_config(cfg):
    E = cfg["C"]
    bev_h, bev_w = cfg["bev"]
    return dict(
        type='BackwardProjection', bev_h=bev_h, bev_w=bev_w, in_channels=E,
        out_channels=E, pc_range=cfg["pc_range"],
        transformer=dict(
            type='BEVFormer', use_cams_embeds=False, embed_dims=E,
            encoder=dict(
                type='bevformer_encoder', num_layers=1,
                pc_range=cfg["pc_range"], grid_config=cfg["grid_bev"],
                data_config=dict(input_size=cfg["input_size"]),
                return_intermediate=False,
                transformerlayers=dict(
                    type='BEVFormerEncoderLayer',
                    attn_cfgs=[
                        dict(type='MultiScaleDeformableAttention',
                             embed_dims=E, dropout=0.0, num_levels=1),
                        dict(type='DA_SpatialCrossAttention',
                             pc_range=cfg["pc_range"], dbound=cfg["dbound"],
                             dropout=0.0,
                             deformable_attention=dict(
                                 type='DA_MSDeformableAttention',
                                 embed_dims=E, num_points=8, num_levels=1),
                             embed_dims=E)],
                    ffn_cfgs=dict(type='FFN', embed_dims=E,
                                  feedforward_channels=4 * E, ffn_drop=0.0,
                                  act_cfg=dict(type='ReLU', inplace=True)),
                    feedforward_channels=4 * E, ffn_dropout=0.0,
                    operation_order=('self_attn', 'norm', 'cross_attn', 'norm',
                                     'ffn', 'norm')))),
        positional_encoding=dict(type='CustormLearnedPositionalEncoding',
                                 num_feats=E // 2, row_num_embed=bev_h,
                                 col_num_embed=bev_w))


class Workload:
    """Modules + one frame of seeded synthetic inputs on `device`."""

    def __init__(self, device, seed, frames=1):
        from fbbev_b200 import synthetic
        from fbbev_b200.registry import build_head, build_neck
        cfg = CFG
        self.cfg = cfg
        self.device = device
        self.frames = frames
        H, W = (s // cfg["downsample"] for s in cfg["input_size"])
        self.vt = build_neck(dict(
            type='LSSViewTransformerFunction3D',
            grid_config=synthetic.GRID_CONFIGS[cfg["grid"]],
            input_size=cfg["input_size"], downsample=cfg["downsample"]))
        torch.manual_seed(1234)  # identical random-init weights on every rank
        bp = build_head(bp_config(cfg))
        bp.init_weights()
        g = torch.Generator().manual_seed(4321)
        with torch.no_grad():
            for p in bp.parameters():  # non-trivial offsets / weights
                p.add_(torch.randn(p.shape, generator=g) * 0.02)
        self.bp = bp.to(device).eval()
        self.D = self.vt.D
        cam = synthetic.make_cam_params(frames, cfg["n_cams"],
                                        cfg["input_size"], jitter=1.0,
                                        seed=seed)
        depth, feat = synthetic.make_depth_feat(frames, cfg["n_cams"], self.D,
                                                H, W, cfg["C"], seed=seed)
        g = torch.Generator().manual_seed(seed + 7)
        lss = torch.randn(frames, cfg["C"], *cfg["bev"], generator=g) * 0.1
        self.host = dict(cam=[t.pin_memory() if device != "cpu" else t
                              for t in cam],
                         depth=_pin(depth, device), feat=_pin(feat, device),
                         lss=_pin(lss, device))
        self.to_device()
        gs = [int(v) for v in self.vt.grid_size]
        self.voxels_per_frame = gs[0] * gs[1] * gs[2]

    def to_device(self):
        h = self.host
        self.cam = [t.to(self.device, non_blocking=True) for t in h["cam"]]
        self.depth = h["depth"].to(self.device, non_blocking=True)
        self.feat = h["feat"].to(self.device, non_blocking=True)
        self.lss = h["lss"].to(self.device, non_blocking=True)

    def h2d_bytes(self):
        h = self.host
        return int(sum(t.numel() * t.element_size() for t in
                       list(h["cam"]) + [h["depth"], h["feat"], h["lss"]]))

    @torch.no_grad()
    def step(self):
        bev = self.vt(self.cam, self.feat, self.depth)          # (B,C,Y,X,Z)
        ref = self.bp([self.feat], None, lss_bev=self.lss,
                      cam_params=self.cam, pred_img_depth=self.depth)
        return bev, ref

    def capture(self):
        """Capture one step (both plugin forwards) into a CUDA graph: the path
        has no host synchronisation, so the ~60 launches of a step replay as
        one submission.  Inputs are read from the same device buffers
        (``to_device`` copies into them in place)."""
        self._static = dict(cam=self.cam, depth=self.depth, feat=self.feat,
                            lss=self.lss)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self.step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.graph_out = self.step()
        torch.cuda.synchronize()
        return self.graph

    def to_device_inplace(self):
        """H2D into the buffers the captured graph reads."""
        h, st = self.host, self._static
        for d, s in zip(st["cam"], h["cam"]):
            d.copy_(s, non_blocking=True)
        st["depth"].copy_(h["depth"], non_blocking=True)
        st["feat"].copy_(h["feat"], non_blocking=True)
        st["lss"].copy_(h["lss"], non_blocking=True)


def _pin(t, device):
    return t.pin_memory() if device != "cpu" else t


def e2e_streamed(w, steps, flush, barrier):
    """K end-to-end steps as a pipeline; returns ms per step (device clock).

    compute stream : wait inputs(i) -> graph replay -> copy results into
                     staging[i % 2] (after the D2H of step i-2 released it)
    H2D stream     : inputs of step i+1 into the graph's input buffers once
                     step i has finished reading them
    D2H stream     : staging[i % 2] -> pinned host[i % 2]
    """
    dev = w.graph_out[0].device
    s_c = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    stage = [[torch.empty_like(t) for t in w.graph_out] for _ in range(2)]
    host = [[torch.empty(t.shape, dtype=t.dtype).pin_memory()
             for t in w.graph_out] for _ in range(2)]

    def run(n):
        h2d_done = [torch.cuda.Event() for _ in range(n + 1)]
        comp_done = [torch.cuda.Event() for _ in range(n)]
        ready = [torch.cuda.Event() for _ in range(n)]
        d2h_done = [torch.cuda.Event() for _ in range(n)]
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0.record(s_c)
        s_in.wait_stream(s_c)
        with torch.cuda.stream(s_in):
            w.to_device_inplace()
            h2d_done[0].record(s_in)
        for i in range(n):
            s_c.wait_event(h2d_done[i])
            w.graph.replay()
            comp_done[i].record(s_c)
            with torch.cuda.stream(s_in):          # inputs of the next step
                s_in.wait_event(comp_done[i])
                w.to_device_inplace()
                h2d_done[i + 1].record(s_in)
            if i >= 2:
                s_c.wait_event(d2h_done[i - 2])
            for d, r in zip(stage[i % 2], w.graph_out):
                d.copy_(r, non_blocking=True)
            ready[i].record(s_c)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ready[i])
                for h, d in zip(host[i % 2], stage[i % 2]):
                    h.copy_(d, non_blocking=True)
                d2h_done[i].record(s_out)
        s_c.wait_stream(s_out)
        s_c.wait_stream(s_in)
        t1.record(s_c)
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / n

    run(3)
    barrier()
    ms = run(steps)
    barrier()
    # the last step's results really are on the host
    assert torch.equal(host[(steps - 1) % 2][1], w.graph_out[1].cpu())
    return ms


# ------------------------------------------------------------------ clocks --
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,"
         "clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader",
                 "-lms", "20", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1].split()[0]))
                mx.append(float(r[2].split()[0]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class KernelTimer:
    """CUDA-event bracket around the dense pooling kernel (ops KERNEL_HOOK)."""

    def __init__(self):
        self.pairs = []
        self.enabled = False

    def before(self):
        if self.enabled:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self._s = e

    def after(self):
        if self.enabled:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.pairs.append((self._s, e))

    def mean_us(self):
        if not self.pairs:
            return None
        return 1e3 * sum(s.elapsed_time(e) for s, e in self.pairs) / len(
            self.pairs)


# ------------------------------------------------------------ CPU baseline --
class CpuReference:
    """The reference algorithm on the host cores (oracle port).

    F: voxel_pooling_prepare_v2 (C, single thread) + bev_pool_v2 as shipped
       (zero-fill + interval kernel + permute; C, OpenMP over intervals).
    B: BackwardProjection with the reference's re-batching cross-attention
       algorithm (spatial_cross_attention_depth.py:156-216) and MSDA through the
       C oracle; dense layers through torch CPU.
    Geometry through torch CPU ops, as the reference does."""

    def __init__(self, w_cpu):
        from oracle import backward_ref, cpu
        self.cpu = cpu
        self.backward_ref = backward_ref
        self.w = w_cpu
        # all host cores this process may use (torchrun exports
        # OMP_NUM_THREADS=1, which would make the baseline single-threaded)
        try:
            n_cores = len(os.sched_getaffinity(0))
        except AttributeError:
            n_cores = os.cpu_count() or 1
        self._n_cores = n_cores
        cpu.set_num_threads(n_cores)
        self.threads = cpu.num_threads()
        torch.set_num_threads(self.threads)
        gs = [int(v) for v in w_cpu.vt.grid_size]
        self.shape = (w_cpu.frames, gs[2], gs[1], gs[0], w_cpu.cfg["C"])
        n = int(np.prod(self.shape))
        self.scratch = np.empty(n, np.float32)
        self.out = np.empty((w_cpu.frames, w_cpu.cfg["C"], gs[2], gs[1],
                             gs[0]), np.float32)

    def calibrate(self):
        """Pick the thread count the CPU arm is fastest with: all logical
        cores is not it on a 2-way SMT host (measured: 3.5 s per step with 128
        threads against ~0.5 s with 64 on the same box)."""
        best = None
        for n in sorted({self._n_cores, max(1, self._n_cores // 2),
                         max(1, self._n_cores // 4)}, reverse=True):
            self.cpu.set_num_threads(n)
            torch.set_num_threads(n)
            self.step()
            t0 = time.perf_counter()
            self.step()
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, n)
        self.threads = best[1]
        self.cpu.set_num_threads(self.threads)
        torch.set_num_threads(self.threads)
        return self.threads

    @torch.no_grad()
    def step(self):
        w = self.w
        coor = w.vt.get_lidar_coor(*w.cam)
        rb, rd, rf, st, ln = self.cpu.voxel_prepare(
            coor.numpy(), w.vt.grid_lower_bound.numpy(),
            w.vt.grid_interval.numpy(), w.vt.grid_size.numpy())
        feat_nhwc = w.feat.permute(0, 1, 3, 4, 2).contiguous().numpy()
        self.cpu.bev_pool_v2(w.depth.numpy(), feat_nhwc, rd, rf, rb,
                             self.shape, st, ln, scratch=self.scratch,
                             out=self.out)
        ref = self.backward_ref.backward_projection_cpu(
            w.bp, [w.feat], w.lss, w.cam, w.depth)
        return self.out, ref


# --------------------------------------------------------------------- main --
def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import torch.distributed as dist
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend, rank=rank, world_size=world)
    return world, rank, local


def run_reference(args, world, rank):
    """--impl reference: CPU implementation of the same step, rank 0 only."""
    if rank != 0:
        return
    w = Workload("cpu", seed=0, frames=1)
    ref = CpuReference(w)
    ref.calibrate()
    for _ in range(args.warmup):
        ref.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.step()
    dt = (time.perf_counter() - t0) / args.steps
    value = w.voxels_per_frame / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": 1,
                   "note": "reference algorithm on host cores (CPU port: the "
                           "reference has no CPU kernel for this path)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ref.threads,
                         "cores_available": ref._n_cores, "kind": "port",
                         "sample": "full step (1 frame: geometry + prepare + "
                                   "bev_pool_v2 + BackwardProjection)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true",
                    help="time eager plugin calls instead of a CUDA graph "
                         "replay of them")
    ap.add_argument("--eager-only", action="store_true",
                    help="profiling aid: stop after the eager pass (what ncu "
                         "captures) and print only its numbers")
    args = ap.parse_args()
    if args.impl == "reference":
        # CPU arm: rank 0 alone works, the other ranks of a torchrun launch
        # exit at once; no process group is needed
        run_reference(args, int(os.environ.get("WORLD_SIZE", "1")),
                      int(os.environ.get("RANK", "0")))
        return
    world, rank, local = dist_setup(args.gpus)

    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device"
    import torch.distributed as dist
    from fbbev_b200 import _lib
    from fbbev_b200.ops import bev_pool_v2 as pool_ops
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    L = _lib.lib()  # fails loudly if the CUDA library is missing
    w = Workload(dev, seed=rank, frames=1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ktimer = KernelTimer()
    pool_ops.KERNEL_HOOK = ktimer

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warm = max(args.warmup, 3)
    for _ in range(warm):
        w.step()
    barrier()

    # ---- eager pass: K steps through the plugin calls, L2 flushed between
    # steps; also brackets the dense pooling kernel with CUDA events ---------
    launches0 = L.fbbev_debug_launch_count()
    ktimer.enabled = True
    events = []
    barrier()
    for _ in range(args.steps):
        flush.zero_()
        s = torch.cuda.Event(enable_timing=True)
        e = torch.cuda.Event(enable_timing=True)
        s.record()
        w.step()
        e.record()
        events.append((s, e))
    barrier()
    ktimer.enabled = False
    launches = (L.fbbev_debug_launch_count() - launches0) / args.steps
    eager_ms = sum(s.elapsed_time(e) for s, e in events) / args.steps
    pool_us = ktimer.mean_us()
    if args.eager_only:
        if rank == 0:
            print(json.dumps({"eager_ms_per_step": eager_ms,
                              "pool_us": pool_us, "gpu_launches": launches}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- headline pass: the same K steps as CUDA-graph replays --------------
    use_graph = not args.no_graph
    if use_graph:
        try:
            w.capture()
            for _ in range(warm):
                w.graph.replay()
        except Exception as ex:  # capture unsupported -> eager numbers stand
            use_graph = False
            print(f"[bench] CUDA graph capture failed: {ex}", file=sys.stderr)
    run_step = w.graph.replay if use_graph else w.step
    sampler = ClockSampler(local)
    events = []
    if rank == 0:
        sampler.start()
    barrier()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        s = torch.cuda.Event(enable_timing=True)
        e = torch.cuda.Event(enable_timing=True)
        s.record()
        run_step()
        e.record()
        events.append((s, e))
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = sum(s.elapsed_time(e) for s, e in events) / args.steps

    # ---- end to end: pinned host inputs -> plugin calls -> host results ----
    bev, ref = w.step()
    host_bev = torch.empty(bev.shape, dtype=bev.dtype).pin_memory()
    host_ref = torch.empty(ref.shape, dtype=ref.dtype).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step():
        if use_graph:
            w.to_device_inplace()                       # H2D of every input
            w.graph.replay()
            b, r = w.graph_out
        else:
            w.to_device()
            b, r = w.step()
        host_bev.copy_(b, non_blocking=True)            # D2H of both results
        host_ref.copy_(r, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    for _ in range(2):
        e2e_step()
    barrier()
    ev = []
    for _ in range(e2e_steps):
        flush.zero_()
        s = torch.cuda.Event(enable_timing=True)
        e = torch.cuda.Event(enable_timing=True)
        s.record()
        e2e_step()
        e.record()
        ev.append((s, e))
    barrier()
    e2e_latency_ms = sum(s.elapsed_time(e) for s, e in ev) / e2e_steps
    e2e_ms, e2e_mode = e2e_latency_ms, "one step at a time (copy in, compute, copy out)"
    d2h = int(host_bev.numel() * 4 + host_ref.numel() * 4)
    h2d = w.h2d_bytes()

    # the same K steps as a stream: three CUDA streams, step i's results leave
    # over PCIe while step i+1 computes and step i+2's inputs arrive.  Every
    # step still copies all of its inputs in and all of its results out.
    if use_graph:
        try:
            e2e_ms = e2e_streamed(w, e2e_steps, flush, barrier)
            e2e_mode = ("streamed: H2D / compute / D2H of consecutive steps "
                        "overlap on three streams (results double-buffered; "
                        "no L2 flush, 217.6 MB of results per step exceed L2)")
        except Exception as ex:
            print(f"[bench] streamed e2e failed: {ex}", file=sys.stderr)

    # ---- max over ranks ----------------------------------------------------
    t = torch.tensor([step_ms, e2e_ms, pool_us or 0.0, eager_ms,
                      e2e_latency_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, e2e_ms, pool_us, eager_ms, e2e_latency_ms = (
        float(v) for v in t.tolist())

    extra = {}
    if world > 1:
        # optional exchange, reported separately (not on the data path): gather
        # of the refined 2-D BEV of every rank's frame
        from fbbev_b200.sharding import gather_bev
        for _ in range(3):
            gather_bev(ref)
        barrier()
        s = torch.cuda.Event(enable_timing=True)
        e = torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            gather_bev(ref)
        e.record()
        barrier()
        tg = torch.tensor([s.elapsed_time(e) / 10], device=dev)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        extra["bev_gather"] = {"ms": float(tg), "bytes_per_rank":
                               int(ref.numel() * 4), "collective":
                               "nccl all_gather_into_tensor"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    frames = world * w.frames
    voxels = frames * w.voxels_per_frame
    value = voxels / (step_ms * 1e-3)

    # ---- roofline of the dominant kernel (dense pooling) --------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured" if "hbm_gbs" in peaks else "fallback"
    idx = w.vt.prepare_index(w.vt.get_lidar_coor(*w.cam))
    n_kept, n_int = (int(v) for v in idx.counts.tolist())
    alg_bytes = 4 * (w.depth.numel() + w.feat.numel() + 3 * n_kept +
                     2 * n_int + w.voxels_per_frame * w.cfg["C"] * w.frames)
    achieved = alg_bytes / (pool_us * 1e-6) / 1e9 if pool_us else None
    roofline = {"bound": "hbm", "kernel": "interval_sums_kernel + dense_write_kernel "
                "(dense lift-splat pooling, both launches of "
                "fbbev_bev_pool_v2_fwd_dense_planned)",
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if achieved else None,
                "peak_source": peak_src, "traffic": None,
                "algorithmic_bytes": int(alg_bytes), "kernel_us": pool_us,
                "timed": "CUDA events around the launch, inside the eager "
                         "pass of the same K steps (L2 flushed per step)"}
    prof = os.path.join(ROOT, "profiles", "pool_dense_traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof))["dram_bytes_per_launch"]
        except Exception:
            pass

    # ---- reference CUDA kernel on this GPU, for context ---------------------
    try:
        extra["reference_cuda"] = time_reference_cuda(w, idx, flush)
    except Exception as ex:  # oracle/_ref not built
        extra["reference_cuda"] = {"unavailable": str(ex)[:120]}

    # ---- CPU baseline beside it (rank 0, N = 1) ------------------------------
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        wc = Workload("cpu", seed=0, frames=1)
        cr = CpuReference(wc)
        cr.calibrate()
        n, t0 = 0, time.perf_counter()
        while n < 3 or (time.perf_counter() - t0 < 10.0 and n < 50):
            cr.step()
            n += 1
        dt = (time.perf_counter() - t0) / n
        cpu_baseline = {"value": wc.voxels_per_frame / dt, "unit": UNIT,
                        "cores": cr.threads, "cores_available": cr._n_cores,
                        "kind": "port",
                        "sample": f"{n} full steps of the same workload "
                                  f"({dt * 1e3:.1f} ms each)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_gpu": w.frames,
                   "step": "LSSViewTransformerFunction3D.forward + "
                           "BackwardProjection.forward",
                   "submission": "cuda graph replay of the two plugin calls"
                                 if use_graph else "eager plugin calls",
                   "eager_ms_per_step": eager_ms,
                   "l2": "flushed between timed steps (256 MiB memset)",
                   "parallelism": f"frames sharded over {world} rank(s), no "
                                  "data-path collective",
                   "wall_s_timed_region": wall},
        "clocks": clocks,
        "e2e": {"value": voxels / (e2e_ms * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms, "mode": e2e_mode,
                "latency_ms_one_step": e2e_latency_ms},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    line.update(extra)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def time_reference_cuda(w, idx, flush):
    """The reference's own bev_pool_cuda.cu (oracle/_ref), same inputs."""
    from oracle import ref_cuda
    if not ref_cuda.available():
        raise RuntimeError("oracle/_ref/libbev_pool_ref.so not built")
    rb, rd, rf, st, ln = idx.trimmed()
    feat = w.feat.permute(0, 1, 3, 4, 2).contiguous()
    shape = w.vt._bev_feat_shape(w.depth, feat)

    def wall(fn, iters=20):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        return statistics.median(ts) * 1e6
    zo = torch.zeros(shape, device=w.device)
    k_us = wall(lambda: ref_cuda.bev_pool_v2_kernel(w.depth, feat, rd, rf, rb,
                                                    st, ln, zo))
    op_us = wall(lambda: ref_cuda.bev_pool_v2(w.depth, feat, rd, rf, rb, shape,
                                              st, ln))
    return {"bev_pool_v2_kernel_us": k_us, "bev_pool_v2_op_as_shipped_us": op_us,
            "timing": "host wall clock around a synchronised launch (the "
                      "reference launches on the legacy default stream)"}


if __name__ == "__main__":
    main()
