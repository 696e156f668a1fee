"""Quick device timing of the forward pooling path (development aid)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fbbev_b200 import synthetic, _lib
from fbbev_b200.ops import bev_pool_v2 as ops
from fbbev_b200.view_transformation.forward_projection import LSSViewTransformerFunction3D

dev = "cuda:0"
name = sys.argv[1] if len(sys.argv) > 1 else "fbocc_200"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = {"fbocc_200": ("fbocc_200", (256, 704), 16, 6), "shipped": ("fbocc_shipped", (256, 704), 16, 6),
       "unit_128": ("unit_128", (256, 704), 4, 1), "fbocc_400": ("fbocc_400", (512, 1408), 16, 6)}[name]
vt = LSSViewTransformerFunction3D(synthetic.GRID_CONFIGS[cfg[0]], cfg[1], cfg[2])
cam = synthetic.make_cam_params(B, cfg[3], cfg[1], device=dev, jitter=1.0 if B > 1 else 0)
H, W = cfg[1][0] // cfg[2], cfg[1][1] // cfg[2]
depth, feat = synthetic.make_depth_feat(B, cfg[3], vt.D, H, W, 80, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, iters=30, warm=5):
    for _ in range(warm): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]

coor = vt.get_lidar_coor(*cam)
idx = vt.prepare_index(coor)
rb, rd, rf, st, ln = idx.trimmed()
feat_nhwc = feat.permute(0, 1, 3, 4, 2).contiguous()
shape = vt._bev_feat_shape(depth, feat_nhwc)
nvox = shape[0] * shape[1] * shape[2] * shape[3]
print(f"{name} B={B}: n_pts={coor.numel()//3} kept={len(rb)} intervals={len(st)} voxels={nvox}")
res = {}
res["geometry(torch)"] = timeit(lambda: vt.get_lidar_coor(*cam))
res["prepare"] = timeit(lambda: vt.prepare_index(coor))
res["pool_dense(plan+kernel)"] = timeit(lambda: ops.bev_pool_v2(depth, feat_nhwc, rd, rf, rb, shape, st, ln))
L = _lib.lib()
ws_bytes = L.fbbev_bev_pool_v2_dense_workspace_bytes(shape[0], shape[1]*shape[2]*shape[3], len(st), len(rb), 80)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
out = torch.empty((shape[0], 80, shape[1], shape[2], shape[3]), device=dev)
sp = _lib.stream_ptr(torch.device(dev))
L.fbbev_bev_pool_v2_plan(_lib.ptr(rb), _lib.ptr(st), _lib.ptr(ln), len(st), None, len(rb), 80, shape[0], shape[1]*shape[2]*shape[3], _lib.ptr(ws), ws_bytes, sp)
def planned():
    L.fbbev_bev_pool_v2_fwd_dense_planned(_lib.ptr(depth), _lib.ptr(feat_nhwc), _lib.ptr(rd), _lib.ptr(rf), _lib.ptr(rb), _lib.ptr(st), _lib.ptr(ln), len(st), len(rb), 80, shape[0], shape[1]*shape[2]*shape[3], _lib.ptr(out), _lib.ptr(ws), ws_bytes, sp)
res["pool_dense(kernel only)"] = timeit(planned)
res["pool_interval(ref layout, +zeros)"] = timeit(lambda: ops.QuickCumsumCuda.apply(depth, feat_nhwc, rd, rf, rb, shape, st, ln))
res["memset_only(out.zero_)"] = timeit(lambda: out.zero_())
res["plugin forward (geom+prep+pool)"] = timeit(lambda: vt(cam, feat, depth))
try:
    from oracle import ref_cuda
    if ref_cuda.available():
        zo = torch.zeros(shape, device=dev)
        def refk():
            ref_cuda.bev_pool_v2_kernel(depth, feat_nhwc, rd, rf, rb, st, ln, zo)
        # reference launches on the legacy default stream: time with sync wall clock
        import time
        def wall(fn, iters=30):
            for _ in range(5): fn()
            torch.cuda.synchronize(); ts=[]
            for _ in range(iters):
                flush.zero_(); torch.cuda.synchronize(); t=time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter()-t)*1e6)
            ts.sort(); return ts[len(ts)//2], ts[0]
        res["REF kernel only (wall, sync)"] = wall(refk)
        res["REF op as shipped zeros+kernel+permute (wall, sync)"] = wall(lambda: ref_cuda.bev_pool_v2(depth, feat_nhwc, rd, rf, rb, shape, st, ln))
        res["OURS kernel only (wall, sync)"] = wall(planned)
except Exception as ex:
    print("ref unavailable", ex)
alg = 4 * (depth.numel() + feat.numel() + 3 * len(rb) + 2 * len(st) + out.numel())
for k, (med, best) in res.items():
    print(f"{k:55s} median {med:9.1f} us   best {best:9.1f} us")
k = res["pool_dense(kernel only)"][0]
print(f"algorithmic bytes {alg/1e6:.1f} MB -> {alg/k/1e3:.1f} GB/s ; {nvox/k/1e3:.2f} Gvoxel/s")
