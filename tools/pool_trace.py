"""Timeline of interval_sums_kernel (K1 of the dense pooling) for two CTAs: the
first one and one of the last wave (clock64 of lane 0 / warp 0, relative to the
kernel entry of that CTA).

Needs a variant build with the hooks compiled in:

    make -C fb-bev_b200/csrc OBJDIR=../../build/var_PTRACE \
         OUT=../../build/var_PTRACE/libfbbev_b200.so EXTRA=-DPOOL_TRACE
    FBBEV_LIB=$PWD/build/var_PTRACE/libfbbev_b200.so python tools/pool_trace.py

Tags: 0 entry, 1 meta word arrived, 2 index words + slice table arrived,
3 depth gathered, 4 run starts known, 5.. feat batch k folded (two points per
lane each), 13 heads exchanged, 14 rows stored (issued).
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from fbbev_b200 import _lib, synthetic  # noqa: E402
from fbbev_b200.view_transformation.forward_projection import \
    LSSViewTransformerFunction3D  # noqa: E402

dev = "cuda:0"
vt = LSSViewTransformerFunction3D(synthetic.GRID_CONFIGS["fbocc_200"], (256, 704), 16)
cam = synthetic.make_cam_params(1, 6, (256, 704), device=dev)
depth, feat = synthetic.make_depth_feat(1, 6, vt.D, 16, 44, 80, device=dev)
L = _lib.lib()
fn = L.fbbev_debug_pool_trace
fn.restype = ctypes.c_int
buf = (ctypes.c_longlong * 32)()
with torch.no_grad():
    for _ in range(4):
        vt(cam, feat, depth)
    torch.cuda.synchronize()
assert fn(buf) == 0
for slot, name in ((0, "first CTA"), (1, "a CTA of the last wave")):
    t0 = buf[16 * slot]
    print(name)
    for tag in range(1, 16):
        t = buf[16 * slot + tag]
        if t:
            print(f"  tag {tag:2d}  {(t - t0) / 1.965e3:7.2f} us")
