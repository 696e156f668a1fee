"""Timeline of CTA 0 of linear_tf32_kernel (clock64 per warp role).

Needs a variant build with the trace hooks compiled in:

    make -C fb-bev_b200/csrc OBJDIR=../../build/var_TRACE \
         OUT=../../build/var_TRACE/libfbbev_b200.so EXTRA=-DLIN_TRACE
    FBBEV_LIB=$PWD/build/var_TRACE/libfbbev_b200.so python tools/lin_trace.py

Tags: 1 set-up done, 100+w loads of round w issued, 200+i / 300+i stage of item
i free / filled, 400+i MMA sees item i, 500+t accumulator of tile t committed,
600..1000+t epilogue of tile t (wait, accumulator ready, passes done,
normalised, stored), 2 kernel end.
"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fbbev_b200 import _lib
from fbbev_b200.ops.linear import linear_fused
dev = "cuda"; m, k, n = 40000, 80, 80
x = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev) / k ** .5; b = torch.randn(n, device=dev)
r = torch.randn(m, n, device=dev); g = torch.ones(n, device=dev); be = torch.zeros(n, device=dev)
L = _lib.lib()
buf = (ctypes.c_longlong * 512)()
fn = L.fbbev_debug_linear_trace; fn.restype = ctypes.c_int
with torch.no_grad():
    for i in range(4):
        y = linear_fused(x, w, b, residual=r, ln_weight=g, ln_bias=be)
        torch.cuda.synchronize()
        nrec = fn(buf, 1)
ev = sorted((buf[2*i+1], buf[2*i]) for i in range(min(nrec, 256)))
t0 = ev[0][0]
for t, tag in ev:
    print(f"{(t - t0)/1.965e3:8.2f} us  tag {tag}")
