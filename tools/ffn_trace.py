"""Timeline of CTA 0 of ffn_tf32_kernel (clock64 per warp role).

    make -C fb-bev_b200/csrc OBJDIR=../../build/var_FTRACE \\
         OUT=../../build/var_FTRACE/libfbbev_b200.so EXTRA=-DFFN_TRACE
    FBBEV_LIB=$PWD/build/var_FTRACE/libfbbev_b200.so python tools/ffn_trace.py

Tags: 1 set-up done; 100/110/120+t loader of tile t (loads issued, X buffer
free, filled); 1000+i producer issues weight stage i; 200/210/220+t MMA warp
(tile start, X ready, all issued); 2000+i MMA sees weight stage i, 3000+i its
hidden K-block too; 4000/5000/6000+u convert of K-block u (chunk ready, lo stage
free, done); 300/310/320+t finish (slab free, Y ready, stored); 330 all stored.
"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fbbev_b200 import _lib
from fbbev_b200.ops.linear import ffn_fused
dev = "cuda"; m, e, h = 40000, 80, 320
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(m, e, device=dev, generator=g)
w1 = torch.randn(h, e, device=dev, generator=g) / e ** .5
w2 = torch.randn(e, h, device=dev, generator=g) / h ** .5
b1 = torch.randn(h, device=dev, generator=g); b2 = torch.randn(e, device=dev, generator=g)
gm = torch.ones(e, device=dev); bt = torch.zeros(e, device=dev)
L = _lib.lib()
CAP = 200
buf = (ctypes.c_longlong * (5 * 2 * CAP))()
cnt = (ctypes.c_int * 5)()
fn = L.fbbev_debug_ffn_trace; fn.restype = ctypes.c_int
with torch.no_grad():
    for i in range(4):
        y = ffn_fused(x, w1, b1, w2, b2, residual=x, ln_weight=gm, ln_bias=bt)
        torch.cuda.synchronize()
    fn(buf, cnt)
ev = sorted((buf[l * 2 * CAP + 2 * i + 1], buf[l * 2 * CAP + 2 * i])
            for l in range(5) for i in range(min(cnt[l], CAP)))
t0 = ev[0][0]
for t, tag in ev:
    print(f"{(t - t0)/1.965e3:8.2f} us  tag {tag}")
