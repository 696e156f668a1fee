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
