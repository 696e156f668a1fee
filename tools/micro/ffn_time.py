"""Time fbbev_ffn_fwd against the three-launch route (CUDA events, L2 flushed)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from fbbev_b200.ops.linear import ffn_fused, linear_fused

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
m, e, h = 40000, 80, 320
x = torch.randn(m, e, device=dev, generator=g)
w1 = torch.randn(h, e, device=dev, generator=g) / e ** 0.5
w2 = torch.randn(e, h, device=dev, generator=g) / h ** 0.5
b1 = torch.randn(h, device=dev, generator=g)
b2 = torch.randn(e, device=dev, generator=g)
gm, bt = torch.rand(e, device=dev, generator=g) + 0.5, torch.randn(e, device=dev, generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def three():
    hid = linear_fused(x, w1, b1, relu=True)
    return linear_fused(hid, w2, b2, residual=x, ln_weight=gm, ln_bias=bt)


def one():
    return ffn_fused(x, w1, b1, w2, b2, residual=x, ln_weight=gm, ln_bias=bt)


def timeit(fn, n=20):
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


with torch.no_grad():
    y1, y3 = one(), three()
    torch.cuda.synchronize()
    print("max diff one vs three:", (y1 - y3).abs().max().item())
    for name, fn in (("three launches", three), ("fused ffn", one)):
        for _ in range(3):
            fn()
        print(name, "median / min us:", timeit(fn))
