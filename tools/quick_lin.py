import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from fbbev_b200.ops.linear import linear_fused
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
def bench(fn, iters=50):
    for _ in range(5): fn()
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize(); s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / iters
for (m, k, n, relu, res, ln) in [(40000, 80, 80, False, True, True), (40000, 80, 320, True, False, False), (40000, 320, 80, False, True, True), (40000, 80, 64, False, False, False), (40000, 80, 128, False, False, False)]:
    x = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev) / k ** .5; b = torch.randn(n, device=dev)
    r = torch.randn(m, n, device=dev) if res else None
    g = torch.ones(n, device=dev); be = torch.zeros(n, device=dev)
    def ours():
        with torch.no_grad():
            return linear_fused(x, w, b, relu=relu, residual=r, ln_weight=g if ln else None, ln_bias=be if ln else None)
    def ref():
        with torch.no_grad():
            y = F.linear(x, w, b)
            if relu: y = y.relu()
            if res: y = y + r
            if ln: y = F.layer_norm(y, (n,), g, be)
            return y
    err = (ours().double() - ref().double()).abs().max().item()
    print(f"m={m} k={k} n={n} relu={relu} res={res} ln={ln}: ours {bench(ours):.1f} us  torch {bench(ref):.1f} us  maxdiff {err:.2e}")
