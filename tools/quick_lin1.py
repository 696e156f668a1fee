import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fbbev_b200.ops.linear import linear_fused
dev = "cuda"; m, k, n = 40000, 80, 80
x = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev) / k ** .5; b = torch.randn(n, device=dev)
r = torch.randn(m, n, device=dev); g = torch.ones(n, device=dev); be = torch.zeros(n, device=dev)
with torch.no_grad():
    for _ in range(6):
        y = linear_fused(x, w, b, residual=r, ln_weight=g, ln_bias=be)
torch.cuda.synchronize()
