"""Which fp32 summation order does torch's broadcast 3x3 @ 3x1 matmul use on
this GPU?  (development aid for a bit-exact fused geometry kernel)"""
import itertools, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from fbbev_b200 import synthetic
from fbbev_b200.view_transformation.forward_projection import LSSViewTransformerFunction3D
dev = "cuda:0"
torch.manual_seed(0)

def f32(x): return x.to(torch.float32)
def fma(a, b, c):  # emulated fused multiply-add (single rounding) via float64
    return f32(a.double() * b.double() + c.double())
def mul(a, b): return a * b
def add(a, b): return a + b

def variants(M, x):
    """M (..,3,3) broadcastable, x (..,3) -> dict name -> (..,3)"""
    m = [[M[..., i, j] for j in range(3)] for i in range(3)]
    xs = [x[..., j] for j in range(3)]
    out = {}
    for perm in itertools.permutations(range(3)):
        a, b, c = perm
        out[f"fma_seq_{a}{b}{c}"] = torch.stack(
            [fma(m[i][c], xs[c], fma(m[i][b], xs[b], mul(m[i][a], xs[a]))) for i in range(3)], -1)
        out[f"nofma_seq_{a}{b}{c}"] = torch.stack(
            [add(add(mul(m[i][a], xs[a]), mul(m[i][b], xs[b])), mul(m[i][c], xs[c])) for i in range(3)], -1)
        out[f"fma_last_{a}{b}{c}"] = torch.stack(
            [fma(m[i][c], xs[c], add(mul(m[i][a], xs[a]), mul(m[i][b], xs[b]))) for i in range(3)], -1)
    return out

vt = LSSViewTransformerFunction3D(synthetic.GRID_CONFIGS["fbocc_200"], (256, 704), 16)
cam = synthetic.make_cam_params(2, 6, (256, 704), device=dev, jitter=1.0)
rots, trans, intr, post_rots, post_trans, bda = cam
B, N = 2, 6
pts = vt.frustum.to(dev) - post_trans.view(B, N, 1, 1, 1, 3)
A = torch.inverse(post_rots).view(B, N, 1, 1, 1, 3, 3)
y = A.matmul(pts.unsqueeze(-1)).squeeze(-1)
print("matmul #1 (inv(post_rots) @ pts):")
for k, v in variants(A, pts).items():
    eq = (v == y).all(-1).float().mean().item()
    if eq > 0.9: print(f"   {k:18s} match {eq*100:.4f}%")
p2 = torch.cat((y[..., :2] * y[..., 2:3], y[..., 2:3]), -1)
C = rots.matmul(torch.inverse(intr)).view(B, N, 1, 1, 1, 3, 3)
y2 = C.matmul(p2.unsqueeze(-1)).squeeze(-1)
print("matmul #2 (combine @ pts):")
for k, v in variants(C, p2).items():
    eq = (v == y2).all(-1).float().mean().item()
    if eq > 0.9: print(f"   {k:18s} match {eq*100:.4f}%")
y3 = y2 + trans.view(B, N, 1, 1, 1, 3)
Bd = torch.eye(3, device=dev) + 0.01 * torch.randn(B, 3, 3, device=dev)
y4 = Bd.view(B, 1, 1, 1, 1, 3, 3).matmul(y3.unsqueeze(-1)).squeeze(-1)
print("matmul #3 (bda @ pts):")
for k, v in variants(Bd.view(B, 1, 1, 1, 1, 3, 3), y3).items():
    eq = (v == y4).all(-1).float().mean().item()
    if eq > 0.9: print(f"   {k:18s} match {eq*100:.4f}%")
# small matrices: rots @ inv(intrins)
Ci = torch.inverse(intr)
y5 = rots.matmul(Ci)
print("3x3 @ 3x3 (rots @ inv(K)): checking column-wise variants")
for k in ["fma_seq_012", "fma_seq_210", "nofma_seq_012"]:
    cols = [variants(rots, Ci[..., :, j])[k] for j in range(3)]
    v = torch.stack(cols, -1)
    print(f"   {k:18s} match {(v == y5).float().mean().item()*100:.3f}%")
