"""Which fp32 evaluation order do torch's broadcast (..,3,3) @ (..,3,1) matmuls
use on this GPU?  All 18 ways to sum three products with / without FMA are
tried for each of the six per-point products of get_lidar_coor
(view_transformer.py:458-498) and point_sampling (bevformer_encoder.py:92-120),
at several problem sizes (cuBLAS may pick its kernel by size).
Development aid for the bit-exact fused geometry kernels."""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(
    os.path.abspath(__file__)))))
import torch

from fbbev_b200 import synthetic
from fbbev_b200.view_transformation.forward_projection import \
    LSSViewTransformerFunction3D

dev = "cuda:0"


def f32(x):
    return x.to(torch.float32)


def fma(a, b, c):
    return f32(a.double() * b.double() + c.double())


def variants(M, x):
    """M (..,3,3) broadcastable against x (..,3) -> {name: (..,3)}"""
    m = [[M[..., i, j] for j in range(3)] for i in range(3)]
    xs = [x[..., j] for j in range(3)]
    out = {}
    for last in range(3):
        a, b = [k for k in range(3) if k != last]
        for inner in ("pp", "fa", "fb"):      # p_a + p_b | fma(a, p_b) | fma(b, p_a)
            for outer in ("p", "f"):          # s + p_last | fma(last, s)
                rows = []
                for i in range(3):
                    pa, pb = m[i][a] * xs[a], m[i][b] * xs[b]
                    if inner == "pp":
                        s = pa + pb
                    elif inner == "fa":
                        s = fma(m[i][a], xs[a], pb)
                    else:
                        s = fma(m[i][b], xs[b], pa)
                    if outer == "p":
                        r = s + m[i][last] * xs[last]
                    else:
                        r = fma(m[i][last], xs[last], s)
                    rows.append(r)
                out[f"last{last}_{inner}_{outer}"] = torch.stack(rows, -1)
    return out


def report(tag, M, x, y):
    res = []
    for k, v in variants(M, x).items():
        res.append(((v == y).all(-1).float().mean().item(), k))
    res.sort(reverse=True)
    print(f"  {tag}: " + ", ".join(f"{k} {r * 100:.4f}%" for r, k in res[:4]))


def run(B, grid, input_size, ds, bev):
    print(f"== B={B} grid={grid} input={input_size} ds={ds} bev={bev}")
    vt = LSSViewTransformerFunction3D(synthetic.GRID_CONFIGS[grid], input_size,
                                      ds)
    cam = synthetic.make_cam_params(B, 6, input_size, device=dev, jitter=1.0)
    rots, trans, intr, post_rots, post_trans, bda = cam
    g = torch.Generator().manual_seed(3)
    bda = (torch.eye(3) + 0.05 * torch.randn(B, 3, 3, generator=g)).to(dev)
    N = 6
    # ---- get_lidar_coor
    pts = vt.frustum.to(dev) - post_trans.view(B, N, 1, 1, 1, 3)
    A = torch.inverse(post_rots).view(B, N, 1, 1, 1, 3, 3)
    y = A.matmul(pts.unsqueeze(-1)).squeeze(-1)
    report("F1 inv(post_rots)@pts", A, pts, y)
    p2 = torch.cat((y[..., :2] * y[..., 2:3], y[..., 2:3]), -1)
    C = rots.matmul(torch.inverse(intr)).view(B, N, 1, 1, 1, 3, 3)
    y2 = C.matmul(p2.unsqueeze(-1)).squeeze(-1)
    report("F2 cam2ego@pts       ", C, p2, y2)
    y3 = y2 + trans.view(B, N, 1, 1, 1, 3)
    Bd = bda.view(B, 1, 1, 1, 1, 3, 3)
    y4 = Bd.matmul(y3.unsqueeze(-1)).squeeze(-1)
    report("F3 bda@pts           ", Bd, y3, y4)
    # ---- point_sampling
    h, w = bev
    xs = torch.arange(-40, 40, 80.0 / w) + 40.0 / w
    ys = torch.arange(-40, 40, 80.0 / h) + 40.0 / h
    zs = torch.arange(-1, 5.4, 1.6) + 0.8
    Yg, Xg, Zg = torch.meshgrid([ys, xs, zs], indexing='ij')
    ref = torch.stack([Xg, Yg, Zg], -1).to(dev)
    q = ref[None, None].repeat(B, N, 1, 1, 1, 1)
    Ib = torch.inverse(bda).view(B, 1, 1, 1, 1, 3, 3)
    z1 = Ib.matmul(q.unsqueeze(-1)).squeeze(-1)
    report("B1 inv(bda)@ref      ", Ib, q, z1)
    z1 = z1 - trans.view(B, N, 1, 1, 1, 3)
    E2 = rots.matmul(torch.inverse(intr)).inverse().view(B, N, 1, 1, 1, 3, 3)
    z2 = E2.matmul(z1.unsqueeze(-1)).squeeze(-1)
    report("B2 ego2cam@pts       ", E2, z1, z2)
    zz = z2[..., 2:3]
    z3 = torch.cat([z2[..., 0:2] / torch.maximum(zz, torch.ones_like(zz) * 1e-5),
                    zz], 5)
    Pr = post_rots.view(B, N, 1, 1, 1, 3, 3)
    z4 = Pr.matmul(z3.unsqueeze(-1)).squeeze(-1)
    report("B3 post_rots@cam     ", Pr, z3, z4)
    # ---- the small 3x3 @ 3x3 products
    Ci = torch.inverse(intr)
    y5 = rots.matmul(Ci)
    res = {}
    for j in range(3):
        for k, v in variants(rots, Ci[..., :, j]).items():
            res.setdefault(k, []).append(v)
    best = sorted(((torch.stack(v, -1) == y5).float().mean().item(), k)
                  for k, v in res.items())[::-1][:3]
    print("  S  rots@inv(K)       : " + ", ".join(
        f"{k} {r * 100:.3f}%" for r, k in best))


if __name__ == "__main__":
    torch.manual_seed(0)
    run(1, "fbocc_200", (256, 704), 16, (100, 100))
    run(2, "fbocc_200", (256, 704), 16, (200, 200))
    run(16, "fbocc_200", (256, 704), 16, (100, 100))
    run(1, "fbocc_400", (512, 1408), 16, (200, 200))
    run(2, "fbocc_shipped", (128, 352), 16, (12, 12))
