"""GPU parity of the history warp kernel (fbbev_history_warp) and of
TemporalFusion.fuse_history against the reference's own fuse_history
(tests/golden/t_fuse_history.npz) -- tolerance 1e-4 -- plus size-independent
properties at the full FB-OCC size (16 x 80 channels of 8 x 100 x 100)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from test_temporal_cpu import build_fusion, run_steps

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ATOL = 1e-4


def test_fuse_history_vs_reference_golden():
    g = load_golden("t_fuse_history")
    tf = build_fusion(g, DEV)
    for step, (out, hist, sweep) in enumerate(run_steps(tf, g, DEV)):
        np.testing.assert_allclose(out.cpu().numpy(), g[f"out{step}"], rtol=0,
                                   atol=ATOL)
        np.testing.assert_allclose(hist.cpu().numpy(), g[f"history{step}"],
                                   rtol=0, atol=ATOL)
        np.testing.assert_array_equal(sweep.cpu().numpy(), g[f"sweep{step}"])


@pytest.mark.parametrize("n,mc,zhw", [(2, 5, (3, 7, 6)), (1, 33, (8, 20, 17)),
                                      (3, 16, (1, 9, 9))])
def test_history_warp_vs_oracle(n, mc, zhw):
    from fbbev_b200.view_transformation.temporal_fusion import history_warp
    from oracle.history_ref import history_warp_cpu
    g = torch.Generator().manual_seed(n * 100 + mc)
    hist = torch.randn(n, mc, *zhw, generator=g)
    flow = torch.eye(4).repeat(n, 1, 1)
    ang = (torch.rand(n, generator=g) - 0.5) * 0.6
    flow[:, 0, 0], flow[:, 0, 1] = ang.cos(), -ang.sin()
    flow[:, 1, 0], flow[:, 1, 1] = ang.sin(), ang.cos()
    flow[:, :3, 3] = (torch.rand(n, 3, generator=g) - 0.5) * 4
    want = history_warp_cpu(hist, flow, torch.zeros(n, mc + 3, *zhw), 3)
    # history as a channel slice of a larger buffer (batch stride > mc*Z*H*W)
    big = torch.randn(n, mc + 2, *zhw, generator=g).to(DEV)
    big[:, :mc] = hist.to(DEV)
    out = torch.zeros(n, mc + 3, *zhw, device=DEV)
    history_warp(big[:, :mc], flow.to(DEV), out, 3)
    np.testing.assert_allclose(out.cpu().numpy(), want.numpy(), rtol=0,
                               atol=2e-5)
    assert float(out[:, :3].abs().max()) == 0.0      # other channels untouched


def test_history_warp_full_size_properties():
    """FB-OCC size: identity flow reproduces the history and an integer shift
    is a shifted copy with zero padding -- up to the rounding of the
    normalise / un-normalise round trip the reference's grid_sample has as well
    ((x / (W-1) * 2 - 1 + 1) / 2 * (W-1) is x only to ~1e-5) --; the
    current-frame slot of the concatenation buffer is left alone."""
    from fbbev_b200.view_transformation.temporal_fusion import history_warp
    n, T, C, Z, H, W = 1, 16, 80, 8, 100, 100
    g = torch.Generator(device=DEV).manual_seed(3)
    hist = torch.randn(n, T * C, Z, H, W, device=DEV, generator=g)
    out = torch.full((n, (T + 1) * C, Z, H, W), 7.0, device=DEV)
    eye = torch.eye(4, device=DEV)[None]
    history_warp(hist, eye, out, C)
    assert float((out[:, C:] - hist).abs().max()) <= ATOL
    assert float((out[:, :C] - 7.0).abs().max()) == 0.0
    shift = eye.clone()
    shift[0, 0, 3], shift[0, 1, 3], shift[0, 2, 3] = 3.0, -2.0, 1.0
    history_warp(hist, shift, out, C)
    want = torch.zeros_like(hist)
    # out[z, y, x] = hist[z + 1, y - 2, x + 3]
    want[:, :, :Z - 1, 2:, :W - 3] = hist[:, :, 1:, :H - 2, 3:]
    assert float((out[:, C:] - want).abs().max()) <= ATOL
