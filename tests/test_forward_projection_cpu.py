"""CPU tests of the host-side mirror of LSSViewTransformerFunction3D: everything
that is plain PyTorch (grid infos, frustum, geometry) must equal the reference's
own Python bit for bit (golden fixtures made by tests/golden/gen_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import load_golden

CASES = ["f_small_6cam", "f_unit_1cam"]


def build(g, **kw):
    from fbbev_b200.registry import build_neck
    cfg = dict(
        type='LSSViewTransformerFunction3D',
        grid_config=dict(x=list(g["grid_x"]), y=list(g["grid_y"]),
                         z=list(g["grid_z"]), depth=list(g["grid_depth"])),
        input_size=tuple(int(v) for v in g["input_size"]),
        downsample=int(g["downsample"]), **kw)
    return build_neck(cfg)


@pytest.mark.parametrize("case", CASES)
def test_grid_infos_frustum_and_buffers(case):
    g = load_golden(case)
    vt = build(g)
    for name in ("frustum", "grid_size", "grid_interval", "grid_lower_bound",
                 "dx", "bx", "nx"):
        np.testing.assert_array_equal(getattr(vt, name).detach().numpy(),
                                      g[name], err_msg=name)
    assert vt.D == g["frustum"].shape[0]
    assert set(vt.state_dict()) == {"dx", "bx", "nx"}


@pytest.mark.parametrize("case", CASES)
def test_get_lidar_coor_bit_exact(case):
    g = load_golden(case)
    vt = build(g)
    cam = [torch.from_numpy(g[k]) for k in
           ("rots", "trans", "intrins", "post_rots", "post_trans", "bda")]
    coor = vt.get_lidar_coor(*cam)
    np.testing.assert_array_equal(coor.numpy(), g["coor"])


def test_grid_size_float32_quirk():
    """(5.4 - -1) / 0.8 = 8.000000000000002 in double -> 8.0f
    (view_transformer.py:384-387)."""
    from fbbev_b200.view_transformation.forward_projection import \
        LSSViewTransformerFunction3D
    from fbbev_b200.synthetic import GRID_CONFIGS
    vt = LSSViewTransformerFunction3D(GRID_CONFIGS["fbocc_shipped"],
                                      (256, 704), 16)
    assert vt.grid_size.tolist() == [100.0, 100.0, 8.0]
    assert vt.D == 80 and tuple(vt.frustum.shape) == (80, 16, 44, 3)
    vt = LSSViewTransformerFunction3D(GRID_CONFIGS["fbocc_200"], (256, 704),
                                      16)
    assert [int(v) for v in vt.grid_size] == [200, 200, 16]


def test_inv3x3_many_equals_per_matrix_inverse():
    """One batched inverse call for several stacks == torch.inverse on each
    (the fused-geometry host code shares the call)."""
    from fbbev_b200.view_transformation.forward_projection import inv3x3_many
    g = torch.Generator().manual_seed(0)
    a = torch.randn(2, 6, 3, 3, generator=g) + 3 * torch.eye(3)
    b = torch.randn(2, 3, 3, generator=g) + 3 * torch.eye(3)
    c = torch.randn(5, 3, 3, generator=g) + 3 * torch.eye(3)
    ia, ib, ic = inv3x3_many(a, b, c)
    assert ia.shape == a.shape and ib.shape == b.shape and ic.shape == c.shape
    for got, m in ((ia, a), (ib, b), (ic, c)):
        assert torch.equal(got, torch.inverse(m))
    only, = inv3x3_many(a)
    assert torch.equal(only, torch.inverse(a))
