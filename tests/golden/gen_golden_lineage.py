"""Golden fixtures for the callers either side of the pooling op (SURVEY.md
section 8 f3 / f4), produced by the reference's OWN classes executed on CPU:

l_lss_*      NECKS.LSSViewTransformer / LSSViewTransformer2 /
             LSSViewTransformerBEVDepth (mmdet3d/models/necks/view_transformer.py
             :15-326, :331-724, :1000-1105), accelerate off and on, including the
             depth-threshold sparsification (:556-557, :657-678)
l_cm_tail    the tail of CM_DepthNet.forward (mmdet3d/models/fbbev/modules/
             depth_net.py:346-363): inputs of context_conv / output of
             depth_conv captured with hooks, outputs as returned

    GOLDEN_ONLY=lineage python tests/golden/gen_golden.py

The reference files are loaded where they lie (ref_import.py); nothing is
copied.  mmdet's BasicBlock (third-party, absent) only appears inside the body
of the depth nets, which is NOT what these fixtures pin: a plain two-conv
residual block stands in for it, and the fixtures record the body's OUTPUT.
"""
import ast
import os
import sys
import types
import warnings

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

import ref_import
from fbbev_b200 import synthetic


def _np(t):
    return t.detach().cpu().numpy()


class _BasicBlock(nn.Module):
    """Stand-in for mmdet.models.backbones.resnet.BasicBlock (see module doc)."""

    def __init__(self, inplanes, planes, stride=1, downsample=None, **kw):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)

    def forward(self, x):
        out = F.relu(self.bn1(self.conv1(x)))
        return F.relu(self.bn2(self.conv2(out)) + x)


def load_necks():
    ref_import.install_stubs()
    sys.modules['mmdet.models.backbones.resnet'].BasicBlock = _BasicBlock
    ref_import._mod('ref_models')
    sys.modules['ref_models'].__path__ = []
    ref_import._mod('ref_models.builder', NECKS=ref_import.REGS['NECKS'])
    mods = ref_import._load_pkg(
        'ref_models.necks',
        os.path.join(ref_import.REF, 'mmdet3d', 'models', 'necks'),
        ['view_transformer'])
    return mods['view_transformer']


def load_cm_depth_net():
    """Mlp / SELayer / CM_DepthNet compiled out of depth_net.py's AST (the
    module imports cv2, matplotlib, torchvision ... at the top)."""
    path = os.path.join(ref_import.REF, 'mmdet3d', 'models', 'fbbev',
                        'modules', 'depth_net.py')
    tree = ast.parse(open(path).read())
    want = {'Mlp', 'SELayer', 'CM_DepthNet'}
    nodes = [n for n in tree.body
             if isinstance(n, ast.ClassDef) and n.name in want]
    assert {n.name for n in nodes} == want
    for n in nodes:
        n.decorator_list = []
    ident = ref_import._identity_decorator_factory
    ns = {'torch': torch, 'nn': nn, 'F': F, 'force_fp32': ident,
          'auto_fp16': ident, 'BaseModule': ref_import.BaseModule,
          'BasicBlock': _BasicBlock, 'cp': types.SimpleNamespace()}
    exec(compile(ast.Module(body=nodes, type_ignores=[]), path, 'exec'), ns)
    return ns


def _randomise_bn(module, gen):
    for m in module.modules():
        if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
            m.running_mean.copy_(torch.rand(m.running_mean.shape,
                                            generator=gen) * 0.4 - 0.2)
            m.running_var.copy_(torch.rand(m.running_var.shape,
                                           generator=gen) + 0.5)


def gen_lineage(save):
    vt = load_necks()
    warnings.filterwarnings('ignore', message='torch.range is deprecated')
    grid = dict(x=[-40, 40, 4.0], y=[-40, 40, 4.0], z=[-1, 5.4, 3.2],
                depth=[2.0, 42.0, 4.0])
    input_size, ds = (64, 176), 16
    H, W = input_size[0] // ds, input_size[1] // ds
    B, N, Cin, Cout = 2, 6, 16, 8

    def run(name, cls_name, seed, depthnet=False):
        cls = getattr(vt, cls_name)
        cam = synthetic.make_cam_params(B, N, input_size=input_size,
                                        jitter=1.0, seed=seed)
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(B, N, Cin, H, W, generator=g)
        x2 = torch.randn(B, N, Cin, H, W, generator=g)
        arrays = {}
        kw = dict(grid_config=grid, input_size=input_size, downsample=ds,
                  in_channels=Cin, out_channels=Cout)
        torch.manual_seed(seed)
        mods = {}
        for acc in (False, True):
            if depthnet:
                m = cls(depthnet_cfg=dict(use_dcn=False, use_aspp=False),
                        accelerate=acc, **kw)
            else:
                m = cls(accelerate=acc, **kw)
            mods[acc] = m
        # same weights in both; peaky depth logits so that the 0.01 threshold
        # actually removes points (D = 10 bins)
        ref_sd = mods[False].state_dict()
        if not depthnet:
            ref_sd['depth_net.weight'] = ref_sd['depth_net.weight'] * 6.0
        mods[False].load_state_dict(ref_sd)
        mods[True].load_state_dict(ref_sd)
        for m in mods.values():
            _randomise_bn(m, torch.Generator().manual_seed(seed + 1))
            m.eval()
        mlp = None
        with torch.no_grad():
            for acc, m in mods.items():
                inp = [x] + list(cam)
                inp2 = [x2] + list(cam)
                if depthnet:
                    mlp = m.get_mlp_input(*cam)
                    inp.append(mlp)
                    inp2.append(mlp)
                    if not acc:
                        B_, N_ = x.shape[:2]
                        scale = 6.0
                        # record the body's output (see module doc): D logits
                        # scaled so the threshold bites
                        orig = m.depth_net.forward

                        def scaled(xx, mm, _o=orig):
                            y = _o(xx, mm)
                            return torch.cat([y[:, :m.D] * scale,
                                              y[:, m.D:]], 1)
                        m.depth_net.forward = scaled
                        arrays['net_out'] = _np(scaled(
                            x.view(B_ * N_, Cin, H, W), mlp))
                        arrays['net_out2'] = _np(scaled(
                            x2.view(B_ * N_, Cin, H, W), mlp))
                    else:
                        m.depth_net.forward = mods[False].depth_net.forward
                bev, depth, digit = m(inp, return_depth_digit=True)
                tag = 'acc' if acc else 'plain'
                arrays[f'bev_{tag}'] = _np(bev)
                arrays[f'depth_{tag}'] = _np(depth)
                arrays[f'digit_{tag}'] = _np(digit)
                if acc:   # second call on the cached index, other features
                    bev2, depth2 = m(inp2)
                    arrays['bev_acc_second'] = _np(bev2)
                    arrays['n_kept_geom'] = np.array(len(m.ranks_bev))
                    if hasattr(m, 'kept'):
                        arrays['kept_mask'] = _np(m.kept)
        d = arrays['depth_plain']
        arrays['frac_below_thresh'] = np.array(float((d <= 0.01).mean()))
        if mlp is not None:
            arrays['mlp_input'] = _np(mlp)
        if not depthnet:
            arrays['depth_net.weight'] = _np(ref_sd['depth_net.weight'])
            arrays['depth_net.bias'] = _np(ref_sd['depth_net.bias'])
        save(name, grid_x=np.array(grid['x'], np.float64),
             grid_y=np.array(grid['y'], np.float64),
             grid_z=np.array(grid['z'], np.float64),
             grid_depth=np.array(grid['depth'], np.float64),
             input_size=np.array(input_size), downsample=np.array(ds),
             in_channels=np.array(Cin), out_channels=np.array(Cout),
             x=_np(x), x2=_np(x2), rots=_np(cam[0]), trans=_np(cam[1]),
             intrins=_np(cam[2]), post_rots=_np(cam[3]),
             post_trans=_np(cam[4]), bda=_np(cam[5]), **arrays)
        print(name, 'fraction of points at or below the depth threshold:',
              float(arrays['frac_below_thresh']))

    run('l_lss_v1', 'LSSViewTransformer', 31)
    run('l_lss_v2', 'LSSViewTransformer2', 32)
    run('l_lss_bevdepth', 'LSSViewTransformerBEVDepth', 33, depthnet=True)

    # ---- CM_DepthNet tail ----
    ns = load_cm_depth_net()
    torch.manual_seed(41)
    Bc, Nc, Cin, mid, Cctx, D, Hc, Wc = 2, 3, 12, 16, 8, 10, 4, 11
    net = ns['CM_DepthNet'](in_channels=Cin, context_channels=Cctx,
                            depth_channels=D, mid_channels=mid, use_dcn=False,
                            use_aspp=False)
    _randomise_bn(net, torch.Generator().manual_seed(42))
    net.eval()
    cap = {}
    net.context_conv.register_forward_hook(
        lambda m, i, o: cap.__setitem__('ctx_in', i[0]))
    net.depth_conv.register_forward_hook(
        lambda m, i, o: cap.__setitem__('logits', o))
    g = torch.Generator().manual_seed(43)
    x = torch.randn(Bc, Nc, Cin, Hc, Wc, generator=g)
    mlp_input = torch.randn(Bc, Nc, 27, generator=g)
    with torch.no_grad():
        context, depth = net(x, mlp_input)
    save('l_cm_tail', B=np.array(Bc), N=np.array(Nc),
         ctx_in=_np(cap['ctx_in']), logits=_np(cap['logits']),
         context_conv_weight=_np(net.context_conv.weight),
         context_conv_bias=_np(net.context_conv.bias),
         context=_np(context), depth=_np(depth))
