"""Import the reference's OWN Python for the hot path in this container.

Used only by ``tests/golden/gen_golden.py`` (to produce the committed golden
fixtures) -- never at test run time on the GPU box, where ``/root/reference``
does not exist.

``import mmdet3d`` is impossible here (mmcv / mmdet / mmseg are not installed,
and ``mmdet3d/models/fbbev/custom_ops/__init__.py:1-3`` imports files that are
not in the tree), so this module

1. installs minimal stand-ins for the mmcv / mmdet names the reference files
   import (registries, ``BaseModule``, ``force_fp32`` as identity, and the
   third-party pieces restated in ``oracle/torch_ref.py``), and
2. loads the reference files straight from ``/root/reference`` as members of a
   synthetic package (their own ``__init__.py`` is not executed).

One line of one reference file is rewritten IN MEMORY: the device test at
``spatial_cross_attention_depth.py:577`` (``if torch.cuda.is_available() and
value.is_cuda:``) becomes ``if True:`` so the depth-aware CUDA branch -- the
one FB-OCC actually runs -- executes on CPU, with ``ms_deform_attn_forward``
served by the C oracle.  Nothing is copied into this repository.
"""
import copy
import importlib
import importlib.abc
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REF = os.environ.get("FBBEV_REFERENCE", "/root/reference")
_VT = os.path.join(REF, "mmdet3d/models/fbbev/view_transformation")

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(
    os.path.abspath(__file__)))))
from oracle import torch_ref  # noqa: E402


class Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return _reg(module)
        return _reg

    def build(self, cfg, **default):
        cfg = dict(cfg)
        for k, v in default.items():
            cfg.setdefault(k, v)
        typ = cfg.pop('type')
        return self.module_dict[typ](**cfg)


class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self._is_init = False
        self.init_cfg = copy.deepcopy(init_cfg)

    def init_weights(self):
        for m in self.children():
            if hasattr(m, 'init_weights'):
                m.init_weights()


class ModuleList(BaseModule, nn.ModuleList):
    def __init__(self, modules=None, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.ModuleList.__init__(self, modules)


class Sequential(BaseModule, nn.Sequential):
    def __init__(self, *args, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.Sequential.__init__(self, *args)


class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def _identity_decorator_factory(*a, **k):
    def deco(fn):
        return fn
    return deco


REGS = {n: Registry(n) for n in (
    'ATTENTION', 'TRANSFORMER_LAYER', 'TRANSFORMER_LAYER_SEQUENCE',
    'FEEDFORWARD_NETWORK', 'POSITIONAL_ENCODING', 'TRANSFORMER', 'HEADS',
    'NECKS')}
REGS['ATTENTION'].register_module(
    'MultiScaleDeformableAttention',
    module=torch_ref.MultiScaleDeformableAttention)
REGS['FEEDFORWARD_NETWORK'].register_module('FFN', module=torch_ref.FFN)


class TransformerLayerSequence(BaseModule):
    """mmcv.cnn.bricks.transformer.TransformerLayerSequence (ctor only)."""

    def __init__(self, transformerlayers=None, num_layers=None, init_cfg=None):
        super().__init__(init_cfg)
        if isinstance(transformerlayers, dict):
            transformerlayers = [copy.deepcopy(transformerlayers)
                                 for _ in range(num_layers)]
        self.num_layers = num_layers
        self.layers = ModuleList()
        for i in range(num_layers):
            self.layers.append(
                REGS['TRANSFORMER_LAYER'].build(transformerlayers[i]))
        self.embed_dims = self.layers[0].embed_dims
        self.pre_norm = self.layers[0].pre_norm


def _build_norm_layer(cfg, num_features):
    assert cfg['type'] == 'LN'
    return 'ln', nn.LayerNorm(num_features)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Anything:
    def __getattr__(self, k):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


def install_stubs():
    ext = types.SimpleNamespace(
        ms_deform_attn_forward=torch_ref.ms_deform_attn_forward,
        ms_deform_attn_backward=torch_ref.ms_deform_attn_backward)
    _mod('mmcv', ConfigDict=ConfigDict,
         deprecated_api_warning=_identity_decorator_factory)
    _mod('mmcv.cnn', xavier_init=torch_ref.xavier_init,
         constant_init=torch_ref.constant_init, Linear=nn.Linear,
         build_activation_layer=lambda cfg: nn.ReLU(inplace=True),
         build_norm_layer=_build_norm_layer, build_conv_layer=_Anything(),
         bias_init_with_prob=_Anything())
    _mod('mmcv.cnn.bricks')
    _mod('mmcv.cnn.bricks.registry', **{k: REGS[k] for k in (
        'ATTENTION', 'TRANSFORMER_LAYER', 'TRANSFORMER_LAYER_SEQUENCE',
        'FEEDFORWARD_NETWORK', 'POSITIONAL_ENCODING')})
    _mod('mmcv.cnn.bricks.transformer',
         build_attention=REGS['ATTENTION'].build,
         build_feedforward_network=REGS['FEEDFORWARD_NETWORK'].build,
         build_positional_encoding=REGS['POSITIONAL_ENCODING'].build,
         build_transformer_layer_sequence=REGS[
             'TRANSFORMER_LAYER_SEQUENCE'].build,
         TransformerLayerSequence=TransformerLayerSequence,
         FFN=torch_ref.FFN, POSITIONAL_ENCODING=REGS['POSITIONAL_ENCODING'])
    _mod('mmcv.runner', BaseModule=BaseModule,
         force_fp32=_identity_decorator_factory,
         auto_fp16=_identity_decorator_factory)
    _mod('mmcv.runner.base_module', BaseModule=BaseModule,
         ModuleList=ModuleList, Sequential=Sequential)
    _mod('mmcv.utils', ext_loader=types.SimpleNamespace(
        load_ext=lambda name, fns: ext), TORCH_VERSION=torch.__version__,
        digit_version=lambda v: tuple(int(x) for x in v.split('+')[0].split('.')))
    _mod('mmcv.ops',
         MultiScaleDeformableAttention=torch_ref.MultiScaleDeformableAttention)
    _mod('mmcv.ops.multi_scale_deform_attn',
         multi_scale_deformable_attn_pytorch=torch_ref.
         multi_scale_deformable_attn_pytorch,
         MultiScaleDeformableAttention=torch_ref.MultiScaleDeformableAttention)
    _mod('mmdet')
    _mod('mmdet.core', multi_apply=_Anything(), reduce_mean=_Anything())
    _mod('mmdet.models', HEADS=REGS['HEADS'], build_neck=_Anything())
    _mod('mmdet.models.utils', build_transformer=REGS['TRANSFORMER'].build)
    _mod('mmdet.models.utils.builder', TRANSFORMER=REGS['TRANSFORMER'])
    _mod('mmdet.models.utils.transformer', inverse_sigmoid=_Anything())
    _mod('mmdet.models.dense_heads', DETRHead=object)
    _mod('mmdet.models.backbones')
    _mod('mmdet.models.backbones.resnet', BasicBlock=object)
    _mod('mmdet3d')
    _mod('mmdet3d.core')
    _mod('mmdet3d.core.bbox')
    _mod('mmdet3d.core.bbox.coders', build_bbox_coder=_Anything())
    _mod('mmdet3d.models')
    _mod('mmdet3d.models.builder', NECKS=REGS['NECKS'])
    _mod('mmdet3d.models.fbbev')
    _mod('mmdet3d.models.fbbev.custom_ops')
    _mod('mmdet3d.models.fbbev.custom_ops.bev_pool_v2', bev_pool_v2=None)
    _mod('mmdet3d.models.fbbev.custom_ops.multi_scale_deformable_attn',
         multi_scale_deformable_attn=None)
    _mod('mmdet3d.ops')
    _mod('mmdet3d.ops.bev_pool_v2')
    # bev_pool_v2 (the op) is served by the oracle when the reference's
    # view transformer calls it on CPU
    _mod('mmdet3d.ops.bev_pool_v2.bev_pool', bev_pool_v2=_oracle_bev_pool_v2)


def _oracle_bev_pool_v2(depth, feat, ranks_depth, ranks_feat, ranks_bev,
                        bev_feat_shape, interval_starts, interval_lengths):
    from oracle import cpu
    out = cpu.bev_pool_v2(
        depth.contiguous().numpy(), feat.contiguous().numpy(),
        ranks_depth.numpy(), ranks_feat.numpy(), ranks_bev.numpy(),
        tuple(int(s) for s in bev_feat_shape), interval_starts.numpy(),
        interval_lengths.numpy())
    return torch.from_numpy(out)


_PATCHES = {
    'spatial_cross_attention_depth.py': [
        ('        if torch.cuda.is_available() and value.is_cuda:',
         '        if True:'),
    ],
}


class _PatchingLoader(importlib.abc.SourceLoader):
    def __init__(self, fullname, path):
        self.fullname, self.path = fullname, path

    def get_filename(self, fullname):
        return self.path

    def get_data(self, path):
        with open(path, 'rb') as f:
            data = f.read()
        if not path.endswith('.py'):
            return data
        for old, new in _PATCHES.get(os.path.basename(path), []):
            assert data.count(old.encode()) >= 1, (path, old)
            data = data.replace(old.encode(), new.encode())
        return data


def _load_pkg(pkg_name, directory, modules):
    pkg = types.ModuleType(pkg_name)
    pkg.__path__ = [directory]
    sys.modules[pkg_name] = pkg
    out = {}
    for m in modules:
        full = f'{pkg_name}.{m}'
        path = os.path.join(directory, m + '.py')
        spec = importlib.util.spec_from_loader(
            full, _PatchingLoader(full, path), origin=path)
        mod = importlib.util.module_from_spec(spec)
        mod.__file__ = path
        sys.modules[full] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, m, mod)
        out[m] = mod
    return out


_loaded = None


def load_reference():
    """Returns a namespace with the reference's hot-path modules."""
    global _loaded
    if _loaded is not None:
        return _loaded
    install_stubs()
    fwd = _load_pkg('ref_forward_projection',
                    os.path.join(_VT, 'forward_projection'),
                    ['view_transformer'])
    utils = _load_pkg(
        'ref_bevformer_utils',
        os.path.join(_VT, 'backward_projection', 'bevformer_utils'),
        ['multi_scale_deformable_attn_function',
         'custom_base_transformer_layer', 'spatial_cross_attention_depth',
         'positional_encoding', 'bevformer_encoder', 'bevformer'])
    bp = _load_pkg('ref_backward_projection',
                   os.path.join(_VT, 'backward_projection'),
                   ['backward_projection'])
    _loaded = types.SimpleNamespace(
        view_transformer=fwd['view_transformer'],
        backward_projection=bp['backward_projection'], **utils)
    return _loaded
