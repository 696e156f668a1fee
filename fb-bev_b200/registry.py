"""Registries and build helpers for the plugin classes.

The reference instantiates the view-transformation modules by name through
mmcv / mmdet / mmdet3d registries (``type='LSSViewTransformerFunction3D'`` etc.
in occupancy_configs/fb_occ/fbocc-r50-cbgs_depth_16f_16x4_20e.py:150-215).
That registry + kwargs contract is the plugin API this package preserves:

* when mmcv / mmdet / mmdet3d are importable, the classes are registered into
  THEIR registries (``force=True``), so an unmodified FB-OCC config builds the
  B200 modules instead of the stock ones;
* otherwise (this image has none of them) the local registries below offer the
  same ``register_module`` / ``build`` surface so the classes can be built
  from the very same config dicts.
"""
import copy
import os

import torch.nn as nn


class Registry:
    """Minimal mmcv-style registry (``register_module`` + ``build``)."""

    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def get(self, key):
        return self.module_dict.get(key)

    def register_module(self, name=None, force=False, module=None):
        def _register(cls):
            key = name or cls.__name__
            if key in self.module_dict and not force:
                raise KeyError(f'{key} is already registered in {self.name}')
            self.module_dict[key] = cls
            return cls
        if module is not None:
            return _register(module)
        return _register

    def build(self, cfg, **default_args):
        if cfg is None:
            return None
        if isinstance(cfg, nn.Module):
            return cfg
        cfg = copy.deepcopy(dict(cfg))
        for k, v in default_args.items():
            cfg.setdefault(k, v)
        typ = cfg.pop('type')
        cls = self.module_dict.get(typ) if isinstance(typ, str) else typ
        if cls is None and isinstance(typ, str):
            _load_plugins()
            cls = self.module_dict.get(typ)
        if cls is None:
            raise KeyError(f'{typ} is not in the {self.name} registry')
        return cls(**cfg)


def _load_plugins():
    """Import the modules whose classes register themselves."""
    import importlib
    importlib.import_module(__package__ + '.view_transformation')


NECKS = Registry('neck')
HEADS = Registry('head')
ATTENTION = Registry('attention')
FEEDFORWARD_NETWORK = Registry('feed-forward network')
POSITIONAL_ENCODING = Registry('position encoding')
TRANSFORMER = Registry('transformer')
TRANSFORMER_LAYER = Registry('transformerLayer')
TRANSFORMER_LAYER_SEQUENCE = Registry('transformer-layers sequence')

_LOCAL = dict(NECKS=NECKS, HEADS=HEADS, ATTENTION=ATTENTION,
              FEEDFORWARD_NETWORK=FEEDFORWARD_NETWORK,
              POSITIONAL_ENCODING=POSITIONAL_ENCODING, TRANSFORMER=TRANSFORMER,
              TRANSFORMER_LAYER=TRANSFORMER_LAYER,
              TRANSFORMER_LAYER_SEQUENCE=TRANSFORMER_LAYER_SEQUENCE)


def _upstream_registries():
    """The reference's registries, when its dependencies are installed."""
    found = {}
    try:  # pragma: no cover - mmcv is not available in the build image
        from mmcv.cnn.bricks import registry as r
        for k in ('ATTENTION', 'FEEDFORWARD_NETWORK', 'POSITIONAL_ENCODING',
                  'TRANSFORMER_LAYER', 'TRANSFORMER_LAYER_SEQUENCE'):
            found[k] = getattr(r, k)
    except Exception:
        pass
    try:  # pragma: no cover
        from mmdet.models.utils.builder import TRANSFORMER as t
        found['TRANSFORMER'] = t
    except Exception:
        pass
    try:  # pragma: no cover
        from mmdet3d.models.builder import HEADS as h, NECKS as n
        found['HEADS'], found['NECKS'] = h, n
    except Exception:
        pass
    return found


_UPSTREAM = _upstream_registries()


def register(registry_name, name=None, upstream=True):
    """Class decorator: register locally and, if present, upstream.

    ``upstream=False`` is for the replacements of GENERIC mmcv bricks
    (``MultiScaleDeformableAttention``, ``FFN``): every other model in the
    process resolves those names through mmcv's global registries, so they are
    only overridden there on request (``FBBEV_OVERRIDE_MMCV=1`` or
    :func:`override_upstream_generic`).  The plugin's own modules never need
    it: they build their children through the local registries below."""
    def _deco(cls):
        _LOCAL[registry_name].register_module(name=name, force=True,
                                              module=cls)
        if upstream or os.environ.get('FBBEV_OVERRIDE_MMCV', '0') == '1':
            up = _UPSTREAM.get(registry_name)
            if up is not None:  # pragma: no cover
                up.register_module(name=name, force=True, module=cls)
        else:
            _GENERIC.append((registry_name, name, cls))
        return cls
    return _deco


_GENERIC = []


def override_upstream_generic():
    """Opt in: also replace mmcv's global ``MultiScaleDeformableAttention`` and
    ``FFN`` registrations with this package's classes."""
    for registry_name, name, cls in _GENERIC:
        up = _UPSTREAM.get(registry_name)
        if up is not None:  # pragma: no cover
            up.register_module(name=name, force=True, module=cls)


def build_attention(cfg, **kw):
    return ATTENTION.build(cfg, **kw)


def build_feedforward_network(cfg, **kw):
    return FEEDFORWARD_NETWORK.build(cfg, **kw)


def build_positional_encoding(cfg, **kw):
    return POSITIONAL_ENCODING.build(cfg, **kw)


def build_transformer(cfg, **kw):
    return TRANSFORMER.build(cfg, **kw)


def build_transformer_layer(cfg, **kw):
    return TRANSFORMER_LAYER.build(cfg, **kw)


def build_transformer_layer_sequence(cfg, **kw):
    return TRANSFORMER_LAYER_SEQUENCE.build(cfg, **kw)


def build_neck(cfg, **kw):
    return NECKS.build(cfg, **kw)


def build_head(cfg, **kw):
    return HEADS.build(cfg, **kw)


class BaseModule(nn.Module):
    """Stand-in for ``mmcv.runner.BaseModule`` (init_cfg bookkeeping only)."""

    def __init__(self, init_cfg=None):
        super().__init__()
        self._is_init = False
        self.init_cfg = copy.deepcopy(init_cfg)

    def init_weights(self):
        for m in self.children():
            if hasattr(m, 'init_weights'):
                m.init_weights()
        self._is_init = True
