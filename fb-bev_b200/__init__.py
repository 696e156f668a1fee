"""fb-bev_b200 -- B200 (sm_100a) implementation of the FB-BEV / FB-OCC
forward-backward view-transformation hot path.

Import name: ``fbbev_b200`` (the directory name ``fb-bev_b200`` is not a valid
Python identifier; ``fbbev_b200/__init__.py`` at the repository root aliases
it).

Layout (only what the path needs):
  csrc/                 hand-written CUDA kernels + the C ABI (include/fbbev_b200.h)
  _lib.py               ctypes binding of lib/libfbbev_b200.so (fails loudly if absent)
  ops/                  mirrors of mmdet3d/ops/bev_pool_v2 and the mmcv MSDA function
  view_transformation/  mirrors of mmdet3d/models/fbbev/view_transformation
                        (forward_projection / backward_projection plugin classes)
  registry.py           mmcv-style registries + build helpers (registers into
                        mmcv/mmdet registries when those are importable)
  sharding.py           frame sharding across ranks + the optional BEV gather
  synthetic.py          seeded synthetic rig / inputs for tests and bench
"""
__version__ = "0.1.0"
