"""Importable alias of the ``fb-bev_b200/`` package directory.

``fb-bev_b200`` (the name this repository's layout prescribes) contains a
hyphen and therefore cannot be imported directly; this shim points the
``fbbev_b200`` package's search path at that directory and executes its
``__init__``.  No code lives here.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(
    _os.path.abspath(__file__))), "fb-bev_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
