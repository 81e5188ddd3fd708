"""Importable alias of the package directory ``stable-ts_b200/`` (a hyphen cannot appear in an import name).

``import stable_ts_b200`` resolves every submodule from ``../stable-ts_b200/``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "stable-ts_b200")
__path__ = [_real]

with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
