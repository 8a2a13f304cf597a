"""Importable alias of the ``incubator-mxnet_b200/`` package directory (a hyphen cannot
appear in a Python module name).  ``import mxnet_b200`` executes
``incubator-mxnet_b200/__init__.py`` and resolves sub-modules from that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "incubator-mxnet_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
