"""Importable alias of the ``3d-dual-fusion_b200/`` package directory.

The product package lives in ``3d-dual-fusion_b200/`` (a name Python cannot import); this shim
makes it importable as ``ddf_b200`` by pointing the package search path there.
"""
import os as _os

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "3d-dual-fusion_b200")
__path__.insert(0, _pkg_dir)

with open(_os.path.join(_pkg_dir, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_pkg_dir, "__init__.py"), "exec"))
