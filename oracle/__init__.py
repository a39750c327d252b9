"""ORACLE — CPU restatements of the reference algorithms on the hot path.

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package (ddf_b200).
"""
import ctypes
import os

from . import build as _build

_lib = None


def lib():
    """The compiled C oracle (built on demand with gcc; prebuilt file reused when current)."""
    global _lib
    if _lib is None:
        path = _build.LIB_PATH
        try:
            path = _build.build()
        except Exception:
            if not os.path.exists(path):
                raise
        _lib = ctypes.CDLL(path)
    return _lib
