"""ORACLE package -- test infrastructure only (see DESIGN.md "Oracle").

Nothing under ``sfvos_b200`` / the product package imports from here.  ``build()`` compiles the two plain-C
restatements (``roi_align_ref.c``, ``conv3d_ref.c``) into ``oracle/_build/liboracle.so`` with gcc.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_SRCS = [os.path.join(_HERE, f) for f in ("roi_align_ref.c", "conv3d_ref.c")]
_lib = None


def build(force=False):
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in _SRCS):
        return _SO
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c99", "-o", _SO] + _SRCS + ["-lm"]
    subprocess.check_call(cmd)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib
