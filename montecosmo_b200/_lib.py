"""Loader of the CUDA engine, libmcpm.so (built in-tree by `make -C montecosmo_b200/csrc` / __graft_entry__.build()).

There is no fallback: if the library is missing or does not export the full ABI of include/mcpm.h, importing fails.
"""
import ctypes
import os

from . import _capi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmcpm.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build the sm_100a engine with `make -C montecosmo_b200/csrc` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). montecosmo_b200 has no CPU fallback.")
        _lib = _capi.bind(ctypes.CDLL(LIB_PATH))
        if _lib.mcpm_version() != 100:
            raise ImportError("libmcpm.so version does not match the Python layer")
    return _lib
