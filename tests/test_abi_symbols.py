"""The C-ABI libraries load and export every symbol include/mcpm.h declares (no compute calls: runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mcpm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcpm_\w+)\s*\(", text)))


def test_header_and_ctypes_table_agree():
    from montecosmo_b200 import _capi
    assert set(declared_symbols()) == set(_capi.SIGNATURES), \
        set(declared_symbols()) ^ set(_capi.SIGNATURES)


def test_cuda_library_exports_every_declared_symbol():
    path = os.path.join(ROOT, "montecosmo_b200", "libmcpm.so")
    if not os.path.exists(path):
        pytest.skip("libmcpm.so not built (run __graft_entry__.build())")
    try:
        lib = ctypes.CDLL(path)
    except OSError as e:  # e.g. libcufft missing on a machine without the CUDA toolkit
        pytest.skip(f"cannot dlopen libmcpm.so here: {e}")
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    lib.mcpm_version.restype = ctypes.c_int
    assert lib.mcpm_version() == 100


def test_cpu_port_exports_every_declared_symbol():
    from oracle import cpu_port
    lib = cpu_port.load()
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_package_has_no_cpu_fallback():
    """The product loader must fail loudly without CUDA, and nothing under montecosmo_b200/ may import oracle/."""
    import torch
    pkg = os.path.join(ROOT, "montecosmo_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn
    if not torch.cuda.is_available():
        from montecosmo_b200.ops import TorchCudaAdapter
        with pytest.raises(RuntimeError):
            TorchCudaAdapter()
