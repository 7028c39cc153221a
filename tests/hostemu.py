"""
Host-emulation build of the engine for CPU-only test runs.

The engine's `.cu` sources compile as plain C++ with -DMCPM_HOSTEMU (kernel bodies run as serial loops, FFTs are
delegated to numpy through a hook).  This lets `-m "not gpu"` tests exercise the C-ABI's orchestration -- buffer
wiring, adjoint bookkeeping, normalisation -- against the oracle without a GPU.  It is NOT a product path: the
library is built into tests/_hostemu/, under a different name, and `montecosmo_b200` never looks for it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from montecosmo_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "montecosmo_b200", "csrc")
OUT_DIR = os.path.join(ROOT, "tests", "_hostemu")
OUT = os.path.join(OUT_DIR, "libmcpm_hostemu.so")
SOURCES = ["api.cu", "engine.cu", "paint.cu", "fourier.cu", "fft.cu"]

_HOOK = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int)
_keep = []


def build(force=False):
    srcs = [os.path.join(SRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(SRC, h) for h in ("rt.h", "engine.h", "window.h")] + \
        [os.path.join(ROOT, "include", "mcpm.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["g++", "-x", "c++", "-std=c++17", "-O2", "-fPIC", "-shared", "-DMCPM_HOSTEMU", "-fvisibility=hidden",
           "-Wno-unknown-pragmas", "-o", OUT] + srcs
    subprocess.run(cmd, check=True, cwd=SRC)
    return OUT


def _as(ptr, dtype, shape):
    n = int(np.prod(shape))
    buf = (C.c_byte * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def _r2c(inp, out, nx, ny, nz, batch):
    a = _as(inp, np.float32, (batch, nx, ny, nz))
    o = _as(out, np.complex64, (batch, nx, ny, nz // 2 + 1))
    o[...] = np.fft.rfftn(a.astype(np.float64), axes=(1, 2, 3)).astype(np.complex64)


def _c2r(inp, out, nx, ny, nz, batch):
    a = _as(inp, np.complex64, (batch, nx, ny, nz // 2 + 1))
    o = _as(out, np.float32, (batch, nx, ny, nz))
    # unnormalised, like cuFFT's C2R
    o[...] = (np.fft.irfftn(a.astype(np.complex128), s=(nx, ny, nz), axes=(1, 2, 3)) * (nx * ny * nz)).astype(np.float32)


def load():
    lib = C.CDLL(build())
    _capi.bind(lib)
    lib.mcpm_hostemu_set_fft.argtypes = [_HOOK, _HOOK]
    lib.mcpm_hostemu_set_fft.restype = None
    r2c, c2r = _HOOK(_r2c), _HOOK(_c2r)
    _keep.extend([r2c, c2r])
    lib.mcpm_hostemu_set_fft(r2c, c2r)
    return lib
