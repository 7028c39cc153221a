"""
Multi-core CPU port of the particle-mesh path  --  CHECKER / CPU BASELINE ONLY, never a product path.

The engine's kernel sources (montecosmo_b200/csrc/*.cu) compile as plain C++ with -DMCPM_HOSTEMU: every kernel body
runs as an OpenMP loop over the same per-element functor the GPU runs, scatter-adds use `omp atomic`, and the 3-D FFTs
are delegated to scipy.fft (pocketfft, all cores) through a hook.  The result, oracle/_build/libmcpm_cpu.so, exports
the same C ABI as libmcpm.so and is used
  * by tests/ (-m "not gpu") to check the C ABI's orchestration and adjoints against the float64 oracle without a GPU;
  * by bench.py's `cpu_baseline` / `--impl reference` legs as the CPU timing of the reference ALGORITHM
    (montecosmo/nbody.py restated; "kind": "port") on this box's host cores.
The independent parity oracle remains oracle/pm_oracle.py (torch float64, pinned to the reference source).
The montecosmo_b200 package never imports this module and never looks for this library.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import scipy.fft as sfft

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "montecosmo_b200", "csrc")
OUT_DIR = os.path.join(ROOT, "oracle", "_build")
OUT = os.path.join(OUT_DIR, "libmcpm_cpu.so")
SOURCES = ["api.cu", "engine.cu", "paint.cu", "fourier.cu", "bias.cu", "fft.cu", "cic4.cu", "cic4_tma.cu", "halo.cu", "brick.cu", "xfft.cu", "yzfft.cu", "xfft_large.cu", "xfft_1024.cu"]
HEADERS = ["rt.h", "engine.h", "window.h", "frame.h", "obs.h", "tma.h", "kspace.h", "xfft_kernel.h"]

_HOOK = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int)
_keep = []
_lib = None


def build(force=False):
    srcs = [os.path.join(SRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(SRC, h) for h in HEADERS] + [os.path.join(ROOT, "include", "mcpm.h")]
    if not force and os.path.exists(OUT):
        try:
            if all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
                return OUT
        except OSError:
            return OUT  # sources absent (deployed copy): use the prebuilt library
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["g++", "-x", "c++", "-std=c++17", "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", "-DMCPM_HOSTEMU",
           "-fvisibility=hidden", "-o", OUT] + srcs
    subprocess.run(cmd, check=True, cwd=SRC)
    return OUT


def _as(ptr, dtype, shape):
    n = int(np.prod(shape))
    buf = (C.c_byte * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def _r2c(inp, out, nx, ny, nz, batch):
    a = _as(inp, np.float32, (batch, nx, ny, nz))
    o = _as(out, np.complex64, (batch, nx, ny, nz // 2 + 1))
    o[...] = sfft.rfftn(a, axes=(1, 2, 3), workers=-1)


def _c2r(inp, out, nx, ny, nz, batch):
    a = _as(inp, np.complex64, (batch, nx, ny, nz // 2 + 1))
    o = _as(out, np.float32, (batch, nx, ny, nz))
    o[...] = sfft.irfftn(a, s=(nx, ny, nz), axes=(1, 2, 3), workers=-1, norm="forward")  # unnormalised, like cuFFT C2R


_SLABHOOK = C.CFUNCTYPE(None, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)


def _slab(kind, inp, out, nx, ny, nz, nb, xl, kyl):
    nzc = nz // 2 + 1
    if kind == 0:  # 2-D R2C over (y, z) of the local planes
        a = _as(inp, np.float32, (nb, xl, ny, nz))
        _as(out, np.complex64, (nb, xl, ny, nzc))[...] = sfft.rfftn(a, axes=(2, 3), workers=-1)
    elif kind == 1:  # 2-D C2R, unnormalised
        a = _as(inp, np.complex64, (nb, xl, ny, nzc))
        _as(out, np.float32, (nb, xl, ny, nz))[...] = sfft.irfftn(a, s=(ny, nz), axes=(2, 3), workers=-1, norm="forward")
    else:  # 1-D C2C along x of the local ky block, in place, unnormalised
        a = _as(inp, np.complex64, (nb, nx, kyl, nzc))
        res = sfft.fft(a, axis=1, workers=-1) if kind == 2 else sfft.ifft(a, axis=1, workers=-1, norm="forward")
        _as(out, np.complex64, (nb, nx, kyl, nzc))[...] = res


def load():
    """The CPU library with prototypes bound and FFT hooks registered."""
    global _lib
    if _lib is None:
        from montecosmo_b200 import _capi  # prototypes only (ctypes signatures of include/mcpm.h)
        lib = C.CDLL(build())
        _capi.bind(lib)
        lib.mcpm_hostemu_set_fft.argtypes = [_HOOK, _HOOK]
        lib.mcpm_hostemu_set_fft.restype = None
        r2c, c2r = _HOOK(_r2c), _HOOK(_c2r)
        _keep.extend([r2c, c2r])
        lib.mcpm_hostemu_set_fft(r2c, c2r)
        lib.mcpm_hostemu_set_slabfft.argtypes = [_SLABHOOK]
        lib.mcpm_hostemu_set_slabfft.restype = None
        sh = _SLABHOOK(_slab)
        _keep.append(sh)
        lib.mcpm_hostemu_set_slabfft(sh)
        _lib = lib
    return _lib


def torch_cpu_adapter():
    """torch-CPU tensors over the CPU library, for driving montecosmo_b200's autograd layer without a GPU
    (the package's own adapter refuses to exist without CUDA)."""
    import torch
    from montecosmo_b200.ops import TorchCudaAdapter

    class TorchCpuAdapter(TorchCudaAdapter):
        def __init__(self):
            self.torch = torch
            self.device = torch.device("cpu")

        def stream(self):
            return 0

    return TorchCpuAdapter()


def cpu_ops():
    from montecosmo_b200.ops import Ops
    return Ops(load(), torch_cpu_adapter())
