"""Array adapters for the test-suite: NumPy over the host-emulation library (CPU), torch over libmcpm.so (GPU)."""
import numpy as np


class NumpyAdapter:
    _DT = {"f32": np.float32, "c64": np.complex64, "f64": np.float64}

    def empty(self, shape, dtype="f32"):
        return np.full(tuple(int(s) for s in shape), np.nan, dtype=self._DT[dtype])  # NaN-poisoned, like fresh HBM

    def zeros(self, shape, dtype="f32"):
        return np.zeros(tuple(int(s) for s in shape), dtype=self._DT[dtype])

    def prepare(self, x, dtype="f32"):
        return np.ascontiguousarray(np.asarray(x), dtype=self._DT[dtype])

    def ptr(self, x):
        return 0 if x is None else x.ctypes.data

    def stream(self):
        return 0

    def shape(self, x):
        return tuple(x.shape)


def to_numpy(x):
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x
    return x.detach().cpu().numpy()


def torch_cpu_adapter():
    from oracle.cpu_port import torch_cpu_adapter as f
    return f()
