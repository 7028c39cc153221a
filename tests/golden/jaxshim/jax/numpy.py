"""NumPy-backed stand-in for `jax.numpy` (float64, eager). Test infrastructure only."""
import sys as _sys
import types as _types
import numpy as _np


class JArray(_np.ndarray):
    """ndarray with the functional-update `.at[idx].op(v)` interface of jax arrays."""

    @property
    def at(self):
        return _At(self)


class _At:
    def __init__(self, a):
        self.a = a

    def __getitem__(self, idx):
        return _AtIdx(self.a, idx)


def _plain(idx):
    if isinstance(idx, tuple):
        return tuple(_plain(i) for i in idx)
    if isinstance(idx, _np.ndarray):
        return _np.asarray(idx)
    return idx


class _AtIdx:
    def __init__(self, a, idx):
        self.a, self.idx = a, _plain(idx)

    def _copy(self):
        return _np.array(self.a, copy=True)

    def add(self, v):
        out = self._copy()
        _np.add.at(out, self.idx, _np.asarray(v))
        return out.view(JArray)

    def set(self, v):
        out = self._copy()
        out[self.idx] = _np.asarray(v)
        return out.view(JArray)

    def multiply(self, v):
        out = self._copy()
        out[self.idx] = out[self.idx] * _np.asarray(v)
        return out.view(JArray)

    def divide(self, v):
        out = self._copy()
        out[self.idx] = out[self.idx] / _np.asarray(v)
        return out.view(JArray)


def _wrap(x):
    if isinstance(x, _np.ndarray) and not isinstance(x, JArray):
        return x.view(JArray)
    if isinstance(x, tuple):
        return tuple(_wrap(i) for i in x)
    if isinstance(x, list):
        return [_wrap(i) for i in x]
    return x


def _wrapped(fn):
    def f(*a, **k):
        return _wrap(fn(*a, **k))
    f.__name__ = getattr(fn, "__name__", "f")
    return f


ndarray = JArray
pi = _np.pi
inf = _np.inf
nan = _np.nan
newaxis = None
float32, float64, int16, int32, int64, complex64, complex128 = (
    _np.float32, _np.float64, _np.int16, _np.int32, _np.int64, _np.complex64, _np.complex128)
bool_ = _np.bool_


def bincount(x, weights=None, minlength=0, length=None):
    """jnp.bincount: `length` fixes the output size (static shape under jit); values beyond it are dropped."""
    n = int(length) if length is not None else int(minlength)
    out = _np.bincount(_np.asarray(x), weights=None if weights is None else _np.asarray(weights), minlength=n)
    return _wrap(out[:n] if length is not None else out)


def clip(a, a_min=None, a_max=None):
    """jnp.clip: either bound may be omitted (nbody.py:859 passes the lower one only)."""
    return _wrap(_np.clip(_np.asarray(a), a_min, a_max))


def unstack(x, axis=0):
    return tuple(_wrap(_np.moveaxis(x, axis, 0)[i]) for i in range(x.shape[axis]))


class _FFT(_types.ModuleType):
    def __getattr__(self, name):
        return _wrapped(getattr(_np.fft, name))


fft = _FFT("jax.numpy.fft")
_sys.modules["jax.numpy.fft"] = fft


def __getattr__(name):
    obj = getattr(_np, name)
    if callable(obj) and not isinstance(obj, type):
        return _wrapped(obj)
    return obj
