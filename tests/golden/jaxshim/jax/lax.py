"""Stand-in for jax.lax: scan as a Python loop."""
import numpy as _np
from .numpy import _wrap
from . import tree as _tree


def scan(f, init, xs, length=None):
    carry = init
    ys = []
    n = length if xs is None else len(_tree.leaves(xs)[0])
    for i in range(n):
        x = None if xs is None else _tree.map(lambda a: _wrap(_np.asarray(a)[i]), xs)
        carry, y = f(carry, x)
        ys.append(y)
    if ys and ys[0] is not None:
        ys = _tree.map(lambda *a: _wrap(_np.stack([_np.asarray(v) for v in a])), *ys)
    else:
        ys = None
    return carry, ys


def optimization_barrier(x):
    return x


def stop_gradient(x):
    return x


def cond(pred, t, f, *ops):
    return t(*ops) if pred else f(*ops)
