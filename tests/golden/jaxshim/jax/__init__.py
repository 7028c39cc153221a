"""NumPy-backed stand-in for the parts of `jax` that montecosmo/{nbody,utils}.py touch. See ../README.md."""
from . import numpy, lax, tree, debug, random  # noqa: F401
from . import scipy  # noqa: F401


def jit(fun=None, *a, **k):
    if fun is None or not callable(fun):
        return lambda f: f
    return fun


def _no(*a, **k):
    raise NotImplementedError("autodiff/vmap are not emulated by the numpy stand-in")


vmap = grad = value_and_grad = pmap = _no


class config:
    @staticmethod
    def update(*a, **k):
        pass
