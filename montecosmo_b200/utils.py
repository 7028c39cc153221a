"""
Mirror of the `montecosmo/utils.py` callables that sit on the engine's path, with the reference's names, arguments and
errors: rg2cgh / cgh2rg (utils.py:785-921), chreshape (975-1013), r2chshape / ch2rshape (769-783), safe_div (21-29).
Arrays are torch tensors on the engine's device; rg2cgh is differentiable (torch.autograd.Function over mcpm_rg2cgh /
mcpm_rg2cgh_vjp), the others are thin wrappers.
"""
import numpy as np
import torch

from . import nbody as _nb
from .nbody import chreshape  # noqa: F401  (re-exported under its reference module)
from .ops import ch2rshape, r2chshape  # noqa: F401


def safe_div(x, y):
    """x / y, 0 where y == 0 (utils.py:21-29)."""
    x, y = torch.as_tensor(x), torch.as_tensor(y)
    ok = y != 0
    return torch.where(ok, x / torch.where(ok, y, torch.ones_like(y)), torch.zeros_like(x / torch.ones_like(y)))


def _rg_scale(shape, norm):
    n = float(np.prod(shape))
    if norm == "backward":
        return (n / 2.0) ** 0.5
    if norm == "ortho":
        return 0.5 ** 0.5
    if norm == "forward":
        return (2.0 * n) ** -0.5
    if norm == "amp":
        raise NotImplementedError("norm='amp' (mask / amplitude partition) is not implemented by the B200 engine")
    raise AssertionError("norm must be either 'backward', 'forward', 'ortho', or 'amp'.")


class _Rg2Cgh(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mesh, scale, transfer):
        ctx.scale, ctx.transfer = scale, transfer
        return _nb.ops().rg2cgh(mesh, scale, transfer)

    @staticmethod
    def backward(ctx, outbar):
        return _nb.ops().rg2cgh_vjp(outbar.contiguous(), ctx.scale, ctx.transfer), None, None


def rg2cgh(mesh, norm="backward", transfer=None):
    """Permute and reweight a real Gaussian tensor (3D) into a complex Gaussian Hermitian tensor, distributed as
    rfftn(N(0,I), norm) (utils.py:888-903).  `transfer` (extension): a real half-spectrum array multiplied in the same
    pass, the `mesh *= transfer` of samp2base_mesh (bricks.py:309)."""
    mesh = _nb._f32(mesh)
    if mesh.dim() != 3 or any(s % 2 for s in mesh.shape):
        raise AssertionError("dimension lengths must be even.")
    transfer = None if transfer is None else _nb._f32(transfer).detach()
    return _Rg2Cgh.apply(mesh, _rg_scale(tuple(mesh.shape), norm), transfer)


def cgh2rg(meshk, norm="backward"):
    """Permute and reweight a complex Gaussian Hermitian tensor into a real Gaussian tensor (utils.py:906-921)."""
    meshk = _nb._c64(meshk)
    shape = ch2rshape(tuple(meshk.shape))
    if any(s % 2 for s in shape):
        raise AssertionError("dimension lengths must be even.")
    return _nb.ops().cgh2rg(meshk.detach(), 1.0 / _rg_scale(shape, norm))
