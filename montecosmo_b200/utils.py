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
    raise AssertionError("norm must be either 'backward', 'forward', 'ortho', or 'amp'.")


# norm = "amp" (utils.py:807-817, 858-868, 915-918): the same permutation applied to an AMPLITUDE mesh -- no sign on the
# mirrored imaginary parts, no sqrt2 on the 8 real modes, no scaling -- so that rg2cgh(mean + amp * N(0,I)) is
# distributed as meank + ampk * rfftn(N(0,I)).  Used at set-up time for masks and preconditioner scales
# (model.py:587, 884, 1147), never per evaluation: a gather through an index map built on the host.
def _amp_index_maps(shape):
    """(src of every real-mesh slot in the flattened half spectrum, src of every half-spectrum element in the flattened
    real mesh), following which Hermitian image the reference keeps (the mirrored one, utils.py:851-861)."""
    nx, ny, nz = shape
    hx, hy, hz = nx // 2, ny // 2, nz // 2
    nzc = hz + 1
    x, y, z = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    face = (z == 0) | (z == hz)
    edge = face & ((y == 0) | (y == hy))
    # real mesh slot -> half-spectrum element
    l = np.where(z > hz, z - hz, z)
    jj = np.where(y > hy, y - hy, y)  # the imaginary-part slots of a face read what their real-part twins read
    ii = np.where(x > hx, x - hx, x)
    si = np.where(edge, np.where((x == 0) | (x == hx), x, nx - ii), np.where(face, (nx - x) % nx, x))
    sj = np.where(edge, y, np.where(face, ny - jj, y))
    to_real = ((si * ny + sj) * nzc + l).reshape(-1)
    # half-spectrum element -> real mesh slot (real part of rg2cgh)
    i, j, l2 = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nzc), indexing="ij")
    face2 = (l2 == 0) | (l2 == hz)
    edge2 = face2 & ((j == 0) | (j == hy))
    ri = np.where(edge2, np.where(i > hx, nx - i, i), np.where(face2 & (j > hy), (nx - i) % nx, i))
    rj = np.where(face2 & ~edge2 & (j > hy), ny - j, j)
    to_k = ((ri * ny + rj) * nz + l2).reshape(-1)
    return to_real, to_k


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
    if norm == "amp":  # real output: the amplitude of every mode
        shape = tuple(mesh.shape)
        idx = torch.as_tensor(_amp_index_maps(shape)[1], device=mesh.device)
        return mesh.reshape(-1)[idx].reshape(r2chshape(shape))
    transfer = None if transfer is None else _nb._f32(transfer).detach()
    return _Rg2Cgh.apply(mesh, _rg_scale(tuple(mesh.shape), norm), transfer)


def cgh2rg(meshk, norm="backward"):
    """Permute and reweight a complex Gaussian Hermitian tensor into a real Gaussian tensor (utils.py:906-921)."""
    if norm == "amp":  # the same amplitude to the real and the imaginary part of every mode (utils.py:915-918)
        amp = torch.as_tensor(meshk)
        amp = _nb._f32(amp.real if torch.is_complex(amp) else amp)
        shape = ch2rshape(tuple(amp.shape))
        if any(s % 2 for s in shape):
            raise AssertionError("dimension lengths must be even.")
        idx = torch.as_tensor(_amp_index_maps(shape)[0], device=amp.device)
        return amp.reshape(-1)[idx].reshape(shape)
    meshk = _nb._c64(meshk)
    shape = ch2rshape(tuple(meshk.shape))
    if any(s % 2 for s in shape):
        raise AssertionError("dimension lengths must be even.")
    return _nb.ops().cgh2rg(meshk.detach(), 1.0 / _rg_scale(shape, norm))
