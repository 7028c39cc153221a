"""
Mirror of the `montecosmo/bricks.py` callables next to the engine's path: regular_pos (593-603), the
Lagrangian bias expansion (327-452) and the flat-sky redshift-space shift in cell units (781-792).

lagrangian_bias composes engine operators -- irfftn (mcpm_irfftn) and read (mcpm_read, differentiable in mesh and
positions) -- with pointwise torch expressions for the Fourier multipliers and the shear invariants, so torch.autograd
differentiates it end to end.  Primordial non-Gaussianity terms (png_type is not None) are not implemented.
"""
import numpy as np
import torch

from . import cosmo as _cosmo
from . import nbody as _nb

_BIAS_KEYS = ("b1", "b2", "bs2", "b3", "bds2", "bs3", "bn2", "bnpar")


def regular_pos(mesh_shape, ptcl_shape=None):
    """Particle lattice in cell units, C order (bricks.py:593-603)."""
    ptcl_shape = mesh_shape if ptcl_shape is None else ptcl_shape
    ax = [np.arange(p, dtype=np.float32) * (m / p) for m, p in zip(mesh_shape, ptcl_shape)]
    return _nb.ops().A.prepare(np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3))


def lagrangian_bias(cosmo, pos, a, box_size, lin_mesh, bias, png=None, png_type=None, kpow=None, read_order: int = 2):
    """Lagrangian bias expansion weights (bricks.py:327-452):
    w = 1 + b1 d + b2 (d^2 - <d^2>)/2 + bs2 (s^2 - <s^2>) + b3 (d^3 - 3<d^2> d)/6 + bds2 d s^2 + bs3 s^3 + bn2 lap d,
    and the higher-derivative velocity term dvel = bnpar grad d.  Returns (weights, dvel, phi = 0)."""
    if png_type is not None:
        raise NotImplementedError("primordial non-Gaussianity bias terms are not implemented by the B200 engine")
    b = {k: bias.get(k, 0.0) if isinstance(bias, dict) else 0.0 for k in _BIAS_KEYS}
    lin_mesh = _nb._c64(lin_mesh)
    pos = _nb._f32(pos)
    dev = lin_mesh.device
    mesh_shape = _nb.ch2rshape(tuple(lin_mesh.shape))
    growth = torch.as_tensor(_cosmo.a2g(cosmo, a), dtype=torch.float64)
    g = growth.to(device=dev, dtype=torch.float32).reshape(-1)  # scalar, or one value per particle
    g = g if g.numel() > 1 else g.reshape(())

    kvec = [torch.as_tensor(k.astype(np.float32), device=dev) for k in _nb.rfftk(mesh_shape, box_size)]  # h/Mpc
    k2 = sum(k**2 for k in kvec)
    inv_k2 = torch.where(k2 == 0, torch.zeros_like(k2), 1.0 / torch.where(k2 == 0, torch.ones_like(k2), k2))

    def rd(mesh):
        return _nb.read(pos, mesh, read_order)

    delta_pos = rd(_nb.irfftn(lin_mesh)) * g
    weights = 1.0 + b["b1"] * delta_pos
    delta2_pos = delta_pos**2
    sigma2 = delta2_pos.mean()
    delta2_pos = delta2_pos - sigma2
    weights = weights + b["b2"] * delta2_pos / 2

    # tidal shear s_ij = (k_i k_j / k^2 - delta_ij / 3) delta: five transforms, the sixth from tracelessness
    sh = {}
    for i in range(2):
        sh[(i, i)] = _nb.irfftn((kvec[i] * kvec[i] * inv_k2 - 1.0 / 3.0) * lin_mesh)
        for j in range(i + 1, 3):
            sh[(i, j)] = _nb.irfftn((kvec[i] * kvec[j] * inv_k2) * lin_mesh)
    sa, sb = sh[(0, 0)], sh[(1, 1)]
    sc = -(sa + sb)
    sd, se, sf = sh[(0, 1)], sh[(0, 2)], sh[(1, 2)]
    shear2 = sa**2 + sb**2 + sc**2 + 2 * (sd**2 + se**2 + sf**2)
    shear2_pos = rd(shear2) * g**2 - 2.0 / 3.0 * sigma2
    weights = weights + b["bs2"] * shear2_pos
    weights = weights + b["b3"] * (delta_pos**3 - 3 * sigma2 * delta_pos) / 6
    weights = weights + b["bds2"] * delta_pos * shear2_pos
    shear3 = 3 * (sa * (sb * sc - sf**2) - sd * (sd * sc - se * sf) + se * (sd * sf - sb * se))
    weights = weights + b["bs3"] * rd(shear3) * g**3
    weights = weights + b["bn2"] * rd(_nb.irfftn(-k2 * lin_mesh)) * g

    grads = [rd(_nb.irfftn((1j * k) * lin_mesh)) for k in kvec]  # h/Mpc
    gcol = g if g.dim() == 0 else g.reshape(-1, 1)
    dvel = b["bnpar"] * torch.stack(grads, dim=-1) * gcol
    return weights, dvel, 0.0


def rsd(cosmo, a, vel, los=(0.0, 0.0, 1.0)):
    """Flat-sky redshift-space displacement in cell units, dpos = (vel . los) D f los (bricks.py:781-792 with the
    growth-time velocity of nbody_bf)."""
    los = torch.as_tensor(np.asarray(los, dtype=np.float32), device=vel.device)
    coef = float(_cosmo.a2g(cosmo, a) * _cosmo.a2f(cosmo, a))
    return (vel * los).sum(-1, keepdim=True) * coef * los
