"""
CPU oracle of the benchmark log-density (TEST INFRASTRUCTURE ONLY; same rules as pm_oracle.py).

Restates, in torch-CPU float64 on top of pm_oracle, the chain that `montecosmo_b200/model.py:FieldModel` runs on the
GPU: model.py:640-679 (prior, precond='real' or 'fourier', bricks.py:290-320), bricks.py:152-157 (white2lin), bricks.py:358-362 (linear Lagrangian
bias, NGP read), nbody.py:634-667 / 967-1002 (lpt / nbody_bf), bricks.py:781-792 (flat-sky RSD in cell units),
model.py:802-809 (nufft paint, irfftn), and a Gaussian likelihood.  Autograd gives the reference gradient.
"""
import numpy as np
import torch

from . import pm_oracle as O


def evolve(white, transfer, cosmo, mesh_shape, evolution="nbody", n_steps=5, a_start=0.0, a_obs=1.0, lpt_order=2,
           paint_order=2, interlace_order=2, paint_deconv=True, paint_shape=None, b1=1.0, rsd=True,
           los=(0.0, 0.0, 1.0), precond="real"):
    mesh_shape = tuple(mesh_shape)
    paint_shape = mesh_shape if paint_shape is None else tuple(paint_shape)
    # samp2base_mesh, bricks.py:300-309: sample in real space (rfftn) or in Fourier space (rg2cgh)
    dk = (torch.fft.rfftn(white) if precond == "real" else O.rg2cgh(white)) * O._t(transfer)
    q = O.regular_pos(mesh_shape)
    weights = 1.0
    if b1 != 0.0:
        delta = torch.fft.irfftn(dk, s=mesh_shape)
        weights = 1.0 + b1 * O.a2g(cosmo, a_obs) * O.read(q, delta, 1)
    if evolution == "lpt":
        dpos, vel = O.lpt(cosmo, dk, q, a_obs, lpt_order, 1)
        pos = q + dpos
    else:
        pos, vel = O.nbody_bf(cosmo, dk, q, a_start, a_obs, n_steps, paint_order, lpt_order, paint_deconv=False)
        pos, vel = pos[-1], vel[-1]
    if rsd:
        l = O._t(np.asarray(los))
        pos = pos + (vel * l).sum(-1, keepdim=True) * (O.a2g(cosmo, a_obs) * O.a2f(cosmo, a_obs)) * l
    gxy = O.nufft(pos, mesh_shape, paint_shape, weights, paint_order, interlace_order, paint_deconv=paint_deconv)
    if paint_shape != mesh_shape:
        gxy = O.chreshape(gxy, O.r2chshape(paint_shape))
    return torch.fft.irfftn(gxy, s=paint_shape)


def logpdf(white, obs, transfer, cosmo, mesh_shape, sigma_obs=1.0, **kw):
    gxy = evolve(white, transfer, cosmo, mesh_shape, **kw)
    return -0.5 * ((gxy - O._t(obs)) ** 2).sum() / sigma_obs**2 - 0.5 * (white**2).sum()


def value_and_force(white, obs, transfer, cosmo, mesh_shape, **kw):
    w = O._t(white).clone().requires_grad_(True)
    lp = logpdf(w, obs, transfer, cosmo, mesh_shape, **kw)
    (g,) = torch.autograd.grad(lp, w)
    return lp.detach(), g
