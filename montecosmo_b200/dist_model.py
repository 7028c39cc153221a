"""
The field-level log-density of `model.FieldModel` on the slab-decomposed engine (`dist.SlabPM`) -- SURVEY section 8e,
BASELINE configs C4 / C5: prior -> 2LPT -> BullFrog steps -> bias weights -> flat-sky RSD -> interlaced, deconvolved
final paint -> Gaussian likelihood, and its gradient w.r.t. the white field, with every mesh split along x over the
ranks of one NVLink domain.  Reference chain: montecosmo/model.py:683-837 (evolve), 840-933 (likelihood), with the
callables of nbody.py cited in model.py / dist.py.

Each rank holds the x-planes [x0, x0 + xl) of the white field and of the observed mesh, and the particles whose lattice
site lies there (dist.py).  Exchange steps on top of SlabPM's: one distributed rFFT of the white field, one inverse for
the bias weights, `interlace_order` halo-reduced paints + one batched distributed rFFT + one inverse for the final mesh,
and an all-reduce of the scalar log-density.  The reverse sweep is hand-chained (no autograd tape across ranks): each
stage is the transpose of its forward, with the Hermitian weights w' written out where a half spectrum is a free
variable.

Conventions: the distributed transforms are unnormalised in both directions, so every 1/N is folded into the Fourier
pass or the axpby next to it.  Spectra that feed a C2R are Hermitian-projected on the kz = 0 / Nyquist planes after the
x-transform (SlabPM.irfftn(project=True)), which is what jnp.fft.irfftn does implicitly and cuFFT's 2-D C2R does not.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import cosmo as _cosmo
from .model import linear_power_table
from .nbody import rfftk


class SlabFieldModel:
    def __init__(self, pm, box_size, n_steps=5, a_start=0.0, a_obs=1.0, lpt_order=2, paint_order=2, interlace_order=2,
                 paint_deconv=True, b1=1.0, rsd=True, los=(0.0, 0.0, 1.0), sigma_obs=1.0, cosmology=None, kpow=None):
        self.pm, self.o, self.A, self.lib = pm, pm.o, pm.A, pm.lib
        self.mesh_shape = (pm.nx, pm.ny, pm.nz)
        self.box_size = tuple(float(b) for b in box_size)
        self.n_steps, self.a_start, self.a_obs, self.lpt_order = int(n_steps), a_start, a_obs, int(lpt_order)
        self.paint_order, self.interlace_order, self.paint_deconv = int(paint_order), int(interlace_order), paint_deconv
        self.b1, self.rsd, self.los, self.sigma_obs = float(b1), bool(rsd), tuple(float(x) for x in los), float(sigma_obs)
        self.cosmology = cosmology if cosmology is not None else _cosmo.Cosmology()
        self.kpow = kpow if kpow is not None else linear_power_table(self.cosmology)
        self.transfer = self.A.prepare(self._transfer_block())

    def _transfer_block(self):
        """My ky rows of FieldModel.transfer_mesh (bricks.py:96-106, 150): sqrt(P(k) N / V), [nx, kyl, nzc]."""
        pm = self.pm
        ks, pows = self.kpow
        kx, ky, kz = rfftk(self.mesh_shape, self.box_size)
        ky = ky[:, pm.y0:pm.y0 + pm.kyl, :]
        kmesh = np.sqrt(kx**2 + ky**2 + kz**2)
        pmesh = np.interp(kmesh.reshape(-1), ks, pows, left=0.0, right=0.0).reshape(kmesh.shape)
        return np.sqrt(pmesh * (np.prod(self.mesh_shape) / np.prod(self.box_size))).astype(np.float32)

    # ------------------------------------------------------------------------------------------------ pieces
    def _call(self, name, *args):
        self.pm._call(name, *args)

    def _paint_ext(self, pos, weights, shift):
        """Weighted paint of my particles into a fresh halo-extended mesh, halos reduced; returns the owned planes."""
        pm, A, st = self.pm, self.A, self.pm._st()
        rho = A.zeros((pm.ext, pm.ny, pm.nz))
        one = (C.c_float * 3)(1.0, 1.0, 1.0)
        wp = 0 if weights is None else weights.data_ptr()
        if not (self.paint_order == 2 and pm.brick and self.lib.mcpm_paint_brick(
                st, pm.xl, pm.ny, pm.nz, pos.data_ptr(), wp, 1.0, float(shift), pos.shape[0], pm.ext, pm.ny, pm.nz,
                rho.data_ptr()) == 0):
            self._call("mcpm_paint", st, pos.data_ptr(), wp, 1.0, pos.shape[0], pm.ext, pm.ny, pm.nz, self.paint_order,
                       one, float(shift), rho.data_ptr(), 1)
        pm.halo_reduce(rho)
        return rho[pm.H:pm.H + pm.xl]

    def _deconv_order(self):
        return self.paint_order if self.paint_deconv else 0

    # ------------------------------------------------------------------------------------------------ evaluation
    def value_and_force(self, white, obs):
        """(logpdf, d logpdf / d white) for my planes of the white field [xl, ny, nz] and of the observed mesh; the
        log-density is the global value (all-reduced), identical on every rank.  Same quantity as
        FieldModel.value_and_force (model.py:350-363 wrappers)."""
        pm, o, A, st = self.pm, self.o, self.A, self.pm._st()
        c = self.cosmology
        N, cells, m = pm.N, pm.xl * pm.ny * pm.nz, self.interlace_order
        white, obs = A.prepare(white), A.prepare(obs)
        D = float(_cosmo.a2g(c, self.a_obs))
        # prior: delta_k = rfftn(white) * transfer (bricks.py:303-309, 152-157)
        dk = o.scale_spectrum(pm.rfftn(white.unsqueeze(0))[0], self.transfer)
        # linear Lagrangian bias weights 1 + b1 D delta_L(q) (bricks.py:358-362): the lattice is the mesh, NGP read = value
        weights = None
        if self.b1 != 0.0:
            dl = pm.irfftn(dk.unsqueeze(0), project=True)[0]
            weights = o.axpby(dl.reshape(-1), self.b1 * D / N, None, 0.0, 1.0)
        pos, vel, tape = pm.nbody_forward(dk, c, self.a_start, self.a_obs, self.n_steps, self.lpt_order)
        coef = float(_cosmo.a2g(c, self.a_obs) * _cosmo.a2f(c, self.a_obs))
        if self.rsd:  # bricks.py:781-792 in cell units
            pos = o.rsd_shift(pos, vel, self.los, coef)
            pm._guard(pos)
            pm.check_guard()
        # interlaced, deconvolved final paint (nbody.py:513-577; Jacobian 1: particles == cells) and back to real space
        painted = torch.stack([self._paint_ext(pos, weights, i / m) for i in range(m)])
        pk = pm.rfftn(painted)
        gk = A.empty((pm.nx, pm.kyl, pm.nzc), "c64")
        self._call("mcpm_interlace_combine_slab", st, pk.data_ptr(), gk.data_ptr(), m, pm.nx, pm.ny, pm.nz, pm.kyl,
                   pm.y0, 1.0 / N, self._deconv_order())
        gxy = pm.irfftn(gk.unsqueeze(0), overwrite=True, project=True)[0]
        # Gaussian likelihood + N(0,1) prior (model.py:840-933 restricted to a constant noise level)
        inv_var = 1.0 / self.sigma_obs**2
        r = o.axpby(gxy, 1.0, obs, -1.0)
        lp = o.dot(r, r).reshape(()) * (-0.5 * inv_var) + o.dot(white, white).reshape(()) * -0.5
        if pm.P > 1:
            dist.all_reduce(lp, group=pm.group)

        # ---- reverse sweep -------------------------------------------------------------------------------------------
        gbar = o.axpby(r, -inv_var)
        gkb = pm.rfftn(gbar.unsqueeze(0))[0]
        mk = A.empty((m, pm.nx, pm.kyl, pm.nzc), "c64")
        self._call("mcpm_interlace_combine_T_slab", st, gkb.data_ptr(), mk.data_ptr(), m, pm.nx, pm.ny, pm.nz, pm.kyl,
                   pm.y0, 1.0 / N, self._deconv_order(), 0, 1.0)
        mbar = pm.irfftn(mk, overwrite=True, project=True)  # [m, xl, ny, nz] cotangents of the painted meshes
        n = pos.shape[0]
        posbar = A.empty((n, 3))
        wbar = A.empty((n,)) if weights is not None else None
        one = (C.c_float * 3)(1.0, 1.0, 1.0)
        for i in range(m):
            ext = A.empty((pm.ext, pm.ny, pm.nz))
            ext[pm.H:pm.H + pm.xl] = mbar[i]
            pm.halo_gather(ext)
            self._call("mcpm_paint_vjp", st, pos.data_ptr(), 0 if weights is None else weights.data_ptr(), 1.0,
                       ext.data_ptr(), n, pm.ext, pm.ny, pm.nz, self.paint_order, one, float(i / m), posbar.data_ptr(),
                       0 if wbar is None else wbar.data_ptr(), int(i > 0))
        velbar = o.rsd_shift_vjp(posbar, self.los, coef) if self.rsd else A.zeros((n, 3))
        dkbar = pm.nbody_backward(tape, posbar, velbar)
        if weights is not None:  # dl = C2R(dk): its transpose is R2C times the Hermitian weights
            dlbar = o.axpby(wbar, self.b1 * D / N).reshape(1, pm.xl, pm.ny, pm.nz)
            dlk = pm.rfftn(dlbar)[0]
            self._call("mcpm_half_weight_axpy", st, dlk.data_ptr(), dkbar.data_ptr(), pm.nx * pm.kyl * pm.nzc, pm.nz, 1.0,
                       0, 1)
        wkbar = o.scale_spectrum(dkbar, self.transfer)
        self._call("mcpm_half_weight_axpy", st, wkbar.data_ptr(), wkbar.data_ptr(), pm.nx * pm.kyl * pm.nzc, pm.nz, 1.0, 1,
                   0)
        whitebar = pm.irfftn(wkbar.unsqueeze(0), overwrite=True, project=True)[0]
        return lp, o.axpby(whitebar, 1.0, white, -1.0)

    def force(self, white, obs):
        return self.value_and_force(white, obs)[1]

    def predict(self, white):
        """My planes of the 1 + delta_obs mesh (FieldModel.evolve)."""
        pm, o, A, st = self.pm, self.o, self.A, self.pm._st()
        c, N, m = self.cosmology, self.pm.N, self.interlace_order
        white = A.prepare(white)
        dk = o.scale_spectrum(pm.rfftn(white.unsqueeze(0))[0], self.transfer)
        weights = None
        if self.b1 != 0.0:
            dl = pm.irfftn(dk.unsqueeze(0), project=True)[0]
            weights = o.axpby(dl.reshape(-1), self.b1 * float(_cosmo.a2g(c, self.a_obs)) / N, None, 0.0, 1.0)
        pos, vel, _ = pm.nbody_forward(dk, c, self.a_start, self.a_obs, self.n_steps, self.lpt_order)
        if self.rsd:
            pos = o.rsd_shift(pos, vel, self.los, float(_cosmo.a2g(c, self.a_obs) * _cosmo.a2f(c, self.a_obs)))
            pm._guard(pos)
            pm.check_guard()
        pk = pm.rfftn(torch.stack([self._paint_ext(pos, weights, i / m) for i in range(m)]))
        gk = A.empty((pm.nx, pm.kyl, pm.nzc), "c64")
        self._call("mcpm_interlace_combine_slab", st, pk.data_ptr(), gk.data_ptr(), m, pm.nx, pm.ny, pm.nz, pm.kyl,
                   pm.y0, 1.0 / N, self._deconv_order())
        return pm.irfftn(gk.unsqueeze(0), overwrite=True, project=True)[0]
