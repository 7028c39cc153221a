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
from .nbody import rfftk, scale_shape


class SlabFieldModel:
    def __init__(self, pm, box_size, evolution="nbody", n_steps=5, a_start=0.0, a_obs=1.0, lpt_order=2, paint_order=2, interlace_order=2,
                 paint_deconv=True, paint_oversamp=1.0, b1=1.0, rsd=True, los=(0.0, 0.0, 1.0), sigma_obs=1.0,
                 cosmology=None, kpow=None):
        if evolution not in ("lpt", "nbody"):
            raise ValueError(f"unknown evolution {evolution}")
        self.evolution = evolution
        if int(paint_order) != 2:
            # model.evolve hands the same paint_order to nbody_bf (model.py:771) and SlabPM's step kernels are CIC
            raise NotImplementedError("the slab-decomposed step loop is CIC only (paint_order = 2)")
        self.pm, self.o, self.A, self.lib = pm, pm.o, pm.A, pm.lib
        self.mesh_shape = (pm.nx, pm.ny, pm.nz)
        self.box_size = tuple(float(b) for b in box_size)
        self.n_steps, self.a_start, self.a_obs, self.lpt_order = int(n_steps), a_start, a_obs, int(lpt_order)
        self.paint_order, self.interlace_order, self.paint_deconv = int(paint_order), int(interlace_order), paint_deconv
        self.b1, self.rsd, self.los, self.sigma_obs = float(b1), bool(rsd), tuple(float(x) for x in los), float(sigma_obs)
        self.cosmology = cosmology if cosmology is not None else _cosmo.Cosmology()
        self.kpow = kpow if kpow is not None else linear_power_table(self.cosmology)
        self.transfer = self.A.prepare(self._transfer_block())
        # a paint mesh finer than the evolution mesh (model.py:802-809 with paint_oversamp, BASELINE C5): a second slab
        # geometry on the same ranks for the paints and their transforms, and the distributed Fourier crop back
        self.paint_shape = scale_shape(self.mesh_shape, paint_oversamp)
        self.pm2, self.resize, self.ratio = pm, None, (1.0, 1.0, 1.0)
        self.frame2 = pm.frame  # my lattice in the cells of the paint mesh
        if self.paint_shape != self.mesh_shape:
            from .dist import SlabPM, SlabResize
            self.ratio = tuple(p / m for p, m in zip(self.paint_shape, self.mesh_shape))
            h2 = self.ratio[0] * pm.H
            if abs(h2 - round(h2)) > 1e-9 or self.paint_shape[0] % pm.P or self.paint_shape[1] % pm.P:
                raise ValueError("paint mesh: nx, ny must divide over the ranks and halo * paint/mesh ratio must be whole")
            self.pm2 = SlabPM(pm.o, self.paint_shape, halo=int(round(h2)), group=pm.group, p2p=False)
            self.resize = SlabResize(self.pm2, pm)
            from ._capi import frame as _frame
            self.frame2 = _frame((pm.xl, pm.ny, pm.nz), span=(self.pm2.xl, self.pm2.ny, self.pm2.nz),
                                 origin=(self.pm2.H, 0, 0))

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
        """Weighted paint of my particles into a fresh halo-extended PAINT mesh (positions scaled by the paint / mesh
        ratio in-kernel, nbody.py:569), halos reduced; returns the owned planes."""
        pm, p2, A, st = self.pm, self.pm2, self.A, self.pm._st()
        rho = A.zeros((p2.ext, p2.ny, p2.nz))
        sc = (C.c_float * 3)(*self.ratio)
        wp = 0 if weights is None else weights.data_ptr()
        fr = C.byref(self.frame2)
        if not (self.resize is None and self.paint_order == 2 and pm.brick and pm._brick_ok(self.lib.mcpm_paint_brick_f(
                st, fr, pm.xl, pm.ny, pm.nz, pos.data_ptr(), wp, 1.0, float(shift), pos.shape[0], pm.ext, pm.ny, pm.nz,
                rho.data_ptr()))):
            self._call("mcpm_paint_f", st, fr, pos.data_ptr(), wp, 1.0, pos.shape[0], p2.ext, p2.ny, p2.nz,
                       self.paint_order, sc, float(shift), rho.data_ptr(), 1)
        p2.halo_reduce(rho)
        return rho[p2.H:p2.H + p2.xl]

    def _final_mesh(self, pos, weights):
        """Interlaced, deconvolved final paint (nbody.py:513-577) and back to real space: my planes of 1 + delta_obs at
        the evolution shape (particles == cells: Jacobian 1 overall, model.py:806)."""
        pm, p2, A, st, m = self.pm, self.pm2, self.A, self.pm._st(), self.interlace_order
        pk = p2.rfftn(torch.stack([self._paint_ext(pos, weights, i / m) for i in range(m)]))
        gk = A.empty((p2.nx, p2.kyl, p2.nzc), "c64")
        jac = float(np.prod(self.ratio))
        self._call("mcpm_interlace_combine_slab", st, pk.data_ptr(), gk.data_ptr(), m, p2.nx, p2.ny, p2.nz, p2.kyl,
                   p2.y0, jac / pm.N, self._deconv_order())
        if self.resize is not None:
            gk = self.resize.forward(gk)
        return pm.irfftn(gk.unsqueeze(0), overwrite=True, project=True)[0]

    def _final_mesh_vjp(self, pos, weights, gbar):
        """Transpose of _final_mesh: cotangent of my planes of the final mesh -> (posbar, weightsbar)."""
        pm, p2, A, st, m = self.pm, self.pm2, self.A, self.pm._st(), self.interlace_order
        gkb = pm.rfftn(gbar.unsqueeze(0))[0]
        jac = float(np.prod(self.ratio))
        hw = 0
        if self.resize is not None:  # cotangent of the free final spectrum (C2R^T = R2C x w'), through the crop's transpose
            self._call("mcpm_half_weight_axpy", st, gkb.data_ptr(), gkb.data_ptr(), pm.nx * pm.kyl * pm.nzc, pm.nz, 1.0, 0,
                       0)
            gkb = self.resize.backward(gkb)
            hw = 1
        mk = A.empty((m, p2.nx, p2.kyl, p2.nzc), "c64")
        self._call("mcpm_interlace_combine_T_slab", st, gkb.data_ptr(), mk.data_ptr(), m, p2.nx, p2.ny, p2.nz, p2.kyl,
                   p2.y0, jac / pm.N, self._deconv_order(), hw, 1.0)
        mbar = p2.irfftn(mk, overwrite=True, project=True)  # [m, xl2, ny2, nz2] cotangents of the painted meshes
        n = pos.shape[0]
        posbar = A.empty((n, 3))
        wbar = A.empty((n,)) if weights is not None else None
        sc = (C.c_float * 3)(*self.ratio)
        for i in range(m):
            ext = A.empty((p2.ext, p2.ny, p2.nz))
            ext[p2.H:p2.H + p2.xl] = mbar[i]
            p2.halo_gather(ext)
            self._call("mcpm_paint_vjp_f", st, C.byref(self.frame2), pos.data_ptr(),
                       0 if weights is None else weights.data_ptr(), 1.0, ext.data_ptr(), n, p2.ext, p2.ny, p2.nz,
                       self.paint_order, sc, float(i / m), posbar.data_ptr(), 0 if wbar is None else wbar.data_ptr(),
                       int(i > 0))
        return posbar, wbar

    def _deconv_order(self):
        return self.paint_order if self.paint_deconv else 0

    def _evolve(self, dk):
        """lpt at a_obs (model.py:763) or lpt at a_start + BullFrog steps (model.py:771-773): local (pos, vel, tape)."""
        c = self.cosmology
        if self.evolution == "lpt":
            d1, d2, dv2 = (float(f(c, self.a_obs)) for f in (_cosmo.a2g, _cosmo.a2g2, _cosmo.a2dg2dg))
            pos, vel, ltape = self.pm.lpt_forward(dk, d1, d2, dv2, self.lpt_order)
            self.pm._guard(pos)
            self.pm.check_guard()
            return pos, vel, ("lpt", ltape)
        pos, vel, tape = self.pm.nbody_forward(dk, c, self.a_start, self.a_obs, self.n_steps, self.lpt_order)
        return pos, vel, ("nbody", tape)

    def _evolve_vjp(self, tape, posbar, velbar):
        kind, t = tape
        return self.pm.lpt_backward(t, posbar, velbar) if kind == "lpt" else self.pm.nbody_backward(t, posbar, velbar)

    # ------------------------------------------------------------------------------------------------ evaluation
    def value_and_force(self, white, obs):
        """(logpdf, d logpdf / d white) for my planes of the white field [xl, ny, nz] and of the observed mesh; the
        log-density is the global value (all-reduced), identical on every rank.  Same quantity as
        FieldModel.value_and_force (model.py:350-363 wrappers)."""
        pm, o, A, st = self.pm, self.o, self.A, self.pm._st()
        c = self.cosmology
        N = pm.N
        white, obs = A.prepare(white), A.prepare(obs)
        D = float(_cosmo.a2g(c, self.a_obs))
        # prior: delta_k = rfftn(white) * transfer (bricks.py:303-309, 152-157)
        dk = o.scale_spectrum(pm.rfftn(white.unsqueeze(0))[0], self.transfer)
        # linear Lagrangian bias weights 1 + b1 D delta_L(q) (bricks.py:358-362): the lattice is the mesh, NGP read = value
        weights = None
        if self.b1 != 0.0:
            dl = pm.irfftn(dk.unsqueeze(0), project=True)[0]
            weights = o.axpby(dl.reshape(-1), self.b1 * D / N, None, 0.0, 1.0)
        pos, vel, tape = self._evolve(dk)
        coef = float(_cosmo.a2g(c, self.a_obs) * _cosmo.a2f(c, self.a_obs))
        if self.rsd:  # bricks.py:781-792 in cell units
            pos = o.rsd_shift(pos, vel, self.los, coef)
            pm._guard(pos)
            pm.check_guard()
        gxy = self._final_mesh(pos, weights)
        # Gaussian likelihood + N(0,1) prior (model.py:840-933 restricted to a constant noise level)
        inv_var = 1.0 / self.sigma_obs**2
        r = o.axpby(gxy, 1.0, obs, -1.0)
        lp = o.dot(r, r).reshape(()) * (-0.5 * inv_var) + o.dot(white, white).reshape(()) * -0.5
        if pm.P > 1:
            dist.all_reduce(lp, group=pm.group)

        # ---- reverse sweep -------------------------------------------------------------------------------------------
        gbar = o.axpby(r, -inv_var)
        posbar, wbar = self._final_mesh_vjp(pos, weights, gbar)
        n = pos.shape[0]
        velbar = o.rsd_shift_vjp(posbar, self.los, coef) if self.rsd else A.zeros((n, 3))
        dkbar = self._evolve_vjp(tape, posbar, velbar)
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
        pm, o, A = self.pm, self.o, self.A
        c, N = self.cosmology, self.pm.N
        white = A.prepare(white)
        dk = o.scale_spectrum(pm.rfftn(white.unsqueeze(0))[0], self.transfer)
        weights = None
        if self.b1 != 0.0:
            dl = pm.irfftn(dk.unsqueeze(0), project=True)[0]
            weights = o.axpby(dl.reshape(-1), self.b1 * float(_cosmo.a2g(c, self.a_obs)) / N, None, 0.0, 1.0)
        pos, vel, _ = self._evolve(dk)
        if self.rsd:
            pos = o.rsd_shift(pos, vel, self.los, float(_cosmo.a2g(c, self.a_obs) * _cosmo.a2f(c, self.a_obs)))
            pm._guard(pos)
            pm.check_guard()
        return self._final_mesh(pos, weights)
