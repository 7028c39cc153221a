"""
Drop-in mirror of the hot-path callables of `montecosmo/nbody.py` on the B200 engine.

Same names, argument meaning and error behaviour as the reference (file:line cited per function); arrays are torch
CUDA tensors (float32 / complex64) instead of jax arrays, and `torch.autograd.Function` plays the role of
`jax.custom_vjp`: `torch.autograd.grad` flows through `paint`, `read`, `nufft`, `pm_forces`, `lpt` and `nbody_bf` via the
engine's hand-written adjoints (mcpm_*_vjp).  All arithmetic on arrays happens in libmcpm.so; torch only owns memory,
streams and the autograd tape.  There is no CPU path: importing this module without the built library or without a
CUDA device raises.

Host-side helpers that return small NumPy kernel arrays (`rfftk`, `*_hat`) are kept for API parity; the engine
recomputes them in-kernel and never reads such arrays.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, cosmo as _cosmo
from .ops import Ops, TorchCudaAdapter, ch2rshape, r2chshape

INF = float("inf")
_OPS = None


def ops() -> Ops:
    """The process-wide operator table (loads libmcpm.so; raises without it or without CUDA)."""
    global _OPS
    if _OPS is None:
        _OPS = Ops(_lib.load(), TorchCudaAdapter())
    return _OPS


def _f32(x):
    return ops().A.prepare(x, "f32")


def _c64(x):
    return ops().A.prepare(x, "c64")


def _kb_kcut(kernel_type, oversamp):
    """Window family selector for the engine (nbody.py:319-324, 381-386): 0.0 for 'rectangular', the Kaiser-Bessel
    cutoff optim_kcut(oversamp) for 'kaiser_bessel' (the mcpm_*_kb entry points)."""
    if kernel_type == "rectangular":
        return 0.0
    if kernel_type == "kaiser_bessel":
        return float(0.98 * np.pi * (2 - 1 / oversamp))
    raise ValueError(f"Unknown kernel type: {kernel_type}")


# ----------------------------------------------------------------------------------------------------------------
# host-side kernel helpers (nbody.py:50-334), NumPy, for API parity
# ----------------------------------------------------------------------------------------------------------------
def scale_shape(shape, scale=1.0):
    """utils.py:1163-1168."""
    return tuple(int(o) for o in 2 * np.rint(np.multiply(shape, scale) / 2).astype(int))


def rfftk(shape, box_size=None):
    """nbody.py:50-77."""
    dim = len(shape)
    scales = dim * (2 * np.pi,) if box_size is None else tuple(2 * np.pi * s / b for s, b in zip(shape, box_size))
    out = []
    for ax, (s, sc) in enumerate(zip(shape, scales)):
        k = (np.fft.fftfreq(s) if ax < dim - 1 else np.fft.rfftfreq(s)) * sc
        shp = [1] * dim
        shp[ax] = -1
        out.append(k.reshape(shp))
    return tuple(out)


def fftk(shape, box_size=None):
    """Wavevectors for fftn (nbody.py:78-103)."""
    dim = len(shape)
    scales = dim * (2 * np.pi,) if box_size is None else tuple(2 * np.pi * s / b for s, b in zip(shape, box_size))
    out = []
    for ax, (s, sc) in enumerate(zip(shape, scales)):
        shp = [1] * dim
        shp[ax] = -1
        out.append((np.fft.fftfreq(s) * sc).reshape(shp))
    return tuple(out)


def top_hat(kvec, kcut=np.inf):
    """Sharp low-pass: True where |k| < kcut (nbody.py:191-217)."""
    if kcut == np.inf:
        return 1.0
    return sum(k**2 for k in kvec) < kcut**2


def invlaplace_hat(kvec, fd_order=np.inf):
    """nbody.py:109-133."""
    if fd_order == 2:
        kk = sum((np.cos(k) - 1) * 2 for k in kvec)
    elif fd_order == 4:
        kk = sum((np.cos(2 * k) - 16 * np.cos(k) + 15) / 6 for k in kvec)
    elif fd_order == np.inf:
        kk = sum(k**2 for k in kvec)
    else:
        raise ValueError("Only orders 2, 4, and inf are supported.")
    nz = np.where(kk == 0, 1, kk)
    return -np.where(kk == 0, 0, 1 / nz)


def gradient_hat(kvec, direction, fd_order=np.inf):
    """nbody.py:136-163."""
    k = kvec[direction]
    if fd_order == 2:
        k = np.sin(k)
    elif fd_order == 4:
        k = (8 * np.sin(k) - np.sin(2 * k)) / 6
    elif fd_order != np.inf:
        raise ValueError("Only orders 2, 4, and inf are supported.")
    return 1j * k


def gaussian_hat(kvec, kcut=np.inf):
    """nbody.py:166-188."""
    if kcut == np.inf:
        return 1.0
    return np.exp(-sum(k**2 for k in kvec) * (2 * np.pi / kcut) ** 2 / 2)


def rectangular_hat(kvec, order=2):
    """nbody.py:249-277."""
    out = 1.0
    for k in kvec:
        out = out * np.sinc(k / (2 * np.pi)) ** order
    return out


# ----------------------------------------------------------------------------------------------------------------
# autograd wrappers over the C ABI
# ----------------------------------------------------------------------------------------------------------------
def rectangular(s, order):
    """1-D mass-assignment window on |s| (nbody.py:220-246): 1 NGP, 2 CIC, 3 TSC, 4 PCS.  Host-side NumPy, like the
    reference's kernel helpers; the engine evaluates the same polynomials in window.h."""
    s = np.abs(np.asarray(s, dtype=np.float64))
    if order == 0:
        return np.full(np.shape(s)[-1:], np.inf)
    if order == 1:
        return np.full(np.shape(s)[-1:], 1.0)
    if order == 2:
        return 1 - s
    if order == 3:
        return (s <= 0.5) * (0.75 - s**2) + (0.5 < s) / 2 * (1.5 - s) ** 2
    if order == 4:
        return (s <= 1) / 6 * (4 - 6 * s**2 + 3 * s**3) + (1 < s) / 6 * (2 - s) ** 3
    raise IndexError("order must be in 0..4")


def optim_kcut(oversamp, safety=0.98):
    """Optimal wavenumber cutoff of the Kaiser-Bessel window (nbody.py:357-363)."""
    return safety * np.pi * (2 - 1 / oversamp)


def kaiser_bessel(s, order, kcut):
    """Kaiser-Bessel window (nbody.py:280-290), host-side NumPy; the engine evaluates it in window.h (KbWin)."""
    s = np.asarray(s, dtype=np.float64) * 2 / order
    kcut = kcut * order / 2
    return np.i0(kcut * (1 - s**2) ** 0.5) / (order * np.sinh(kcut) / kcut)


def kaiser_bessel_hat(kvec, order, kcut):
    """Fourier transform of the Kaiser-Bessel window (nbody.py:293-312), host-side NumPy."""
    def kernel(k, kc):
        k = k * order / 2
        kc = kc * order / 2
        dist = np.abs(kc**2 - k**2) ** 0.5
        with np.errstate(divide="ignore", invalid="ignore"):
            out = np.where(np.abs(k) <= kc, np.sinh(dist) / dist, np.sin(dist) / dist)
        return out / (np.sinh(kc) / kc)
    out = 1.0
    for k in kvec:
        out = out * kernel(k, kcut)
    return out


class _Paint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, weights, shape, wscalar, order, scale, shift, kb=0.0):
        ctx.save_for_backward(pos, weights)
        ctx.cfg = (shape, wscalar, order, scale, shift, kb)
        return ops().paint(pos, shape, weights, wscalar, order, scale, shift, kb_kcut=kb)

    @staticmethod
    def backward(ctx, mbar):
        pos, weights = ctx.saved_tensors
        shape, wscalar, order, scale, shift, kb = ctx.cfg
        need_p, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1] and weights is not None
        pb, wb = ops().paint_vjp(pos, mbar.contiguous(), weights, wscalar, order, scale, shift, need_p, need_w,
                                 kb_kcut=kb)
        return pb, wb, None, None, None, None, None, None


class _Read(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, mesh, order, scale, shift, kb=0.0):
        ctx.save_for_backward(pos, mesh)
        ctx.cfg = (order, scale, shift, kb)
        return ops().read(pos, mesh, order, scale, shift, kb_kcut=kb)

    @staticmethod
    def backward(ctx, obar):
        pos, mesh = ctx.saved_tensors
        order, scale, shift, kb = ctx.cfg
        obar = obar.contiguous()
        pb = mb = None
        if ctx.needs_input_grad[0]:
            ob = obar.reshape(pos.shape[0], -1)
            if mesh.dim() == 4 and mesh.shape[0] > 4:  # mcpm_read_grad takes up to 4 meshes: chunks, summed
                pb = sum(ops().read_grad(pos, mesh[i:i + 4], ob[:, i:i + 4].contiguous(), order, scale, shift, kb_kcut=kb)
                         for i in range(0, mesh.shape[0], 4))
            else:
                pb = ops().read_grad(pos, mesh, ob, order, scale, shift, kb_kcut=kb)
        if ctx.needs_input_grad[1]:
            if mesh.dim() == 3:
                mb = ops().paint(pos, tuple(mesh.shape), obar, 1.0, order, scale, shift, kb_kcut=kb)
            else:
                mb = torch.stack([ops().paint(pos, tuple(mesh.shape[1:]), obar[:, i].contiguous(), 1.0, order, scale,
                                              shift, kb_kcut=kb) for i in range(mesh.shape[0])])
        return pb, mb, None, None, None, None


class _Rfftn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mesh):
        return ops().rfftn(mesh)

    @staticmethod
    def backward(ctx, kbar):
        return ops().irfftn(ops().hermitian_weights(kbar.contiguous(), 0), overwrite=True)


class _Irfftn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, meshk):
        return ops().irfftn(meshk)

    @staticmethod
    def backward(ctx, mbar):
        return ops().hermitian_weights(ops().rfftn(mbar.contiguous()), 1)


class _Deconv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, meshk, order, kb=0.0):
        ctx.cfg = (order, kb)
        return ops().deconv(meshk, order, kb_kcut=kb)

    @staticmethod
    def backward(ctx, kbar):
        return ops().deconv(kbar.contiguous(), ctx.cfg[0], kb_kcut=ctx.cfg[1]), None, None


class _Chreshape(torch.autograd.Function):
    @staticmethod
    def forward(ctx, meshk, cshape):
        ctx.in_cshape = tuple(meshk.shape)
        return ops().chreshape(meshk, cshape)

    @staticmethod
    def backward(ctx, kbar):
        return ops().chreshape_vjp(kbar.contiguous(), ctx.in_cshape), None


class _ScaleSpectrum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, meshk, transfer):
        ctx.save_for_backward(transfer)
        return ops().scale_spectrum(meshk, transfer)

    @staticmethod
    def backward(ctx, kbar):
        (transfer,) = ctx.saved_tensors
        return ops().scale_spectrum(kbar.contiguous(), transfer), None


class _NufftPaint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, weights, paint_shape, wscalar, scale, paint_order, interlace_order, paint_deconv, kb=0.0,
                lattice=None, vel=None, los=None, coef=0.0):
        ctx.save_for_backward(pos, weights, vel)
        ctx.cfg = (paint_shape, wscalar, scale, paint_order, interlace_order, paint_deconv, kb, lattice, los, coef)
        return ops().nufft_paint(pos, paint_shape, weights, wscalar, scale, paint_order, interlace_order, paint_deconv,
                                 kb_kcut=kb, lattice=lattice, rsd=None if vel is None else (vel, los, coef))

    @staticmethod
    def backward(ctx, kbar):
        pos, weights, vel = ctx.saved_tensors
        paint_shape, wscalar, scale, paint_order, interlace_order, paint_deconv, kb, lattice, los, coef = ctx.cfg
        need_p, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1] and weights is not None
        if vel is None:
            pb, wb = ops().nufft_paint_vjp(pos, kbar.contiguous(), paint_shape, weights, wscalar, scale, paint_order,
                                           interlace_order, paint_deconv, need_p, need_w, kb_kcut=kb, lattice=lattice)
            vb = None
        else:  # the shift's transpose rides in the same gather: velbar = coef (xbar . los) los
            pb, wb, vb = ops().nufft_paint_vjp(pos, kbar.contiguous(), paint_shape, weights, wscalar, scale, paint_order,
                                               interlace_order, paint_deconv, need_p, need_w, kb_kcut=kb, lattice=lattice,
                                               rsd=(vel, los, coef))
        return pb, wb, None, None, None, None, None, None, None, None, vb, None, None


class _NufftObserved(torch.autograd.Function):
    """nufft of the observed positions: the general observation chain inside the paint (mcpm_nufft_obs).  Differentiable
    in the positions, velocities, velocity bias, weights, the scalars (D f at a_obs, Alcock-Paczynski factors) and the
    nodes of the two radius tables -- through which the cosmology's gradient flows back to the caller's tables."""

    @staticmethod
    def forward(ctx, pos, vel, dvel, weights, par, tab_gf, tab_ap, static, paint_shape, wscalar, scale, paint_order,
                interlace_order, paint_deconv, kb, lattice):
        obs = dict(static)
        obs["gf"], obs["a_par"], obs["a_perp"] = (float(v) for v in par.detach().cpu())
        obs["tab_gf"] = None if tab_gf is None else _f32(tab_gf.detach())
        obs["tab_ap"] = None if tab_ap is None else _f32(tab_ap.detach())
        obs["dvel"] = dvel
        ctx.save_for_backward(pos, vel, dvel, weights, obs["tab_gf"], obs["tab_ap"])
        ctx.obs = {k: v for k, v in obs.items() if k not in ("tab_gf", "tab_ap", "dvel")}
        ctx.cfg = (paint_shape, wscalar, scale, paint_order, interlace_order, paint_deconv, kb, lattice)
        ctx.like = (par, tab_gf, tab_ap)
        return ops().nufft_observed(pos, vel, obs, paint_shape, weights, wscalar, scale, paint_order, interlace_order,
                                    paint_deconv, kb, lattice)

    @staticmethod
    def backward(ctx, kbar):
        pos, vel, dvel, weights, tg, ta = ctx.saved_tensors
        obs = dict(ctx.obs, tab_gf=tg, tab_ap=ta, dvel=dvel)
        par, tab_gf, tab_ap = ctx.like
        want_par = any(ctx.needs_input_grad[4:7])
        pb, vb, db, wb, parbar = ops().nufft_observed_vjp(pos, vel, obs, kbar.contiguous(), *ctx.cfg[:1], weights,
                                                          *ctx.cfg[1:], want_par=want_par)
        back = lambda g, like: None if (g is None or like is None) else g.to(device=like.device, dtype=like.dtype)
        nt = 0 if tg is None and ta is None else int((tg if tg is not None else ta).shape[0])
        gpar = ggf = gap = None
        if parbar is not None:
            gpar = back(parbar[:3], par)
            ggf = back(parbar[3:3 + nt], tab_gf)
            gap = back(parbar[3 + nt:3 + 2 * nt], tab_ap)
        return (pb, vb, db, wb, gpar, ggf, gap) + (None,) * 9


class _RadialTables(torch.autograd.Function):
    """Functions of the comoving distance at the particles (the light cone): one pass over the positions evaluating every
    table; differentiable in the table nodes (the cosmology) and in the positions."""

    @staticmethod
    def forward(ctx, pos, tabs, geom):
        tabs32 = _f32(tabs.detach())
        ctx.save_for_backward(pos, tabs32)
        ctx.geom, ctx.like = geom, tabs
        return ops().radial_tables(pos, geom, tabs32)

    @staticmethod
    def backward(ctx, outbar):
        pos, tabs32 = ctx.saved_tensors
        pb, tb = ops().radial_tables_vjp(pos, ctx.geom, tabs32, outbar.contiguous(), want_pos=ctx.needs_input_grad[0])
        return pb, tb.to(device=ctx.like.device, dtype=ctx.like.dtype), None


def radial_tables(pos, tabs, geom):
    """[Np, K]: the K tables `tabs` [K, nt] (on the uniform radius grid of `geom`: r0, dr) interpolated at the distance
    of every particle from the observer -- curved sky -- or along the line of sight -- flat (bricks.py:747-766); `geom`
    from bricks.observation / bricks.lightcone_functions."""
    return _RadialTables.apply(_f32(pos), tabs, geom)


class _PmForcesPaint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, shape, order, paint_deconv, lap_fd, grad_fd, kcut):
        forces, fm = ops().pm_forces(pos, shape, order, paint_deconv, lap_fd, grad_fd, kcut, want_meshes=True)
        ctx.save_for_backward(pos, fm)
        ctx.cfg = (order, paint_deconv, lap_fd, grad_fd, kcut)
        return forces

    @staticmethod
    def backward(ctx, fbar):
        pos, fm = ctx.saved_tensors
        order, paint_deconv, lap_fd, grad_fd, kcut = ctx.cfg
        return ops().pm_forces_vjp(pos, fbar.contiguous(), fm, order, paint_deconv, lap_fd, grad_fd, kcut), None, \
            None, None, None, None, None


class _Lpt(torch.autograd.Function):
    """(delta_k, coef[3] float64 host) -> (dpos, vel); VJP to delta_k and the coefficients (nbody.py:634-667)."""

    @staticmethod
    def forward(ctx, dk, coef, pos, lpt_order, read_order, lap_fd, grad_fd, add_pos=False):
        c = [float(v) for v in coef.detach().cpu()]
        dpos, vel, tape = ops().lpt(dk, pos, c[0], c[1], c[2], lpt_order, read_order, lap_fd, grad_fd, tape=True)
        if add_pos:  # return the displaced positions pos + dpos (same cotangent as dpos)
            if pos is None:
                raise ValueError("lpt on the mesh cells (pos=None) returns displacements")
            dpos = ops().axpby(dpos, 1.0, pos, 1.0)
        ctx.cfg = (c, lpt_order, read_order, lap_fd, grad_fd, tuple(dk.shape))
        ctx.pos, ctx.tape = pos, tape
        ctx.coef_meta = (coef.device, coef.dtype)
        return dpos, vel

    @staticmethod
    def backward(ctx, dpb, vlb):
        c, lpt_order, read_order, lap_fd, grad_fd, cshape = ctx.cfg
        like = dpb if dpb is not None else vlb
        dpb = torch.zeros_like(like) if dpb is None else dpb.contiguous()
        vlb = torch.zeros_like(like) if vlb is None else vlb.contiguous()
        want_coef = ctx.needs_input_grad[1]
        out = ops().lpt_vjp(ctx.pos, cshape, c[0], c[1], c[2], dpb, vlb, ctx.tape, lpt_order, read_order, lap_fd,
                            grad_fd, want_coef=want_coef)
        dkbar, cb = out if want_coef else (out, None)
        if cb is not None:
            cb = cb.to(device=ctx.coef_meta[0], dtype=ctx.coef_meta[1])
        ctx.tape = None
        return dkbar, cb, None, None, None, None, None, None


class _NbodySteps(torch.autograd.Function):
    """BullFrog DKD loop (nbody.py:933-951, 999) with per-step coefficients [n_steps, 4] (float64, host)."""

    @staticmethod
    def forward(ctx, pos, vel, coefs, shape, order, paint_deconv, lap_fd, grad_fd, lattice=None, tape_forces=True):
        co = coefs.detach().cpu().to(torch.float64).numpy()
        al, be, pre, post = (co[:, i].tolist() for i in range(4))
        want_coef = coefs.requires_grad
        pos, vel0 = pos.clone(), vel.contiguous()
        vel = vel0.clone()
        tape = ops().nbody_steps(pos, vel, shape, al, be, pre, post, order, paint_deconv, lap_fd, grad_fd, tape=True,
                                 tape_vel=want_coef, lattice=lattice, tape_forces=tape_forces)
        ctx.cfg = (al, be, pre, post, shape, order, paint_deconv, lap_fd, grad_fd, want_coef, lattice)
        ctx.tape, ctx.v0 = tape, (vel0 if want_coef else None)
        ctx.coef_meta = (coefs.device, coefs.dtype)
        return pos, vel

    @staticmethod
    def backward(ctx, pb, vb):
        al, be, pre, post, shape, order, paint_deconv, lap_fd, grad_fd, want_coef, lattice = ctx.cfg
        xk = ctx.tape[0]
        pb = torch.zeros_like(xk[0]) if pb is None else pb.clone().contiguous()
        vb = torch.zeros_like(xk[0]) if vb is None else vb.clone().contiguous()
        cb = ops().nbody_steps_vjp(pb, vb, shape, al, be, pre, post, ctx.tape, order, paint_deconv, lap_fd, grad_fd,
                                   v0=ctx.v0, want_coef=want_coef, lattice=lattice)
        if cb is not None:
            cb = cb.to(device=ctx.coef_meta[0], dtype=ctx.coef_meta[1])
        ctx.tape = ctx.v0 = None
        return pb, vb, cb, None, None, None, None, None, None, None


# ----------------------------------------------------------------------------------------------------------------
# public API, reference names
# ----------------------------------------------------------------------------------------------------------------
def _split_weights(weights):
    """weights scalar -> (None, float); tensor -> (tensor, 1.0)."""
    if isinstance(weights, (int, float)):
        return None, float(weights)
    w = _f32(weights)
    if w.dim() == 0:
        return None, float(w)
    return w, 1.0


def paint(pos, shape: tuple, weights=1.0, order: int = 2, kernel_type="rectangular", oversamp=1.0):
    """Paint the positions onto a mesh of given shape (nbody.py:365-396)."""
    kb = _kb_kcut(kernel_type, oversamp)
    w, ws = _split_weights(weights)
    return _Paint.apply(_f32(pos), w, tuple(int(s) for s in shape), ws, int(order), None, 0.0, kb)


def read(pos, mesh, order: int = 2, kernel_type="rectangular", oversamp=1.0):
    """Read the value at the positions from the mesh (nbody.py:398-427)."""
    kb = _kb_kcut(kernel_type, oversamp)
    return _Read.apply(_f32(pos), _f32(mesh), int(order), None, 0.0, kb)


def rfftn(mesh):
    """jnp.fft.rfftn on the engine's cuFFT plans."""
    return _Rfftn.apply(_f32(mesh))


def irfftn(meshk):
    """jnp.fft.irfftn (last real side = 2*(m-1), utils.py:769-776)."""
    return _Irfftn.apply(_c64(meshk))


def chreshape(mesh, shape):
    """Hermitian- and mean-preserving Fourier crop / pad (utils.py:975-1013); `shape` is the complex target shape."""
    return _Chreshape.apply(_c64(mesh), tuple(int(s) for s in shape))


def deconv_paint(mesh, order: int = 2, kernel_type="rectangular", oversamp=1.0):
    """Deconvolve the mesh by the paint kernel (nbody.py:315-334); real meshes go through rfftn / irfftn."""
    kb = _kb_kcut(kernel_type, oversamp)
    if not torch.is_complex(torch.as_tensor(mesh)):
        return irfftn(_Deconv.apply(rfftn(mesh), int(order), kb))
    return _Deconv.apply(_c64(mesh), int(order), kb)


def interlace(pos, shape: tuple, weights=1.0, paint_order: int = 2, interlace_order: int = 2,
              kernel_type="rectangular", paint_oversamp: float = 1.0):
    """Equal-spacing interlacing (nbody.py:513-529)."""
    kb = _kb_kcut(kernel_type, paint_oversamp)
    w, ws = _split_weights(weights)
    return _NufftPaint.apply(_f32(pos), w, tuple(int(s) for s in shape), ws, None, int(paint_order),
                             int(interlace_order), False, kb)


def nufft(pos, final_shape: tuple, paint_shape=None, weights=1.0, paint_order: int = 2, interlace_order: int = 2,
          kernel_type="rectangular", paint_deconv=True, lattice=None, rsd=None):
    """Non-uniform FFT with oversampling, deconvolution and interlacing (nbody.py:532-577).

    `rsd` (extension) = (vel, los, coef): paint the redshift-space positions pos + (vel . los) * coef * los of the flat-sky
    observer (bricks.py:781-792, what model.py:780-809 hands to nufft) with the shift applied inside the paint kernels --
    one pass over the particles fewer in each direction; differentiable in vel as well.

    `lattice` (extension): `pos` holds displacements (in final_shape cells) from the sites of the regular lattice of that
    shape spanning the mesh (regular_pos, bricks.py:593-603) instead of absolute positions -- what
    nbody_bf(..., relative=True) returns; float32 then resolves the CIC fraction to ~1e-7 cell at any mesh size."""
    final_shape = tuple(int(s) for s in final_shape)
    if paint_shape is None:
        paint_shape, paint_oversamp = final_shape, 1.0
    elif isinstance(paint_shape, float):
        paint_oversamp = paint_shape
        paint_shape = scale_shape(final_shape, paint_oversamp)
    elif isinstance(paint_shape, (tuple, np.ndarray)):
        paint_shape = tuple(int(s) for s in paint_shape)
        paint_oversamp = float(np.exp(np.log(np.divide(final_shape, paint_shape)).mean()))  # as nbody.py:566
    else:
        raise ValueError("paint_shape must be None, a float, or a tuple/ndarray")
    kb = _kb_kcut(kernel_type, paint_oversamp)
    scale = tuple(float(p) / float(f) for p, f in zip(paint_shape, final_shape))
    w, ws = _split_weights(weights)
    vel, los, coef = (None, None, 0.0) if rsd is None else (_f32(rsd[0]), tuple(float(x) for x in rsd[1]), float(rsd[2]))
    mesh = _NufftPaint.apply(_f32(pos), w, paint_shape, ws, scale, int(paint_order), int(interlace_order),
                             bool(paint_deconv), kb, None if lattice is None else tuple(int(s) for s in lattice),
                             vel, los, coef)
    if final_shape != paint_shape:
        mesh = chreshape(mesh, r2chshape(final_shape))
    return mesh


def nufft_observed(pos, vel, final_shape: tuple, obs: dict, paint_shape=None, weights=1.0, dvel=None, paint_order: int = 2,
                   interlace_order: int = 2, kernel_type="rectangular", paint_deconv=True, lattice=None, pos_shape=None):
    """nufft (nbody.py:532-577) of the positions as the observer sees them, the chain model.py:780-799 builds between
    the evolution and the paint -- cell2phys_pos, los_scalefactor_pos, rsd (with the velocity bias `dvel`), ap_auto |
    ap_param, phys2cell_pos -- applied inside the paint kernels: no transformed position array, no elementwise pass.

    `obs` comes from bricks.observation(...): the static geometry plus the differentiable tensors `par` (D f at a_obs,
    a_par, a_perp) and `tab_gf` / `tab_ap` (light-cone growth and ap_auto rescaling on a uniform radius grid).
    `pos` and `vel` are in the cells the geometry was built for: those of `pos_shape` (model.py: evol_shape; default
    final_shape), which phys2cell_pos converts to final_shape cells (model.py:796, init_shape) -- here one scale factor
    of the paint; the paint lands on `paint_shape` and is brought to `final_shape`, as in nufft."""
    final_shape = tuple(int(s) for s in final_shape)
    if paint_shape is None:
        paint_shape, paint_oversamp = final_shape, 1.0
    elif isinstance(paint_shape, float):
        paint_oversamp = paint_shape
        paint_shape = scale_shape(final_shape, paint_oversamp)
    else:
        paint_shape = tuple(int(s) for s in paint_shape)
        paint_oversamp = float(np.exp(np.log(np.divide(final_shape, paint_shape)).mean()))
    kb = _kb_kcut(kernel_type, paint_oversamp)
    pos_shape = final_shape if pos_shape is None else tuple(int(s) for s in pos_shape)
    scale = tuple(float(p) / float(f) for p, f in zip(paint_shape, pos_shape))
    w, ws = _split_weights(weights)
    ws = ws * float(np.divide(pos_shape, final_shape).prod())  # the Jacobian is final -> paint units (nbody.py:571)
    static = {k: v for k, v in obs.items() if k not in ("par", "tab_gf", "tab_ap")}
    mesh = _NufftObserved.apply(_f32(pos), None if vel is None else _f32(vel), None if dvel is None else _f32(dvel), w,
                                obs["par"], obs.get("tab_gf"), obs.get("tab_ap"), static, paint_shape, ws, scale,
                                int(paint_order), int(interlace_order), bool(paint_deconv), kb,
                                None if lattice is None else tuple(int(s) for s in lattice))
    if final_shape != paint_shape:
        mesh = chreshape(mesh, r2chshape(final_shape))
    return mesh


def _coef_tensor(*vals):
    return torch.stack([_cosmo._t(v).reshape(()) for v in vals])


def pm_forces(pos, mesh, read_order: int = 2, paint_deconv: bool = False, grad_fd=np.inf, lap_fd=np.inf, kcut=np.inf):
    """PM forces -grad lap^-1 delta at pos (nbody.py:583-604).  `mesh` is a shape tuple (paint pos) or delta_k."""
    pos = _f32(pos)
    if isinstance(mesh, tuple):
        return _PmForcesPaint.apply(pos, tuple(int(s) for s in mesh), int(read_order), bool(paint_deconv), lap_fd,
                                    grad_fd, kcut)
    if kcut != np.inf:  # spectrum input with a long-range filter: not differentiable here, plain forward
        return ops().pm_forces_mesh(pos, _c64(mesh), int(read_order), lap_fd, grad_fd, kcut)
    # vel of a first-order lpt with unit coefficients is exactly pm_forces(pos, delta_k)
    _, vel = _Lpt.apply(_c64(mesh), _coef_tensor(0.0, 0.0, 0.0), pos, 1, int(read_order), lap_fd, grad_fd)
    return vel


def pm_forces2(pos, mesh, read_order: int = 2, grad_fd=np.inf, lap_fd=np.inf):
    """2LPT source force (nbody.py:607-631)."""
    # dpos = d1 F1 - d2 F2 with d1 = 0, d2 = -1
    dpos, _ = _Lpt.apply(_c64(mesh), _coef_tensor(0.0, -1.0, 0.0), _f32(pos), 2, int(read_order), lap_fd, grad_fd)
    return dpos


def lpt(cosmo, init_mesh, pos, a, lpt_order: int = 2, read_order: int = 2, grad_fd=np.inf, lap_fd=np.inf,
        _displaced=False, growth=None):
    """First or second order LPT displacement at scale factor `a` (nbody.py:634-667).  `a` is a scalar, or one scale
    factor per particle ([Np] or [Np, 1], the light-cone use of nbody.py:651-653): the two force fields then come from
    the engine and the per-particle growth factors multiply them here.  `growth` (extension) = the per-particle
    (a2g, a2g2, a2dg2dg) as [Np] device arrays, e.g. from bricks.lightcone_functions: `a` is then not looked at."""
    if lpt_order not in (1, 2):
        raise ValueError("lpt_order must be 1 or 2")
    init_mesh = torch.as_tensor(init_mesh)
    if not torch.is_complex(init_mesh):
        init_mesh = rfftn(init_mesh)
    if pos is None:  # extension: the particles sit on the cells of the mesh (regular_pos(mesh_shape)); scalar `a` only
        coef = _coef_tensor(_cosmo.a2g(cosmo, a), _cosmo.a2g2(cosmo, a), _cosmo.a2dg2dg(cosmo, a))
        return _Lpt.apply(_c64(init_mesh), coef, None, int(lpt_order), int(read_order), lap_fd, grad_fd, False)
    if growth is not None:
        force1 = pm_forces(pos, init_mesh, read_order, grad_fd=grad_fd, lap_fd=lap_fd)
        g1, g2, dg2 = (_f32(g).reshape(-1, 1) for g in growth)
        dpos, vel = g1 * force1, force1
        if lpt_order == 2:
            force2 = pm_forces2(pos, init_mesh, read_order, grad_fd=grad_fd, lap_fd=lap_fd)
            dpos, vel = dpos - g2 * force2, vel - dg2 * force2
        return (dpos + _f32(pos), vel) if _displaced else (dpos, vel)
    if np.ndim(a) != 0:
        a_col = torch.as_tensor(a, dtype=torch.float64).detach().cpu().reshape(-1, 1)
        force1 = pm_forces(pos, init_mesh, read_order, grad_fd=grad_fd, lap_fd=lap_fd)
        if a_col.shape[0] != force1.shape[0]:
            raise ValueError("a must be a scalar or hold one scale factor per particle")
        col = lambda t: t.to(device=force1.device, dtype=force1.dtype)
        dpos, vel = col(_cosmo.a2g(cosmo, a_col)) * force1, force1
        if lpt_order == 2:
            force2 = pm_forces2(pos, init_mesh, read_order, grad_fd=grad_fd, lap_fd=lap_fd)
            dpos = dpos - col(_cosmo.a2g2(cosmo, a_col)) * force2
            vel = vel - col(_cosmo.a2dg2dg(cosmo, a_col)) * force2
        return (dpos + _f32(pos), vel) if _displaced else (dpos, vel)
    coef = _coef_tensor(_cosmo.a2g(cosmo, a), _cosmo.a2g2(cosmo, a), _cosmo.a2dg2dg(cosmo, a))
    return _Lpt.apply(_c64(init_mesh), coef, _f32(pos), int(lpt_order), int(read_order), lap_fd, grad_fd, _displaced)


def save_y(t, y, args):
    """Default save function of nbody_bf (nbody.py:964-965)."""
    return y


def _save_plan(ts, g0, dg, n_steps):
    """For every save time, the step whose dense output holds it and the position inside it: diffrax 0.5.0 saves
    `SaveAt(ts=...)` from the solver's interpolation, which for Euler is the straight line between the two ends of a
    step.  t in (t_i, t_i+1] belongs to step i (t <= g0 to step 0, t > g1 to the last step, linearly continued)."""
    plan = []
    for t in ts:
        u = (t - g0) / dg
        uf = float(u)
        if abs(uf - round(uf)) < 1e-9 and 0 <= round(uf) <= n_steps:  # on a step boundary: that state itself
            plan.append((max(int(round(uf)) - 1, 0), 1.0 if round(uf) >= 1 else 0.0))
        else:
            i = min(max(int(np.ceil(uf)) - 1, 0), n_steps - 1)
            plan.append((i, u - i))
    return plan


def nbody_bf(cosmo, init_mesh, pos, a0=0.0, a1=1.0, n_steps=5, paint_order: int = 2, lpt_order: int = 2,
             paint_deconv=False, grad_fd=np.inf, lap_fd=np.inf, snapshots=None, fn=save_y, ptcl_shape="auto",
             integrator="bullfrog", relative=None, tape_forces=True):
    """N-body simulation with the BullFrog solver (nbody.py:967-1002): lpt at a0, then n_steps DKD steps in growth time.

    Returns (pos, vel), each [S, Np, 3].  `snapshots` as in the reference: None or an int <= 1 saves the final state
    (S = 1); an int S saves at linspace(g0, g1, S) in growth time; a sequence of scale factors saves at a2g(cosmo, .).
    Save times inside a step get the solver's dense output (linear between the step's ends, as diffrax's Euler), and
    `fn(t, (pos, vel), None)` maps every saved state (any tuple / tensor result is stacked along a new leading axis);
    with snapshots=None the reference ignores `fn`, and so does this.

    `integrator` (extension): "bullfrog" (the live alpha_bf, nbody.py:907-919) or "fastpm" (alpha_fpm, nbody.py:921-931,
    which the reference keeps next to it with the switch commented out): same drift-kick-drift loop, other kick weights.

    `ptcl_shape` (extension) is a performance hint only: the lattice shape of `pos` (regular_pos order).  "auto" assumes
    the mesh shape when the particle count matches it; None disables the brick-tiled kernels.

    `relative` (extension): `pos` must be regular_pos(mesh_shape, ptcl_shape); the loop then carries DISPLACEMENTS from
    those lattice sites instead of absolute positions.  The reference carries the float64 sum q + dpos
    (nbody.py:984-985); a float32 sum resolves only 1.5e-5 cell at x ~ 256, which is what limited grad(log-density)
    parity in round 1 -- see mcpm_engine_set_relative in include/mcpm.h.  True: the call also RETURNS displacements
    (add `pos` to get positions; what FieldModel hands on to nufft(..., lattice=...)).  None (default): displacements
    inside the loop whenever the caller declares the lattice (`ptcl_shape` given as a tuple), positions returned as the
    reference does; with ptcl_shape="auto" nothing is assumed about `pos` and the loop carries absolute positions.
    False: absolute positions throughout (round 1's arithmetic).

    `tape_forces` (extension; memory only, same result): False keeps positions but not the per-step force meshes for the
    backward pass (16 bytes per cell and step) and recomputes them there -- the memory / recompute trade of the
    reference's checkpointed diffrax adjoint (nbody.py:999).
    """
    fn = save_y if fn is None else fn
    n_steps = int(n_steps)
    init_mesh = _c64(init_mesh)
    pos = _f32(pos)
    mesh_shape = ch2rshape(tuple(init_mesh.shape))
    declared = ptcl_shape != "auto" and ptcl_shape is not None
    if ptcl_shape == "auto":
        ptcl_shape = mesh_shape if pos.shape[0] == int(np.prod(mesh_shape)) else None
    ops().set_lattice(mesh_shape, ptcl_shape)
    if relative and ptcl_shape is None:
        raise ValueError("relative=True needs the particle lattice (ptcl_shape)")
    inner_rel = bool(relative) or (relative is None and declared)
    lattice = tuple(int(s) for s in ptcl_shape) if inner_rel else None
    if lattice is not None and int(np.prod(lattice)) != pos.shape[0]:
        raise ValueError("ptcl_shape does not match the number of particles")
    # displacements from a lattice that IS the mesh: lpt's NGP reads at the sites are the cell values (pos=None fast path)
    on_cells = inner_rel and lattice == tuple(mesh_shape)
    x, vel = lpt(cosmo, init_mesh, None if on_cells else pos, a0, lpt_order, 1, grad_fd, lap_fd,
                 _displaced=not inner_rel)
    al, be, pre, post, _, _ = _cosmo.bullfrog_coefficients(cosmo, a0, a1, n_steps, integrator)
    coefs = torch.stack([al, be, pre, post], dim=1)
    g0, g1 = _cosmo.a2g(cosmo, a0), _cosmo.a2g(cosmo, a1)
    dg = (g1 - g0) / n_steps
    if snapshots is None:
        ts, fn = None, save_y
    elif isinstance(snapshots, (int, np.integer)):
        ts = None if snapshots <= 1 else [g0 + (g1 - g0) * (i / (snapshots - 1)) for i in range(int(snapshots))]
    else:
        ts = list(_cosmo.a2g(cosmo, torch.as_tensor(np.asarray(snapshots, dtype=np.float64))).reshape(-1))
    plan = [(n_steps - 1, 1.0)] if ts is None else _save_plan(ts, g0, dg, n_steps)
    # step boundaries whose states are needed, reached segment by segment (one engine call each)
    need = set()
    for i, th in plan:
        need.update((i + 1,) if th == 1.0 and isinstance(th, float) else (i,) if isinstance(th, float) else (i, i + 1))
    states, s = {0: (x, vel)}, 0
    for b in sorted(need):
        if b > s:
            x, vel = _NbodySteps.apply(x, vel, coefs[s:b], mesh_shape, int(paint_order), bool(paint_deconv), lap_fd,
                                       grad_fd, lattice, bool(tape_forces))
            s = b
            states[b] = (x, vel)
    if inner_rel and not relative:  # back to absolute positions at the boundary, as the reference returns them
        states = {b: (xs + pos, vs) for b, (xs, vs) in states.items()}
    outs = []
    for k, (i, th) in enumerate(plan):
        if isinstance(th, float):
            y = states[i + 1] if th == 1.0 else states[i]
        else:
            thf = th.to(device=x.device, dtype=x.dtype)
            y = tuple(a + (b - a) * thf for a, b in zip(states[i], states[i + 1]))
        outs.append(fn(g1 if ts is None else ts[k], y, None))
    if isinstance(outs[0], (tuple, list)):
        return tuple(torch.stack([o[j] for o in outs]) for j in range(len(outs[0])))
    return torch.stack(outs)


def lpt_fpm(cosmo, init_mesh, pos, a, lpt_order: int = 1, paint_order: int = 2, grad_fd=np.inf, lap_fd=np.inf):
    """FastPM-convention LPT (nbody.py:1030-1073): displacement dq = D1 F1 - D2 F2 and momentum p = a^2 E (f1 D1 F1 -
    f2 D2 F2), the initial state of the legacy scale-factor-time solvers.  Composed from the engine's pm_forces /
    pm_forces2 (the reference's inline Hessian sum is pm_forces2's source, nbody.py:615-627)."""
    init_mesh = torch.as_tensor(init_mesh)
    if not torch.is_complex(init_mesh):
        init_mesh = rfftn(init_mesh)
    a_t = _cosmo._t(a).reshape(-1)
    f1 = pm_forces(pos, init_mesh, paint_order, grad_fd=grad_fd, lap_fd=lap_fd)
    col = lambda t: t.to(device=f1.device, dtype=f1.dtype).reshape(-1, 1)
    E = _cosmo.Esqr(cosmo, a_t) ** 0.5
    dq = col(_cosmo.a2g(cosmo, a_t)) * f1
    p = col(a_t**2 * _cosmo.a2f(cosmo, a_t) * E) * dq
    if lpt_order == 2:
        f2 = pm_forces2(pos, init_mesh, paint_order, grad_fd=grad_fd, lap_fd=lap_fd)
        dq2 = col(_cosmo.a2g2(cosmo, a_t)) * f2
        dq = dq - dq2
        p = p - col(a_t**2 * _cosmo.a2f2(cosmo, a_t) * E) * dq2
    return dq, p


def diffrax_vf(cosmo, mesh_shape, paint_order, grad_fd=np.inf, lap_fd=np.inf):
    """N-body vector field in scale-factor time for generic ODE solvers (nbody.py:1076-1092):
    `vector_field(a, (pos, vel), args) -> (dpos, dvel)` with dpos = vel / (a^3 E), dvel = 1.5 Omega_m F / (a^2 E), F the
    particle-mesh force from the engine (differentiable in pos).  The adaptive Tsit5 driver around it (nbody_tsit5,
    1109-1138) is diffrax's and is not restated here."""
    mesh_shape = tuple(int(s) for s in mesh_shape)

    def vector_field(a, state, args=None):
        pos, vel = _f32(state[0]), _f32(state[1])
        forces = pm_forces(pos, mesh_shape, paint_order, grad_fd=grad_fd, lap_fd=lap_fd)
        a_t = _cosmo._t(a)
        E = float(_cosmo.Esqr(cosmo, a_t) ** 0.5)
        af = float(a_t)
        return vel * (1.0 / (af**3 * E)), forces * (1.5 * float(cosmo.Omega_m) / (af**2 * E))

    return vector_field


def bullfrog_vf(cosmo, dg, mesh_shape: tuple, paint_order: int = 2, paint_deconv=False, grad_fd=np.inf, lap_fd=np.inf):
    """BullFrog vector field (nbody.py:902-960): `vector_field(g0, state, args)` runs one drift-kick-drift step of size
    `dg` from growth time `g0` on `state = (pos, vel)` and returns `(new - old) / dg` for both, the increment an explicit
    Euler step of size `dg` turns back into the step (nbody.py:999).  One engine call per evaluation, differentiable
    in the state and, through alpha_bf, in the cosmology."""
    mesh_shape = tuple(int(s) for s in mesh_shape)

    def vector_field(g0, state, args=None):
        pos, vel = _f32(state[0]), _f32(state[1])
        g0_t, dg_t = _cosmo._t(g0), _cosmo._t(dg)
        alpha = _cosmo.alpha_bf(cosmo, g0_t, dg_t)
        coefs = torch.stack([alpha.reshape(()), ((1 - alpha) / (g0_t + dg_t / 2)).reshape(()), (dg_t / 2).reshape(()),
                             (dg_t / 2).reshape(())]).reshape(1, 4)
        new_pos, new_vel = _NbodySteps.apply(pos, vel, coefs, mesh_shape, int(paint_order), bool(paint_deconv), lap_fd,
                                             grad_fd)
        inv = 1.0 / float(dg_t)
        return (new_pos - pos) * inv, (new_vel - vel) * inv

    return vector_field


def nbody_bf_scan(cosmo, init_mesh, pos, a, n_steps=5, paint_order: int = 2, grad_fd=np.inf, lap_fd=np.inf,
                  snapshots=None):
    """No-diffrax BullFrog loop (nbody.py:1005-1029): n_steps equal steps in growth time from g = 0 to g(a), starting
    from `pos` with vel = pm_forces(pos, init_mesh) -- no LPT displacement, unlike nbody_bf.  The reference scans
    `bullfrog_vf`'s vector field as if it were a step function (its commented-out `step`); what is implemented here is
    that intended loop, state <- state + dg * vector_field(g0, state), as one engine call.  Returns (pos, vel) each
    [1, Np, 3]."""
    if snapshots is not None:
        raise NotImplementedError("nbody_bf_scan returns the final state only, as the reference does")
    n_steps = int(n_steps)
    init_mesh = _c64(init_mesh)
    pos = _f32(pos)
    mesh_shape = ch2rshape(tuple(init_mesh.shape))
    vel = pm_forces(pos, init_mesh, int(paint_order), grad_fd=grad_fd, lap_fd=lap_fd)
    dg = _cosmo.a2g(cosmo, a) / n_steps
    rows = []
    for i in range(n_steps):
        g0 = i * dg
        alpha = _cosmo.alpha_bf(cosmo, g0, dg)
        rows.append(torch.stack([alpha.reshape(()), ((1 - alpha) / (g0 + dg / 2)).reshape(()), (dg / 2).reshape(()),
                                 (dg / 2).reshape(())]))
    x, v = _NbodySteps.apply(pos, vel, torch.stack(rows), mesh_shape, int(paint_order), False, lap_fd, grad_fd)
    return x.unsqueeze(0), v.unsqueeze(0)


# growth helpers re-exported under the reference names (nbody.py:750-808)
a2g, a2g2, a2f, a2f2, a2dg2dg = _cosmo.a2g, _cosmo.a2g2, _cosmo.a2f, _cosmo.a2f2, _cosmo.a2dg2dg
g2a, g2g2, g2f, g2f2, g2dg2dg = _cosmo.g2a, _cosmo.g2g2, _cosmo.g2f, _cosmo.g2f2, _cosmo.g2dg2dg
alpha_bf = _cosmo.alpha_bf  # nbody.py:907-919
alpha_fpm = _cosmo.alpha_fpm  # nbody.py:921-931
a2chi, chi2a, k2ell, ell2k = _cosmo.a2chi, _cosmo.chi2a, _cosmo.k2ell, _cosmo.ell2k  # distances, nbody.py:810-896
