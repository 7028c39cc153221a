"""
CPU oracle for the particle-mesh hot path of hsimonfroy/montecosmo  --  TEST INFRASTRUCTURE ONLY.

This file restates, in torch-CPU float64 (so that autograd supplies reference gradients), the algorithm of
`montecosmo/nbody.py` and the Fourier helpers of `montecosmo/utils.py`.  Every function cites the reference
file:line it follows.  Nothing in `montecosmo_b200/` may import it: only `tests/`, `__graft_entry__.smoke()` and
the `cpu_baseline` / `--impl reference` legs of `bench.py` do, and only as the checker / reported baseline.

Pinning status.  The reference repository holds no golden vectors or runnable tests for this path
and JAX cannot be installed in this image (F6).  The oracle is therefore pinned against the reference's OWN SOURCE
(`/root/reference/montecosmo/{nbody,utils}.py`, unmodified) executed in float64 under a NumPy stand-in for
jax / jax_cosmo / diffrax (`tests/golden/jaxshim`, generator `tests/golden/make_golden.py`, fixtures
`tests/golden/*.npz`).  It is NOT pinned against outputs of real JAX/XLA: "parity pinned to reference source under a
NumPy stand-in; unpinned against JAX itself".

Third-party arithmetic restated here (absent from /root/reference):
  * jax_cosmo 0.1.0 (montenv.yml:360): background.{w, f_de, Esqr, Omega_m_a, Omega_de_a}, scipy.ode.odeint (RK4).
  * diffrax 0.5.0 (montenv.yml:352): diffeqsolve(Euler, constant step) == y <- y + vf(t, y) * dt.
"""
from __future__ import annotations

import itertools
import math

import numpy as np
import torch

F64 = torch.float64
C128 = torch.complex128


def _t(x, dtype=F64):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x), dtype=dtype)


# ----------------------------------------------------------------------------------------------------------------
# shapes (utils.py:769-783, 1163-1168)
# ----------------------------------------------------------------------------------------------------------------
def ch2rshape(shape):
    """utils.py:769-776."""
    return (*shape[:-1], 2 * (shape[-1] - 1))


def r2chshape(shape):
    """utils.py:778-782."""
    return (*shape[:-1], shape[-1] // 2 + 1)


def scale_shape(shape, scale=1.0):
    """utils.py:1163-1168: even-rounded scaled shape."""
    out = 2 * np.rint(np.multiply(shape, scale) / 2).astype(int)
    return tuple(int(o) for o in out)


def safe_div(x, y):
    """utils.py:21-29: x / y with 0 where y == 0."""
    if isinstance(x, torch.Tensor) or isinstance(y, torch.Tensor):
        x, y = _t(x) if not isinstance(x, torch.Tensor) else x, _t(y) if not isinstance(y, torch.Tensor) else y
        y_nz = torch.where(y == 0, torch.ones_like(y), y)
        return torch.where(y == 0, torch.zeros_like(x / y_nz), x / y_nz)
    y = np.asarray(y, dtype=float)
    y_nz = np.where(y == 0, 1, y)
    return np.where(y == 0, 0, x / y_nz)


# ----------------------------------------------------------------------------------------------------------------
# Fourier kernels (nbody.py:50-334), all in cell units unless box_size is given
# ----------------------------------------------------------------------------------------------------------------
def rfftk(shape, box_size=None):
    """nbody.py:50-77: fftfreq on all axes but the last (rfftfreq); Nyquist is -pi on x,y and +pi on z."""
    dim = len(shape)
    scales = dim * (2 * np.pi,) if box_size is None else tuple(2 * np.pi * s / b for s, b in zip(shape, box_size))
    kvec = []
    for ax, (s, sc) in enumerate(zip(shape, scales)):
        k = (np.fft.fftfreq(s) if ax < dim - 1 else np.fft.rfftfreq(s)) * sc
        shp = [1] * dim
        shp[ax] = -1
        kvec.append(k.reshape(shp))
    return tuple(kvec)


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
    return -safe_div(1, kk)


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
    """nbody.py:166-188 (JaxPM's long-range filter)."""
    if kcut == np.inf:
        return 1.0
    kk = sum(k**2 for k in kvec)
    rcut = 2 * np.pi / kcut
    return np.exp(-kk * rcut**2 / 2)


def rectangular_hat(kvec, order=2):
    """nbody.py:249-277: prod_j sinc(k_j / 2pi)^order."""
    out = 1.0
    for k in kvec:
        out = out * np.sinc(k / (2 * np.pi)) ** order
    return out


def optim_kcut(oversamp, safety=0.98):
    """nbody.py:357-363."""
    return safety * np.pi * (2 - 1 / oversamp)


def kaiser_bessel_hat(kvec, order, kcut):
    """nbody.py:293-312."""
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


# ----------------------------------------------------------------------------------------------------------------
# mass assignment (nbody.py:220-246, 280-290, 365-427)
# ----------------------------------------------------------------------------------------------------------------
def rectangular(s, order):
    """nbody.py:220-246 on |s|; order 1 NGP, 2 CIC, 3 TSC, 4 PCS."""
    s = s.abs()
    if order == 1:
        return torch.ones_like(s)
    if order == 2:
        return 1 - s
    if order == 3:
        return (s <= 0.5).to(s.dtype) * (0.75 - s**2) + (0.5 < s).to(s.dtype) / 2 * (1.5 - s) ** 2
    if order == 4:
        return (s <= 1).to(s.dtype) / 6 * (4 - 6 * s**2 + 3 * s**3) + (1 < s).to(s.dtype) / 6 * (2 - s) ** 3
    raise ValueError("order must be 1..4")


def kaiser_bessel(s, order, kcut):
    """nbody.py:280-290."""
    s = s * 2 / order
    kcut = kcut * order / 2
    out = torch.special.i0(kcut * (1 - s**2) ** 0.5)
    return out / (order * math.sinh(kcut) / kcut)


def _assignment(pos, shape, order, kernel_type, oversamp):
    """Shared index/weight generator of nbody.py:369-388 / 402-418: yields (wrapped idx [Np,3] int64, ker [Np])."""
    shape_t = torch.as_tensor(shape, dtype=torch.int64)
    id0 = (torch.round(pos) if order % 2 else torch.floor(pos)).detach().to(torch.int64)  # round = half-to-even
    ishifts = np.arange(order) - (order - 1) // 2
    if kernel_type == "rectangular":
        kernel = lambda s: rectangular(s, order)
    elif kernel_type == "kaiser_bessel":
        kernel = lambda s: kaiser_bessel(s, order, optim_kcut(oversamp))
    else:
        raise ValueError(f"Unknown kernel type: {kernel_type}")
    for ish in itertools.product(*(len(shape) * (ishifts,))):
        idx = id0 + torch.as_tensor(ish, dtype=torch.int64)
        ker = kernel(idx.to(pos.dtype) - pos).prod(-1)  # window on the UNWRAPPED index (nbody.py:388)
        yield torch.remainder(idx, shape_t), ker  # python-style modulo


# Memory-lean twins of paint / read for the full-size fixtures (tests/golden/make_full_size_fixture.py): identical
# forward arithmetic, but the autograd graph of the order^3 neighbour loop is not kept -- the backward pass rebuilds it one
# neighbour at a time (torch.autograd on that neighbour's window alone), so a 256^3 gradient fits in RAM.  Switched on
# with LEAN_ASSIGNMENT = True; tests/test_oracle_golden.py checks both forms give the same gradients.
LEAN_ASSIGNMENT = False


class _LeanPaint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, weights, shape, order, kernel_type, oversamp):
        ctx.save_for_backward(pos, weights)
        ctx.cfg = (shape, order, kernel_type, oversamp)
        mesh = torch.zeros(shape, dtype=pos.dtype)
        for idx, ker in _assignment(pos, shape, order, kernel_type, oversamp):
            mesh.index_put_(tuple(idx.unbind(-1)), (weights * ker).expand(pos.shape[0]), accumulate=True)
        return mesh

    @staticmethod
    def backward(ctx, gmesh):
        pos, weights = ctx.saved_tensors
        shape, order, kernel_type, oversamp = ctx.cfg
        gpos, gw = torch.zeros_like(pos), torch.zeros(pos.shape[0], dtype=pos.dtype)
        with torch.enable_grad():
            p = pos.detach().requires_grad_(True)
            for idx, ker in _assignment(p, shape, order, kernel_type, oversamp):
                gm = gmesh[tuple(idx.unbind(-1))]
                gw += gm * ker.detach()
                if ker.requires_grad:  # NGP's window is constant
                    gpos += torch.autograd.grad(ker, p, gm * weights)[0]
        gw = gw.sum() if weights.dim() == 0 else gw
        return gpos, gw, None, None, None, None


class _LeanRead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, mesh, order, kernel_type, oversamp):
        ctx.save_for_backward(pos, mesh)
        ctx.cfg = (order, kernel_type, oversamp)
        out = torch.zeros(pos.shape[0], dtype=mesh.dtype)
        for idx, ker in _assignment(pos, mesh.shape, order, kernel_type, oversamp):
            out += mesh[tuple(idx.unbind(-1))] * ker
        return out

    @staticmethod
    def backward(ctx, gout):
        pos, mesh = ctx.saved_tensors
        order, kernel_type, oversamp = ctx.cfg
        gpos, gmesh = torch.zeros_like(pos), torch.zeros_like(mesh)
        with torch.enable_grad():
            p = pos.detach().requires_grad_(True)
            for idx, ker in _assignment(p, mesh.shape, order, kernel_type, oversamp):
                ii = tuple(idx.unbind(-1))
                gmesh.index_put_(ii, gout * ker.detach(), accumulate=True)
                if ker.requires_grad:
                    gpos += torch.autograd.grad(ker, p, gout * mesh[ii])[0]
        return gpos, gmesh, None, None, None


def paint(pos, shape, weights=1.0, order=2, kernel_type="rectangular", oversamp=1.0):
    """nbody.py:365-396: scatter-add of weights * prod_j W over order^3 neighbours."""
    pos = _t(pos)
    if LEAN_ASSIGNMENT:
        return _LeanPaint.apply(pos, _t(weights, pos.dtype), tuple(int(s) for s in shape), order, kernel_type, oversamp)
    mesh = torch.zeros(tuple(int(s) for s in shape), dtype=pos.dtype)
    weights = _t(weights, pos.dtype)
    for idx, ker in _assignment(pos, shape, order, kernel_type, oversamp):
        val = (weights * ker).expand(pos.shape[0])
        mesh = mesh.index_put(tuple(idx.unbind(-1)), val, accumulate=True)
    return mesh


def read(pos, mesh, order=2, kernel_type="rectangular", oversamp=1.0):
    """nbody.py:398-427: gather transpose of paint."""
    pos = _t(pos)
    mesh = _t(mesh, pos.dtype)
    if LEAN_ASSIGNMENT:
        return _LeanRead.apply(pos, mesh, order, kernel_type, oversamp)
    out = torch.zeros(pos.shape[0], dtype=mesh.dtype)
    for idx, ker in _assignment(pos, mesh.shape, order, kernel_type, oversamp):
        out = out + mesh[tuple(idx.unbind(-1))] * ker
    return out


# ----------------------------------------------------------------------------------------------------------------
# Fourier reshape (utils.py:924-1013)
# ----------------------------------------------------------------------------------------------------------------
def _chreshape(mesh, shape):
    """utils.py:924-958: centre, crop or zero-pad each axis, de-centre, scale by the ratio of real sizes."""
    scale = np.divide(ch2rshape(shape), ch2rshape(mesh.shape)).prod()
    nd = mesh.dim()
    for ax in range(nd - 1):
        mesh = torch.roll(mesh, mesh.shape[ax] // 2, ax)
    sl = []
    for ax, (ms, s) in enumerate(zip(mesh.shape, shape)):
        trunc = max(ms - s, 0)
        if ax < nd - 1:
            trunc //= 2
            sl.append(slice(trunc, None if trunc == 0 else -trunc))
        else:
            sl.append(slice(0, None if trunc == 0 else -trunc))
    mesh = mesh[tuple(sl)]
    pads = []
    for ax, (ms, s) in enumerate(zip(mesh.shape, shape)):
        pad = max(s - ms, 0)
        pads.append((pad // 2, pad // 2) if ax < nd - 1 else (0, pad))
    flat = [p for pair in reversed(pads) for p in pair]  # torch pads from the last axis backwards
    mesh = torch.nn.functional.pad(mesh, flat)
    for ax in range(nd - 1):
        mesh = torch.roll(mesh, (-mesh.shape[ax]) // 2, ax)
    return mesh * scale


def hermitian_symmetric(arr):
    """utils.py:962-972: conj(arr[-k]) for every axis."""
    dims = tuple(range(arr.dim()))
    out = torch.flip(arr, dims).conj()
    for ax in dims:
        out = torch.roll(out, 1, ax)
    return out


_RG_NORMS = ("backward", "ortho", "forward", "amp")


def _rg_norm_factor(shape, norm):
    """utils.py:821-829 / 872-880: the factor rg2cgh divides by (cgh2rg multiplies by)."""
    n = float(np.prod(shape))
    if norm not in _RG_NORMS:
        raise AssertionError("norm must be either 'backward', 'forward', 'ortho', or 'amp'.")
    return {"backward": (2.0 / n) ** 0.5, "ortho": 2.0 ** 0.5, "forward": (2.0 * n) ** 0.5, "amp": 1.0}[norm]


def _rg_slices(shape, part):
    hx, hy, hz = (s // 2 for s in shape)
    if part == "imag":
        return slice(hx + 1, None), slice(hy + 1, None), slice(hz + 1, None)
    return slice(1, hx), slice(1, hy), slice(1, hz)


def _rg2cgh_part(mesh, part, norm):
    """utils.py:785-831.  One part (real or imaginary) of the Hermitian half spectrum, gathered from the real mesh:
    interior planes, then the two self-conjugate planes kz in {0, Nyquist} (rows, mirrored rows with sign), their
    self-conjugate rows ky in {0, Nyquist} (same along x), and the 8 purely real modes (x sqrt2)."""
    shape = tuple(mesh.shape)
    assert all(s % 2 == 0 for s in shape), "dimension lengths must be even."
    hx, hy, hz = (s // 2 for s in shape)
    sx, sy, sz = _rg_slices(shape, part)
    flip = part == "imag" and norm != "amp"
    out = torch.zeros(r2chshape(shape), dtype=mesh.dtype)
    out[:, :, 1:-1] = mesh[:, :, sz]
    for k in (0, hz):
        out[:, 1:hy, k] = mesh[:, sy, k]
        out[1:, hy + 1:, k] = torch.flip(mesh[1:, sy, k], dims=(0, 1))
        out[0, hy + 1:, k] = torch.flip(mesh[0, sy, k], dims=(0,))
        if flip:
            out[:, hy + 1:, k] = -out[:, hy + 1:, k]
        for j in (0, hy):
            out[1:hx, j, k] = mesh[sx, j, k]
            out[hx + 1:, j, k] = torch.flip(mesh[sx, j, k], dims=(0,))
            if flip:
                out[hx + 1:, j, k] = -out[hx + 1:, j, k]
            for i in (0, hx):
                if part == "real":
                    out[i, j, k] = mesh[i, j, k] * (2.0 ** 0.5 if norm != "amp" else 1.0)
    return out / _rg_norm_factor(shape, norm)


def rg2cgh(mesh, norm="backward"):
    """utils.py:888-903: real Gaussian mesh -> complex Gaussian Hermitian half spectrum, distributed as rfftn(N(0,I))."""
    mesh = _t(mesh)
    real = _rg2cgh_part(mesh, "real", norm)
    if norm == "amp":
        return real
    return torch.complex(real, _rg2cgh_part(mesh, "imag", norm))


def _cgh2rg_part(vals, part, norm):
    """utils.py:836-882: scatter one part of the half spectrum back onto its partition of the real mesh.  Where a real
    number has two Hermitian images the later assignment (the mirrored one) is the one the reference keeps."""
    shape = ch2rshape(tuple(vals.shape))
    assert all(s % 2 == 0 for s in shape), "dimension lengths must be even."
    hx, hy, hz = (s // 2 for s in shape)
    sx, sy, sz = _rg_slices(shape, part)
    flip = part == "imag" and norm != "amp"
    mesh = torch.zeros(shape, dtype=vals.dtype)
    mesh[:, :, sz] = vals[:, :, 1:-1]
    for k in (0, hz):
        mesh[:, sy, k] = vals[:, 1:hy, k]
        mesh[1:, sy, k] = torch.flip(vals[1:, hy + 1:, k], dims=(0, 1))
        mesh[0, sy, k] = torch.flip(vals[0, hy + 1:, k], dims=(0,))
        if flip:
            mesh[:, sy, k] = -mesh[:, sy, k]
        for j in (0, hy):
            mesh[sx, j, k] = vals[1:hx, j, k]
            mesh[sx, j, k] = torch.flip(vals[hx + 1:, j, k], dims=(0,))
            if flip:
                mesh[sx, j, k] = -mesh[sx, j, k]
            for i in (0, hx):
                if part == "real":
                    mesh[i, j, k] = vals[i, j, k] / (2.0 ** 0.5 if norm != "amp" else 1.0)
    return mesh * _rg_norm_factor(shape, norm)


def cgh2rg(meshk, norm="backward"):
    """utils.py:906-921: the inverse of rg2cgh (for "amp": the same amplitude on both parts)."""
    meshk = _t(meshk, C128) if not isinstance(meshk, torch.Tensor) else meshk
    re = meshk.real if torch.is_complex(meshk) else meshk
    im = re if norm == "amp" else meshk.imag
    return _cgh2rg_part(re, "real", norm) + _cgh2rg_part(im, "imag", norm)


def chreshape(mesh, shape):
    """utils.py:975-1013: Hermitian- and mean-preserving Fourier crop / pad with Nyquist-plane fix-ups."""
    mesh = _t(mesh, C128).clone()
    nd = mesh.dim()
    in_shape = tuple(mesh.shape)
    for ax in reversed(range(nd)):
        ms, s = in_shape[ax], shape[ax]
        if s < ms:
            pre = (slice(None),) * ax
            if ax < nd - 1:
                neg, posi = pre + (-s // 2,), pre + (s // 2,)
                new = (mesh[posi] + mesh[neg]) / 2**0.5
                mesh = mesh.clone()
                mesh[neg] = new
            else:
                posi = pre + (s - 1,)
                plane = mesh[posi]
                new = (plane + hermitian_symmetric(plane)) / 2**0.5
                mesh = mesh.clone()
                mesh[posi] = new
    out = _chreshape(mesh, shape)
    for ax in range(nd):
        ms, s = in_shape[ax], shape[ax]
        if s > ms:
            pre = (slice(None),) * ax
            out = out.clone()
            if ax < nd - 1:
                neg, posi = pre + (-ms // 2,), pre + (ms // 2,)
                out[neg] = out[neg] / 2**0.5
                out[posi] = out[neg]
            else:
                posi = pre + (ms - 1,)
                out[posi] = out[posi] / 2**0.5
    return out


# ----------------------------------------------------------------------------------------------------------------
# NUFFT (nbody.py:315-334, 513-577)
# ----------------------------------------------------------------------------------------------------------------
def deconv_paint(mesh, order=2, kernel_type="rectangular", oversamp=1.0):
    """nbody.py:315-334."""
    if kernel_type == "rectangular":
        kernel = lambda kvec: rectangular_hat(kvec, order)
    elif kernel_type == "kaiser_bessel":
        kernel = lambda kvec: kaiser_bessel_hat(kvec, order, optim_kcut(oversamp))
    else:
        raise ValueError(f"Unknown kernel type: {kernel_type}")
    if not torch.is_complex(mesh):
        kvec = rfftk(mesh.shape)
        out = torch.fft.rfftn(mesh) / _t(kernel(kvec))
        return torch.fft.irfftn(out, s=tuple(mesh.shape))
    kvec = rfftk(ch2rshape(mesh.shape))
    return mesh / _t(kernel(kvec))


def interlace(pos, shape, weights=1.0, paint_order=2, interlace_order=2, kernel_type="rectangular",
              paint_oversamp=1.0):
    """nbody.py:513-529: mean over shifts s of rfftn(paint(pos + s)) * exp(i s (kx + ky + kz))."""
    kvec = rfftk(shape)
    ksum = _t(sum(kvec))
    mesh = torch.zeros(r2chshape(tuple(shape)), dtype=C128)
    for i in range(interlace_order):
        shift = i / interlace_order
        m = paint(pos + shift, shape, weights, paint_order, kernel_type, paint_oversamp)
        mesh = mesh + torch.fft.rfftn(m) * torch.exp(1j * shift * ksum) / interlace_order
    return mesh


def nufft(pos, final_shape, paint_shape=None, weights=1.0, paint_order=2, interlace_order=2,
          kernel_type="rectangular", paint_deconv=True):
    """nbody.py:532-577."""
    pos = _t(pos)
    if paint_shape is None:
        paint_shape, paint_oversamp = final_shape, 1.0
    elif isinstance(paint_shape, float):
        paint_oversamp = paint_shape
        paint_shape = scale_shape(final_shape, paint_oversamp)
    else:
        paint_oversamp = float(np.exp(np.log(np.divide(final_shape, paint_shape)).mean()))  # sic, nbody.py:565
    pos = pos * _t(np.divide(paint_shape, final_shape))
    mesh = interlace(pos, paint_shape, weights, paint_order, interlace_order, kernel_type, paint_oversamp)
    mesh = mesh * float(np.divide(paint_shape, final_shape).prod())
    if paint_deconv:
        mesh = deconv_paint(mesh, paint_order, kernel_type, paint_oversamp)
    if tuple(final_shape) != tuple(paint_shape):
        mesh = chreshape(mesh, r2chshape(tuple(final_shape)))
    return mesh


# ----------------------------------------------------------------------------------------------------------------
# forces and LPT (nbody.py:583-667)
# ----------------------------------------------------------------------------------------------------------------
def _irfftn(meshk):
    return torch.fft.irfftn(meshk, s=ch2rshape(tuple(meshk.shape)))


def pm_forces(pos, mesh, read_order=2, paint_deconv=False, grad_fd=np.inf, lap_fd=np.inf, kcut=np.inf):
    """nbody.py:583-604: F = -grad lap^-1 delta, read at pos.  `mesh` is a shape tuple (paint pos) or delta_k."""
    pos = _t(pos)
    if isinstance(mesh, tuple):
        mesh = torch.fft.rfftn(paint(pos, mesh, order=read_order))
        if paint_deconv:
            mesh = mesh / _t(rectangular_hat(rfftk(ch2rshape(mesh.shape)), order=read_order) ** 2)
    kvec = rfftk(ch2rshape(tuple(mesh.shape)))
    pot = mesh * _t(invlaplace_hat(kvec, lap_fd))
    if kcut != np.inf:
        pot = pot * _t(gaussian_hat(kvec, kcut))
    return torch.stack([read(pos, _irfftn(-_t(gradient_hat(kvec, i, grad_fd), C128) * pot), read_order)
                        for i in range(len(kvec))], dim=-1)


def pm_forces2(pos, mesh, read_order=2, grad_fd=np.inf, lap_fd=np.inf):
    """nbody.py:607-631: 2LPT source sum_{i<j}(phi_ii phi_jj - phi_ij^2) then pm_forces on it."""
    kvec = rfftk(ch2rshape(tuple(mesh.shape)))
    pot = mesh * _t(invlaplace_hat(kvec, lap_fd))
    delta2, hesses = 0.0, 0.0
    for i in range(len(kvec)):
        gi = _t(gradient_hat(kvec, i, grad_fd), C128)
        hess_ii = _irfftn(gi**2 * pot)
        delta2 = delta2 + hess_ii * hesses
        hesses = hesses + hess_ii
        for j in range(i + 1, len(kvec)):
            gj = _t(gradient_hat(kvec, j, grad_fd), C128)
            delta2 = delta2 - _irfftn(gi * gj * pot) ** 2
    return pm_forces(pos, torch.fft.rfftn(delta2), read_order, grad_fd=grad_fd, lap_fd=lap_fd)


def lpt(cosmo, init_mesh, pos, a, lpt_order=2, read_order=2, grad_fd=np.inf, lap_fd=np.inf):
    """nbody.py:634-667: dpos = D1 F1 - D2 F2 ; vel = F1 - (D2 f2)/(D1 f1) F2."""
    init_mesh = _t(init_mesh, C128) if not isinstance(init_mesh, torch.Tensor) else init_mesh
    if not torch.is_complex(init_mesh):
        init_mesh = torch.fft.rfftn(init_mesh)
    force1 = pm_forces(pos, init_mesh, read_order, grad_fd=grad_fd, lap_fd=lap_fd)
    dpos = a2g(cosmo, a) * force1
    vel = force1
    if lpt_order == 2:
        force2 = pm_forces2(pos, init_mesh, read_order, grad_fd=grad_fd, lap_fd=lap_fd)
        dpos = dpos - a2g2(cosmo, a) * force2
        vel = vel - a2dg2dg(cosmo, a) * force2
    return dpos, vel


# ----------------------------------------------------------------------------------------------------------------
# cosmology background (jax_cosmo 0.1.0) and growth tables (nbody.py:675-808)
# ----------------------------------------------------------------------------------------------------------------
class Cosmology:
    """jax_cosmo.core.Cosmology fields; AbacusSummit0 defaults (bricks.py:40-50)."""

    def __init__(self, Omega_c=0.26447041, Omega_b=0.04930169, h=0.6736, n_s=0.9649, sigma8=0.8076353990239834,
                 Omega_k=0.0, w0=-1.0, wa=0.0):
        self.Omega_c, self.Omega_b, self.h, self.n_s = Omega_c, Omega_b, h, n_s
        self.sigma8, self.Omega_k, self.w0, self.wa = sigma8, Omega_k, w0, wa
        self._workspace = {}

    @property
    def Omega_m(self):
        return self.Omega_b + self.Omega_c

    @property
    def Omega_de(self):
        return 1.0 - self.Omega_k - self.Omega_m


def _w(c, a):
    return c.w0 + (1.0 - a) * c.wa


def _f_de(c, a):
    eps = float(np.finfo(np.float32).eps)
    return -3.0 * (1.0 + c.w0) + 3.0 * c.wa * ((a - 1.0) / torch.log(a - eps) - 1.0)


def Esqr(c, a):
    a = _t(a)
    return c.Omega_m * a**-3 + c.Omega_k * a**-2 + c.Omega_de * a ** _f_de(c, a)


def Omega_m_a(c, a):
    a = _t(a)
    return c.Omega_m * a**-3 / Esqr(c, a)


def Omega_de_a(c, a):
    a = _t(a)
    return c.Omega_de * a ** _f_de(c, a) / Esqr(c, a)


growth_log10_amin = -3.0
growth_steps = 128


def _odeint_rk4(fn, y0, t):
    """jax_cosmo.scipy.ode.odeint: one classical RK4 step per interval of the nodes t (first step has h = 0)."""
    y, t_prev, out = y0, t[0], []
    for ti in t:
        h = ti - t_prev
        k1 = fn(y, t_prev)
        k2 = fn(y + h * k1 / 2, t_prev + h / 2)
        k3 = fn(y + h * k2 / 2, t_prev + h / 2)
        k4 = fn(y + k3 * h, ti)
        y = y + 1.0 / 6.0 * h * (k1 + 2 * k2 + 2 * k3 + k4)
        t_prev = ti
        out.append(y)
    return torch.stack(out)


def growth_table(cosmo):
    """nbody.py:679-748: RK4 table of (D1, D2) and derivatives on 128 log-spaced a in [1e-3, 1], normalised at a=1."""
    if "growth" in cosmo._workspace:
        return cosmo._workspace["growth"]
    atab = _t(np.logspace(growth_log10_amin, 0.0, growth_steps))

    def derivs(y, x):
        q = (2.0 - (Omega_m_a(cosmo, x) + (1.0 + 3.0 * _w(cosmo, x)) * Omega_de_a(cosmo, x)) / 2) / x
        r = 1.5 * Omega_m_a(cosmo, x) / x**2
        g1, g2 = y[0, 0], y[0, 1]
        f1, f2 = y[1, 0], y[1, 1]
        return torch.stack([torch.stack([f1, f2]),
                            torch.stack([-q * f1 + r * g1, -q * f2 + r * g2 - r * g1**2])])

    a0 = atab[0]
    y0 = torch.stack([torch.stack([a0, -3.0 / 7 * a0**2]), torch.stack([torch.ones_like(a0), -6.0 / 7 * a0])])
    y = _odeint_rk4(derivs, y0, atab)
    y1, y2 = y[:, 0, 0], y[:, 0, 1]
    gtab, g2tab = y1 / y1[-1], y2 / y2[-1]
    ftab = y[:, 1, 0] / y1[-1] * atab / gtab
    f2tab = y[:, 1, 1] / y2[-1] * atab / g2tab
    cache = {"a": atab, "g": gtab, "f": ftab, "g2": g2tab, "f2": f2tab}
    cosmo._workspace["growth"] = cache
    return cache


def _interp(x, xp, fp):
    """jnp.interp: piecewise linear, clamped to the end values."""
    x = _t(x)
    xs = x.reshape(-1)
    i = torch.clamp(torch.searchsorted(xp.detach().contiguous(), xs.detach().contiguous(), right=True) - 1,
                    0, len(xp) - 2)
    t = (xs - xp[i]) / (xp[i + 1] - xp[i])
    out = fp[i] + t * (fp[i + 1] - fp[i])
    out = torch.where(xs <= xp[0], fp[0].expand_as(out), out)
    out = torch.where(xs >= xp[-1], fp[-1].expand_as(out), out)
    return out.reshape(x.shape)


def a2g(c, a):
    t = growth_table(c)
    return _interp(a, t["a"], t["g"])  # nbody.py:750-754


def a2g2(c, a):
    t = growth_table(c)
    return _interp(a, t["a"], t["g2"]) * -3 / 7  # nbody.py:756-761


def a2f(c, a):
    t = growth_table(c)
    return _interp(a, t["a"], t["f"])  # nbody.py:763-767


def a2f2(c, a):
    t = growth_table(c)
    return _interp(a, t["a"], t["f2"])  # nbody.py:769-773


def a2dg2dg(c, a):
    return safe_div(a2g2(c, a) * a2f2(c, a), a2g(c, a) * a2f(c, a))  # nbody.py:775-777


def g2a(c, g):
    t = growth_table(c)
    return _interp(g, t["g"], t["a"])  # nbody.py:781-785


def g2g2(c, g):
    t = growth_table(c)
    return _interp(g, t["g"], t["g2"]) * -3 / 7  # nbody.py:787-792


def g2f(c, g):
    t = growth_table(c)
    return _interp(g, t["g"], t["f"])  # nbody.py:794-798


def g2f2(c, g):
    t = growth_table(c)
    return _interp(g, t["g"], t["f2"])  # nbody.py:800-804


def g2dg2dg(c, g):
    return safe_div(g2g2(c, g) * g2f2(c, g), _t(g) * g2f(c, g))  # nbody.py:806-808


# ----------------------------------------------------------------------------------------------------------------
# BullFrog solver (nbody.py:902-1002)
# ----------------------------------------------------------------------------------------------------------------
def alpha_bf(cosmo, g0, dg):
    """nbody.py:907-919."""
    g1 = g0 + dg / 2
    g2 = g0 + dg
    d0, d2 = g2dg2dg(cosmo, g0), g2dg2dg(cosmo, g2)
    lin_ratio = (g2g2(cosmo, g0) + d0 * dg / 2) / g1 - g1
    return (d2 - lin_ratio) / (d0 - lin_ratio)


def lpt_fpm(cosmo, init_mesh, pos, a, lpt_order=1, paint_order=2, grad_fd=np.inf, lap_fd=np.inf):
    """nbody.py:1030-1073.  The inline Hessian accumulation of the reference (sum over i of h_ii * (h_00 + .. + h_i-1,i-1)
    minus the squared off-diagonals) is restated literally rather than through pm_forces2."""
    a = _t(a).reshape(-1)
    E = Esqr(cosmo, a) ** 0.5
    init_mesh = init_mesh if isinstance(init_mesh, torch.Tensor) else _t(init_mesh, C128)
    mesh_shape = ch2rshape(tuple(init_mesh.shape))
    pos = _t(pos)
    f1 = pm_forces(pos, init_mesh, paint_order, grad_fd=grad_fd, lap_fd=lap_fd)
    dq = a2g(cosmo, a) * f1
    p = a ** 2 * a2f(cosmo, a) * E * dq
    if lpt_order == 2:
        kvec = rfftk(mesh_shape)
        pot = init_mesh * _t(invlaplace_hat(kvec, lap_fd))
        delta2, acc = 0.0, 0.0
        for i in range(3):
            hii = torch.fft.irfftn(_t(gradient_hat(kvec, i, grad_fd) ** 2, C128) * pot, s=mesh_shape)
            delta2 = delta2 + hii * acc
            acc = acc + hii
            for j in range(i + 1, 3):
                hij = _t(gradient_hat(kvec, i, grad_fd) * gradient_hat(kvec, j, grad_fd), C128)
                delta2 = delta2 - torch.fft.irfftn(hij * pot, s=mesh_shape) ** 2
        f2 = pm_forces(pos, torch.fft.rfftn(delta2), paint_order, grad_fd=grad_fd, lap_fd=lap_fd)
        dq2 = a2g2(cosmo, a) * f2
        dq = dq - dq2
        p = p - a ** 2 * a2f2(cosmo, a) * E * dq2
    return dq, p


def diffrax_vf(cosmo, mesh_shape, paint_order, grad_fd=np.inf, lap_fd=np.inf):
    """nbody.py:1076-1092."""
    def vector_field(a, state, args=None):
        pos, vel = state
        forces = pm_forces(pos, tuple(mesh_shape), paint_order, grad_fd=grad_fd, lap_fd=lap_fd) * 1.5 * cosmo.Omega_m
        E = Esqr(cosmo, _t(a)) ** 0.5
        return vel / (a ** 3 * E), forces / (a ** 2 * E)
    return vector_field


RH = 2997.92458  # jax_cosmo.constants.rh (h^-1 Mpc)


def distance_table(cosmo, log10_amin=-3.0, steps=256):
    """nbody.py:816-857: chi(a) on `steps` log-spaced a, jax_cosmo's dchioverda = rh / (a^2 E(a)) integrated in ln a by
    its fixed-step RK4 odeint, then chi(a) = chi_tab[-1] - chi_tab."""
    atab = _t(np.logspace(log10_amin, 0.0, steps))

    def dchioverdlna(y, x):
        xa = torch.exp(x)
        return RH / (xa ** 2 * Esqr(cosmo, xa) ** 0.5) * xa

    chitab = _odeint_rk4(dchioverdlna, torch.zeros((), dtype=F64), torch.log(atab))
    return {"a": atab, "chi": chitab[-1] - chitab}


def a2chi(cosmo, a):
    """nbody.py:816-857."""
    t = distance_table(cosmo)
    return torch.clamp(_interp(a, t["a"], t["chi"]), min=0.0)


def chi2a(cosmo, chi):
    """nbody.py:860-884 (chi decreases with a: both tables reversed)."""
    t = distance_table(cosmo)
    return _interp(chi, torch.flip(t["chi"], dims=(0,)), torch.flip(t["a"], dims=(0,)))


def alpha_fpm(cosmo, g0, dg):
    """nbody.py:921-931 (FastPM growth-time coefficient, defined next to alpha_bf; the reference's kick uses alpha_bf)."""
    g0 = _t(g0)
    g2 = g0 + dg
    a0, a2 = g2a(cosmo, g0), g2a(cosmo, g2)
    coeff0 = Esqr(cosmo, a0) ** 0.5 * g0 * g2f(cosmo, g0) * a0 ** 2
    coeff2 = Esqr(cosmo, a2) ** 0.5 * g2 * g2f(cosmo, g2) * a2 ** 2
    return coeff0 / coeff2


def bullfrog_step(cosmo, state, g0, dg, mesh_shape, paint_order=2, paint_deconv=False,
                  grad_fd=np.inf, lap_fd=np.inf):
    """One drift-kick-drift step, nbody.py:933-951."""
    pos, vel = state
    pos = pos + vel * (dg / 2)
    g1 = g0 + dg / 2
    forces = pm_forces(pos, tuple(mesh_shape), paint_order, paint_deconv=paint_deconv, grad_fd=grad_fd, lap_fd=lap_fd)
    alpha = alpha_bf(cosmo, g0, dg)
    vel = alpha * vel + (1 - alpha) * forces / g1
    pos = pos + vel * (dg / 2)
    return pos, vel


def nbody_bf(cosmo, init_mesh, pos, a0=0.0, a1=1.0, n_steps=5, paint_order=2, lpt_order=2, paint_deconv=False,
             grad_fd=np.inf, lap_fd=np.inf, snapshots=None, fn=None, checkpoint=False):
    """nbody.py:967-1002 with diffrax 0.5.0's Euler restated: y <- y + ((new - y) / dg) * dt, last dt clipped onto g1.
    `checkpoint` (oracle-only): recompute every step in the backward pass instead of keeping its autograd graph, the
    memory policy diffrax's default adjoint applies in the reference (nbody.py:999) -- same numbers, 256^3 fits in RAM.

    Returns (pos, vel), each [S, Np, 3]; S = 1 (final state) unless `snapshots` asks for more (nbody.py:987-996): an int
    S > 1 saves at linspace(g0, g1, S), a sequence of scale factors at a2g(cosmo, .).  diffrax fills `SaveAt(ts=...)`
    from the solver's dense output, which for Euler is the straight line between the two ends of the step that holds
    t; `fn(t, y, args)` (default: identity) maps every saved state.
    """
    n_steps = int(n_steps)
    init_mesh = _t(init_mesh, C128) if not isinstance(init_mesh, torch.Tensor) else init_mesh
    pos = _t(pos)
    g0, g1 = a2g(cosmo, a0), a2g(cosmo, a1)
    dg = (g1 - g0) / n_steps
    mesh_shape = ch2rshape(tuple(init_mesh.shape))
    dpos, vel = lpt(cosmo, init_mesh, pos=pos, a=a0, lpt_order=lpt_order, read_order=1, grad_fd=grad_fd, lap_fd=lap_fd)
    state = (pos + dpos, vel)
    traj, times = [state], [g0]
    t = g0
    for n in range(n_steps):
        tn = g1 if n == n_steps - 1 else t + dg
        dt = tn - t
        if checkpoint and torch.is_grad_enabled():
            from torch.utils.checkpoint import checkpoint as _ckpt
            new = _ckpt(lambda p_, v_, t_=t: bullfrog_step(cosmo, (p_, v_), t_, dg, mesh_shape, paint_order, paint_deconv,
                                                           grad_fd, lap_fd), *state, use_reentrant=False)
        else:
            new = bullfrog_step(cosmo, state, t, dg, mesh_shape, paint_order, paint_deconv, grad_fd, lap_fd)
        state = tuple(y + ((nw - y) / dg) * dt for y, nw in zip(state, new))
        t = tn
        traj.append(state)
        times.append(t)
    if snapshots is None:
        return state[0][None], state[1][None]
    fn = (lambda t, y, args: y) if fn is None else fn
    if isinstance(snapshots, (int, np.integer)):
        if snapshots <= 1:
            outs = [fn(g1, state, None)]
        else:
            ts = [g0 + (g1 - g0) * (i / (snapshots - 1)) for i in range(int(snapshots))]
    else:
        ts = list(a2g(cosmo, _t(np.asarray(snapshots, dtype=np.float64))).reshape(-1))
    if not (isinstance(snapshots, (int, np.integer)) and snapshots <= 1):
        tt = np.array([float(x) for x in times])
        outs = []
        for s in ts:
            i = int(np.clip(np.searchsorted(tt, float(s), side="left") - 1, 0, n_steps - 1))  # t in (t_i, t_i+1]
            th = (s - times[i]) / (times[i + 1] - times[i])
            outs.append(fn(s, tuple(a + th * (b - a) for a, b in zip(traj[i], traj[i + 1])), None))
    if isinstance(outs[0], (tuple, list)):
        return tuple(torch.stack([o[j] for o in outs]) for j in range(len(outs[0])))
    return torch.stack(outs)


# ----------------------------------------------------------------------------------------------------------------
# helpers around the path (bricks.py:593-603, 781-792 flat-sky; utils.py:785-921)
# ----------------------------------------------------------------------------------------------------------------
def regular_pos(mesh_shape, ptcl_shape=None):
    """bricks.py:593-603."""
    ptcl_shape = mesh_shape if ptcl_shape is None else ptcl_shape
    ax = [np.linspace(0, m, p, endpoint=False) for m, p in zip(mesh_shape, ptcl_shape)]
    return _t(np.stack(np.meshgrid(*ax, indexing="ij"), axis=-1).reshape(-1, 3))
