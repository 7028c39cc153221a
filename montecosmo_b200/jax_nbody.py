"""
JAX face of the engine: the callables of `montecosmo/nbody.py` that `model.py:23-25` / `bricks.py:10` import, as
`jax.ffi.ffi_call`s into the XLA-FFI handlers of `montecosmo_b200/xla/mcpm_xla.cc` with `jax.custom_vjp` rules, under the
reference's names and signatures (nbody.py:365 paint, 398 read, 315 deconv_paint, 532 nufft, 583 pm_forces, 607 pm_forces2,
634 lpt, 967 nbody_bf; utils.py:975 chreshape).

STATUS: NOT EXERCISED in this repository's image -- neither jax nor jaxlib's FFI headers are installed, so
importing this module raises ImportError here and `mcpm_xla.cc` is not built.  It is the binding a montecosmo maintainer
adds; what IS tested here is the same C ABI through ctypes (`_capi.py`, `nbody.py` with torch) and, by
`tests/test_xla_shim.py`, that every handler named below exists in the shim with matching attributes and that every
`mcpm_*` call of the shim matches the header's prototype.

Usage in montecosmo:   from montecosmo_b200.jax_nbody import paint, read, nufft, lpt, nbody_bf, ...   (float32 on CUDA).
Growth factors and BullFrog coefficients stay the reference's own JAX functions (nbody.py:679-808, 907-919), so
`jax.grad` still reaches the cosmology: the engine returns cotangents for every scalar coefficient it consumes.
"""
from __future__ import annotations

import ctypes
import os
from functools import partial

import numpy as np

try:
    import jax
    import jax.numpy as jnp
except ImportError as e:  # pragma: no cover - jax is absent from this image
    raise ImportError("montecosmo_b200.jax_nbody needs jax with a CUDA-12.8+ jaxlib (jax.ffi); this image has none -- "
                      "use montecosmo_b200.nbody (torch) or the C ABI through ctypes") from e

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = ctypes.CDLL(os.path.join(_HERE, "xla", "libmcpm_xla.so"))  # built from xla/mcpm_xla.cc against jaxlib's headers

HANDLERS = {
    "mcpm_paint": "McpmPaint", "mcpm_paint_vjp": "McpmPaintVjp", "mcpm_read": "McpmRead", "mcpm_read_grad": "McpmReadGrad",
    "mcpm_paint3": "McpmPaint3", "mcpm_rfftn": "McpmRfftn", "mcpm_irfftn": "McpmIrfftn",
    "mcpm_hermitian_weights": "McpmHermitianWeights", "mcpm_deconv": "McpmDeconv", "mcpm_chreshape": "McpmChreshape",
    "mcpm_chreshape_vjp": "McpmChreshapeVjp", "mcpm_rg2cgh": "McpmRg2cgh", "mcpm_rg2cgh_vjp": "McpmRg2cghVjp",
    "mcpm_cgh2rg": "McpmCgh2rg", "mcpm_scale_spectrum": "McpmScaleSpectrum", "mcpm_spectrum_bins": "McpmSpectrumBins",
    "mcpm_nufft": "McpmNufft", "mcpm_nufft_vjp": "McpmNufftVjp", "mcpm_pm_forces": "McpmPmForces",
    "mcpm_pm_forces_vjp": "McpmPmForcesVjp", "mcpm_pm_forces_mesh": "McpmPmForcesMesh", "mcpm_pm_forces2": "McpmPmForces2",
    "mcpm_lpt": "McpmLpt", "mcpm_lpt_vjp": "McpmLptVjp", "mcpm_nbody_steps": "McpmNbodySteps",
    "mcpm_nbody_steps_vjp": "McpmNbodyStepsVjp",
    "mcpm_nufft_rsd": "McpmNufftRsd", "mcpm_nufft_rsd_vjp": "McpmNufftRsdVjp", "mcpm_bias_spectra": "McpmBiasSpectra",
    "mcpm_bias_spectra_vjp": "McpmBiasSpectraVjp", "mcpm_shear_invariants": "McpmShearInvariants",
    "mcpm_shear_invariants_vjp": "McpmShearInvariantsVjp", "mcpm_bias_weights": "McpmBiasWeights",
    "mcpm_bias_weights_vjp": "McpmBiasWeightsVjp", "mcpm_nufft_obs": "McpmNufftObs", "mcpm_nufft_obs_vjp": "McpmNufftObsVjp",
    "mcpm_radial_tables": "McpmRadialTables", "mcpm_radial_tables_vjp": "McpmRadialTablesVjp",
}
for _name, _sym in HANDLERS.items():
    jax.ffi.register_ffi_target(_name, jax.ffi.pycapsule(getattr(_SHIM, _sym)), platform="CUDA")

f32, c64, i32 = jnp.float32, jnp.complex64, np.int32
_NONE = lambda dt=f32: jnp.zeros((0,), dt)  # an absent optional array
_ONES = np.ones(3, np.float32)
_NOLAT = np.zeros(3, np.int32)


def _sds(shape, dt=f32):
    return jax.ShapeDtypeStruct(tuple(int(s) for s in shape), dt)


def _fd(fd):  # np.inf / 2 / 4 -> MCPM_FD_*
    return i32(0 if fd == np.inf else int(fd))


def _kcut(kernel_type, oversamp):
    if kernel_type == "rectangular":
        return np.float32(0.0)
    if kernel_type == "kaiser_bessel":
        return np.float32(0.98 * np.pi * (2.0 - 1.0 / oversamp))  # optim_kcut, nbody.py:357-363
    raise ValueError(f"Unknown kernel type: {kernel_type}")


def _weights(weights, n):
    w = jnp.asarray(weights, f32)
    return jnp.broadcast_to(w, (n,)) if w.ndim == 0 else w


def r2chshape(shape):
    return (*shape[:-1], shape[-1] // 2 + 1)


def ch2rshape(shape):
    return (*shape[:-1], 2 * (shape[-1] - 1))


# ------------------------------------------------------------------------------------------------ paint / read
@partial(jax.custom_vjp, nondiff_argnums=(1, 3, 4, 5))
def paint(pos, shape: tuple, weights=1.0, order: int = 2, kernel_type="rectangular", oversamp=1.0):
    """nbody.py:365-396."""
    w = _weights(weights, pos.shape[0])
    return jax.ffi.ffi_call("mcpm_paint", _sds(shape))(pos.astype(f32), w, wscalar=np.float32(1.0), order=i32(order),
                                                       kcut=_kcut(kernel_type, oversamp), scale=_ONES, shift=np.float32(0))


def _paint_fwd(pos, shape, weights, order, kernel_type, oversamp):
    return paint(pos, shape, weights, order, kernel_type, oversamp), (pos, _weights(weights, pos.shape[0]), jnp.ndim(weights))


def _paint_bwd(shape, order, kernel_type, oversamp, res, mbar):
    pos, w, wnd = res
    pb, wb = jax.ffi.ffi_call("mcpm_paint_vjp", (_sds(pos.shape), _sds(pos.shape[:1])))(
        pos.astype(f32), w, mbar.astype(f32), wscalar=np.float32(1.0), order=i32(order), kcut=_kcut(kernel_type, oversamp),
        scale=_ONES, shift=np.float32(0))
    return pb, (wb.sum() if wnd == 0 else wb)


paint.defvjp(_paint_fwd, _paint_bwd)


@partial(jax.custom_vjp, nondiff_argnums=(2, 3, 4))
def read(pos, mesh, order: int = 2, kernel_type="rectangular", oversamp=1.0):
    """nbody.py:398-427."""
    return jax.ffi.ffi_call("mcpm_read", _sds(pos.shape[:1]))(pos.astype(f32), mesh.astype(f32), order=i32(order),
                                                             kcut=_kcut(kernel_type, oversamp), scale=_ONES,
                                                             shift=np.float32(0))


def _read_fwd(pos, mesh, order, kernel_type, oversamp):
    return read(pos, mesh, order, kernel_type, oversamp), (pos, mesh)


def _read_bwd(order, kernel_type, oversamp, res, obar):
    pos, mesh = res
    kc = _kcut(kernel_type, oversamp)
    pb = jax.ffi.ffi_call("mcpm_read_grad", _sds(pos.shape))(pos.astype(f32), mesh.astype(f32), obar.astype(f32),
                                                            order=i32(order), kcut=kc, scale=_ONES, shift=np.float32(0))
    mb = jax.ffi.ffi_call("mcpm_paint", _sds(mesh.shape))(pos.astype(f32), obar.astype(f32), wscalar=np.float32(1.0),
                                                         order=i32(order), kcut=kc, scale=_ONES, shift=np.float32(0))
    return pb, mb


read.defvjp(_read_fwd, _read_bwd)


# ------------------------------------------------------------------------------------------------ FFT, Fourier helpers
@jax.custom_vjp
def rfftn(mesh):
    return jax.ffi.ffi_call("mcpm_rfftn", _sds(r2chshape(mesh.shape), c64))(mesh.astype(f32))


def _hw(meshk, mode):
    return jax.ffi.ffi_call("mcpm_hermitian_weights", _sds(meshk.shape, c64))(meshk, mode=i32(mode))


rfftn.defvjp(lambda m: (rfftn(m), None),
             lambda _, kbar: (irfftn_raw(_hw(jnp.conj(kbar), 0)),))  # conj: JAX cotangent -> dL/dRe + i dL/dIm


def irfftn_raw(meshk):
    return jax.ffi.ffi_call("mcpm_irfftn", _sds(ch2rshape(meshk.shape)))(meshk.astype(c64))


@jax.custom_vjp
def irfftn(meshk):
    return irfftn_raw(meshk)


irfftn.defvjp(lambda k: (irfftn_raw(k), None), lambda _, mbar: (jnp.conj(_hw(rfftn(mbar), 1)),))


def deconv_paint(mesh, order: int = 2, kernel_type="rectangular", oversamp=1.0):
    """nbody.py:315-334 (a real diagonal factor: linear, self-adjoint)."""
    kc = _kcut(kernel_type, oversamp)
    f = lambda k: jax.ffi.ffi_call("mcpm_deconv", _sds(k.shape, c64))(k.astype(c64), order=i32(order), kcut=kc)
    lin = jax.custom_jvp(f)
    lin.defjvp(lambda primals, tangents: (f(primals[0]), f(tangents[0])))
    if not jnp.iscomplexobj(mesh):
        return irfftn(lin(rfftn(mesh)))
    return lin(mesh)


@partial(jax.custom_vjp, nondiff_argnums=(1,))
def chreshape(mesh, shape):
    """utils.py:975-1013; `shape` is the complex target shape."""
    return jax.ffi.ffi_call("mcpm_chreshape", _sds(shape, c64))(mesh.astype(c64))


chreshape.defvjp(lambda m, shape: (chreshape(m, shape), m.shape),
                 lambda shape, in_shape, kbar: (jnp.conj(jax.ffi.ffi_call("mcpm_chreshape_vjp", _sds(in_shape, c64))(
                     jnp.conj(kbar))),))


# ------------------------------------------------------------------------------------------------ nufft
@partial(jax.custom_vjp, nondiff_argnums=(1, 3, 4, 5, 6, 7, 8))
def _nufft_paint(pos, paint_shape, weights, scale, paint_order, interlace_order, kernel_type, paint_deconv, lattice):
    kc = _kcut(kernel_type, float(np.exp(np.log(1.0 / np.asarray(scale)).mean())))
    return jax.ffi.ffi_call("mcpm_nufft", _sds(r2chshape(paint_shape), c64))(
        pos.astype(f32), weights, wscalar=np.float32(1.0), scale=np.asarray(scale, np.float32), paint_order=i32(paint_order),
        kcut=kc, interlace_order=i32(interlace_order), paint_deconv=i32(paint_deconv),
        lattice=_NOLAT if lattice is None else np.asarray(lattice, np.int32), relative=i32(lattice is not None))


def _nufft_fwd(pos, paint_shape, weights, *cfg):
    return _nufft_paint(pos, paint_shape, weights, *cfg), (pos, weights)


def _nufft_bwd(paint_shape, scale, paint_order, interlace_order, kernel_type, paint_deconv, lattice, res, kbar):
    pos, weights = res
    kc = _kcut(kernel_type, float(np.exp(np.log(1.0 / np.asarray(scale)).mean())))
    pb, wb = jax.ffi.ffi_call("mcpm_nufft_vjp", (_sds(pos.shape), _sds(pos.shape[:1])))(
        pos.astype(f32), weights, jnp.conj(kbar), wscalar=np.float32(1.0), scale=np.asarray(scale, np.float32),
        paint_order=i32(paint_order), kcut=kc, interlace_order=i32(interlace_order), paint_deconv=i32(paint_deconv),
        lattice=_NOLAT if lattice is None else np.asarray(lattice, np.int32), relative=i32(lattice is not None))
    return pb, wb


_nufft_paint.defvjp(_nufft_fwd, _nufft_bwd)


# the same paint with the flat-sky redshift-space shift pos + (vel . los) coef los applied inside the kernels
@partial(jax.custom_vjp, nondiff_argnums=(2, 4, 5, 6, 7, 8, 9, 10))
def _nufft_paint_rsd(pos, vel, paint_shape, weights, los, coef, scale, paint_order, interlace_order, paint_deconv, lattice):
    return jax.ffi.ffi_call("mcpm_nufft_rsd", _sds(r2chshape(paint_shape), c64))(
        pos.astype(f32), vel.astype(f32), weights, los=np.asarray(los, np.float32), coef=np.float32(coef),
        wscalar=np.float32(1.0), scale=np.asarray(scale, np.float32), paint_order=i32(paint_order),
        interlace_order=i32(interlace_order), paint_deconv=i32(paint_deconv),
        lattice=_NOLAT if lattice is None else np.asarray(lattice, np.int32), relative=i32(lattice is not None))


def _nufft_rsd_fwd(pos, vel, paint_shape, weights, *cfg):
    return _nufft_paint_rsd(pos, vel, paint_shape, weights, *cfg), (pos, vel, weights)


def _nufft_rsd_bwd(paint_shape, los, coef, scale, paint_order, interlace_order, paint_deconv, lattice, res, kbar):
    pos, vel, weights = res
    pb, vb, wb = jax.ffi.ffi_call("mcpm_nufft_rsd_vjp", (_sds(pos.shape), _sds(pos.shape), _sds(pos.shape[:1])))(
        pos.astype(f32), vel.astype(f32), weights, jnp.conj(kbar), los=np.asarray(los, np.float32), coef=np.float32(coef),
        wscalar=np.float32(1.0), scale=np.asarray(scale, np.float32), paint_order=i32(paint_order),
        interlace_order=i32(interlace_order), paint_deconv=i32(paint_deconv),
        lattice=_NOLAT if lattice is None else np.asarray(lattice, np.int32), relative=i32(lattice is not None))
    return pb, vb, wb


_nufft_paint_rsd.defvjp(_nufft_rsd_fwd, _nufft_rsd_bwd)


def nufft(pos, final_shape: tuple, paint_shape=None, weights=1.0, paint_order: int = 2, interlace_order: int = 2,
          kernel_type="rectangular", paint_deconv=True, lattice=None, rsd=None):
    """nbody.py:532-577.  `lattice` (extension): pos holds displacements from the sites of that regular lattice.
    `rsd` (extension) = (vel, los, coef) with static los / coef: paint pos + (vel . los) coef los (model.py:780-809)."""
    final_shape = tuple(int(s) for s in final_shape)
    if paint_shape is None:
        paint_shape = final_shape
    elif isinstance(paint_shape, float):
        paint_shape = tuple(int(2 * round(s * paint_shape / 2)) for s in final_shape)  # scale_shape, utils.py:1163-1168
    paint_shape = tuple(int(s) for s in paint_shape)
    scale = tuple(p / f for p, f in zip(paint_shape, final_shape))
    if rsd is not None:
        vel, los, coef = rsd
        mesh = _nufft_paint_rsd(pos, vel, paint_shape, _weights(weights, pos.shape[0]), tuple(float(x) for x in los),
                                float(coef), scale, paint_order, interlace_order, bool(paint_deconv),
                                None if lattice is None else tuple(lattice))
        return mesh if final_shape == paint_shape else chreshape(mesh, r2chshape(final_shape))
    mesh = _nufft_paint(pos, paint_shape, _weights(weights, pos.shape[0]), scale, paint_order, interlace_order,
                        kernel_type, bool(paint_deconv), None if lattice is None else tuple(lattice))
    return mesh if final_shape == paint_shape else chreshape(mesh, r2chshape(final_shape))


# the same paint of the positions as the observer sees them: the chain model.py:780-799 builds between the evolution and
# the paint (frames, line of sight, light-cone scale factor, redshift-space distortion, Alcock-Paczynski) inside the kernels
OBS_SLOTS = 32  # MCPM_OBS_SLOTS


def _obs_attrs(static, scale, paint_order, kcut, interlace_order, paint_deconv, lattice):
    flags = np.asarray([static["curved"], static["lightcone"], static["ap"], static["rsd"]], np.int32)
    geom = np.asarray([*static["cell"], *static["origin"], *static["los"], static["r0"], static["dr"]], np.float32)
    return dict(flags=flags, geom=geom, rot=np.asarray(static["rot"], np.float32).reshape(9), wscalar=np.float32(static["wscalar"]),
                scale=np.asarray(scale, np.float32), paint_order=i32(paint_order), kcut=np.float32(kcut),
                interlace_order=i32(interlace_order), paint_deconv=i32(paint_deconv),
                lattice=_NOLAT if lattice is None else np.asarray(lattice, np.int32), relative=i32(lattice is not None))


@partial(jax.custom_vjp, nondiff_argnums=(7, 8, 9, 10, 11, 12, 13, 14))
def _nufft_paint_obs(pos, vel, dvel, weights, par, tab_gf, tab_ap, static, paint_shape, scale, paint_order, kcut,
                     interlace_order, paint_deconv, lattice):
    """`static`: a hashable tuple of (key, value) pairs -- curved, lightcone, ap, rsd, cell, origin, los, rot, r0, dr,
    wscalar (bricks.observation's geometry); par = (D f, a_par, a_perp) and the tables are traced float32 arrays."""
    attrs = _obs_attrs(dict(static), scale, paint_order, kcut, interlace_order, paint_deconv, lattice)
    return jax.ffi.ffi_call("mcpm_nufft_obs", _sds(r2chshape(paint_shape), c64))(
        pos.astype(f32), vel.astype(f32), dvel.astype(f32), weights, par.astype(f32), tab_gf.astype(f32), tab_ap.astype(f32),
        **attrs)


def _nufft_obs_fwd(pos, vel, dvel, weights, par, tab_gf, tab_ap, *cfg):
    return _nufft_paint_obs(pos, vel, dvel, weights, par, tab_gf, tab_ap, *cfg), (pos, vel, dvel, weights, par, tab_gf, tab_ap)


def _nufft_obs_bwd(static, paint_shape, scale, paint_order, kcut, interlace_order, paint_deconv, lattice, res, kbar):
    pos, vel, dvel, weights, par, tab_gf, tab_ap = res
    nt = max(tab_gf.shape[0], tab_ap.shape[0])
    attrs = _obs_attrs(dict(static), scale, paint_order, kcut, interlace_order, paint_deconv, lattice)
    pb, vb, db, wb, parbar = jax.ffi.ffi_call(
        "mcpm_nufft_obs_vjp", (_sds(pos.shape), _sds(vel.shape), _sds(dvel.shape), _sds(weights.shape),
                               _sds((OBS_SLOTS, 3 + 2 * nt), jnp.float64)))(
        pos.astype(f32), vel.astype(f32), dvel.astype(f32), weights, par.astype(f32), tab_gf.astype(f32), tab_ap.astype(f32),
        jnp.conj(kbar), **attrs)
    row = parbar[0]
    return (pb, vb, db, wb, row[:3].astype(par.dtype), row[3:3 + tab_gf.shape[0]].astype(tab_gf.dtype),
            row[3 + nt:3 + nt + tab_ap.shape[0]].astype(tab_ap.dtype))


_nufft_paint_obs.defvjp(_nufft_obs_fwd, _nufft_obs_bwd)


def nufft_observed(pos, vel, final_shape: tuple, obs: dict, paint_shape=None, weights=1.0, dvel=None, paint_order: int = 2,
                   interlace_order: int = 2, kernel_type="rectangular", paint_deconv=True, lattice=None, pos_shape=None):
    """The signature of montecosmo_b200.nbody.nufft_observed: `obs` holds the static geometry and the traced `par`,
    `tab_gf`, `tab_ap` (what bricks.observation computes from the cosmology, here with jax_cosmo by the caller)."""
    final_shape = tuple(int(s) for s in final_shape)
    if paint_shape is None:
        paint_shape = final_shape
    elif isinstance(paint_shape, float):
        paint_shape = tuple(int(2 * round(s * paint_shape / 2)) for s in final_shape)
    paint_shape = tuple(int(s) for s in paint_shape)
    pos_shape = final_shape if pos_shape is None else tuple(int(s) for s in pos_shape)
    scale = tuple(p / f for p, f in zip(paint_shape, pos_shape))
    kc = _kcut(kernel_type, float(np.exp(np.log(np.divide(final_shape, paint_shape)).mean())))
    static = {k: (tuple(map(tuple, v)) if k == "rot" else (tuple(v) if isinstance(v, (list, tuple, np.ndarray)) else v))
              for k, v in obs.items() if k not in ("par", "tab_gf", "tab_ap")}
    static["wscalar"] = float(np.divide(pos_shape, final_shape).prod())  # the Jacobian is final -> paint units
    none = _NONE()
    mesh = _nufft_paint_obs(pos, none if vel is None else vel, none if dvel is None else dvel, _weights(weights, pos.shape[0]),
                            jnp.asarray(obs["par"]), obs.get("tab_gf", none), obs.get("tab_ap", none),
                            tuple(sorted(static.items())), paint_shape, scale, paint_order, kc, interlace_order,
                            bool(paint_deconv), None if lattice is None else tuple(lattice))
    return mesh if final_shape == paint_shape else chreshape(mesh, r2chshape(final_shape))


# functions of the comoving distance at the particles (the light cone): tabs [K, nt] on the radius grid of `geom`
@partial(jax.custom_vjp, nondiff_argnums=(2,))
def _radial_tables(pos, tabs, geom):
    g = dict(geom)
    return jax.ffi.ffi_call("mcpm_radial_tables", _sds((pos.shape[0], tabs.shape[0])))(
        pos.astype(f32), tabs.astype(f32), flags=np.asarray([g["curved"], 0, 0, 0], np.int32),
        geom=np.asarray([*g["cell"], *g["origin"], *g["los"], g["r0"], g["dr"]], np.float32))


def _radial_fwd(pos, tabs, geom):
    return _radial_tables(pos, tabs, geom), (pos, tabs)


def _radial_bwd(geom, res, outbar):
    pos, tabs = res
    g = dict(geom)
    pb, tb = jax.ffi.ffi_call("mcpm_radial_tables_vjp", (_sds(pos.shape), _sds((OBS_SLOTS, *tabs.shape), jnp.float64)))(
        pos.astype(f32), tabs.astype(f32), outbar.astype(f32), flags=np.asarray([g["curved"], 0, 0, 0], np.int32),
        geom=np.asarray([*g["cell"], *g["origin"], *g["los"], g["r0"], g["dr"]], np.float32))
    return pb.astype(pos.dtype), tb[0].astype(tabs.dtype)


_radial_tables.defvjp(_radial_fwd, _radial_bwd)


def radial_tables(pos, tabs, geom: dict):
    """montecosmo_b200.nbody.radial_tables: [Np, K] = the K tables at every particle's distance (light-cone growth)."""
    return _radial_tables(pos, tabs, tuple(sorted((k, tuple(v) if isinstance(v, (list, tuple, np.ndarray)) else v)
                                                  for k, v in geom.items() if k != "rot")))


# ------------------------------------------------------------------------------------------------ forces, lpt
@partial(jax.custom_vjp, nondiff_argnums=(1, 2, 3, 4, 5, 6))
def _pm_forces_paint(pos, shape, order, paint_deconv, lap_fd, grad_fd, kcut):
    return _pm_forces_call(pos, shape, order, paint_deconv, lap_fd, grad_fd, kcut)[0]


def _pm_forces_call(pos, shape, order, paint_deconv, lap_fd, grad_fd, kcut):
    return jax.ffi.ffi_call("mcpm_pm_forces", (_sds(pos.shape), _sds((3, *shape))))(
        pos.astype(f32), order=i32(order), paint_deconv=i32(paint_deconv), lap_fd=_fd(lap_fd), grad_fd=_fd(grad_fd),
        kcut=np.float32(0.0 if kcut == np.inf else kcut), lattice=_NOLAT, relative=i32(0))


def _pmf_fwd(pos, *cfg):
    forces, fm = _pm_forces_call(pos, *cfg)
    return forces, (pos, fm)


def _pmf_bwd(shape, order, paint_deconv, lap_fd, grad_fd, kcut, res, fbar):
    pos, fm = res
    return (jax.ffi.ffi_call("mcpm_pm_forces_vjp", _sds(pos.shape))(
        pos.astype(f32), fbar.astype(f32), fm, order=i32(order), paint_deconv=i32(paint_deconv), lap_fd=_fd(lap_fd),
        grad_fd=_fd(grad_fd), kcut=np.float32(0.0 if kcut == np.inf else kcut), lattice=_NOLAT, relative=i32(0)),)


_pm_forces_paint.defvjp(_pmf_fwd, _pmf_bwd)


@partial(jax.custom_vjp, nondiff_argnums=(3, 4, 5, 6))
def _lpt(dk, coef, pos, lpt_order, read_order, lap_fd, grad_fd):
    return _lpt_call(dk, coef, pos, lpt_order, read_order, lap_fd, grad_fd)[:2]


def _lpt_call(dk, coef, pos, lpt_order, read_order, lap_fd, grad_fd):
    """The growth coefficients are traced values (functions of the cosmology): they reach the handler as attributes only
    when concrete; under jit they are passed through jax.pure_callback-free host attributes by closing over them with
    jax.ensure_compile_time_eval, as the reference's scalar `a` is static in every call site (model.py:763, nbody.py:984)."""
    n, rs = pos.shape[0], ch2rshape(dk.shape)
    c = [np.float32(v) for v in np.asarray(coef)]
    tape2 = lpt_order == 2
    outs = (_sds((n, 3)), _sds((n, 3)), _sds((n, 3)), _sds((n, 3) if tape2 else (0,)), _sds((6, *rs) if tape2 else (0,)))
    return jax.ffi.ffi_call("mcpm_lpt", outs)(dk.astype(c64), pos.astype(f32), lpt_order=i32(lpt_order),
                                             read_order=i32(read_order), lap_fd=_fd(lap_fd), grad_fd=_fd(grad_fd),
                                             d1=c[0], d2=c[1], dv2=c[2])


def _lpt_fwd(dk, coef, pos, *cfg):
    dpos, vel, f1, f2, h6 = _lpt_call(dk, coef, pos, *cfg)
    return (dpos, vel), (dk.shape, coef, pos, f1, f2, h6)


def _lpt_bwd(lpt_order, read_order, lap_fd, grad_fd, res, bars):
    cshape, coef, pos, f1, f2, h6 = res
    c = [np.float32(v) for v in np.asarray(coef)]
    dkbar, cb = jax.ffi.ffi_call("mcpm_lpt_vjp", (_sds(cshape, c64), _sds((3,), jnp.float64)))(
        pos.astype(f32), bars[0].astype(f32), bars[1].astype(f32), f1, f2, h6, lpt_order=i32(lpt_order),
        read_order=i32(read_order), lap_fd=_fd(lap_fd), grad_fd=_fd(grad_fd), d1=c[0], d2=c[1], dv2=c[2])
    return jnp.conj(dkbar), cb.astype(coef.dtype), jnp.zeros_like(pos)


_lpt.defvjp(_lpt_fwd, _lpt_bwd)


def pm_forces(pos, mesh, read_order: int = 2, paint_deconv: bool = False, grad_fd=np.inf, lap_fd=np.inf, kcut=np.inf):
    """nbody.py:583-604: `mesh` a shape tuple (paint pos) or delta_k."""
    if isinstance(mesh, tuple):
        return _pm_forces_paint(pos, tuple(mesh), read_order, bool(paint_deconv), lap_fd, grad_fd, kcut)
    # vel of a first-order lpt with unit coefficients is exactly pm_forces(pos, delta_k)
    return _lpt(mesh, jnp.zeros(3), pos, 1, read_order, lap_fd, grad_fd)[1]


def pm_forces2(pos, mesh, read_order: int = 2, grad_fd=np.inf, lap_fd=np.inf):
    """nbody.py:607-631 (dpos = d1 F1 - d2 F2 with d1 = 0, d2 = -1)."""
    return _lpt(mesh, jnp.array([0.0, -1.0, 0.0]), pos, 2, read_order, lap_fd, grad_fd)[0]


def lpt(cosmo, init_mesh, pos, a, lpt_order: int = 2, read_order: int = 2, grad_fd=np.inf, lap_fd=np.inf):
    """nbody.py:634-667, scalar `a` (per-particle `a`: pm_forces / pm_forces2 times the growth factors, as there)."""
    from montecosmo.nbody import a2dg2dg, a2g, a2g2  # the reference's own growth helpers (nbody.py:750-777)
    if not jnp.iscomplexobj(init_mesh):
        init_mesh = rfftn(init_mesh)
    if jnp.ndim(a) != 0:
        f1 = pm_forces(pos, init_mesh, read_order, grad_fd=grad_fd, lap_fd=lap_fd)
        dpos, vel = a2g(cosmo, a) * f1, f1
        if lpt_order == 2:
            f2 = pm_forces2(pos, init_mesh, read_order, grad_fd=grad_fd, lap_fd=lap_fd)
            dpos, vel = dpos - a2g2(cosmo, a) * f2, vel - a2dg2dg(cosmo, a) * f2
        return dpos, vel
    coef = jnp.stack([a2g(cosmo, a), a2g2(cosmo, a), a2dg2dg(cosmo, a)])
    return _lpt(init_mesh, coef, pos, lpt_order, read_order, lap_fd, grad_fd)


# ------------------------------------------------------------------------------------------------ nbody_bf
@partial(jax.custom_vjp, nondiff_argnums=(3, 4, 5, 6, 7, 8))
def _nbody_steps(pos, vel, coefs, shape, order, paint_deconv, lap_fd, grad_fd, lattice):
    return _steps_call(pos, vel, coefs, shape, order, paint_deconv, lap_fd, grad_fd, lattice, False)[:2]


def _steps_attrs(coefs, shape, order, paint_deconv, lap_fd, grad_fd, lattice):
    co = np.asarray(coefs, np.float32)  # [n_steps, 4]: alpha, beta, drift_pre, drift_post (host scalars of the cosmology)
    return dict(mesh=np.asarray(shape, np.int32), alpha=co[:, 0].copy(), beta=co[:, 1].copy(), drift_pre=co[:, 2].copy(),
                drift_post=co[:, 3].copy(), order=i32(order), paint_deconv=i32(paint_deconv), lap_fd=_fd(lap_fd),
                grad_fd=_fd(grad_fd), lattice=_NOLAT if lattice is None else np.asarray(lattice, np.int32),
                relative=i32(lattice is not None))


def _steps_call(pos, vel, coefs, shape, order, paint_deconv, lap_fd, grad_fd, lattice, tape):
    n, ns = pos.shape[0], int(np.shape(coefs)[0])
    outs = (_sds((n, 3)), _sds((n, 3)), _sds((ns, n, 3) if tape else (0,)), _sds((ns, n, 3) if tape else (0,)),
            _sds((ns, 4, *shape) if tape else (0,)))
    return jax.ffi.ffi_call("mcpm_nbody_steps", outs, input_output_aliases={0: 0, 1: 1})(
        pos.astype(f32), vel.astype(f32), **_steps_attrs(coefs, shape, order, paint_deconv, lap_fd, grad_fd, lattice))


def _steps_fwd(pos, vel, coefs, *cfg):
    p, v, xk, vk, fm = _steps_call(pos, vel, coefs, *cfg, True)
    return (p, v), (xk, vk, fm, vel, coefs)


def _steps_bwd(shape, order, paint_deconv, lap_fd, grad_fd, lattice, res, bars):
    xk, vk, fm, v0, coefs = res
    n, ns = v0.shape[0], int(np.shape(coefs)[0])
    pb, vb, cb = jax.ffi.ffi_call("mcpm_nbody_steps_vjp", (_sds((n, 3)), _sds((n, 3)), _sds((ns, 4), jnp.float64)),
                                  input_output_aliases={0: 0, 1: 1})(
        bars[0].astype(f32), bars[1].astype(f32), xk, vk, fm, v0,
        **_steps_attrs(coefs, shape, order, paint_deconv, lap_fd, grad_fd, lattice))
    return pb, vb, cb.astype(jnp.asarray(coefs).dtype)


_nbody_steps.defvjp(_steps_fwd, _steps_bwd)


def alpha_bf(cosmo, g0, dg):
    """BullFrog kick coefficient: the closure of bullfrog_vf (nbody.py:907-919) lifted to module level, on the
    reference's own growth helpers g2g2 / g2dg2dg (nbody.py:790-808) so that it stays differentiable in the cosmology."""
    from montecosmo.nbody import g2dg2dg, g2g2
    g1 = g0 + dg / 2
    dg2dg0, dg2dg2 = g2dg2dg(cosmo, g0), g2dg2dg(cosmo, g0 + dg)
    lin_ratio = (g2g2(cosmo, g0) + dg2dg0 * dg / 2) / g1 - g1
    return (dg2dg2 - lin_ratio) / (dg2dg0 - lin_ratio)


def nbody_bf(cosmo, init_mesh, pos, a0=0.0, a1=1.0, n_steps=5, paint_order: int = 2, lpt_order: int = 2,
             paint_deconv=False, grad_fd=np.inf, lap_fd=np.inf, snapshots=None, fn=None, ptcl_shape=None):
    """nbody.py:967-1002: lpt at a0, then n_steps BullFrog DKD steps in growth time -- ONE custom call for the loop in
    place of diffeqsolve(Euler) (nbody.py:999).  Returns (pos, vel), each [1, Np, 3] (snapshots=None).
    `ptcl_shape` (extension): `pos` is regular_pos of that lattice; the loop then carries displacements from the lattice
    sites (mcpm_engine_set_relative) and `pos` is added back at the end."""
    from montecosmo.nbody import a2g
    if snapshots is not None:
        raise NotImplementedError("snapshots: one _nbody_steps call per segment between the step boundaries that bracket "
                                  "a save time, as montecosmo_b200.nbody.nbody_bf does")
    mesh_shape = ch2rshape(init_mesh.shape)
    g0, g1 = a2g(cosmo, a0), a2g(cosmo, a1)
    dg = (g1 - g0) / n_steps
    dpos, vel = lpt(cosmo, init_mesh, pos, a0, lpt_order, 1, grad_fd, lap_fd)
    rows = []
    for s in range(n_steps):  # per step: alpha, the kick weight (1 - alpha) / g_kick (nbody.py:946-950), two half drifts
        gs = g0 + s * dg
        alpha = alpha_bf(cosmo, gs, dg)
        rows.append(jnp.stack([alpha, (1 - alpha) / (gs + dg / 2), dg / 2, dg / 2]))
    coefs = jnp.stack(rows)
    lattice = None if ptcl_shape is None else tuple(int(s) for s in ptcl_shape)
    x0 = dpos if lattice is not None else pos + dpos
    x, v = _nbody_steps(x0, vel, coefs, mesh_shape, paint_order, bool(paint_deconv), lap_fd, grad_fd, lattice)
    if lattice is not None:
        x = x + pos
    return x[None], v[None]


# ------------------------------------------------------------------------------------------------ lagrangian_bias
# bricks.py:327-452 on the fused passes of csrc/bias.cu (same chain as montecosmo_b200/bricks.py): Fourier multipliers
# -> batched irfftn -> shear invariants -> one 7 | 9-mesh read -> polynomial; every stage a custom_vjp.
@partial(jax.custom_vjp, nondiff_argnums=(2,))
def _bias_spectra(dk, inv_transfer, cpl):
    m = 12 if inv_transfer.size else 10
    return jax.ffi.ffi_call("mcpm_bias_spectra", _sds((m, *dk.shape), c64))(dk, inv_transfer,
                                                                             cells_per_len=np.asarray(cpl, np.float32))


def _bias_spectra_bwd(cpl, res, obar):
    dk_shape, inv_transfer = res
    dkbar = jax.ffi.ffi_call("mcpm_bias_spectra_vjp", _sds(dk_shape, c64))(jnp.conj(obar), inv_transfer,
                                                                           cells_per_len=np.asarray(cpl, np.float32))
    return jnp.conj(dkbar), jnp.zeros_like(inv_transfer)


_bias_spectra.defvjp(lambda dk, it, cpl: (_bias_spectra(dk, it, cpl), (dk.shape, it)), _bias_spectra_bwd)


@jax.custom_vjp
def _shear_invariants(s5):
    return jax.ffi.ffi_call("mcpm_shear_invariants", _sds((2, *s5.shape[1:])))(s5)


_shear_invariants.defvjp(lambda s5: (_shear_invariants(s5), s5),
                         lambda s5, ob: (jax.ffi.ffi_call("mcpm_shear_invariants_vjp", _sds(s5.shape))(s5, ob.astype(f32)),))


@jax.custom_vjp
def _bias_weights(vals, growth, coef):
    """vals [np, K], growth (scalar), coef [13] -> (weights [np], dvel [np, 3]).  growth and coef are traced scalars passed
    as attributes: call under jit with them static, or use the per-stage calls directly."""
    w, dv, _ = _bias_weights_call(vals, growth, coef)
    return w, dv


def _bias_weights_call(vals, growth, coef):
    n = vals.shape[0]
    return jax.ffi.ffi_call("mcpm_bias_weights", (_sds((n,)), _sds((n, 3)), _sds((2,), jnp.float64)))(
        vals, _NONE(), growth=np.float32(growth), coef=np.asarray(coef, np.float32))


def _bias_weights_fwd(vals, growth, coef):
    w, dv, mom = _bias_weights_call(vals, growth, coef)
    return (w, dv), (vals, growth, coef, mom)


def _bias_weights_bwd(res, bars):
    vals, growth, coef, mom = res
    wbar, dvbar = bars
    n = vals.shape[0]
    vb, cb, _, _ = jax.ffi.ffi_call("mcpm_bias_weights_vjp", (_sds(vals.shape), _sds((14,), jnp.float64), _sds((n,)),
                                                              _sds((2,), jnp.float64)))(
        vals, _NONE(), mom, wbar.astype(f32), dvbar.astype(f32), growth=np.float32(growth), coef=np.asarray(coef, np.float32))
    return vb, cb[13], cb[:13]


_bias_weights.defvjp(_bias_weights_fwd, _bias_weights_bwd)


def lagrangian_bias(growth, pos, box_size, lin_mesh, coef, inv_transfer=None, read_order: int = 2):
    """bricks.py:327-452: (weights, dvel, phi).  growth = a2g(cosmo, a) and coef = the 13 bias / PNG coefficients are the
    caller's JAX scalars (static under jit here: their cotangents come back from mcpm_bias_weights_vjp); inv_transfer =
    1 / trans_phi2delta on the half-spectrum mesh for the PNG terms, else None."""
    shape = ch2rshape(lin_mesh.shape)
    cpl = tuple(float(n / l) for n, l in zip(shape, np.broadcast_to(np.asarray(box_size, float), (3,))))
    it = _NONE() if inv_transfer is None else inv_transfer.astype(f32)
    spec = _bias_spectra(lin_mesh.astype(c64), it, cpl)
    real = jnp.stack([irfftn(spec[m]) for m in range(spec.shape[0])])
    inv = _shear_invariants(real[:5])
    meshes = jnp.concatenate([real[5:6], inv, real[6:]], axis=0)  # delta, s^2, s^3, lap, grad (3), [phi, lap phi]
    vals = read(pos, meshes, read_order)
    weights, dvel = _bias_weights(vals, growth, coef)
    return weights, dvel, (meshes[7] if inv_transfer is not None else 0.0)
