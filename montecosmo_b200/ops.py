"""
Thin Python layer over the C ABI (include/mcpm.h): one method per entry point, arrays in / arrays out.

`Ops` is parametrised by an *array adapter* (allocation, pointer, stream).  The package instantiates it with the torch
CUDA adapter below -- device memory and streams are torch's, the arithmetic is libmcpm's.  The test-suite instantiates
the same class with a NumPy adapter over the host-emulation build (tests/hostemu.py) to check orchestration on CPU.
"""
import ctypes as C
import math

from . import _capi
from ._capi import check, fd_code, host_floats

INF = float("inf")


def r2chshape(shape):
    """Real shape -> half-spectrum shape (utils.py:778-782)."""
    return (*shape[:-1], shape[-1] // 2 + 1)


def ch2rshape(shape):
    """Half-spectrum shape -> real shape, last real side even (utils.py:769-776)."""
    return (*shape[:-1], 2 * (shape[-1] - 1))


class TorchCudaAdapter:
    """Device memory and streams from torch; float32 / complex64 contiguous CUDA tensors."""

    def __init__(self, device=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("montecosmo_b200 needs a CUDA device: there is no CPU fallback")
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)

    def empty(self, shape, dtype="f32"):
        t = self.torch
        dt = {"f32": t.float32, "c64": t.complex64, "f64": t.float64}[dtype]
        return t.empty(tuple(int(s) for s in shape), dtype=dt, device=self.device)

    def zeros(self, shape, dtype="f32"):
        return self.empty(shape, dtype).zero_()

    def prepare(self, x, dtype="f32"):
        t = self.torch
        dt = {"f32": t.float32, "c64": t.complex64, "f64": t.float64}[dtype]
        x = t.as_tensor(x, device=self.device)
        if x.dtype != dt:
            x = x.to(dt)
        return x.contiguous()

    def ptr(self, x):
        return 0 if x is None else x.data_ptr()

    def stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def shape(self, x):
        return tuple(x.shape)


class Engine:
    """Owner of one mcpm_engine (cuFFT plans + scratch for one mesh shape)."""

    def __init__(self, lib, shape):
        self.lib, self.shape = lib, tuple(int(s) for s in shape)
        h = C.c_void_p()
        check(lib, lib.mcpm_engine_create(*self.shape, C.byref(h)))
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.mcpm_engine_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class Ops:
    def __init__(self, lib, adapter):
        self.lib, self.A = lib, adapter
        self._engines = {}

    # ------------------------------------------------------------------------------------------------ helpers
    def engine(self, shape):
        shape = tuple(int(s) for s in shape)
        if shape not in self._engines:
            self._engines[shape] = Engine(self.lib, shape)
        return self._engines[shape]

    def set_lattice(self, mesh_shape, ptcl_shape):
        """Performance hint for the engine of `mesh_shape` (mcpm_engine_set_lattice); ptcl_shape None clears it."""
        p = (0, 0, 0) if ptcl_shape is None else tuple(int(s) for s in ptcl_shape)
        eng = self._frame(mesh_shape, None)
        self._call("mcpm_engine_set_lattice", eng.handle, *p)

    def _frame(self, shape, lattice):
        """Select what the `pos` arrays of the next composite call on the engine of `shape` mean: absolute positions
        (lattice None) or displacements from the sites of the regular lattice `lattice` (mcpm_engine_set_relative)."""
        eng = self.engine(shape)
        want = None if lattice is None else tuple(int(s) for s in lattice)
        if getattr(eng, "rel_state", None) == want:
            return eng
        if want is not None:
            self._call("mcpm_engine_set_lattice", eng.handle, *want)
        self._call("mcpm_engine_set_relative", eng.handle, int(want is not None))
        eng.rel_state = want
        return eng

    def set_fused_fft(self, mesh_shape, on):
        """Select the fused x-transform path (mcpm_engine_set_fused_fft); raises McpmError where unsupported."""
        self._call("mcpm_engine_set_fused_fft", self.engine(mesh_shape).handle, int(bool(on)))

    def _xf(self, scale, shift):
        return host_floats((1.0, 1.0, 1.0) if scale is None else scale), float(shift)

    def _call(self, name, *args):
        check(self.lib, getattr(self.lib, name)(*args))

    # ------------------------------------------------------------------------------------------------ assignment
    @staticmethod
    def _win(name, order, kb_kcut):
        """Entry point and order arguments for the window family: kb_kcut > 0 selects the Kaiser-Bessel variants."""
        return (name + "_kb", (order, float(kb_kcut))) if kb_kcut else (name, (order,))

    def paint(self, pos, shape, weights=None, wscalar=1.0, order=2, scale=None, shift=0.0, out=None, accumulate=False,
              kb_kcut=0.0):
        A = self.A
        pos = A.prepare(pos)
        weights = None if weights is None else A.prepare(weights)
        mesh = A.empty(shape) if out is None else out
        sc, sh = self._xf(scale, shift)
        fn, oa = self._win("mcpm_paint", order, kb_kcut)
        self._call(fn, A.stream(), A.ptr(pos), A.ptr(weights), wscalar, A.shape(pos)[0], *shape, *oa, sc,
                   sh, A.ptr(mesh), int(accumulate))
        return mesh

    def read(self, pos, mesh, order=2, scale=None, shift=0.0, kb_kcut=0.0):
        """mesh [nx,ny,nz] -> [np];  mesh [m,nx,ny,nz] -> [np,m]."""
        A = self.A
        pos, mesh = A.prepare(pos), A.prepare(mesh)
        ms = A.shape(mesh)
        nm = 1 if len(ms) == 3 else ms[0]
        n = A.shape(pos)[0]
        out = A.empty((n,) if len(ms) == 3 else (n, nm))
        sc, sh = self._xf(scale, shift)
        fn, oa = self._win("mcpm_read", order, kb_kcut)
        self._call(fn, A.stream(), A.ptr(pos), A.ptr(mesh), nm, n, *ms[-3:], *oa, sc, sh, A.ptr(out))
        return out

    def read_grad(self, pos, mesh, cot=None, order=2, scale=None, shift=0.0, kb_kcut=0.0):
        A = self.A
        pos, mesh = A.prepare(pos), A.prepare(mesh)
        cot = None if cot is None else A.prepare(cot)
        ms = A.shape(mesh)
        nm = 1 if len(ms) == 3 else ms[0]
        n = A.shape(pos)[0]
        out = A.empty((n, 3))
        sc, sh = self._xf(scale, shift)
        fn, oa = self._win("mcpm_read_grad", order, kb_kcut)
        self._call(fn, A.stream(), A.ptr(pos), A.ptr(mesh), nm, A.ptr(cot), n, *ms[-3:], *oa, sc, sh, A.ptr(out), 0)
        return out

    def paint_vjp(self, pos, mesh_bar, weights=None, wscalar=1.0, order=2, scale=None, shift=0.0, want_pos=True,
                  want_weights=True, kb_kcut=0.0):
        A = self.A
        pos, mesh_bar = A.prepare(pos), A.prepare(mesh_bar)
        weights = None if weights is None else A.prepare(weights)
        n = A.shape(pos)[0]
        pb = A.empty((n, 3)) if want_pos else None
        wb = A.empty((n,)) if want_weights else None
        sc, sh = self._xf(scale, shift)
        fn, oa = self._win("mcpm_paint_vjp", order, kb_kcut)
        self._call(fn, A.stream(), A.ptr(pos), A.ptr(weights), wscalar, A.ptr(mesh_bar), n, *A.shape(mesh_bar), *oa, sc,
                   sh, A.ptr(pb), A.ptr(wb), 0)
        return pb, wb

    def paint3(self, pos, vals3, shape, vscale=1.0, order=2):
        A = self.A
        pos, vals3 = A.prepare(pos), A.prepare(vals3)
        out = A.empty((3, *shape))
        self._call("mcpm_paint3", A.stream(), A.ptr(pos), A.ptr(vals3), vscale, A.shape(pos)[0], *shape, order,
                   A.ptr(out), 0)
        return out

    # ------------------------------------------------------------------------------------------------ FFT, Fourier passes
    def rfftn(self, mesh, out=None):
        """`out`: a contiguous complex block of the right size to write into (e.g. a slice of a larger batch)."""
        A = self.A
        mesh = A.prepare(mesh)
        ms = A.shape(mesh)
        batch = 1 if len(ms) == 3 else ms[0]
        if out is None:
            out = A.empty((*ms[:-3], *r2chshape(ms[-3:])), "c64")
        self._call("mcpm_rfftn", self.engine(ms[-3:]).handle, A.stream(), A.ptr(mesh), A.ptr(out), batch)
        return out

    def irfftn(self, meshk, overwrite=False, out=None):
        A = self.A
        meshk = A.prepare(meshk, "c64")
        if not overwrite:
            meshk = meshk.clone() if hasattr(meshk, "clone") else meshk.copy()
        ms = A.shape(meshk)
        batch = 1 if len(ms) == 3 else ms[0]
        rs = ch2rshape(ms[-3:])
        if out is None:
            out = A.empty((*ms[:-3], *rs))
        self._call("mcpm_irfftn", self.engine(rs).handle, A.stream(), A.ptr(meshk), A.ptr(out), batch)
        return out

    def hermitian_project(self, meshk):
        """In place: the Hermitian projection jnp.fft.irfftn applies implicitly (mcpm_hermitian_project)."""
        A = self.A
        ms = A.shape(meshk)
        batch = 1 if len(ms) == 3 else ms[0]
        self._call("mcpm_hermitian_project", A.stream(), A.ptr(meshk), *ch2rshape(ms[-3:]), batch)
        return meshk

    def force_spectra(self, dk, lap_fd=INF, grad_fd=INF, kcut=INF, deconv_order=0):
        A = self.A
        dk = A.prepare(dk, "c64")
        rs = ch2rshape(A.shape(dk))
        out = A.empty((3, *A.shape(dk)), "c64")
        self._call("mcpm_force_spectra", A.stream(), A.ptr(dk), A.ptr(out), *rs, fd_code(lap_fd), fd_code(grad_fd),
                   0.0 if kcut == INF else kcut, deconv_order)
        return out

    def force_spectra_T(self, in3, lap_fd=INF, grad_fd=INF, kcut=INF, deconv_order=0, half_weights=False):
        A = self.A
        in3 = A.prepare(in3, "c64")
        cs = A.shape(in3)[1:]
        out = A.empty(cs, "c64")
        self._call("mcpm_force_spectra_T", A.stream(), A.ptr(in3), A.ptr(out), *ch2rshape(cs), fd_code(lap_fd),
                   fd_code(grad_fd), 0.0 if kcut == INF else kcut, deconv_order, int(half_weights), 0)
        return out

    def hessian_spectra(self, dk, lap_fd=INF, grad_fd=INF):
        A = self.A
        dk = A.prepare(dk, "c64")
        out = A.empty((6, *A.shape(dk)), "c64")
        self._call("mcpm_hessian_spectra", A.stream(), A.ptr(dk), A.ptr(out), *ch2rshape(A.shape(dk)),
                   fd_code(lap_fd), fd_code(grad_fd))
        return out

    def hessian_spectra_T(self, in6, lap_fd=INF, grad_fd=INF, half_weights=False):
        A = self.A
        in6 = A.prepare(in6, "c64")
        cs = A.shape(in6)[1:]
        out = A.empty(cs, "c64")
        self._call("mcpm_hessian_spectra_T", A.stream(), A.ptr(in6), A.ptr(out), *ch2rshape(cs), fd_code(lap_fd),
                   fd_code(grad_fd), int(half_weights), 0)
        return out

    def lpt2_source(self, h6):
        A = self.A
        h6 = A.prepare(h6)
        out = A.empty(A.shape(h6)[1:])
        self._call("mcpm_lpt2_source", A.stream(), A.ptr(h6), A.ptr(out), math.prod(A.shape(h6)[1:]))
        return out

    def lpt2_source_vjp(self, h6, d2bar):
        A = self.A
        h6, d2bar = A.prepare(h6), A.prepare(d2bar)
        out = A.empty(A.shape(h6))
        self._call("mcpm_lpt2_source_vjp", A.stream(), A.ptr(h6), A.ptr(d2bar), A.ptr(out),
                   math.prod(A.shape(h6)[1:]))
        return out

    def deconv(self, meshk, order=2, kb_kcut=0.0):
        A = self.A
        meshk = A.prepare(meshk, "c64")
        out = A.empty(A.shape(meshk), "c64")
        fn, oa = self._win("mcpm_deconv", order, kb_kcut)
        self._call(fn, A.stream(), A.ptr(meshk), A.ptr(out), *ch2rshape(A.shape(meshk)), *oa)
        return out

    def interlace_combine(self, in_m, scale=1.0, deconv_order=0):
        A = self.A
        in_m = A.prepare(in_m, "c64")
        m, cs = A.shape(in_m)[0], A.shape(in_m)[1:]
        out = A.empty(cs, "c64")
        self._call("mcpm_interlace_combine", A.stream(), A.ptr(in_m), A.ptr(out), m, *ch2rshape(cs), scale,
                   deconv_order)
        return out

    def interlace_combine_T(self, inp, m, scale=1.0, deconv_order=0):
        A = self.A
        inp = A.prepare(inp, "c64")
        cs = A.shape(inp)
        out = A.empty((m, *cs), "c64")
        self._call("mcpm_interlace_combine_T", A.stream(), A.ptr(inp), A.ptr(out), m, *ch2rshape(cs), scale,
                   deconv_order)
        return out

    def chreshape(self, meshk, out_cshape):
        A = self.A
        meshk = A.prepare(meshk, "c64")
        out = A.empty(out_cshape, "c64")
        self._call("mcpm_chreshape", A.stream(), A.ptr(meshk), *ch2rshape(A.shape(meshk)), A.ptr(out),
                   *ch2rshape(out_cshape))
        return out

    def scale_spectrum(self, meshk, transfer):
        A = self.A
        meshk, transfer = A.prepare(meshk, "c64"), A.prepare(transfer)
        out = A.empty(A.shape(meshk), "c64")
        self._call("mcpm_scale_spectrum", A.stream(), A.ptr(meshk), A.ptr(transfer), A.ptr(out),
                   math.prod(A.shape(meshk)))
        return out

    # ------------------------------------------------------------------------------------------------ composites
    def pm_forces(self, pos, shape, order=2, paint_deconv=False, lap_fd=INF, grad_fd=INF, kcut=INF, want_meshes=False,
                  lattice=None):
        A = self.A
        pos = A.prepare(pos)
        n = A.shape(pos)[0]
        forces = A.empty((n, 3))
        fm = A.empty((3, *shape)) if want_meshes else None
        self._call("mcpm_pm_forces", self._frame(shape, lattice).handle, A.stream(), A.ptr(pos), n, order, int(paint_deconv),
                   fd_code(lap_fd), fd_code(grad_fd), 0.0 if kcut == INF else kcut, A.ptr(fm), A.ptr(forces))
        return (forces, fm) if want_meshes else forces

    def pm_forces_vjp(self, pos, fbar, fmesh3, order=2, paint_deconv=False, lap_fd=INF, grad_fd=INF, kcut=INF,
                      lattice=None):
        A = self.A
        pos, fbar, fmesh3 = A.prepare(pos), A.prepare(fbar), A.prepare(fmesh3)
        n = A.shape(pos)[0]
        shape = A.shape(fmesh3)[1:]
        out = A.empty((n, 3))
        self._call("mcpm_pm_forces_vjp", self._frame(shape, lattice).handle, A.stream(), A.ptr(pos), A.ptr(fbar),
                   A.ptr(fmesh3), n, order, int(paint_deconv), fd_code(lap_fd), fd_code(grad_fd),
                   0.0 if kcut == INF else kcut, A.ptr(out), 0)
        return out

    def pm_forces_mesh(self, pos, dk, order=2, lap_fd=INF, grad_fd=INF, kcut=INF):
        A = self.A
        pos, dk = A.prepare(pos), A.prepare(dk, "c64")
        n = A.shape(pos)[0]
        out = A.empty((n, 3))
        self._call("mcpm_pm_forces_mesh", self._frame(ch2rshape(A.shape(dk)), None).handle, A.stream(), A.ptr(pos),
                   A.ptr(dk), n, order, fd_code(lap_fd), fd_code(grad_fd), 0.0 if kcut == INF else kcut, A.ptr(out))
        return out

    def pm_forces2(self, pos, dk, order=2, lap_fd=INF, grad_fd=INF):
        A = self.A
        pos, dk = A.prepare(pos), A.prepare(dk, "c64")
        n = A.shape(pos)[0]
        out = A.empty((n, 3))
        self._call("mcpm_pm_forces2", self._frame(ch2rshape(A.shape(dk)), None).handle, A.stream(), A.ptr(pos), A.ptr(dk),
                   n, order, fd_code(lap_fd), fd_code(grad_fd), A.ptr(out), 0)
        return out

    def lpt(self, dk, pos, d1, d2, dv2, lpt_order=2, read_order=2, lap_fd=INF, grad_fd=INF, tape=False):
        """pos None: the particles sit on the cells of the mesh (regular_pos(mesh_shape))."""
        A = self.A
        dk = A.prepare(dk, "c64")
        rs = ch2rshape(A.shape(dk))
        pos = None if pos is None else A.prepare(pos)
        n = math.prod(rs) if pos is None else A.shape(pos)[0]
        dpos, vel = A.empty((n, 3)), A.empty((n, 3))
        f1 = A.empty((n, 3)) if tape else None
        f2 = A.empty((n, 3)) if tape and lpt_order == 2 else None
        h6 = A.empty((6, *rs)) if tape and lpt_order == 2 else None
        self._call("mcpm_lpt", self._frame(rs, None).handle, A.stream(), A.ptr(dk), A.ptr(pos), n, lpt_order, read_order,
                   fd_code(lap_fd), fd_code(grad_fd), d1, d2, dv2, A.ptr(dpos), A.ptr(vel), A.ptr(f1), A.ptr(f2),
                   A.ptr(h6))
        return (dpos, vel, (f1, f2, h6)) if tape else (dpos, vel)

    def lpt_vjp(self, pos, cshape, d1, d2, dv2, dposbar, velbar, tape, lpt_order=2, read_order=2, lap_fd=INF,
                grad_fd=INF, want_coef=False):
        A = self.A
        dposbar, velbar = A.prepare(dposbar), A.prepare(velbar)
        pos = None if pos is None else A.prepare(pos)
        f1, f2, h6 = tape
        n = A.shape(dposbar)[0]
        rs = ch2rshape(cshape)
        dkbar = A.empty(cshape, "c64")
        coef = A.zeros((3,), "f64") if want_coef else None
        self._call("mcpm_lpt_vjp", self._frame(rs, None).handle, A.stream(), A.ptr(pos), n, lpt_order, read_order,
                   fd_code(lap_fd), fd_code(grad_fd), d1, d2, dv2, A.ptr(dposbar), A.ptr(velbar), A.ptr(f1),
                   A.ptr(f2), A.ptr(h6), A.ptr(dkbar), A.ptr(coef), 0)
        return (dkbar, coef) if want_coef else dkbar

    def nbody_steps(self, pos, vel, shape, alpha, beta, drift_pre, drift_post, order=2, paint_deconv=False,
                    lap_fd=INF, grad_fd=INF, tape=False, tape_vel=False, lattice=None, tape_forces=True):
        """In place on (pos, vel) -- pass fresh copies.  Returns the tape (xk, vk, fm) when asked.  `lattice`: pos (and
        the xk tape) hold displacements from the sites of that regular lattice (mcpm_engine_set_relative)."""
        A = self.A
        n = A.shape(pos)[0]
        ns = len(alpha)
        xk = A.empty((ns, n, 3)) if tape else None
        vk = A.empty((ns, n, 3)) if tape and tape_vel else None
        # tape_forces=False: 16 bytes per cell and step less tape; the reverse sweep recomputes each step's force meshes
        fm = A.empty((ns, 4, *shape)) if tape and tape_forces else None
        self._call("mcpm_nbody_steps", self._frame(shape, lattice).handle, A.stream(), A.ptr(pos), A.ptr(vel), n, ns,
                   host_floats(alpha), host_floats(beta), host_floats(drift_pre), host_floats(drift_post), order,
                   int(paint_deconv), fd_code(lap_fd), fd_code(grad_fd), A.ptr(xk), A.ptr(vk), A.ptr(fm))
        return (xk, vk, fm)

    def nbody_steps_vjp(self, posbar, velbar, shape, alpha, beta, drift_pre, drift_post, tape, order=2,
                        paint_deconv=False, lap_fd=INF, grad_fd=INF, v0=None, want_coef=False, lattice=None):
        """In place on (posbar, velbar)."""
        A = self.A
        xk, vk, fm = tape
        n = A.shape(posbar)[0]
        ns = len(alpha)
        coef = A.zeros((ns, 4), "f64") if want_coef else None
        self._call("mcpm_nbody_steps_vjp", self._frame(shape, lattice).handle, A.stream(), A.ptr(posbar), A.ptr(velbar), n, ns,
                   host_floats(alpha), host_floats(beta), host_floats(drift_pre), host_floats(drift_post), order,
                   int(paint_deconv), fd_code(lap_fd), fd_code(grad_fd), A.ptr(xk), A.ptr(vk), A.ptr(fm), A.ptr(v0),
                   A.ptr(coef))
        return coef

    def nufft_paint(self, pos, paint_shape, weights=None, wscalar=1.0, scale=None, paint_order=2, interlace_order=2,
                    paint_deconv=True, kb_kcut=0.0, lattice=None, rsd=None):
        """rsd = (vel, los, coef): deposit at pos + (vel . los) * coef * los, the shift applied inside the paint kernels
        (mcpm_nufft_rsd; rectangular windows)."""
        A = self.A
        pos = A.prepare(pos)
        weights = None if weights is None else A.prepare(weights)
        out = A.empty(r2chshape(paint_shape), "c64")
        sc, _ = self._xf(scale, 0.0)
        if rsd is not None:
            if kb_kcut:
                raise ValueError("the fused redshift-space shift is built for the rectangular windows")
            vel, los, coef = rsd
            vel = A.prepare(vel)
            self._call("mcpm_nufft_rsd", self._frame(paint_shape, lattice).handle, A.stream(), A.ptr(pos), A.ptr(vel),
                       host_floats(los), float(coef), A.ptr(weights), wscalar, A.shape(pos)[0], sc, paint_order,
                       interlace_order, int(paint_deconv), A.ptr(out))
            return out
        fn, oa = self._win("mcpm_nufft", paint_order, kb_kcut)
        self._call(fn, self._frame(paint_shape, lattice).handle, A.stream(), A.ptr(pos), A.ptr(weights), wscalar,
                   A.shape(pos)[0], sc, *oa, interlace_order, int(paint_deconv), A.ptr(out))
        return out

    def nufft_paint_vjp(self, pos, outbar, paint_shape, weights=None, wscalar=1.0, scale=None, paint_order=2,
                        interlace_order=2, paint_deconv=True, want_pos=True, want_weights=True, kb_kcut=0.0,
                        lattice=None, rsd=None):
        """-> (posbar, weightsbar), or with rsd = (vel, los, coef): (posbar, weightsbar, velbar)."""
        A = self.A
        pos, outbar = A.prepare(pos), A.prepare(outbar, "c64")
        weights = None if weights is None else A.prepare(weights)
        n = A.shape(pos)[0]
        pb = A.empty((n, 3)) if want_pos else None
        wb = A.empty((n,)) if want_weights else None
        sc, _ = self._xf(scale, 0.0)
        if rsd is not None:
            vel, los, coef = rsd
            vel = A.prepare(vel)
            vb = A.empty((n, 3))
            self._call("mcpm_nufft_rsd_vjp", self._frame(paint_shape, lattice).handle, A.stream(), A.ptr(pos), A.ptr(vel),
                       host_floats(los), float(coef), A.ptr(weights), wscalar, n, sc, paint_order, interlace_order,
                       int(paint_deconv), A.ptr(outbar), A.ptr(pb), A.ptr(vb), A.ptr(wb))
            return pb, wb, vb
        fn, oa = self._win("mcpm_nufft_vjp", paint_order, kb_kcut)
        self._call(fn, self._frame(paint_shape, lattice).handle, A.stream(), A.ptr(pos), A.ptr(weights), wscalar,
                   n, sc, *oa, interlace_order, int(paint_deconv), A.ptr(outbar), A.ptr(pb), A.ptr(wb))
        return pb, wb

    def _obs_struct(self, obs):
        """mcpm_obs from a dict: curved, lightcone, ap, rsd, cell[3], origin[3], los[3], gf, a_par, a_perp, r0, dr,
        tab_gf / tab_ap (float32 arrays of equal length, or None), dvel ([np,3] or None), rot (3 x 3 or None).  Returns
        the struct and the prepared arrays (to keep alive across the call)."""
        from ._capi import McpmObs
        A = self.A
        o = McpmObs()
        o.curved, o.lightcone = int(bool(obs.get("curved", False))), int(bool(obs.get("lightcone", False)))
        o.ap, o.rsd = int(obs.get("ap", 0)), int(bool(obs.get("rsd", True)))
        for name, default in (("cell", (1.0,) * 3), ("origin", (0.0,) * 3), ("los", (0.0,) * 3)):
            getattr(o, name)[:] = [float(v) for v in obs.get(name, default)]
        o.gf, o.a_par, o.a_perp = (float(obs.get(k, d)) for k, d in (("gf", 0.0), ("a_par", 1.0), ("a_perp", 1.0)))
        o.r0, o.dr = float(obs.get("r0", 0.0)), float(obs.get("dr", 0.0))
        keep = [None if obs.get(k) is None else A.prepare(obs[k]) for k in ("tab_gf", "tab_ap", "dvel")]
        nts = {A.shape(t)[0] for t in keep[:2] if t is not None}
        if len(nts) > 1:
            raise ValueError("tab_gf and tab_ap must have the same number of nodes")
        o.nt = nts.pop() if nts else 0
        o.tab_gf, o.tab_ap, o.dvel = (A.ptr(t) for t in keep)
        rot = obs.get("rot")
        o.rot[:] = [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0] if rot is None else [float(v) for r in rot for v in r]
        return o, keep

    def nufft_observed(self, pos, vel, obs, paint_shape, weights=None, wscalar=1.0, scale=None, paint_order=2,
                       interlace_order=2, paint_deconv=True, kb_kcut=0.0, lattice=None):
        """nufft of the OBSERVED positions, the general observation chain of model.py:780-799 applied inside the paint
        kernels (mcpm_nufft_obs; `obs` as _obs_struct takes it)."""
        A = self.A
        pos = A.prepare(pos)
        vel = None if vel is None else A.prepare(vel)
        weights = None if weights is None else A.prepare(weights)
        out = A.empty(r2chshape(paint_shape), "c64")
        sc, _ = self._xf(scale, 0.0)
        o, keep = self._obs_struct(obs)
        self._call("mcpm_nufft_obs", self._frame(paint_shape, lattice).handle, A.stream(), A.ptr(pos), A.ptr(vel),
                   C.addressof(o), A.ptr(weights), wscalar, A.shape(pos)[0], sc, paint_order, float(kb_kcut),
                   interlace_order, int(paint_deconv), A.ptr(out))
        del keep
        return out

    def nufft_observed_vjp(self, pos, vel, obs, outbar, paint_shape, weights=None, wscalar=1.0, scale=None,
                           paint_order=2, interlace_order=2, paint_deconv=True, kb_kcut=0.0, lattice=None,
                           want_par=True):
        """-> posbar, velbar (None without rsd), dvelbar (None without dvel), weightsbar, parbar: float64
        [3 + 2 nt] cotangents of (gf, a_par, a_perp, tab_gf nodes, tab_ap nodes), or None."""
        from ._capi import OBS_SLOTS
        A = self.A
        pos, outbar = A.prepare(pos), A.prepare(outbar, "c64")
        vel = None if vel is None else A.prepare(vel)
        weights = None if weights is None else A.prepare(weights)
        n = A.shape(pos)[0]
        sc, _ = self._xf(scale, 0.0)
        o, keep = self._obs_struct(obs)
        pb, wb = A.empty((n, 3)), (A.empty((n,)) if weights is not None else None)
        vb = A.empty((n, 3)) if o.rsd else None
        db = A.empty((n, 3)) if (o.rsd and keep[2] is not None) else None
        par = A.empty((OBS_SLOTS, 3 + 2 * o.nt), "f64") if want_par else None
        self._call("mcpm_nufft_obs_vjp", self._frame(paint_shape, lattice).handle, A.stream(), A.ptr(pos), A.ptr(vel),
                   C.addressof(o), A.ptr(weights), wscalar, n, sc, paint_order, float(kb_kcut), interlace_order,
                   int(paint_deconv), A.ptr(outbar), A.ptr(pb), A.ptr(vb), A.ptr(db), A.ptr(wb), A.ptr(par))
        del keep
        return pb, vb, db, wb, (None if par is None else par[0])

    def radial_tables(self, pos, geom, tabs):
        """out[p, k] = tab_k(r_p): functions of the comoving distance at the particles (mcpm_radial_tables; `geom` as
        _obs_struct takes it: curved, cell, origin, los, r0, dr; tabs [ntab, nt])."""
        A = self.A
        pos, tabs = A.prepare(pos), A.prepare(tabs)
        ntab, nt = A.shape(tabs)
        o, keep = self._obs_struct(dict(geom, tab_gf=None, tab_ap=None, dvel=None))
        o.nt = nt
        n = A.shape(pos)[0]
        out = A.empty((n, ntab))
        self._call("mcpm_radial_tables", A.stream(), A.ptr(pos), n, C.addressof(o), ntab, A.ptr(tabs), A.ptr(out))
        return out

    def radial_tables_vjp(self, pos, geom, tabs, outbar, want_pos=False):
        """-> (posbar [np,3] or None, tabbar float64 [ntab, nt])."""
        from ._capi import OBS_SLOTS
        A = self.A
        pos, tabs, outbar = A.prepare(pos), A.prepare(tabs), A.prepare(outbar)
        ntab, nt = A.shape(tabs)
        o, keep = self._obs_struct(dict(geom, tab_gf=None, tab_ap=None, dvel=None))
        o.nt = nt
        n = A.shape(pos)[0]
        pb = A.empty((n, 3)) if want_pos else None
        tb = A.empty((OBS_SLOTS, ntab, nt), "f64")
        self._call("mcpm_radial_tables_vjp", A.stream(), A.ptr(pos), n, C.addressof(o), ntab, A.ptr(tabs), A.ptr(outbar),
                   A.ptr(pb), A.ptr(tb))
        return pb, tb[0]

    # ------------------------------------------------------------------------------------------------ glue
    def chreshape_vjp(self, outbar, in_cshape):
        A = self.A
        outbar = A.prepare(outbar, "c64")
        inbar = A.empty(in_cshape, "c64")
        self._call("mcpm_chreshape_vjp", A.stream(), A.ptr(outbar), *ch2rshape(A.shape(outbar)), A.ptr(inbar),
                   *ch2rshape(in_cshape))
        return inbar

    def spectrum_bins(self, m0, m1, box_size, kedges, deconv=(0, 0), ell=0, los=(0.0, 0.0, 0.0)):
        """[4, len(kedges)+1] float64: count, sum |k|, sum Re, sum Im of m0 conj(m1) per np.digitize bin; the power sums
        of multipole `ell` along the unit vector `los` carry (2 ell + 1) L_ell(mu)."""
        A = self.A
        m0 = A.prepare(m0, "c64")
        m1 = None if m1 is None else A.prepare(m1, "c64")
        rs = ch2rshape(A.shape(m0))
        ke = A.prepare(kedges, "f64")
        out = A.zeros((4, len(kedges) + 1), "f64")
        box = (float(box_size[0]), float(box_size[1]), float(box_size[2]))
        if ell == 0:
            self._call("mcpm_spectrum_bins", A.stream(), A.ptr(m0), A.ptr(m1), *rs, *box, A.ptr(ke), len(kedges),
                       int(deconv[0]), int(deconv[1]), A.ptr(out))
        else:
            self._call("mcpm_spectrum_bins_ell", A.stream(), A.ptr(m0), A.ptr(m1), *rs, *box, A.ptr(ke), len(kedges),
                       int(deconv[0]), int(deconv[1]), int(ell), (C.c_double * 3)(*(float(x) for x in los)), A.ptr(out))
        return out

    def rg2cgh(self, mesh, scale, transfer=None):
        A = self.A
        mesh = A.prepare(mesh)
        rs = A.shape(mesh)
        out = A.empty(r2chshape(rs), "c64")
        transfer = None if transfer is None else A.prepare(transfer)
        self._call("mcpm_rg2cgh", A.stream(), A.ptr(mesh), A.ptr(out), *rs, float(scale), A.ptr(transfer))
        return out

    def rg2cgh_vjp(self, outbar, scale, transfer=None):
        A = self.A
        outbar = A.prepare(outbar, "c64")
        rs = ch2rshape(A.shape(outbar))
        out = A.empty(rs)
        transfer = None if transfer is None else A.prepare(transfer)
        self._call("mcpm_rg2cgh_vjp", A.stream(), A.ptr(outbar), A.ptr(out), *rs, float(scale), A.ptr(transfer))
        return out

    def cgh2rg(self, meshk, inv_scale):
        A = self.A
        meshk = A.prepare(meshk, "c64")
        rs = ch2rshape(A.shape(meshk))
        out = A.empty(rs)
        self._call("mcpm_cgh2rg", A.stream(), A.ptr(meshk), A.ptr(out), *rs, float(inv_scale))
        return out

    def hermitian_weights(self, meshk, mode, inplace=False):
        """3-D, or a batch [M, nx, ny, nzc] (one launch per member); inplace: overwrite meshk."""
        A = self.A
        meshk = A.prepare(meshk, "c64")
        out = meshk if inplace else A.empty(A.shape(meshk), "c64")
        ms = A.shape(meshk)
        rs = ch2rshape(ms[-3:])
        if len(ms) == 3:
            self._call("mcpm_hermitian_weights", A.stream(), A.ptr(meshk), A.ptr(out), *rs, mode)
        else:
            for m in range(ms[0]):
                self._call("mcpm_hermitian_weights", A.stream(), A.ptr(meshk[m]), A.ptr(out[m]), *rs, mode)
        return out

    def axpby(self, x, a=1.0, y=None, b=0.0, c=0.0):
        A = self.A
        x = A.prepare(x)
        y = None if y is None else A.prepare(y)
        out = A.empty(A.shape(x))
        self._call("mcpm_axpby", A.stream(), A.ptr(x), a, A.ptr(y), b, c, math.prod(A.shape(x)), A.ptr(out))
        return out

    def dot(self, a, b):
        """float64 device scalar sum(a*b)."""
        A = self.A
        a, b = A.prepare(a), A.prepare(b)
        out = A.zeros((1,), "f64")
        self._call("mcpm_dot", A.stream(), A.ptr(a), A.ptr(b), math.prod(A.shape(a)), A.ptr(out))
        return out

    # ---- Lagrangian bias expansion (bias.cu; bricks.py:327-452)
    def bias_spectra(self, delta_k, cells_per_len, inv_transfer=None):
        """[M, nx, ny, nzc] spectra (M = 10, or 12 with an inverse transfer) from one pass over delta_k."""
        A = self.A
        delta_k = A.prepare(delta_k, "c64")
        nx, ny, nzc = A.shape(delta_k)
        nz = 2 * (nzc - 1)
        it = None if inv_transfer is None else A.prepare(inv_transfer)
        out = A.empty((12 if it is not None else 10, nx, ny, nzc), "c64")
        self._call("mcpm_bias_spectra", A.stream(), A.ptr(delta_k), nx, ny, nz, host_floats(cells_per_len), A.ptr(it),
                   A.ptr(out))
        return out

    def bias_spectra_vjp(self, outbar, cells_per_len, inv_transfer=None):
        A = self.A
        outbar = A.prepare(outbar, "c64")
        _, nx, ny, nzc = A.shape(outbar)
        it = None if inv_transfer is None else A.prepare(inv_transfer)
        out = A.empty((nx, ny, nzc), "c64")
        self._call("mcpm_bias_spectra_vjp", A.stream(), A.ptr(outbar), nx, ny, 2 * (nzc - 1), host_floats(cells_per_len),
                   A.ptr(it), A.ptr(out), 0)
        return out

    def shear_invariants(self, s5, out=None):
        A = self.A
        s5 = A.prepare(s5)
        shape = A.shape(s5)[1:]
        out = A.empty((2, *shape)) if out is None else out
        self._call("mcpm_shear_invariants", A.stream(), A.ptr(s5), math.prod(shape), A.ptr(out))
        return out

    def shear_invariants_vjp(self, s5, out2bar, out=None):
        A = self.A
        s5, out2bar = A.prepare(s5), A.prepare(out2bar)
        out = A.empty(A.shape(s5)) if out is None else out
        self._call("mcpm_shear_invariants_vjp", A.stream(), A.ptr(s5), A.ptr(out2bar), math.prod(A.shape(s5)[1:]), A.ptr(out))
        return out

    def bias_weights(self, vals, growth, coef, want_dvel=True):
        """weights [np], dvel [np, 3], moments (float64 [2], kept for the reverse pass); growth: float or [np] array."""
        A = self.A
        vals = A.prepare(vals)
        n, K = A.shape(vals)
        garr = None if isinstance(growth, float) else A.prepare(growth)
        gs = growth if garr is None else 0.0
        mom = A.zeros((2,), "f64")
        self._call("mcpm_bias_moments", A.stream(), A.ptr(vals), K, gs, A.ptr(garr), n, A.ptr(mom))
        w, dv = A.empty((n,)), (A.empty((n, 3)) if want_dvel else None)
        self._call("mcpm_bias_weights", A.stream(), A.ptr(vals), K, gs, A.ptr(garr), host_floats(coef), A.ptr(mom), n,
                   A.ptr(w), A.ptr(dv))
        return w, dv, mom

    def bias_weights_vjp(self, vals, growth, coef, mom, wbar, dvelbar=None):
        """-> valsbar [np, K], coefbar float64 [14] (13 coefficients + scalar growth), gbar [np] or None."""
        A = self.A
        vals, wbar = A.prepare(vals), A.prepare(wbar)
        n, K = A.shape(vals)
        garr = None if isinstance(growth, float) else A.prepare(growth)
        gs = growth if garr is None else 0.0
        dvb = None if dvelbar is None else A.prepare(dvelbar)
        msum, cb = A.zeros((2,), "f64"), A.zeros((14,), "f64")
        vb = A.empty((n, K))
        gb = A.empty((n,)) if garr is not None else None
        self._call("mcpm_bias_weights_vjp", A.stream(), A.ptr(vals), K, gs, A.ptr(garr), host_floats(coef), A.ptr(mom),
                   A.ptr(wbar), A.ptr(dvb), n, A.ptr(msum), A.ptr(vb), A.ptr(cb), A.ptr(gb))
        return vb, cb, gb

    def rsd_shift(self, pos, vel, los, coef):
        A = self.A
        pos, vel = A.prepare(pos), A.prepare(vel)
        out = A.empty(A.shape(pos))
        self._call("mcpm_rsd_shift", A.stream(), A.ptr(pos), A.ptr(vel), host_floats(los), coef, A.shape(pos)[0],
                   A.ptr(out))
        return out

    def rsd_shift_vjp(self, posbar, los, coef):
        A = self.A
        posbar = A.prepare(posbar)
        out = A.empty(A.shape(posbar))
        self._call("mcpm_rsd_shift_vjp", A.stream(), A.ptr(posbar), host_floats(los), coef, A.shape(posbar)[0],
                   A.ptr(out), 0)
        return out
