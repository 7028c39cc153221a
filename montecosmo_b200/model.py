"""
Field-level log-density on the engine: the prior -> evolve -> likelihood chain of `FieldLevelModel`
(montecosmo/model.py:631-634, 640-679, 683-837, 840-933) restricted to the configuration BASELINE.json benchmarks:

  * white noise sampled in real space, `precond='real'` (bricks.py:303-305): delta_k = rfftn(white) * transfer,
    transfer = sqrt(P_lin(k) * N / V) from a tabulated linear power (white2lin, bricks.py:96-106, 152-157);
  * particles on the regular lattice (bricks.py:593-603); linear Lagrangian bias weights 1 + b1 D delta_L(q)
    (first term of lagrangian_bias, bricks.py:358-362, NGP read);
  * evolution 'lpt' (model.py:763) or 'nbody' = 2LPT + BullFrog steps (model.py:771-773, paint_deconv=False there);
  * flat-sky redshift-space shift along `los` (bricks.py:781-792 in cell units), a_obs scalar;
  * interlaced, deconvolved NUFFT paint with the bias weights (model.py:802-809) -> 1 + delta_obs mesh;
  * Gaussian likelihood of an observed mesh with constant noise, N(0,1) prior on the white field.

`logpdf` / `force` keep the reference's wrapper names (model.py:350-363).  Everything elementwise is an engine kernel
(mcpm_axpby, mcpm_dot, mcpm_rsd_shift, ...); torch supplies memory and the autograd tape only.
"""
from __future__ import annotations

import numpy as np
import torch

from . import cosmo as _cosmo
from . import nbody as nb
from .ops import r2chshape


class _Axpby(torch.autograd.Function):
    """out = a*x + c with host constants a, c."""

    @staticmethod
    def forward(ctx, x, a, c):
        ctx.a = a
        return nb.ops().axpby(x, a, None, 0.0, c)

    @staticmethod
    def backward(ctx, g):
        return nb.ops().axpby(g.contiguous(), ctx.a), None, None


class _RsdShift(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, vel, los, coef):
        ctx.cfg = (los, coef)
        return nb.ops().rsd_shift(pos, vel, los, coef)

    @staticmethod
    def backward(ctx, g):
        los, coef = ctx.cfg
        g = g.contiguous()
        return g, nb.ops().rsd_shift_vjp(g, los, coef), None, None


class _GaussLogp(torch.autograd.Function):
    """-0.5 * sum((x - mu)^2) * inv_var as a float64 device scalar; mu may be None (zero mean)."""

    @staticmethod
    def forward(ctx, x, mu, inv_var):
        o = nb.ops()
        r = x if mu is None else o.axpby(x, 1.0, mu, -1.0)
        ctx.save_for_backward(r)
        ctx.inv_var = inv_var
        return o.dot(r, r).reshape(()) * (-0.5 * inv_var)

    @staticmethod
    def backward(ctx, g):
        (r,) = ctx.saved_tensors
        # g is a float64 device scalar (1.0 in practice): scale on the device -- no D2H read, so logpdf().backward(), what a
        # generic sampler calls, neither synchronises the stream nor breaks a CUDA-graph capture
        return r * (g * (-ctx.inv_var)).to(r.dtype), None, None


def eisenstein_hu_nowiggle(k, cosmo):
    """Zero-baryon-oscillation transfer function of Eisenstein & Hu (1998, ApJ 496, 605, eqs. 26-31).

    Own restatement of the published fit, used only to tabulate a realistic linear power for synthetic benchmarks;
    it is not pinned to jax_cosmo.power (which the reference calls at bricks.py:74).  k in h/Mpc.
    """
    h = float(cosmo.h)
    om, ob = float(cosmo.Omega_m), float(cosmo.Omega_b)
    omh2, obh2, fb = om * h**2, ob * h**2, ob / om
    theta = 2.7255 / 2.7
    s = 44.5 * np.log(9.83 / omh2) / np.sqrt(1 + 10 * obh2**0.75)  # Mpc
    alpha = 1 - 0.328 * np.log(431 * omh2) * fb + 0.38 * np.log(22.3 * omh2) * fb**2
    kmpc = k * h
    gamma = om * h * (alpha + (1 - alpha) / (1 + (0.43 * kmpc * s) ** 4))
    q = k * theta**2 / gamma
    L0 = np.log(2 * np.e + 1.8 * q)
    C0 = 14.2 + 731.0 / (1 + 62.5 * q)
    return L0 / (L0 + C0 * q**2)


def linear_power_table(cosmo, n_interp=256):
    """(k, P_lin(k, a=1)) on logspace(-4, 1, n) h/Mpc (the grid of bricks.py:73), sigma8-normalised."""
    ks = np.logspace(-4, 1, n_interp)
    kk = np.logspace(-5, 2, 4096)
    p = kk ** float(cosmo.n_s) * eisenstein_hu_nowiggle(kk, cosmo) ** 2
    x = kk * 8.0
    wth = 3 * (np.sin(x) - x * np.cos(x)) / x**3
    sig2 = np.trapezoid(kk**3 * p * wth**2 / (2 * np.pi**2), np.log(kk))
    amp = float(cosmo.sigma8) ** 2 / sig2
    return ks, amp * ks ** float(cosmo.n_s) * eisenstein_hu_nowiggle(ks, cosmo) ** 2


class FieldModel:
    def __init__(self, mesh_shape=(64, 64, 64), box_size=(640.0, 640.0, 640.0), evolution="nbody", n_steps=5,
                 a_start=0.0, a_obs=1.0, lpt_order=2, paint_order=2, interlace_order=2, paint_deconv=True,
                 paint_oversamp=1.0, b1=1.0, rsd=True, los=(0.0, 0.0, 1.0), sigma_obs=1.0, cosmology=None, kpow=None,
                 precond="real", out_shape="paint", relative=True, tape_forces=True):
        if precond not in ("real", "fourier"):
            raise ValueError("precond must be 'real' or 'fourier'")
        if out_shape not in ("paint", "mesh"):
            raise ValueError("out_shape must be 'paint' or 'mesh'")
        # 'paint': the mesh is handed on at the paint shape, as evolve does (model.py:806-807); 'mesh': at the evolution
        # shape, where the reference's likelihood brings it back anyway when final_shape == init_shape and there is no
        # selection function (model.py:855) -- what the slab-decomposed model (dist_model.py) evaluates
        self.precond, self.out_shape = precond, out_shape
        # relative: carry particles as float32 displacements from their lattice sites (mcpm_engine_set_relative) instead
        # of float32 absolute positions; False reproduces round 1's arithmetic (kept for the A/B in the parity report)
        self.relative = bool(relative)
        # tape_forces=False: the BullFrog loop tapes positions only and recomputes each step's force meshes in the backward
        # pass (one extra paint + forward Fourier pass per step) instead of keeping 16 bytes per cell and step
        self.tape_forces = bool(tape_forces)
        self.mesh_shape = tuple(int(s) for s in mesh_shape)
        self.box_size = tuple(float(b) for b in box_size)
        self.evolution, self.n_steps, self.a_start, self.a_obs = evolution, int(n_steps), a_start, a_obs
        self.lpt_order, self.paint_order, self.interlace_order = lpt_order, paint_order, interlace_order
        self.paint_deconv, self.b1, self.rsd, self.los = paint_deconv, float(b1), rsd, tuple(float(x) for x in los)
        self.paint_shape = nb.scale_shape(self.mesh_shape, paint_oversamp)
        self.sigma_obs = float(sigma_obs)
        self.cosmology = cosmology if cosmology is not None else _cosmo.Cosmology()
        self.kpow = kpow if kpow is not None else linear_power_table(self.cosmology)
        o = nb.ops()
        self.transfer = o.A.prepare(self.transfer_mesh())
        ax = [np.arange(s, dtype=np.float32) for s in self.mesh_shape]
        self.q = o.A.prepare(np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3))  # regular_pos

    # -- host-side constant (bricks.py:96-106 with a (k, P) table; white_noise scaling bricks.py:150) -----------------
    def transfer_mesh(self):
        ks, pows = self.kpow
        kvec = nb.rfftk(self.mesh_shape, self.box_size)
        kmesh = np.sqrt(sum(k**2 for k in kvec))
        pmesh = np.interp(kmesh.reshape(-1), ks, pows, left=0.0, right=0.0).reshape(kmesh.shape)
        return np.sqrt(pmesh * (np.prod(self.mesh_shape) / np.prod(self.box_size))).astype(np.float32)

    # -- prior -> evolve ----------------------------------------------------------------------------------------------
    def linear_field(self, white):
        """delta_k(a=1) from the real white field: samp2base_mesh (bricks.py:290-320) with precond 'real' (rfftn) or
        'fourier' (rg2cgh, the transfer multiply fused into the permutation pass) + white2lin."""
        if self.precond == "fourier":
            from . import utils as _utils
            return _utils.rg2cgh(white, "backward", self.transfer)
        return nb._ScaleSpectrum.apply(nb.rfftn(white), self.transfer)

    def evolve(self, white):
        c = self.cosmology  # growth tables are cached per parameter values (cosmo.growth_table), cf. model.py:762,769
        dk = self.linear_field(white)
        weights = 1.0
        if self.b1 != 0.0:
            # read(q, delta, order=1) on the lattice that is the mesh: the NGP value at a cell centre is the cell itself
            delta_q = nb.irfftn(dk).reshape(-1)
            weights = _Axpby.apply(delta_q, self.b1 * float(_cosmo.a2g(c, self.a_obs)), 1.0)
        # Particles are carried as float32 displacements from their lattice sites, never as the sum q + displacement
        # (nbody_bf's `relative`, nufft's `lattice`): the reference's float64 sum has no float32 equivalent at 256^3.
        rel = self.relative
        if self.evolution == "lpt":
            pos, vel = nb.lpt(c, dk, None if rel else self.q, self.a_obs, self.lpt_order, 1, _displaced=not rel)
        elif self.evolution == "nbody":
            pos, vel = nb.nbody_bf(c, dk, self.q, self.a_start, self.a_obs, self.n_steps, self.paint_order,
                                   self.lpt_order, paint_deconv=False, ptcl_shape=self.mesh_shape, relative=rel,
                                   tape_forces=self.tape_forces)
            pos, vel = pos[-1], vel[-1]
        else:
            raise ValueError(f"unknown evolution {self.evolution}")
        rsd = None
        if self.rsd:  # flat-sky shift (a shift: the same rule serves displacements), applied inside the paint kernels
            rsd = (vel, self.los, float(_cosmo.a2g(c, self.a_obs) * _cosmo.a2f(c, self.a_obs)))
        gxy = nb.nufft(pos, self.mesh_shape, self.paint_shape, weights, self.paint_order, self.interlace_order,
                       paint_deconv=self.paint_deconv, lattice=self.mesh_shape if rel else None, rsd=rsd)
        if self.paint_shape != self.mesh_shape and self.out_shape == "paint":
            gxy = nb.chreshape(gxy, r2chshape(self.paint_shape))
        return nb.irfftn(gxy)  # 1 + delta_obs at the paint shape (particles == cells: Jacobian 1, model.py:806)

    predict = evolve

    # -- wrappers (model.py:350-363) ----------------------------------------------------------------------------------
    def logpdf(self, white, obs):
        """log prior N(0,1) on the white field + Gaussian log-likelihood of `obs`, up to constants."""
        white = nb._f32(white)
        gxy = self.evolve(white)
        return _GaussLogp.apply(gxy, nb._f32(obs), 1.0 / self.sigma_obs**2) + _GaussLogp.apply(white, None, 1.0)

    def potential(self, white, obs):
        return -self.logpdf(white, obs)

    def value_and_force(self, white, obs):
        """(logpdf, d logpdf / d white) with the likelihood seeded by hand, so that nothing synchronises the stream."""
        o = nb.ops()
        white = nb._f32(white).detach().requires_grad_(True)
        gxy = self.evolve(white)
        inv_var = 1.0 / self.sigma_obs**2
        r = o.axpby(gxy.detach(), 1.0, nb._f32(obs), -1.0)
        (gw,) = torch.autograd.grad(gxy, white, o.axpby(r, -inv_var))
        wd = white.detach()
        lp = o.dot(r, r).reshape(()) * (-0.5 * inv_var) + o.dot(wd, wd).reshape(()) * -0.5
        return lp, o.axpby(gw, 1.0, wd, -1.0)

    def force(self, white, obs):
        return self.value_and_force(white, obs)[1]

    def graphed_value_and_force(self, obs, warmup=2):
        """Capture one evaluation -- every kernel of lpt, the BullFrog loop, the paints, the FFTs and the whole reverse
        sweep -- in a CUDA graph and return `fn(white) -> (logpdf, force)` that replays it (a sampler calls the same
        evaluation thousands of times: one graph launch instead of ~250 kernel launches and the Python around them).
        The returned tensors are the graph's static outputs: consume or copy them before the next call.  Nothing on the
        engine's call path allocates or synchronises, so the capture needs no special casing; cuFFT plans and scratch are
        created during the warm-up evaluations."""
        dev = nb.ops().A.device
        static_white = torch.zeros(self.mesh_shape, device=dev, dtype=torch.float32)
        static_obs = nb._f32(obs).detach().clone()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self.value_and_force(static_white, static_obs)
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.value_and_force(static_white, static_obs)

        def run(white):
            static_white.copy_(nb._f32(white), non_blocking=True)
            graph.replay()
            return out
        # the graph reads these buffers on every replay: they must live as long as the callable does
        run.graph, run.static_white, run.static_obs, run.out = graph, static_white, static_obs, out
        return run


class FieldLevelModel(FieldModel):
    """The general `evolve` of the reference (model.py:683-837: 'kaiser', and 'lpt' / 'nbody' with Lagrangian bias), beyond
    the benchmarked configuration of FieldModel: initial, evolution and paint meshes of different shapes, any particle
    lattice, the full Lagrangian bias expansion with its velocity term and primordial non-Gaussianity, a box placed
    and rotated with respect to the observer, curved or flat sky, light-cone scale factors (a_obs=None: read from the
    comoving distance of every particle; 'lpt' only, as in the reference), redshift-space distortions and
    Alcock-Paczynski rescaling.  Every step is a mirrored callable (nbody.py / bricks.py / utils.py names), so
    torch.autograd differentiates the chain end to end through the engine's adjoints.
    """

    def __init__(self, mesh_shape=(64, 64, 64), box_size=(640.0, 640.0, 640.0), evol_oversamp=1.0, ptcl_oversamp=1.0,
                 box_center=(0.0, 0.0, 0.0), box_rot=None, curved_sky=False, bias=None, png=None, png_type=None,
                 ap_auto=None, cosmo_fid=None, kernel_type="rectangular", fused_observation=True, **kw):
        super().__init__(mesh_shape, box_size, **kw)
        self.fused_observation = bool(fused_observation)
        self.init_shape = self.mesh_shape
        self.evol_shape = nb.scale_shape(self.init_shape, evol_oversamp)
        self.ptcl_shape = nb.scale_shape(self.evol_shape, ptcl_oversamp)
        self.box_center, self.box_rot, self.curved_sky = tuple(float(c) for c in box_center), box_rot, bool(curved_sky)
        self.bias = dict(b1=self.b1) if bias is None else dict(bias)
        self.png, self.png_type = png, png_type
        self.ap_auto, self.cosmo_fid, self.kernel_type = ap_auto, cosmo_fid, kernel_type

    def evolve(self, white, ap=None):
        from . import bricks as B
        c = self.cosmology
        geo = (self.box_center, self.box_rot, self.box_size)
        init_mesh = self.linear_field(white)  # white2lin at the initial shape (bricks.py:152-157)
        if self.evol_shape != self.init_shape:
            init_mesh = nb.chreshape(init_mesh, r2chshape(self.evol_shape))
        if self.evolution == "kaiser":  # model.py:690-699: growth, Eulerian bias, RSD and PNG all linear
            los, a = B.los_scalefactor_mesh(*geo, self.evol_shape, c, self.a_obs, self.curved_sky)
            cell_los = B._rot(los, self.box_rot, inverse=True)
            fnl_bp = 0.0 if self.png is None else self.png.get("fNL_bp", 0.0)
            gxy = B.kaiser_model(c, a, init_mesh, self.box_size, B.b1_L2E(self.bias.get("b1", 0.0)), fnl_bp,
                                 self.png_type, cell_los, self.kpow_sigma8_1() if self.png_type is not None else None)
            if self.evol_shape != self.init_shape:  # back to the final shape (model.py:733-736)
                gxy = nb.irfftn(nb.chreshape(nb.rfftn(gxy), r2chshape(self.init_shape)))
            return gxy
        pos = B.regular_pos(self.evol_shape, self.ptcl_shape)
        growth = None
        if self.a_obs is None and self.fused_observation:
            # the light cone: growth factors at every particle's own scale factor, one engine pass over the lattice
            # (bricks.lightcone_functions) instead of a per-particle scale factor array looked up on the host
            a = None
            growth = B.lightcone_functions(c, pos, *geo, self.evol_shape, self.curved_sky,
                                           (_cosmo.a2g, _cosmo.a2g2, _cosmo.a2dg2dg))
        else:
            _, a = B.los_scalefactor_pos(pos, *geo, self.evol_shape, c, self.a_obs, self.curved_sky)
        weights, dvel, _ = B.lagrangian_bias(c, pos, a, self.box_size, init_mesh, self.bias, self.png, self.png_type,
                                             self.kpow_sigma8_1() if self.png_type is not None else None, read_order=1,
                                             growth=None if growth is None else growth[0])
        if self.png_type is not None:
            init_mesh = B.add_png(c, self.png["fNL"], init_mesh, self.box_size, self.kpow_sigma8_1())
            init_mesh = nb.chreshape(nb.chreshape(init_mesh, r2chshape(self.init_shape)), r2chshape(self.evol_shape))
        # With the observation chain inside the paint the particles can stay float32 DISPLACEMENTS from their lattice
        # sites all the way (nufft_observed's `lattice`): the absolute sum q + dpos costs 1e-5 cell at the faces where
        # the CIC derivative jumps (tools/obs_accuracy.py at 64^3: gradient 1.2e-3 -> 5e-4 against float64)
        lattice = None
        if self.evolution == "lpt":
            dpos, vel = nb.lpt(c, init_mesh, pos, a, self.lpt_order, 1, growth=growth)
            if self.fused_observation:
                pos, lattice = dpos, self.ptcl_shape
            else:
                pos = pos + dpos
        elif self.evolution == "nbody":
            if self.a_obs is None or np.ndim(a) != 0:
                raise AssertionError("N-body light-cone not implemented yet")  # model.py:768
            declared = self.ptcl_shape == self.evol_shape
            rel = True if (self.fused_observation and declared) else None
            pos, vel = nb.nbody_bf(c, init_mesh, pos, self.a_start, a, self.n_steps, self.paint_order, self.lpt_order,
                                   paint_deconv=False, ptcl_shape=self.ptcl_shape if declared else None, relative=rel)
            pos, vel = pos[-1], vel[-1]
            lattice = self.ptcl_shape if rel else None
        else:
            raise ValueError(f"unknown evolution {self.evolution}")
        paint_shape = self.paint_shape if self.paint_shape != self.init_shape else None
        if self.fused_observation:
            # cell -> physical, line of sight and scale factor, redshift-space distortion, Alcock-Paczynski, physical ->
            # cell (model.py:780-799) applied inside the paint kernels: the cosmology enters through the descriptor's
            # scalars and radius tables, no position array is transformed in between
            obs = B.observation(c, *geo, self.evol_shape, self.a_obs, self.curved_sky, self.rsd, self.ap_auto,
                                self.cosmo_fid, ap)
            gxy = nb.nufft_observed(pos, vel if self.rsd else None, self.init_shape, obs, paint_shape, weights,
                                    dvel if (self.rsd and torch.is_tensor(dvel)) else None, self.paint_order,
                                    self.interlace_order, self.kernel_type, self.paint_deconv, lattice=lattice,
                                    pos_shape=self.evol_shape)
        else:  # the same chain as elementwise passes over the particle arrays (cross-check)
            los, a = B.los_scalefactor_pos(pos, *geo, self.evol_shape, c, self.a_obs, self.curved_sky)
            pos = B.cell2phys_pos(pos, *geo, self.evol_shape)
            if self.rsd:
                pos = pos + B.rsd(c, vel, los, a, self.box_rot, self.box_size, self.evol_shape, dvel)
            if self.ap_auto is not None:
                pos = B.ap_auto(pos, los, c, self.cosmo_fid, self.curved_sky) if self.ap_auto else \
                    B.ap_param(pos, los, ap, self.curved_sky)
            pos = B.phys2cell_pos(pos, *geo, self.init_shape)
            gxy = nb.nufft(pos, self.init_shape, paint_shape, weights, self.paint_order, self.interlace_order,
                           self.kernel_type, self.paint_deconv)
        gxy = gxy * float(np.divide(self.init_shape, self.ptcl_shape).prod())  # particle units -> mesh units
        if self.paint_shape != self.init_shape and self.out_shape == "paint":
            gxy = nb.chreshape(gxy, r2chshape(self.paint_shape))
        return nb.irfftn(gxy)

    predict = evolve

    def kpow_sigma8_1(self):
        """The tabulated power normalised to sigma8 = 1, as bricks.lin_power's `kpow` argument wants it."""
        ks, pows = self.kpow
        return ks, pows / float(self.cosmology.sigma8) ** 2
