"""
Generate golden vectors for the particle-mesh hot path by executing the UNMODIFIED reference source
(`/root/reference/montecosmo/{nbody,utils}.py`) in float64 under the NumPy stand-in for JAX in `jaxshim/`.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Writes tests/golden/*.npz (inputs and outputs together, float64 / complex128).  The GPU box never runs this.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MCPM_REFERENCE", "/root/reference")


def load_reference():
    sys.path.insert(0, os.path.join(HERE, "jaxshim"))
    sys.path.insert(0, REF)
    from montecosmo import nbody, utils  # noqa: E402  (the reference itself)
    from jax_cosmo import Cosmology  # noqa: E402  (stand-in container)
    import jax.numpy as jnp  # noqa: E402
    return nbody, utils, Cosmology, jnp


def load_reference_next_rows():
    """montecosmo/{metrics,bricks}.py, the "next" rows of SURVEY 8f (import after load_reference)."""
    from montecosmo import bricks, metrics  # noqa: E402
    return metrics, bricks


ABACUS = dict(Omega_c=0.26447041, Omega_b=0.04930169, h=0.6736, n_s=0.9649, sigma8=0.8076353990239834,
              Omega_k=0.0, w0=-1.0, wa=0.0)  # bricks.py:40-50
OTHER = dict(Omega_c=0.21, Omega_b=0.05, h=0.7, n_s=0.96, sigma8=0.8, Omega_k=0.0, w0=-1.0, wa=0.0)


def lattice(mesh_shape, ptcl_shape=None):
    ptcl_shape = mesh_shape if ptcl_shape is None else ptcl_shape
    ax = [np.linspace(0, m, p, endpoint=False) for m, p in zip(mesh_shape, ptcl_shape)]
    return np.stack(np.meshgrid(*ax, indexing="ij"), axis=-1).reshape(-1, 3)


def gaussian_field_k(rng, shape, amp=0.05, slope=-1.5):
    """A smooth random Hermitian spectrum: rfftn of white noise times a power-law amplitude (test input only)."""
    w = rng.normal(size=shape)
    kx, ky, kz = np.meshgrid(np.fft.fftfreq(shape[0]), np.fft.fftfreq(shape[1]), np.fft.rfftfreq(shape[2]),
                             indexing="ij")
    kk = np.sqrt(kx**2 + ky**2 + kz**2)
    kk[0, 0, 0] = 1.0
    t = amp * kk**slope
    t[0, 0, 0] = 0.0
    return np.fft.rfftn(w) * t * (np.prod(shape) ** -0.5) * 4.0


def main():
    nbody, utils, Cosmology, jnp = load_reference()
    A = lambda x: np.asarray(x)
    out = {}

    # ---- kernels on a non-cubic mesh ---------------------------------------------------------------------------
    shape = (8, 6, 10)
    kvec = nbody.rfftk(shape)
    d = {"shape": np.array(shape), "kx": kvec[0], "ky": kvec[1], "kz": kvec[2]}
    for fd, tag in [(2, "2"), (4, "4"), (np.inf, "inf")]:
        d[f"invlaplace_{tag}"] = nbody.invlaplace_hat(kvec, fd) * np.ones(utils.r2chshape(shape))
        for i in range(3):
            d[f"gradient{i}_{tag}"] = nbody.gradient_hat(kvec, i, fd) * np.ones(utils.r2chshape(shape))
    d["gaussian_kcut2"] = nbody.gaussian_hat(kvec, 2.0)
    for o in (1, 2, 3, 4):
        d[f"rectangular_hat_{o}"] = nbody.rectangular_hat(kvec, o)
    d["kaiser_bessel_hat_4"] = A(nbody.kaiser_bessel_hat(kvec, 4, A(nbody.optim_kcut(1.5))))
    out["kernels"] = d

    # ---- paint / read, all orders, positions outside the box, scalar and per-particle weights ------------------
    rng = np.random.default_rng(1)
    shape = (12, 10, 14)
    pos = rng.uniform(-20.0, 40.0, (700, 3))
    pos[:40] = np.round(pos[:40])  # exact integers
    pos[40:80] = np.floor(pos[40:80]) + 0.5  # exact half-integers: half-to-even rounding for odd orders
    wts = rng.normal(size=700)
    mesh = rng.normal(size=shape)
    d = {"shape": np.array(shape), "pos": pos, "weights": wts, "mesh": mesh}
    for o in (1, 2, 3, 4):
        d[f"paint_w_{o}"] = A(nbody.paint(jnp.asarray(pos), shape, jnp.asarray(wts), o))
        d[f"paint_1_{o}"] = A(nbody.paint(jnp.asarray(pos), shape, 1.0, o))
        d[f"read_{o}"] = A(nbody.read(jnp.asarray(pos), jnp.asarray(mesh), o))
    d["paint_kb_4"] = A(nbody.paint(jnp.asarray(pos), shape, jnp.asarray(wts), 4, "kaiser_bessel", 1.5))
    d["read_kb_4"] = A(nbody.read(jnp.asarray(pos), jnp.asarray(mesh), 4, "kaiser_bessel", 1.5))
    for o, ov in ((1, 1.0), (2, 2.0), (3, 1.25)):
        d[f"paint_kb_{o}"] = A(nbody.paint(jnp.asarray(pos), shape, jnp.asarray(wts), o, "kaiser_bessel", ov))
        d[f"read_kb_{o}"] = A(nbody.read(jnp.asarray(pos), jnp.asarray(mesh), o, "kaiser_bessel", ov))
    out["paint_read"] = d

    # ---- chreshape: crop and pad, non-cubic ---------------------------------------------------------------------
    rng = np.random.default_rng(2)
    d = {}
    for tag, (src, dst) in {"down": ((12, 10, 8), (8, 6, 6)), "up": ((8, 6, 6), (12, 10, 8)),
                            "mixed": ((8, 12, 6), (10, 8, 10)), "same": ((8, 8, 8), (8, 8, 8))}.items():
        m = np.fft.rfftn(rng.normal(size=src))
        d[f"{tag}_in"] = m
        d[f"{tag}_src"] = np.array(src)
        d[f"{tag}_dst"] = np.array(dst)
        d[f"{tag}_out"] = A(utils.chreshape(jnp.asarray(m), utils.r2chshape(dst)))
    out["chreshape"] = d

    # ---- interlace / deconv / nufft -----------------------------------------------------------------------------
    rng = np.random.default_rng(3)
    final = (8, 8, 8)
    pos = rng.uniform(-2.0, 10.0, (400, 3))
    wts = rng.uniform(0.5, 1.5, 400)
    d = {"final_shape": np.array(final), "pos": pos, "weights": wts}
    d["interlace_2_2"] = A(nbody.interlace(jnp.asarray(pos), final, jnp.asarray(wts), 2, 2))
    d["interlace_4_3"] = A(nbody.interlace(jnp.asarray(pos), final, jnp.asarray(wts), 4, 3))
    d["nufft_same"] = A(nbody.nufft(jnp.asarray(pos), final, None, jnp.asarray(wts), 2, 2))
    d["nufft_over15"] = A(nbody.nufft(jnp.asarray(pos), final, 1.5, jnp.asarray(wts), 2, 2))
    d["nufft_over15_nodeconv_o3"] = A(nbody.nufft(jnp.asarray(pos), final, 1.5, 1.0, 3, 2, paint_deconv=False))
    d["nufft_tuple"] = A(nbody.nufft(jnp.asarray(pos), final, (12, 10, 14), jnp.asarray(wts), 2, 2))
    rmesh = rng.normal(size=final)
    d["deconv_real_in"] = rmesh
    d["deconv_real_2"] = A(nbody.deconv_paint(jnp.asarray(rmesh), 2))
    cm = np.fft.rfftn(rmesh)
    d["deconv_cplx_3"] = A(nbody.deconv_paint(jnp.asarray(cm.copy()), 3))  # the reference divides in place
    # Kaiser-Bessel window through the same chain (nbody.py:280-312, 321-322, 383-384)
    d["deconv_kb_real_4"] = A(nbody.deconv_paint(jnp.asarray(rmesh), 4, "kaiser_bessel", 1.5))
    d["deconv_kb_cplx_2"] = A(nbody.deconv_paint(jnp.asarray(cm.copy()), 2, "kaiser_bessel", 2.0))
    d["interlace_kb_4_2"] = A(nbody.interlace(jnp.asarray(pos), final, jnp.asarray(wts), 4, 2, "kaiser_bessel", 1.5))
    d["nufft_kb_over15"] = A(nbody.nufft(jnp.asarray(pos), final, 1.5, jnp.asarray(wts), 4, 2, "kaiser_bessel"))
    d["nufft_kb_tuple_o2"] = A(nbody.nufft(jnp.asarray(pos), final, (12, 10, 14), jnp.asarray(wts), 2, 2, "kaiser_bessel"))
    out["nufft"] = d

    # ---- growth tables ------------------------------------------------------------------------------------------
    d = {}
    for tag, par in (("abacus", ABACUS), ("other", OTHER)):
        c = Cosmology(**par)
        a = np.array([0.0, 1e-3, 0.0123, 0.1, 0.3333, 0.5, 0.9, 1.0])
        d[f"{tag}_a"] = a
        for name in ("a2g", "a2g2", "a2f", "a2f2", "a2dg2dg"):
            d[f"{tag}_{name}"] = A(getattr(nbody, name)(c, a))
        g = A(nbody.a2g(c, a))
        for name in ("g2a", "g2g2", "g2f", "g2f2", "g2dg2dg"):
            d[f"{tag}_{name}"] = A(getattr(nbody, name)(c, g))
        d[f"{tag}_params"] = np.array([par[k] for k in ("Omega_c", "Omega_b", "h", "n_s", "sigma8")])
        # distances (nbody.py:810-896)
        c._workspace = {}
        d[f"{tag}_a2chi"] = A(nbody.a2chi(c, a))
        chi = np.array([0.0, 10.0, 500.0, 2500.0, 6000.0])
        d[f"{tag}_chi"] = chi
        d[f"{tag}_chi2a"] = A(nbody.chi2a(c, chi))
        d[f"{tag}_k2ell"] = A(nbody.k2ell(c, 0.5, np.array([0.01, 0.1])))
        d[f"{tag}_ell2k"] = A(nbody.ell2k(c, 0.5, np.array([10.0, 1000.0])))
    out["growth"] = d

    # ---- forces, LPT, N-body at 16^3 (and a 12x10x14 non-cubic force case) -------------------------------------
    rng = np.random.default_rng(4)
    shape = (16, 16, 16)
    dk = gaussian_field_k(rng, shape)
    q = lattice(shape)
    pos = q + rng.normal(scale=0.7, size=q.shape)
    c = Cosmology(**ABACUS)
    d = {"shape": np.array(shape), "delta_k": dk, "pos": pos}
    d["pm_forces_paint"] = A(nbody.pm_forces(jnp.asarray(pos), shape, 2))
    d["pm_forces_paint_deconv_kcut"] = A(nbody.pm_forces(jnp.asarray(pos), shape, 2, paint_deconv=True, kcut=2.5))
    d["pm_forces_paint_o3_fd"] = A(nbody.pm_forces(jnp.asarray(pos), shape, 3, grad_fd=4, lap_fd=2))
    d["pm_forces_mesh"] = A(nbody.pm_forces(jnp.asarray(pos), jnp.asarray(dk), 2))
    d["pm_forces_mesh_ngp_lattice"] = A(nbody.pm_forces(jnp.asarray(q), jnp.asarray(dk), 1))
    d["pm_forces2"] = A(nbody.pm_forces2(jnp.asarray(pos), jnp.asarray(dk), 2))
    for order in (1, 2):
        c._workspace = {}
        dp, vl = nbody.lpt(c, jnp.asarray(dk), jnp.asarray(q), 0.3, order, 1)
        d[f"lpt{order}_a0.3_dpos"], d[f"lpt{order}_a0.3_vel"] = A(dp), A(vl)
    c._workspace = {}
    dp, vl = nbody.lpt(c, jnp.asarray(dk), jnp.asarray(pos), 0.0, 2, 2)
    d["lpt2_a0_cic_dpos"], d["lpt2_a0_cic_vel"] = A(dp), A(vl)
    # legacy scale-factor-time pieces (nbody.py:1030-1092)
    for order in (1, 2):
        c._workspace = {}
        dq, pp = nbody.lpt_fpm(c, jnp.asarray(dk), jnp.asarray(pos), 0.3, order, 2)
        d[f"lpt_fpm{order}_dq"], d[f"lpt_fpm{order}_p"] = A(dq), A(pp)
    vel0 = rng.normal(scale=0.3, size=pos.shape)
    d["vf_vel"] = vel0
    # the solver's time is a float64 array; a Python float would meet jax_cosmo's float32 epsilon as a weak scalar
    dpv, dvv = nbody.diffrax_vf(c, shape, 2)(jnp.asarray(0.5), (jnp.asarray(pos), jnp.asarray(vel0)), None)
    d["vf_dpos"], d["vf_dvel"] = A(dpv), A(dvv)
    out["forces_lpt"] = d

    d = {"shape": np.array(shape), "delta_k": dk}
    c._workspace = {}
    p, v = nbody.nbody_bf(c, jnp.asarray(dk), jnp.asarray(q), a0=0.0, a1=1.0, n_steps=4)
    d["bf4_pos"], d["bf4_vel"] = A(p), A(v)
    c._workspace = {}
    p, v = nbody.nbody_bf(c, jnp.asarray(dk), jnp.asarray(q), a0=0.1, a1=0.8, n_steps=3, paint_order=3,
                          lpt_order=1, paint_deconv=True, snapshots=4)
    d["bf3_snap_pos"], d["bf3_snap_vel"] = A(p), A(v)
    # save times inside steps (dense output of the Euler solver), scale-factor list, custom save function
    c._workspace = {}
    p, v = nbody.nbody_bf(c, jnp.asarray(dk), jnp.asarray(q), a0=0.1, a1=0.8, n_steps=3, snapshots=3)
    d["bf3_mid_pos"], d["bf3_mid_vel"] = A(p), A(v)
    c._workspace = {}
    d["bf3_alist"] = np.array([0.15, 0.4, 0.8])
    p, v = nbody.nbody_bf(c, jnp.asarray(dk), jnp.asarray(q), a0=0.1, a1=0.8, n_steps=3, snapshots=list(d["bf3_alist"]))
    d["bf3_alist_pos"], d["bf3_alist_vel"] = A(p), A(v)
    c._workspace = {}
    d["bf3_fn_disp"] = A(nbody.nbody_bf(c, jnp.asarray(dk), jnp.asarray(q), a0=0.1, a1=0.8, n_steps=3, snapshots=5,
                                        fn=lambda t, y, args: y[0] - jnp.asarray(q)))
    # BullFrog coefficients per step, as the engine receives them (nbody.py:907-919, 933-938)
    c._workspace = {}
    g0, g1 = float(nbody.a2g(c, 0.0)), float(nbody.a2g(c, 1.0))
    dg = (g1 - g0) / 4
    vf = nbody.bullfrog_vf(c, dg, shape)
    alphas = []
    for n in range(4):
        gg0 = g0 + n * dg
        gm, ge = gg0 + dg / 2, gg0 + dg
        d0, d2 = float(nbody.g2dg2dg(c, gg0)), float(nbody.g2dg2dg(c, ge))
        lin = (float(nbody.g2g2(c, gg0)) + d0 * dg / 2) / gm - gm
        alphas.append((d2 - lin) / (d0 - lin))
    d["bf4_alpha"] = np.array(alphas)
    # alpha_fpm (nbody.py:921-931) is a closure of bullfrog_vf that nothing calls: its formula, evaluated with the
    # reference's own growth helpers and background
    from jax_cosmo import background
    fpm = []
    for n in range(4):
        gg0 = g0 + n * dg
        gg2 = gg0 + dg
        a0_, a2_ = nbody.g2a(c, gg0), nbody.g2a(c, gg2)
        c0 = background.Esqr(c, a0_) ** .5 * gg0 * nbody.g2f(c, gg0) * a0_ ** 2
        c2 = background.Esqr(c, a2_) ** .5 * gg2 * nbody.g2f(c, gg2) * a2_ ** 2
        fpm.append(float(c0 / c2))
    d["bf4_alpha_fpm"] = np.array(fpm)
    d["bf4_g0_dg"] = np.array([g0, dg])
    out["nbody"] = d

    # non-cubic force + small particle count != cells
    rng = np.random.default_rng(5)
    shape = (12, 10, 14)
    dk = gaussian_field_k(rng, shape)
    pos = rng.uniform(-5, 20, (333, 3))
    d = {"shape": np.array(shape), "delta_k": dk, "pos": pos}
    d["pm_forces_paint"] = A(nbody.pm_forces(jnp.asarray(pos), shape, 2))
    d["pm_forces_mesh"] = A(nbody.pm_forces(jnp.asarray(pos), jnp.asarray(dk), 2))
    d["pm_forces2"] = A(nbody.pm_forces2(jnp.asarray(pos), jnp.asarray(dk), 2))
    out["forces_noncubic"] = d

    # ---- white-noise permutation (next row f-2) -----------------------------------------------------------------
    rng = np.random.default_rng(6)
    w = rng.normal(size=(8, 6, 10))
    d = {"white": w, "rg2cgh": A(utils.rg2cgh(jnp.asarray(w)))}
    d["cgh2rg_roundtrip"] = A(utils.cgh2rg(jnp.asarray(d["rg2cgh"])))
    ampk = rng.uniform(0.5, 2.0, size=(8, 6, 6))  # an amplitude per mode, deliberately NOT even in k
    d["ampk"] = ampk
    d["cgh2rg_amp"] = A(utils.cgh2rg(jnp.asarray(ampk + 0j), norm="amp"))
    d["rg2cgh_amp"] = A(utils.rg2cgh(jnp.asarray(w), norm="amp"))
    out["rg2cgh"] = d

    # ---- power-spectrum estimator and Lagrangian bias (next rows f-4, f-1) --------------------------------------
    metrics, bricks = load_reference_next_rows()
    rng = np.random.default_rng(7)
    a = rng.normal(size=(12, 8, 10))
    b = 0.7 * a + 0.5 * rng.normal(size=(12, 8, 10))
    box = (100.0, 80.0, 120.0)
    d = {"mesh0": a, "mesh1": b, "box_size": np.array(box)}
    for tag, kw in [("default", dict(kedges=None, include_corners=True, deconv=2)),
                    ("n5_nocorners", dict(kedges=5, include_corners=False, deconv=0)),
                    ("dk02", dict(kedges=0.2, include_corners=True, deconv=(1, 2)))]:
        kc, km, p = metrics._spectrum(jnp.asarray(a), None, box_size=box, **kw)
        d[f"auto_{tag}_kcount"], d[f"auto_{tag}_kmean"], d[f"auto_{tag}_pow"] = A(kc), A(km), A(p)
        kc, km, p = metrics._spectrum(jnp.asarray(a), jnp.asarray(b), box_size=box, **kw)
        d[f"cross_{tag}_pow"] = A(p)
    # multipoles along box_center / |box_center| (metrics.py:127-128, 165-166); a centred box has mu = 0
    center = (30.0, -20.0, 90.0)
    d["box_center"] = np.array(center)
    kc, km, p = metrics._spectrum(jnp.asarray(a), None, box_size=box, box_center=center, ells=[0, 2, 4], deconv=2)
    for ell in (0, 2, 4):
        d[f"auto_ell{ell}_pow"] = A(p[ell])
    kc, km, p = metrics._spectrum(jnp.asarray(a), jnp.asarray(b), box_size=box, box_center=center, ells=[1, 2], kedges=6)
    d["cross_ell1_pow"], d["cross_ell2_pow"] = A(p[1]), A(p[2])
    kc, km, p = metrics._spectrum(jnp.asarray(a), None, box_size=box, ells=2)
    d["auto_ell2_centred_pow"] = A(p)
    ks, p1, tr, coh = metrics.powtranscoh(jnp.asarray(a), jnp.asarray(b), np.array(box))
    d["ptc_k"], d["ptc_pow1"], d["ptc_trans"], d["ptc_coh"] = A(ks), A(p1), A(tr), A(coh)
    out["spectrum"] = d

    rng = np.random.default_rng(8)
    shape, box = (8, 10, 12), (80.0, 100.0, 96.0)
    dk = np.fft.rfftn(rng.normal(size=shape)) * 0.05
    pos = lattice(shape) + rng.normal(scale=0.4, size=(int(np.prod(shape)), 3))
    bias = dict(b1=0.8, b2=0.3, bs2=-0.2, b3=0.1, bds2=0.05, bs3=-0.07, bn2=0.4, bnpar=0.6)
    png = dict(fNL_bp=0.0, fNL_bpd=0.0, fNL_bpd2=0.0, fNL_bps2=0.0, fNL_bn2p=0.0)
    cosmo = Cosmology(**ABACUS)
    cosmo._workspace = {}
    w, dvel, _ = bricks.lagrangian_bias(cosmo, jnp.asarray(pos), jnp.asarray(0.7), np.array(box), jnp.asarray(dk),
                                        bias, png, png_type=None, read_order=2)
    d = {"shape": np.array(shape), "box_size": np.array(box), "delta_k": dk, "pos": pos, "a": np.array(0.7),
         "weights": A(w), "dvel": A(dvel)}
    d.update({f"bias_{k}": np.array(v) for k, v in bias.items()})
    # primordial non-Gaussianity terms (bricks.py:411-438) with a tabulated (k, P) normalised to sigma8 = 1
    ks = np.logspace(-4, 1, 64)
    pk = 2.0e4 * (ks / 0.02) ** 0.96 / (1 + (ks / 0.02) ** 2) ** 1.7
    png = dict(fNL_bp=0.9, fNL_bpd=-0.4, fNL_bpd2=0.25, fNL_bps2=0.15, fNL_bn2p=-0.3)
    cosmo._workspace = {}
    w, dvel, phi = bricks.lagrangian_bias(cosmo, jnp.asarray(pos), jnp.asarray(0.7), np.array(box), jnp.asarray(dk),
                                          bias, png, png_type="fNL", kpow=(ks, pk), read_order=2)
    d.update({"kpow_k": ks, "kpow_p": pk, "png_weights": A(w), "png_phi": A(phi)})
    d.update({f"png_{k}": np.array(v) for k, v in png.items()})
    cosmo._workspace = {}
    d["add_png_fNL50"] = A(bricks.add_png(cosmo, 50.0, jnp.asarray(dk), np.array(box), kpow=(ks, pk)))
    # power, bias parametrisations and the Eulerian expansion (bricks.py:67-166, 454-585)
    cosmo._workspace = {}
    d["lin_power_mesh"] = A(bricks.lin_power_mesh(cosmo, shape, np.array(box), kpow=(ks, pk)))
    d["trans_phi2delta"] = A(bricks.trans_phi2delta_interp(cosmo, kpow=(ks, pk))(np.array([1e-3, 0.05, 0.7])))
    d["white2lin"] = A(bricks.white2lin(cosmo, jnp.asarray(dk), shape, np.array(box), kpow=(ks, pk)))
    d["lin2white"] = A(bricks.lin2white(cosmo, jnp.asarray(d["white2lin"]), shape, np.array(box), kpow=(ks, pk)))
    phik = np.fft.rfftn(rng.normal(size=shape)) * 0.02
    d["eul_phik"] = phik
    png_e = dict(fNL=20.0, fNL_bp=0.0, fNL_bpd=0.0)
    png_e = bricks.fNL_bias(png_e, bias, p=1.0, png_type="fNL")
    d["fNL_bias"] = np.array([png_e["fNL_bp"], png_e["fNL_bpd"]])
    we, _ = bricks.eulerian_bias(jnp.asarray(dk.copy()), jnp.asarray(phik), np.array(box), bias, png_e, png_type="fNL")
    d["eulerian_weights_png"] = A(we)
    we, _ = bricks.eulerian_bias(jnp.asarray(dk.copy()), jnp.asarray(phik), np.array(box), bias, png_e, png_type=None)
    d["eulerian_weights"] = A(we)
    cm = rng.uniform(0.5, 2.0, size=shape)
    sm = rng.uniform(0.1, 1.0, size=shape)
    d["count_mesh"], d["selec_mesh"] = cm, sm
    d["count2delta"] = A(bricks.count2delta(jnp.asarray(cm), jnp.asarray(sm)))
    out["lagrangian_bias"] = d

    # ---- cell -> physical -> redshift space (bricks.py:628-877, next row f-3) ------------------------------------
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(9)
    shape, box, center = (8, 10, 12), (800.0, 1000.0, 960.0), (900.0, -300.0, 1500.0)
    rot = Rotation.from_rotvec([0.3, -0.5, 0.8])
    pos = rng.uniform(0, 1, (200, 3)) * np.array(shape)
    vel = rng.normal(scale=0.05, size=(200, 3))
    cosmo, fid = Cosmology(**ABACUS), Cosmology(**OTHER)
    d = {"shape": np.array(shape), "box_size": np.array(box), "box_center": np.array(center),
         "rot_matrix": rot.as_matrix(), "pos": pos, "vel": vel}
    phys = A(bricks.cell2phys_pos(jnp.asarray(pos), center, rot, box, shape))
    d["cell2phys_pos"] = phys
    d["phys2cell_roundtrip"] = A(bricks.phys2cell_pos(jnp.asarray(phys), center, rot, box, shape))
    d["cell2phys_vel"] = A(bricks.cell2phys_vel(jnp.asarray(vel), rot, box, shape))
    d["phys2cell_vel"] = A(bricks.phys2cell_vel(jnp.asarray(vel), rot, box, shape))
    d["pos_mesh"] = A(bricks.pos_mesh(center, rot, box, shape))
    for tag, curved in (("curved", True), ("flat", False)):
        d[f"radius_mesh_{tag}"] = A(bricks.radius_mesh(np.array(center), rot, box, shape, curved))
        cosmo._workspace = {}
        los, a = bricks.los_scalefactor_pos(jnp.asarray(pos), np.array(center), rot, box, shape, cosmo, None, curved)
        d[f"los_{tag}"], d[f"a_{tag}"] = A(los) * np.ones((1, 3)), A(a)
        dpos = bricks.rsd(cosmo, jnp.asarray(vel), jnp.asarray(los), jnp.asarray(a), rot, box, shape, dvel=0.01)
        d[f"rsd_{tag}"] = A(dpos)
        fid._workspace = {}
        d[f"ap_auto_{tag}"] = A(bricks.ap_auto(jnp.asarray(phys), jnp.asarray(los), cosmo, fid, curved))
        d[f"ap_param_{tag}"] = A(bricks.ap_param(jnp.asarray(phys.copy()), jnp.asarray(los),
                                                 dict(alpha_iso=1.03, alpha_ap=0.97), curved))
        rpos = np.linalg.norm(phys, axis=-1, keepdims=True) if curved else np.abs((phys * A(los)).sum(-1, keepdims=True))
        d[f"rsd_ap_auto_{tag}"] = A(bricks.rsd_ap_auto(jnp.asarray(phys.copy()), jnp.asarray(d["cell2phys_vel"]),
                                                       jnp.asarray(rpos), jnp.asarray(los), jnp.asarray(a), cosmo, fid,
                                                       curved))
    d["scale_pos"] = A(bricks.scale_pos(jnp.asarray(phys.copy()), jnp.asarray(d["los_flat"][:1]), 1.1, 0.9))
    d["isoap2parperp"] = np.array(bricks.isoap2parperp(1.03, 0.97))
    d["parperp2isoap"] = np.array(bricks.parperp2isoap(1.05, 0.98))
    cosmo._workspace = {}
    red, aa = bricks.redges_and_scalefactors(cosmo, 500.0, 2500.0, 4)
    d["redges"], d["redges_a"] = A(red), A(aa)
    # Kaiser model (bricks.py:170-232): flat sky, flat sky on the light cone, curved sky
    shape, box = (8, 10, 12), (800.0, 1000.0, 960.0)
    dkk = np.fft.rfftn(rng.normal(size=shape)) * 0.05
    dkk[0, 0, 0] = 0.0
    ks = np.logspace(-4, 1, 64)
    pk = 2.0e4 * (ks / 0.02) ** 0.96 / (1 + (ks / 0.02) ** 2) ** 1.7
    flat_los = np.array(center) / np.linalg.norm(center)
    d["kaiser_delta_k"], d["kaiser_shape"], d["kaiser_box"], d["kaiser_los"] = dkk, np.array(shape), np.array(box), flat_los
    d["kaiser_kpow_k"], d["kaiser_kpow_p"] = ks, pk
    cosmo._workspace = {}
    d["kaiser_boost"] = A(bricks.kaiser_boost(cosmo, 0.7, shape, box, 1.8, 0.5, "fNL", flat_los, (ks, pk)))
    d["kaiser_flat"] = A(bricks.kaiser_model(cosmo, 0.7, jnp.asarray(dkk.copy()), box, 1.8, 0.5, "fNL", flat_los, (ks, pk)))
    los_m, a_m = bricks.los_scalefactor_mesh(np.array(center), rot, box, shape, cosmo, None, False)
    d["kaiser_a_mesh_flat"] = A(a_m)
    d["kaiser_lightcone"] = A(bricks.kaiser_model(cosmo, jnp.asarray(a_m), jnp.asarray(dkk.copy()), box, 1.8,
                                                  los=jnp.asarray(los_m)))
    los_c, a_c = bricks.los_scalefactor_mesh(np.array(center), rot, box, shape, cosmo, None, True)
    cell_los = rot.apply(np.asarray(los_c).reshape(-1, 3), inverse=True).reshape(tuple(shape) + (3,))
    d["kaiser_a_mesh_curved"], d["kaiser_cell_los"] = A(a_c), cell_los
    d["kaiser_curved"] = A(bricks.kaiser_model(cosmo, jnp.asarray(a_c), jnp.asarray(dkk.copy()), box, 1.8,
                                               los=jnp.asarray(cell_los)))
    # catalogue registration (bricks.py:879-897, 1026-1100)
    nobj = 300
    cat = {"RA": rng.uniform(20.0, 40.0, nobj), "DEC": rng.uniform(-10.0, 15.0, nobj), "Z": rng.uniform(0.4, 0.6, nobj),
           "WEIGHT": rng.uniform(0.5, 1.5, nobj)}
    cosmo._workspace = {}
    cart = A(bricks.radecz2cart(cosmo, cat))
    csize, ccenter, crotvec = (500.0, 600.0, 550.0), tuple(cart.mean(0)), (0.1, 0.2, -0.3)
    d.update({f"cat_{k}": v for k, v in cat.items()})
    d["cat_cart"], d["cat_box"], d["cat_center"], d["cat_rotvec"] = cart, np.array(csize), np.array(ccenter), np.array(crotvec)
    back = bricks.cart2radecz(cosmo, jnp.asarray(cart))
    d["cat_back"] = np.stack([A(back["RA"]), A(back["DEC"]), A(back["Z"])])
    d["cutsky2count"] = A(bricks.cutsky2count(cat, cosmo, (12, 14, 12), 1.5, csize, ccenter, crotvec))
    sel, msk = bricks.cutsky2selection(cat, cosmo, (6, 8, 6), (12, 14, 12), None, csize, ccenter, crotvec)
    d["cutsky_selection"], d["cutsky_mask"] = A(sel), A(msk)
    ppos = rng.uniform(-250.0, 250.0, (nobj, 3)) + np.array(ccenter)
    pvel = rng.normal(scale=300.0, size=(nobj, 3))
    d["full_pos"], d["full_vel"] = ppos, pvel
    flos = np.array(ccenter) / np.linalg.norm(ccenter)
    chunks = [{"pos": ppos[:100], "vel": pvel[:100], "WEIGHT": cat["WEIGHT"][:100]},
              {"pos": ppos[100:], "vel": pvel[100:], "WEIGHT": cat["WEIGHT"][100:]}]
    d["fullsky2count"] = A(bricks.fullsky2count(chunks, cosmo, 0.65, flos, csize, ccenter, crotvec, (12, 14, 12), None))
    out["observation"] = d

    # ---- the whole observation chain followed by nufft, in the order model.py:780-805 composes it (what the engine
    # applies inside its paint kernels: mcpm_nufft_obs) -- its own generator, so the fixtures above keep their streams
    rng = np.random.default_rng(21)
    shape, paint, box, center = (12, 10, 14), (18, 16, 20), (600.0, 500.0, 700.0), (250.0, -300.0, 1400.0)
    rot = Rotation.from_rotvec([0.3, 0.2, -0.5])
    q = lattice(shape)
    pos = q + rng.normal(scale=0.7, size=q.shape)
    vel = rng.normal(scale=0.3, size=q.shape)
    dvel = rng.normal(scale=2.0, size=q.shape)
    wts = rng.uniform(0.3, 2.0, q.shape[0])
    ap = dict(alpha_iso=1.03, alpha_ap=0.97)
    d = {"shape": np.array(shape), "paint_shape": np.array(paint), "box_size": np.array(box), "box_center": np.array(center),
         "rot_matrix": rot.as_matrix(), "pos": pos, "vel": vel, "dvel": dvel, "weights": wts,
         "alpha_iso": np.array(ap["alpha_iso"]), "alpha_ap": np.array(ap["alpha_ap"]), "a_obs": np.array(0.7)}
    for tag, curved, a_obs, ap_auto in (("curved_lightcone_auto", True, None, True), ("flat_lightcone_auto", False, None, True),
                                        ("curved_scalar_param", True, 0.7, False), ("flat_scalar_param", False, 0.7, False),
                                        ("flat_scalar_plain", False, 0.7, None)):
        cosmo._workspace, fid._workspace = {}, {}
        los, a = bricks.los_scalefactor_pos(jnp.asarray(pos), np.array(center), rot, box, shape, cosmo, a_obs, curved)
        p = bricks.cell2phys_pos(jnp.asarray(pos), center, rot, box, shape)
        p = p + bricks.rsd(cosmo, jnp.asarray(vel), los, a, rot, box, shape, jnp.asarray(dvel))
        if ap_auto is not None:
            p = bricks.ap_auto(p, los, cosmo, fid, curved) if ap_auto else bricks.ap_param(p, los, dict(ap), curved)
        p = bricks.phys2cell_pos(p, center, rot, box, shape)
        d[f"pos_{tag}"] = A(p)
        d[f"nufft_{tag}"] = A(nbody.nufft(p, shape, paint, jnp.asarray(wts), 2, 2, "rectangular", True))
    out["observed_nufft"] = d

    only = set(sys.argv[1:])  # e.g. `make_golden.py observed_nufft`: write that fixture alone
    for name, dd in out.items():
        if only and name not in only:
            continue
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **{k: np.asarray(v) for k, v in dd.items()})
        print(f"{name:18s} {os.path.getsize(path) / 1024:8.1f} KiB  {len(dd)} arrays")


if __name__ == "__main__":
    main()
