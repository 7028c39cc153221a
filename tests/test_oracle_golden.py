"""The CPU oracle (oracle/pm_oracle.py) against golden vectors produced by the reference's own source
(tests/golden/make_golden.py).  float64 both sides: agreement to rounding."""
import numpy as np
import pytest
import torch

from oracle import pm_oracle as O

RTOL, ATOL = 1e-11, 1e-12


def close(a, b, rtol=RTOL, atol=ATOL):
    a = a.detach().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol * max(1.0, np.abs(b).max()))


def test_kernels(golden):
    g = golden("kernels")
    shape = tuple(g["shape"])
    kvec = O.rfftk(shape)
    for k, name in zip(kvec, ("kx", "ky", "kz")):
        close(k, g[name])
    ones = np.ones(O.r2chshape(shape))
    for fd, tag in [(2, "2"), (4, "4"), (np.inf, "inf")]:
        close(O.invlaplace_hat(kvec, fd) * ones, g[f"invlaplace_{tag}"])
        for i in range(3):
            close(O.gradient_hat(kvec, i, fd) * ones, g[f"gradient{i}_{tag}"])
    close(O.gaussian_hat(kvec, 2.0), g["gaussian_kcut2"])
    for o in (1, 2, 3, 4):
        close(O.rectangular_hat(kvec, o), g[f"rectangular_hat_{o}"])
    close(O.kaiser_bessel_hat(kvec, 4, O.optim_kcut(1.5)), g["kaiser_bessel_hat_4"])


@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_paint_read(golden, order):
    g = golden("paint_read")
    shape = tuple(int(s) for s in g["shape"])
    close(O.paint(g["pos"], shape, g["weights"], order), g[f"paint_w_{order}"])
    close(O.paint(g["pos"], shape, 1.0, order), g[f"paint_1_{order}"])
    close(O.read(g["pos"], g["mesh"], order), g[f"read_{order}"])


def test_paint_read_kaiser_bessel(golden):
    g = golden("paint_read")
    shape = tuple(int(s) for s in g["shape"])
    close(O.paint(g["pos"], shape, g["weights"], 4, "kaiser_bessel", 1.5), g["paint_kb_4"])
    close(O.read(g["pos"], g["mesh"], 4, "kaiser_bessel", 1.5), g["read_kb_4"])
    for o, ov in ((1, 1.0), (2, 2.0), (3, 1.25)):
        close(O.paint(g["pos"], shape, g["weights"], o, "kaiser_bessel", ov), g[f"paint_kb_{o}"])
        close(O.read(g["pos"], g["mesh"], o, "kaiser_bessel", ov), g[f"read_kb_{o}"])
    g = golden("nufft")
    final = tuple(int(s) for s in g["final_shape"])
    close(O.deconv_paint(O._t(g["deconv_real_in"]), 4, "kaiser_bessel", 1.5), g["deconv_kb_real_4"])
    close(O.deconv_paint(O._t(np.fft.rfftn(g["deconv_real_in"]), O.C128), 2, "kaiser_bessel", 2.0), g["deconv_kb_cplx_2"])
    close(O.interlace(g["pos"], final, g["weights"], 4, 2, "kaiser_bessel", 1.5), g["interlace_kb_4_2"])
    close(O.nufft(g["pos"], final, 1.5, g["weights"], 4, 2, "kaiser_bessel"), g["nufft_kb_over15"])
    close(O.nufft(g["pos"], final, (12, 10, 14), g["weights"], 2, 2, "kaiser_bessel"), g["nufft_kb_tuple_o2"])


@pytest.mark.parametrize("tag", ["down", "up", "mixed", "same"])
def test_chreshape(golden, tag):
    g = golden("chreshape")
    dst = tuple(int(s) for s in g[f"{tag}_dst"])
    close(O.chreshape(g[f"{tag}_in"], O.r2chshape(dst)), g[f"{tag}_out"])


def test_nufft(golden):
    g = golden("nufft")
    final = tuple(int(s) for s in g["final_shape"])
    pos, w = g["pos"], g["weights"]
    close(O.interlace(O._t(pos), final, w, 2, 2), g["interlace_2_2"])
    close(O.interlace(O._t(pos), final, w, 4, 3), g["interlace_4_3"])
    close(O.nufft(pos, final, None, w, 2, 2), g["nufft_same"])
    close(O.nufft(pos, final, 1.5, w, 2, 2), g["nufft_over15"])
    close(O.nufft(pos, final, 1.5, 1.0, 3, 2, paint_deconv=False), g["nufft_over15_nodeconv_o3"])
    close(O.nufft(pos, final, (12, 10, 14), w, 2, 2), g["nufft_tuple"])
    close(O.deconv_paint(O._t(g["deconv_real_in"]), 2), g["deconv_real_2"])
    close(O.deconv_paint(O._t(np.fft.rfftn(g["deconv_real_in"]), O.C128), 3), g["deconv_cplx_3"])


@pytest.mark.parametrize("tag", ["abacus", "other"])
def test_growth(golden, tag):
    g = golden("growth")
    oc, ob, h, ns, s8 = g[f"{tag}_params"]
    c = O.Cosmology(Omega_c=oc, Omega_b=ob, h=h, n_s=ns, sigma8=s8)
    a = g[f"{tag}_a"]
    for name in ("a2g", "a2g2", "a2f", "a2f2", "a2dg2dg"):
        close(getattr(O, name)(c, a), g[f"{tag}_{name}"], rtol=1e-10)
    gg = g[f"{tag}_a2g"]
    for name in ("g2a", "g2g2", "g2f", "g2f2", "g2dg2dg"):
        close(getattr(O, name)(c, gg), g[f"{tag}_{name}"], rtol=1e-10)
    # distances (nbody.py:810-896), oracle and the engine's host-side mirror (cosmo.py; no device involved)
    from montecosmo_b200 import cosmo as MC
    mc = MC.Cosmology(Omega_c=oc, Omega_b=ob, h=h, n_s=ns, sigma8=s8)
    close(O.a2chi(c, a), g[f"{tag}_a2chi"], rtol=1e-10)
    close(O.chi2a(c, g[f"{tag}_chi"]), g[f"{tag}_chi2a"], rtol=1e-10)
    close(MC.a2chi(mc, a), g[f"{tag}_a2chi"], rtol=1e-10)
    close(MC.chi2a(mc, g[f"{tag}_chi"]), g[f"{tag}_chi2a"], rtol=1e-10)
    close(MC.k2ell(mc, 0.5, np.array([0.01, 0.1])), g[f"{tag}_k2ell"], rtol=1e-10)
    close(MC.ell2k(mc, 0.5, np.array([10.0, 1000.0])), g[f"{tag}_ell2k"], rtol=1e-10)
    for name in ("a2g", "a2g2", "a2f", "a2f2", "a2dg2dg"):
        close(getattr(MC, name)(mc, a), g[f"{tag}_{name}"], rtol=1e-10)


@pytest.mark.parametrize("name", ["forces_lpt", "forces_noncubic"])
def test_forces(golden, name):
    g = golden(name)
    shape = tuple(int(s) for s in g["shape"])
    pos, dk = g["pos"], O._t(g["delta_k"], O.C128)
    close(O.pm_forces(pos, shape, 2), g["pm_forces_paint"], rtol=1e-9)
    close(O.pm_forces(pos, dk, 2), g["pm_forces_mesh"], rtol=1e-9)
    close(O.pm_forces2(pos, dk, 2), g["pm_forces2"], rtol=1e-9)
    if name == "forces_lpt":
        close(O.pm_forces(pos, shape, 2, paint_deconv=True, kcut=2.5), g["pm_forces_paint_deconv_kcut"], rtol=1e-9)
        close(O.pm_forces(pos, shape, 3, grad_fd=4, lap_fd=2), g["pm_forces_paint_o3_fd"], rtol=1e-9)
        q = O.regular_pos(shape)
        close(O.pm_forces(q, dk, 1), g["pm_forces_mesh_ngp_lattice"], rtol=1e-9)


def test_lpt(golden):
    g = golden("forces_lpt")
    shape = tuple(int(s) for s in g["shape"])
    dk = O._t(g["delta_k"], O.C128)
    q = O.regular_pos(shape)
    c = O.Cosmology()
    for order in (1, 2):
        dp, vl = O.lpt(c, dk, q, 0.3, order, 1)
        close(dp, g[f"lpt{order}_a0.3_dpos"], rtol=1e-9)
        close(vl, g[f"lpt{order}_a0.3_vel"], rtol=1e-9)
    dp, vl = O.lpt(c, dk, g["pos"], 0.0, 2, 2)
    close(dp, g["lpt2_a0_cic_dpos"], rtol=1e-9)
    close(vl, g["lpt2_a0_cic_vel"], rtol=1e-9)
    # legacy scale-factor-time pieces (nbody.py:1030-1092)
    for order in (1, 2):
        dq, pp = O.lpt_fpm(c, dk, g["pos"], 0.3, order, 2)
        close(dq, g[f"lpt_fpm{order}_dq"], rtol=1e-9)
        close(pp, g[f"lpt_fpm{order}_p"], rtol=1e-9)
    dpv, dvv = O.diffrax_vf(c, shape, 2)(0.5, (O._t(g["pos"]), O._t(g["vf_vel"])), None)
    close(dpv, g["vf_dpos"], rtol=1e-9)
    close(dvv, g["vf_dvel"], rtol=1e-9)


def test_nbody_bf(golden):
    g = golden("nbody")
    shape = tuple(int(s) for s in g["shape"])
    dk = O._t(g["delta_k"], O.C128)
    q = O.regular_pos(shape)
    c = O.Cosmology()
    p, v = O.nbody_bf(c, dk, q, 0.0, 1.0, 4)
    close(p, g["bf4_pos"], rtol=1e-8, atol=1e-10)
    close(v, g["bf4_vel"], rtol=1e-8, atol=1e-10)
    p, v = O.nbody_bf(c, dk, q, 0.1, 0.8, 3, paint_order=3, lpt_order=1, paint_deconv=True, snapshots=4)
    close(p, g["bf3_snap_pos"], rtol=1e-8, atol=1e-10)
    close(v, g["bf3_snap_vel"], rtol=1e-8, atol=1e-10)
    # save times inside steps, a list of scale factors, a custom save function (nbody.py:987-996)
    p, v = O.nbody_bf(c, dk, q, 0.1, 0.8, 3, snapshots=3)
    close(p, g["bf3_mid_pos"], rtol=1e-8, atol=1e-10)
    close(v, g["bf3_mid_vel"], rtol=1e-8, atol=1e-10)
    p, v = O.nbody_bf(c, dk, q, 0.1, 0.8, 3, snapshots=list(g["bf3_alist"]))
    close(p, g["bf3_alist_pos"], rtol=1e-8, atol=1e-10)
    close(v, g["bf3_alist_vel"], rtol=1e-8, atol=1e-10)
    d = O.nbody_bf(c, dk, q, 0.1, 0.8, 3, snapshots=5, fn=lambda t, y, args: y[0] - q)
    close(d, g["bf3_fn_disp"], rtol=1e-8, atol=1e-10)
    g0, dg = g["bf4_g0_dg"]
    alphas = [float(O.alpha_bf(c, g0 + n * dg, dg)) for n in range(4)]
    close(np.array(alphas), g["bf4_alpha"], rtol=1e-10)
    close(np.array([float(O.alpha_fpm(c, g0 + n * dg, dg)) for n in range(4)]), g["bf4_alpha_fpm"], rtol=1e-10)


def test_oracle_properties():
    """Reference-independent checks listed in SURVEY.md section 4 / 8c."""
    rng = np.random.default_rng(0)
    shape = (8, 10, 6)
    pos = O._t(rng.uniform(-3, 12, (200, 3)))
    w = O._t(rng.normal(size=200))
    m = O._t(rng.normal(size=shape))
    for order in (1, 2, 3, 4):
        # weight conservation (bricks.py:1101-1102) and read = paint^T
        assert abs(float(O.paint(pos, shape, w, order).sum() - w.sum())) < 1e-10
        lhs = float((O.read(pos, m, order) * w).sum())
        rhs = float((m * O.paint(pos, shape, w, order)).sum())
        assert abs(lhs - rhs) < 1e-10
    # NGP / CIC read on lattice returns the mesh values (nbody.py:602)
    q = O.regular_pos(shape)
    for order in (1, 2):
        close(O.read(q, m, order), m.reshape(-1).numpy())
    # chreshape up-then-down is the identity on a band-limited field, and preserves the mean
    mk = torch.fft.rfftn(m)
    up = O.chreshape(mk, O.r2chshape((12, 14, 10)))
    close(O.chreshape(up, O.r2chshape(shape)), mk.numpy(), rtol=1e-10)
    assert abs(float(torch.fft.irfftn(up, s=(12, 14, 10)).mean() - m.mean())) < 1e-12


def test_growth_against_independent_quadrature():
    """The growth helpers are pinned to golden vectors of the reference source run under a stand-in for jax_cosmo that the
    builder wrote too.  Independent of both: for matter + Lambda (w = -1, no radiation -- the physics
    of jax_cosmo's growth ODE) the growing mode is Heath's integral D(a) ~ H(a) int_0^a da' / (a' H(a'))^3.  D(a) / D(1) and
    f = dlnD / dlna from scipy quadrature against the oracle's and the engine-side table's a2g / a2f: 5e-4 (measured: up to
    2.1e-4 -- the restated jax_cosmo table is a 128-point RK4 solve started at a = 1e-2 on the matter-dominated solution,
    which is exact only as a -> 0; a transcription error in the ODE or its normalisation would show at the percent level)."""
    from scipy.integrate import quad
    from montecosmo_b200 import cosmo as PC
    for Oc, Ob, Ok in ((0.26447041, 0.04930169, 0.0), (0.20, 0.05, 0.0), (0.30, 0.05, 0.05)):
        om = Oc + Ob
        ol = 1.0 - om - Ok
        E = lambda a: np.sqrt(om / a**3 + Ok / a**2 + ol)
        Dun = lambda a: E(a) * quad(lambda x: 1.0 / (x * E(x)) ** 3, 0.0, a, epsabs=1e-14, epsrel=1e-12)[0]
        for a in (0.1, 0.3, 0.6, 1.0):
            D = Dun(a) / Dun(1.0)
            h = 1e-4 * a
            f = (np.log(Dun(a + h)) - np.log(Dun(a - h))) / (np.log(a + h) - np.log(a - h))
            co, cp = O.Cosmology(Omega_c=Oc, Omega_b=Ob, Omega_k=Ok), PC.Cosmology(Omega_c=Oc, Omega_b=Ob, Omega_k=Ok)
            for name, g, ff in (("oracle", float(O.a2g(co, a)), float(O.a2f(co, a))),
                                ("engine", float(PC.a2g(cp, a)), float(PC.a2f(cp, a)))):
                assert abs(g / D - 1) < 5e-4 and abs(ff / f - 1) < 5e-4, (name, Oc, Ok, a, g, D, ff, f)


def test_spectrum_estimator(golden):
    """oracle/metrics_oracle.py against montecosmo/metrics.py's own _spectrum / powtranscoh."""
    from oracle import metrics_oracle as MX
    g = golden("spectrum")
    a, b, box = g["mesh0"], g["mesh1"], tuple(g["box_size"])
    for tag, kw in [("default", dict(kedges=None, include_corners=True, deconv=(2, 2))),
                    ("n5_nocorners", dict(kedges=5, include_corners=False, deconv=(0, 0))),
                    ("dk02", dict(kedges=0.2, include_corners=True, deconv=(1, 2)))]:
        kc, km, p = MX.spectrum(a, box_size=box, **kw)
        assert np.array_equal(kc, g[f"auto_{tag}_kcount"])
        close(km, g[f"auto_{tag}_kmean"])
        close(p, g[f"auto_{tag}_pow"])
        close(MX.spectrum(a, b, box_size=box, **kw)[2], g[f"cross_{tag}_pow"])
    center = tuple(g["box_center"])
    p = MX.spectrum(a, box_size=box, box_center=center, ells=[0, 2, 4], deconv=(2, 2))[2]
    for ell in (0, 2, 4):
        close(p[ell], g[f"auto_ell{ell}_pow"], rtol=1e-9, atol=1e-9)
    p = MX.spectrum(a, b, box_size=box, box_center=center, ells=[1, 2], kedges=6)[2]
    close(p[1], g["cross_ell1_pow"], rtol=1e-9, atol=1e-9)
    close(p[2], g["cross_ell2_pow"], rtol=1e-9, atol=1e-9)
    close(MX.spectrum(a, box_size=box, ells=2)[2], g["auto_ell2_centred_pow"], rtol=1e-9, atol=1e-9)
    _, km, p0 = MX.spectrum(a, box_size=box)
    _, _, p1 = MX.spectrum(b, box_size=box)
    _, _, p01 = MX.spectrum(a, b, box_size=box)
    close(km, g["ptc_k"])
    close(p1, g["ptc_pow1"])
    close((p1 / p0) ** 0.5, g["ptc_trans"])
    close(p01 / (p0 * p1) ** 0.5, g["ptc_coh"])


def test_lagrangian_bias(golden):
    """oracle/model_oracle.py:lagrangian_bias against montecosmo/bricks.py:327-452."""
    from oracle import model_oracle as MO
    g = golden("lagrangian_bias")
    bias = {k[5:]: float(v) for k, v in g.items() if k.startswith("bias_")}
    w, dvel = MO.lagrangian_bias(O.Cosmology(), torch.as_tensor(g["pos"]), float(g["a"]), tuple(g["box_size"]),
                                 torch.as_tensor(g["delta_k"]), bias, read_order=2)
    close(w, g["weights"])
    close(dvel, g["dvel"], atol=1e-11)
    # primordial non-Gaussianity terms (bricks.py:411-438) and add_png (129-141), tabulated power
    png = {k[4:]: float(v) for k, v in g.items() if k.startswith("png_fNL")}
    kpow = (g["kpow_k"], g["kpow_p"])
    w, dvel, phi = MO.lagrangian_bias(O.Cosmology(), torch.as_tensor(g["pos"]), float(g["a"]), tuple(g["box_size"]),
                                      torch.as_tensor(g["delta_k"]), bias, 2, png, "fNL", kpow)
    close(w, g["png_weights"])
    close(phi, g["png_phi"], atol=1e-14)
    close(MO.add_png(O.Cosmology(), 50.0, torch.as_tensor(g["delta_k"]), tuple(g["box_size"]), kpow), g["add_png_fNL50"],
          atol=1e-10)


def test_rg2cgh_cgh2rg(golden):
    """oracle rg2cgh / cgh2rg (utils.py:785-921) against the reference source, every use: the Gaussian permutation, its
    inverse, and norm="amp" (amplitude meshes: no sign, no sqrt2, no scaling) on an amplitude that is not even in k."""
    g = golden("rg2cgh")
    k = O.rg2cgh(g["white"])
    close(k, g["rg2cgh"])
    close(O.cgh2rg(k), g["cgh2rg_roundtrip"])
    close(O.cgh2rg(O._t(g["ampk"]), "amp"), g["cgh2rg_amp"], rtol=0, atol=0)
    close(O.rg2cgh(g["white"], "amp"), g["rg2cgh_amp"], rtol=0, atol=0)


def test_observation_chain(golden):
    """oracle/model_oracle.py's float64 restatement of bricks.py:628-813 (frames, lines of sight, light-cone scale
    factors, RSD, automatic Alcock-Paczynski) against the reference source."""
    from scipy.spatial.transform import Rotation
    from oracle import model_oracle as MO
    g, gr = golden("observation"), golden("growth")
    shape, box, center = tuple(int(s) for s in g["shape"]), tuple(g["box_size"]), tuple(g["box_center"])
    rot = Rotation.from_matrix(g["rot_matrix"])
    oc, ob, h, ns, s8 = gr["other_params"]
    cosmo, fid = O.Cosmology(), O.Cosmology(Omega_c=oc, Omega_b=ob, h=h, n_s=ns, sigma8=s8)
    pos, vel = O._t(g["pos"]), O._t(g["vel"])
    phys = MO.cell2phys_pos(pos, center, rot, box, shape)
    close(phys, g["cell2phys_pos"], rtol=1e-10)
    close(MO.phys2cell_pos(phys, center, rot, box, shape), g["phys2cell_roundtrip"], rtol=1e-10)
    for tag, curved in (("curved", True), ("flat", False)):
        los, a = MO.los_scalefactor_pos(pos, center, rot, box, shape, cosmo, None, curved)
        close(los * torch.ones(1, 3, dtype=torch.float64), g[f"los_{tag}"], rtol=1e-10)
        close(a, g[f"a_{tag}"], rtol=1e-10)
        close(MO.rsd(cosmo, vel, los, a, rot, box, shape, dvel=0.01), g[f"rsd_{tag}"], rtol=1e-9)
        close(MO.ap_auto(phys, los, cosmo, fid, curved), g[f"ap_auto_{tag}"], rtol=1e-10)


OBSERVED = (("curved_lightcone_auto", True, None, True), ("flat_lightcone_auto", False, None, True),
            ("curved_scalar_param", True, 0.7, False), ("flat_scalar_param", False, 0.7, False),
            ("flat_scalar_plain", False, 0.7, None))


def test_observed_nufft(golden):
    """The whole observation chain in the order model.py:780-805 composes it -- los_scalefactor_pos, cell2phys_pos, rsd
    with a velocity bias, ap_auto | ap_param, phys2cell_pos -- followed by nufft, float64 restatement against the
    reference source: observed positions 1e-10, half spectra 1e-9."""
    from scipy.spatial.transform import Rotation
    from oracle import model_oracle as MO
    g, gr = golden("observed_nufft"), golden("growth")
    shape, paint = tuple(int(s) for s in g["shape"]), tuple(int(s) for s in g["paint_shape"])
    box, center, rot = tuple(g["box_size"]), tuple(g["box_center"]), Rotation.from_matrix(g["rot_matrix"])
    oc, ob, h, ns, s8 = gr["other_params"]
    cosmo, fid = O.Cosmology(), O.Cosmology(Omega_c=oc, Omega_b=ob, h=h, n_s=ns, sigma8=s8)
    pos, vel, dvel, w = (O._t(g[k]) for k in ("pos", "vel", "dvel", "weights"))
    for tag, curved, a_obs, ap_auto in OBSERVED:
        los, a = MO.los_scalefactor_pos(pos, center, rot, box, shape, cosmo, a_obs, curved)
        p = MO.cell2phys_pos(pos, center, rot, box, shape) + MO.rsd(cosmo, vel, los, a, rot, box, shape, dvel)
        if ap_auto is not None:
            p = MO.ap_auto(p, los, cosmo, fid, curved) if ap_auto else \
                MO.ap_param(p, los, float(g["alpha_iso"]), float(g["alpha_ap"]), curved)
        p = MO.phys2cell_pos(p, center, rot, box, shape)
        close(p, g[f"pos_{tag}"], rtol=1e-10)
        close(O.nufft(p, shape, paint, w, 2, 2, "rectangular", True), g[f"nufft_{tag}"], rtol=1e-9)
