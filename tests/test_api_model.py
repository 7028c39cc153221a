"""
The reference-named Python API (montecosmo_b200/nbody.py) and the field-level model (model.py) against the oracle.

CPU runs drive the autograd layer over the host-emulation library with torch-CPU tensors; the gpu-marked runs use the
product configuration (libmcpm.so + CUDA tensors).  Tolerances follow SURVEY.md 8c: gradient of the log-density
relative L2 <= 1e-3 and cosine >= 0.9999 (float32 engine against the float64 oracle).
"""
import numpy as np
import pytest
import torch

from oracle import model_oracle as MO
from oracle import pm_oracle as O


@pytest.fixture(scope="module", params=["hostemu", pytest.param("cuda", marks=pytest.mark.gpu)])
def nb(request):
    import montecosmo_b200.nbody as nbody
    from montecosmo_b200.ops import Ops
    old = nbody._OPS
    if request.param == "hostemu":
        from tests import hostemu
        from tests.backends import torch_cpu_adapter
        nbody._OPS = Ops(hostemu.load(), torch_cpu_adapter())
    else:
        nbody._OPS = None
        nbody.ops()
    yield nbody
    nbody._OPS = old


def rel(a, b):
    a = a.detach().cpu().numpy().astype(np.complex128 if torch.is_complex(a) else np.float64)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def dev(nb):
    return nb.ops().A.device


def leaf(x, nb=None, dtype=None):
    """Fresh leaf tensor (on the engine's device when nb is given) that requires grad."""
    x = x.detach().clone()
    if dtype is not None:
        x = x.to(dtype)
    if nb is not None:
        x = x.to(dev(nb))
    return x.requires_grad_()


def test_paint_read_autograd(nb):
    rng = np.random.default_rng(0)
    shape = (8, 6, 10)
    pos = torch.tensor(rng.uniform(-2, 12, (300, 3)), dtype=torch.float32)
    w = torch.tensor(rng.uniform(0.5, 1.5, 300), dtype=torch.float32)
    mbar = torch.tensor(rng.normal(size=shape), dtype=torch.float32)
    mesh = torch.tensor(rng.normal(size=shape), dtype=torch.float32)
    for order in (2, 3):
        p, ww, m = leaf(pos, nb), leaf(w, nb), leaf(mesh, nb)
        loss = (nb.paint(p, shape, ww, order) * mbar.to(dev(nb))).sum() + (nb.read(p, m, order) * ww).sum()
        loss.backward()
        po, wo, mo = leaf(pos, dtype=torch.float64), leaf(w, dtype=torch.float64), leaf(mesh, dtype=torch.float64)
        lo = (O.paint(po, shape, wo, order) * mbar.double()).sum() + (O.read(po, mo, order) * wo).sum()
        lo.backward()
        assert abs(float(loss.detach()) - float(lo.detach())) < 1e-4 * abs(float(lo.detach()))
        assert rel(p.grad, po.grad) < 1e-5 and rel(ww.grad, wo.grad) < 1e-5 and rel(m.grad, mo.grad) < 1e-5
    with pytest.raises(ValueError):
        nb.paint(pos, shape, 1.0, 2, kernel_type="nope")


def test_kaiser_bessel_window_autograd(nb, golden):
    """kernel_type='kaiser_bessel' through the reference-named callables (paint, read, deconv_paint, interlace, nufft;
    nbody.py:280-312, 321-322, 383-384, 415-416) with autograd, vs the oracle's autograd and the golden vectors."""
    rng = np.random.default_rng(41)
    shape = (8, 6, 10)
    pos = torch.tensor(rng.uniform(-2, 12, (300, 3)), dtype=torch.float32)
    w = torch.tensor(rng.uniform(0.5, 1.5, 300), dtype=torch.float32)
    mbar = torch.tensor(rng.normal(size=shape), dtype=torch.float32)
    mesh = torch.tensor(rng.normal(size=shape), dtype=torch.float32)
    for order, ov in ((2, 2.0), (4, 1.5)):
        p, ww, m = leaf(pos, nb), leaf(w, nb), leaf(mesh, nb)
        loss = (nb.paint(p, shape, ww, order, "kaiser_bessel", ov) * mbar.to(dev(nb))).sum() \
            + (nb.read(p, m, order, "kaiser_bessel", ov) * ww).sum()
        loss.backward()
        po, wo, mo = leaf(pos, dtype=torch.float64), leaf(w, dtype=torch.float64), leaf(mesh, dtype=torch.float64)
        lo = (O.paint(po, shape, wo, order, "kaiser_bessel", ov) * mbar.double()).sum() \
            + (O.read(po, mo, order, "kaiser_bessel", ov) * wo).sum()
        lo.backward()
        assert abs(float(loss.detach()) - float(lo.detach())) < 1e-4 * abs(float(lo.detach()))
        assert rel(p.grad, po.grad) < 2e-5 and rel(ww.grad, wo.grad) < 2e-5 and rel(m.grad, mo.grad) < 2e-5
    g = golden("nufft")
    final = tuple(int(s) for s in g["final_shape"])
    gp = torch.tensor(g["pos"], dtype=torch.float32, device=dev(nb))
    gw = torch.tensor(g["weights"], dtype=torch.float32, device=dev(nb))
    assert rel(nb.nufft(gp, final, 1.5, gw, 4, 2, "kaiser_bessel"), g["nufft_kb_over15"]) < 2e-5
    assert rel(nb.nufft(gp, final, (12, 10, 14), gw, 2, 2, "kaiser_bessel"), g["nufft_kb_tuple_o2"]) < 2e-5
    assert rel(nb.interlace(gp, final, gw, 4, 2, "kaiser_bessel", 1.5), g["interlace_kb_4_2"]) < 2e-5
    rm = torch.tensor(g["deconv_real_in"], dtype=torch.float32, device=dev(nb))
    assert rel(nb.deconv_paint(rm, 4, "kaiser_bessel", 1.5), g["deconv_kb_real_4"]) < 5e-6
    # gradient through nufft with the Kaiser-Bessel window
    cs = O.r2chshape(final)
    cot = torch.tensor(rng.normal(size=cs) + 1j * rng.normal(size=cs), dtype=torch.complex64)
    p, ww = leaf(gp.cpu(), nb), leaf(gw.cpu(), nb)
    out = nb.nufft(p, final, 1.5, ww, 4, 2, "kaiser_bessel")
    torch.view_as_real(out * cot.to(dev(nb)).conj()).select(-1, 0).sum().backward()
    po, wo = leaf(gp.cpu(), dtype=torch.float64), leaf(gw.cpu(), dtype=torch.float64)
    oo = O.nufft(po, final, 1.5, wo, 4, 2, "kaiser_bessel")
    (oo * cot.to(torch.complex128).conj()).real.sum().backward()
    assert rel(out, oo) < 2e-5 and rel(p.grad, po.grad) < 1e-4 and rel(ww.grad, wo.grad) < 1e-4


def test_fft_chreshape_deconv_autograd(nb):
    rng = np.random.default_rng(1)
    shape, big = (8, 6, 10), (12, 10, 12)
    x = torch.tensor(rng.normal(size=shape), dtype=torch.float32)
    cot = torch.tensor(rng.normal(size=big), dtype=torch.float32)
    xe = leaf(x, nb)
    y = nb.irfftn(nb.chreshape(nb.deconv_paint(nb.rfftn(xe), 2), O.r2chshape(big)))
    (y * cot.to(dev(nb))).sum().backward()
    xo = leaf(x, dtype=torch.float64)
    yo = torch.fft.irfftn(O.chreshape(O.deconv_paint(torch.fft.rfftn(xo), 2), O.r2chshape(big)), s=big)
    (yo * cot.double()).sum().backward()
    assert rel(y, yo) < 1e-5 and rel(xe.grad, xo.grad) < 1e-5
    # down-sampling direction, and real-mesh deconvolution
    xe = leaf(cot, nb)
    y = nb.irfftn(nb.chreshape(nb.rfftn(nb.deconv_paint(xe, 3)), O.r2chshape(shape)))
    (y * x.to(dev(nb))).sum().backward()
    xo = leaf(cot, dtype=torch.float64)
    yo = torch.fft.irfftn(O.chreshape(torch.fft.rfftn(O.deconv_paint(xo, 3)), O.r2chshape(shape)), s=shape)
    (yo * x.double()).sum().backward()
    assert rel(y, yo) < 1e-5 and rel(xe.grad, xo.grad) < 1e-5


def test_nufft_autograd_oversampled(nb):
    rng = np.random.default_rng(2)
    final = (8, 8, 8)
    pos = torch.tensor(rng.uniform(-1, 9, (400, 3)), dtype=torch.float32)
    w = torch.tensor(rng.uniform(0.5, 1.5, 400), dtype=torch.float32)
    cs = O.r2chshape(final)
    cot = torch.tensor(rng.normal(size=cs) + 1j * rng.normal(size=cs), dtype=torch.complex64)
    p, ww = leaf(pos, nb), leaf(w, nb)
    out = nb.nufft(p, final, 1.5, ww, 2, 2)
    torch.view_as_real(out * cot.to(dev(nb)).conj()).select(-1, 0).sum().backward()
    po, wo = leaf(pos, dtype=torch.float64), leaf(w, dtype=torch.float64)
    oo = O.nufft(po, final, 1.5, wo, 2, 2)
    (oo * cot.to(torch.complex128).conj()).real.sum().backward()
    assert rel(out, oo) < 2e-5 and rel(p.grad, po.grad) < 1e-4 and rel(ww.grad, wo.grad) < 1e-4


def test_nbody_bf_matches_golden_and_oracle_grad(nb, golden):
    g = golden("nbody")
    shape = tuple(int(s) for s in g["shape"])
    from montecosmo_b200.cosmo import Cosmology
    dk = torch.tensor(g["delta_k"], dtype=torch.complex64, device=dev(nb)).requires_grad_()
    q = O.regular_pos(shape).float().to(dev(nb))
    pos, vel = nb.nbody_bf(Cosmology(), dk, q, 0.0, 1.0, 4)
    assert pos.shape == (1, q.shape[0], 3)
    assert np.abs(pos[0].detach().cpu().numpy() - g["bf4_pos"][0]).max() < 2e-4
    assert rel(vel[0], g["bf4_vel"][0]) < 2e-4
    # snapshots + other orders (nbody.py:990-995)
    p4, v4 = nb.nbody_bf(Cosmology(), dk.detach(), q, 0.1, 0.8, 3, paint_order=3, lpt_order=1, paint_deconv=True,
                         snapshots=4)
    assert np.abs(p4.cpu().numpy() - g["bf3_snap_pos"]).max() < 2e-4
    assert rel(v4, g["bf3_snap_vel"]) < 2e-4
    # gradient w.r.t. delta_k vs oracle autograd
    rng = np.random.default_rng(3)
    cp = torch.tensor(rng.normal(size=(q.shape[0], 3)), dtype=torch.float32)
    cv = torch.tensor(rng.normal(size=(q.shape[0], 3)), dtype=torch.float32)
    ((pos[0] * cp.to(dev(nb))).sum() + (vel[0] * cv.to(dev(nb))).sum()).backward()
    dko = torch.tensor(g["delta_k"], dtype=torch.complex128).requires_grad_()
    po, vo = O.nbody_bf(O.Cosmology(), dko, O.regular_pos(shape), 0.0, 1.0, 4)
    ((po[0] * cp.double()).sum() + (vo[0] * cv.double()).sum()).backward()
    # a0 = 0 -> a = 1 in 4 steps is strongly non-linear at this amplitude: CIC derivatives are discontinuous at cell
    # faces, so float32 rounding of ABSOLUTE positions moves a few particles across them (5e-3; measured 1.9e-3) ...
    assert rel(dk.grad, dko.grad) < 5e-3
    # ... which is why the loop carries displacements from the lattice sites once the caller declares the lattice
    # (ptcl_shape; mcpm_engine_set_relative): same API, same returned positions.  Measured 5.8e-7 on the CPU port (no
    # particle on the other side of a face) and 1.3e-3 on a B200 (one of the 4096: sqrt(1/4096) x the ~10 % jump of its
    # CIC derivative); 3e-3 admits two such particles, where absolute positions needed 5e-3 above.
    dk2 = leaf(dk)
    pos2, vel2 = nb.nbody_bf(Cosmology(), dk2, q, 0.0, 1.0, 4, ptcl_shape=shape)
    assert np.abs(pos2[0].detach().cpu().numpy() - g["bf4_pos"][0]).max() < 2e-4
    ((pos2[0] * cp.to(dev(nb))).sum() + (vel2[0] * cv.to(dev(nb))).sum()).backward()
    assert rel(dk2.grad, dko.grad) < 3e-3
    # relative=True returns the displacements themselves
    d3, v3 = nb.nbody_bf(Cosmology(), dk.detach(), q, 0.0, 1.0, 4, ptcl_shape=shape, relative=True)
    assert np.abs((d3[0] + q).cpu().numpy() - g["bf4_pos"][0]).max() < 2e-4 and rel(v3[0], g["bf4_vel"][0]) < 2e-4


def test_cosmology_gradient_through_engine(nb):
    """d/dOmega_c of a loss through lpt + 2 BullFrog steps: host growth tables chained to the engine's coefbar."""
    from montecosmo_b200.cosmo import Cosmology
    rng = np.random.default_rng(4)
    shape = (8, 8, 8)
    dk0 = (np.fft.rfftn(rng.normal(size=shape)) * 0.03)
    q = O.regular_pos(shape)
    cp = torch.tensor(rng.normal(size=(q.shape[0], 3)))

    oc = torch.tensor(0.26447041, dtype=torch.float64, requires_grad=True)
    pos, vel = nb.nbody_bf(Cosmology(Omega_c=oc), torch.tensor(dk0, dtype=torch.complex64, device=dev(nb)),
                           q.float().to(dev(nb)), 0.05, 0.9, 2)
    (pos[0] * cp.float().to(dev(nb))).sum().backward()

    oco = torch.tensor(0.26447041, dtype=torch.float64, requires_grad=True)
    po, _ = O.nbody_bf(O.Cosmology(Omega_c=oco), torch.tensor(dk0, dtype=torch.complex128), q, 0.05, 0.9, 2)
    (po[0] * cp).sum().backward()
    assert abs(float(oc.grad) - float(oco.grad)) < 2e-3 * abs(float(oco.grad))


@pytest.mark.parametrize("evolution,n_steps", [("lpt", 0), ("nbody", 3)])
def test_model_logpdf_and_force(nb, evolution, n_steps):
    from montecosmo_b200.model import FieldModel
    rng = np.random.default_rng(5)
    shape = (16, 16, 16)
    precond = "fourier" if evolution == "lpt" else "real"  # both initial-condition parametrisations get a run
    m = FieldModel(shape, (160.0,) * 3, evolution=evolution, n_steps=n_steps, a_obs=0.8, b1=0.7, sigma_obs=0.5,
                   precond=precond)
    white = rng.normal(size=shape).astype(np.float32)
    kw = dict(evolution=evolution, n_steps=n_steps, a_obs=0.8, b1=0.7, precond=precond)
    transfer = m.transfer.cpu().numpy().astype(np.float64)
    with torch.no_grad():
        truth = MO.evolve(torch.tensor(rng.normal(size=shape)), transfer, O.Cosmology(), shape, **kw)
    obs = (truth.numpy() + 0.5 * rng.normal(size=shape)).astype(np.float32)
    lp, g = m.value_and_force(white, obs)
    lpo, go = MO.value_and_force(white.astype(np.float64), obs.astype(np.float64), transfer, O.Cosmology(), shape,
                                 sigma_obs=0.5, **kw)
    assert abs(float(lp) - float(lpo)) < 1e-4 * abs(float(lpo))
    assert rel(g, go) < 1e-3
    gn, gon = g.detach().cpu().numpy().ravel().astype(np.float64), go.numpy().ravel()
    assert gn @ gon / np.linalg.norm(gn) / np.linalg.norm(gon) > 0.9999
    # the autograd-friendly wrapper agrees with the hand-seeded path
    w = leaf(torch.tensor(white), nb)
    lp2 = m.logpdf(w, obs)
    lp2.backward()
    assert abs(float(lp2.detach()) - float(lp)) < 1e-6 * abs(float(lp)) and rel(w.grad, g.detach().cpu().numpy()) < 1e-5
    assert rel(m.force(white, obs), g.detach().cpu().numpy()) < 1e-5


@pytest.mark.parametrize("lpt_order", [1, 2])
def test_lpt_per_particle_scale_factor(nb, lpt_order):
    """Light-cone lpt (nbody.py:651-653: `a` of shape [Np, 1]) against the oracle, value (5e-5) and gradient w.r.t. the
    initial mesh (2e-4); the scalar path must agree with the per-particle path fed a constant array."""
    from montecosmo_b200.cosmo import Cosmology
    rng = np.random.default_rng(17)
    shape = (8, 6, 10)
    dk0 = np.fft.rfftn(rng.normal(size=shape)) * 0.02
    q = O.regular_pos(shape)
    pos = (q + torch.tensor(rng.normal(scale=0.3, size=q.shape))).float()
    a = torch.tensor(rng.uniform(0.2, 1.0, (q.shape[0], 1)))
    cd, cv = (torch.tensor(rng.normal(size=q.shape), dtype=torch.float32) for _ in range(2))
    dk = torch.tensor(dk0, dtype=torch.complex64, device=dev(nb)).requires_grad_()
    dp, vl = nb.lpt(Cosmology(), dk, pos.to(dev(nb)), a, lpt_order, 2)
    ((dp * cd.to(dev(nb))).sum() + (vl * cv.to(dev(nb))).sum()).backward()
    dko = torch.tensor(dk0, dtype=torch.complex128).requires_grad_()
    dpo, vlo = O.lpt(O.Cosmology(), dko, pos.double(), a, lpt_order, 2)
    ((dpo * cd.double()).sum() + (vlo * cv.double()).sum()).backward()
    assert rel(dp, dpo) < 5e-5 and rel(vl, vlo) < 5e-5
    assert rel(dk.grad, dko.grad) < 2e-4
    d0, v0 = nb.lpt(Cosmology(), dk.detach(), pos.to(dev(nb)), 0.5, lpt_order, 2)
    d1, v1 = nb.lpt(Cosmology(), dk.detach(), pos.to(dev(nb)), np.full(q.shape[0], 0.5), lpt_order, 2)
    assert rel(d1, d0) < 1e-6 and rel(v1, v0) < 1e-6
    with pytest.raises(ValueError):
        nb.lpt(Cosmology(), dk.detach(), pos.to(dev(nb)), np.full(7, 0.5), lpt_order, 2)


def test_rg2cgh_golden_roundtrip_and_gradient(nb, golden):
    """utils.rg2cgh / cgh2rg (utils.py:785-921) vs the golden vectors of the reference source (1e-6: a permutation with
    float32 scaling), every norm vs the oracle, the inverse, and the VJP (with a fused transfer) vs oracle autograd."""
    from montecosmo_b200 import utils as U
    g = golden("rg2cgh")
    white = torch.tensor(g["white"], dtype=torch.float32, device=dev(nb))
    out = U.rg2cgh(white)
    assert rel(out, g["rg2cgh"]) < 1e-6
    assert rel(U.cgh2rg(out), g["cgh2rg_roundtrip"]) < 1e-6
    rng = np.random.default_rng(4)
    for shape in [(8, 6, 10), (4, 4, 4), (6, 8, 4)]:
        w = torch.tensor(rng.normal(size=shape), dtype=torch.float32)
        for norm in ("backward", "ortho", "forward"):
            ko = O.rg2cgh(w.double(), norm)
            assert rel(U.rg2cgh(w.to(dev(nb)), norm), ko) < 1e-6
            assert rel(U.cgh2rg(ko.to(torch.complex64).to(dev(nb)), norm), w.double()) < 1e-6
        # the result is a Hermitian half spectrum: irfftn then rfftn gives it back
        k = U.rg2cgh(w.to(dev(nb)))
        assert rel(nb.rfftn(nb.irfftn(k)), k.detach().cpu().numpy()) < 1e-5
        # VJP with the transfer multiply fused in
        cs = O.r2chshape(shape)
        tr = torch.tensor(rng.uniform(0.5, 2.0, cs), dtype=torch.float32)
        cot = torch.tensor(rng.normal(size=cs) + 1j * rng.normal(size=cs), dtype=torch.complex64)
        we = leaf(w, nb)
        oe = U.rg2cgh(we, "backward", tr.to(dev(nb)))
        torch.view_as_real(oe * cot.to(dev(nb)).conj()).select(-1, 0).sum().backward()
        wo = leaf(w, dtype=torch.float64)
        oo = O.rg2cgh(wo) * tr.double()
        (oo * cot.to(torch.complex128).conj()).real.sum().backward()
        assert rel(oe, oo) < 1e-6 and rel(we.grad, wo.grad) < 1e-6
    with pytest.raises(AssertionError):
        U.rg2cgh(torch.zeros(4, 5, 4))
    # norm = "amp" (utils.py:807-817, 858-868, 915-918): golden vectors of the reference source, exact (a permutation)
    ampk = torch.tensor(g["ampk"], dtype=torch.float32, device=dev(nb))
    assert np.array_equal(U.cgh2rg(ampk, "amp").cpu().numpy(), g["cgh2rg_amp"].astype(np.float32))
    assert np.array_equal(U.rg2cgh(white, "amp").cpu().numpy(), g["rg2cgh_amp"].astype(np.float32))
    with pytest.raises(AssertionError):
        U.rg2cgh(torch.zeros(4, 4, 4), norm="nope")


def test_lagrangian_bias_weights_and_gradient(nb, golden):
    """bricks.lagrangian_bias (bricks.py:327-452) against the golden vectors of the reference source and the
    oracle: weights and dvel 5e-5, gradient of a scalar functional w.r.t. the linear mesh 2e-4."""
    from montecosmo_b200 import bricks as B
    from montecosmo_b200.cosmo import Cosmology
    gd = golden("lagrangian_bias")
    gb = {k[5:]: float(v) for k, v in gd.items() if k.startswith("bias_")}
    wg, dvg, _ = B.lagrangian_bias(Cosmology(), torch.tensor(gd["pos"], dtype=torch.float32, device=dev(nb)),
                                   float(gd["a"]), tuple(gd["box_size"]),
                                   torch.tensor(gd["delta_k"], dtype=torch.complex64, device=dev(nb)), gb, read_order=2)
    assert rel(wg, gd["weights"]) < 5e-5 and rel(dvg, gd["dvel"]) < 5e-5
    rng = np.random.default_rng(23)
    shape, box = (8, 10, 12), (80.0, 100.0, 96.0)
    dk0 = np.fft.rfftn(rng.normal(size=shape)) * 0.05
    q = O.regular_pos(shape)
    pos = (q + torch.tensor(rng.normal(scale=0.4, size=q.shape))).float()
    bias = dict(b1=0.8, b2=0.3, bs2=-0.2, b3=0.1, bds2=0.05, bs3=-0.07, bn2=0.4, bnpar=0.6)
    cw = torch.tensor(rng.normal(size=q.shape[0]), dtype=torch.float32)
    cv = torch.tensor(rng.normal(size=q.shape), dtype=torch.float32)
    dk = torch.tensor(dk0, dtype=torch.complex64, device=dev(nb)).requires_grad_()
    w, dvel, phi = B.lagrangian_bias(Cosmology(), pos.to(dev(nb)), 0.7, box, dk, bias, read_order=2)
    ((w * cw.to(dev(nb))).sum() + (dvel * cv.to(dev(nb))).sum()).backward()
    dko = torch.tensor(dk0, dtype=torch.complex128).requires_grad_()
    wo, dvo = MO.lagrangian_bias(O.Cosmology(), pos.double(), 0.7, box, dko, bias, read_order=2)
    ((wo * cw.double()).sum() + (dvo * cv.double()).sum()).backward()
    assert phi == 0.0 and rel(w, wo) < 5e-5 and rel(dvel, dvo) < 5e-5
    assert rel(dk.grad, dko.grad) < 2e-4
    assert rel(B.regular_pos(shape), q.numpy()) == 0.0
    with pytest.raises(NotImplementedError):  # kpow=None would need jax_cosmo's Eisenstein-Hu power
        B.lagrangian_bias(Cosmology(), pos.to(dev(nb)), 0.7, box, dk.detach(), bias, png_type="fNL")
    # primordial non-Gaussianity terms (bricks.py:411-438) and add_png (129-141) with a tabulated power: golden vectors
    # of the reference source, and the gradient w.r.t. the linear mesh against the oracle's autograd
    png = {k[4:]: float(v) for k, v in gd.items() if k.startswith("png_fNL")}
    kpow = (gd["kpow_k"], gd["kpow_p"])
    gpos = torch.tensor(gd["pos"], dtype=torch.float32, device=dev(nb))
    gdk = torch.tensor(gd["delta_k"], dtype=torch.complex64, device=dev(nb)).requires_grad_()
    wp, _, phi = B.lagrangian_bias(Cosmology(), gpos, float(gd["a"]), tuple(gd["box_size"]), gdk, gb, png, "fNL", kpow, 2)
    assert rel(wp, gd["png_weights"]) < 5e-5 and rel(phi, gd["png_phi"]) < 5e-5
    assert rel(B.add_png(Cosmology(), 50.0, gdk.detach(), tuple(gd["box_size"]), kpow), gd["add_png_fNL50"]) < 5e-5
    cw = torch.tensor(rng.normal(size=gpos.shape[0]), dtype=torch.float32)
    (wp * cw.to(dev(nb))).sum().backward()
    dko = torch.tensor(gd["delta_k"]).requires_grad_()
    wo, _, _ = MO.lagrangian_bias(O.Cosmology(), torch.tensor(gd["pos"]), float(gd["a"]), tuple(gd["box_size"]), dko, gb, 2,
                                  png, "fNL", kpow)
    (wo * cw.double()).sum().backward()
    assert rel(gdk.grad, dko.grad) < 2e-4


def test_lagrangian_bias_fused_passes_match_composition(nb, golden):
    """The fused passes of csrc/bias.cu behind bricks.lagrangian_bias against round 1's pointwise-torch composition of the
    same expansion (bricks.lagrangian_bias_composed, itself pinned to the reference source's golden vectors): weights,
    dvel and phi, and the cotangents of EVERY differentiable input -- linear mesh, positions, the 8 bias and 5 PNG
    coefficients, and the growth factor (scalar scale factor and per-particle light-cone scale factors) -- 2e-4.
    bricks.py:327-452."""
    from montecosmo_b200 import bricks as B
    from montecosmo_b200.cosmo import Cosmology
    gd = golden("lagrangian_bias")
    kpow = (gd["kpow_k"], gd["kpow_p"])
    rng = np.random.default_rng(77)
    shape, box = (8, 10, 12), (80.0, 100.0, 96.0)
    dk0 = np.fft.rfftn(rng.normal(size=shape)) * 0.05
    q = O.regular_pos(shape)
    pos0 = (q + torch.tensor(rng.normal(scale=0.4, size=q.shape))).float()
    bias0 = dict(b1=0.8, b2=0.3, bs2=-0.2, b3=0.1, bds2=0.05, bs3=-0.07, bn2=0.4, bnpar=0.6)
    png0 = dict(fNL_bp=0.5, fNL_bpd=-0.3, fNL_bpd2=0.2, fNL_bps2=0.1, fNL_bn2p=-0.4)
    cw = torch.tensor(rng.normal(size=q.shape[0]), dtype=torch.float32, device=dev(nb))
    cv = torch.tensor(rng.normal(size=q.shape), dtype=torch.float32, device=dev(nb))
    cp = torch.tensor(rng.normal(size=shape), dtype=torch.float32, device=dev(nb))
    a_lightcone = torch.tensor(rng.uniform(0.4, 0.9, q.shape[0]))
    for png_type, a in ((None, 0.7), ("fNL", 0.7), ("fNL", a_lightcone)):
        out = {}
        for fn in (B.lagrangian_bias, B.lagrangian_bias_composed):
            dk = torch.tensor(dk0, dtype=torch.complex64, device=dev(nb)).requires_grad_()
            pos = pos0.to(dev(nb)).requires_grad_()
            bias = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in bias0.items()}
            png = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in png0.items()}
            om = torch.tensor(0.2611, dtype=torch.float64, requires_grad=True)  # Omega_c: the growth factor depends on it
            w, dvel, phi = fn(Cosmology(Omega_c=om), pos, a, box, dk, bias, png, png_type, kpow, 2)
            loss = (w * cw).sum() + (dvel * cv).sum() + ((phi * cp).sum() if png_type else 0.0)
            loss.backward()
            coefs = [bias[k].grad for k in bias0] + ([png[k].grad for k in png0] if png_type else [])
            out[fn.__name__] = (w.detach(), dvel.detach(), dk.grad, pos.grad, torch.stack(coefs), om.grad)
        f, c = out["lagrangian_bias"], out["lagrangian_bias_composed"]
        tag = (png_type, "lightcone" if isinstance(a, torch.Tensor) else a)
        assert rel(f[0], c[0]) < 2e-5 and rel(f[1], c[1]) < 2e-5, tag
        assert rel(f[2], c[2]) < 2e-4, tag          # linear mesh
        assert rel(f[3], c[3]) < 2e-4, tag          # positions
        assert rel(f[4], c[4]) < 2e-4, tag          # coefficients
        assert abs(float(f[5]) - float(c[5])) < 2e-4 * abs(float(c[5])), tag  # cosmology through the growth factor


def test_nufft_with_fused_redshift_space_shift(nb):
    """nufft(..., rsd=(vel, los, coef)) -- the flat-sky shift of bricks.py:781-792 applied inside the paint kernels
    (mcpm_nufft_rsd) -- against the two-pass composition mcpm_rsd_shift -> mcpm_nufft it replaces on the evolve path
    (model.py:780-809), and against the float64 oracle: half spectrum 1e-6 (same arithmetic, same association), and the
    cotangents of pos, vel and weights 2e-5.  Absolute positions on a generic mesh (global-atomic kernels, paint mesh
    1.5x the final one) and lattice-relative displacements on a mesh the brick-tiled kernel takes."""
    rng = np.random.default_rng(91)
    los, coef = (0.36, -0.48, 0.8), 0.37
    for shape, paint, lattice in (((8, 12, 12), (12, 18, 18), None), ((32, 24, 64), None, (32, 24, 64))):
        n = int(np.prod(shape))
        q = O.regular_pos(shape)
        disp = torch.tensor(rng.normal(scale=0.6, size=q.shape)).float()
        pos0 = disp if lattice else (q.float() + disp)
        vel0 = torch.tensor(rng.normal(scale=1.5, size=q.shape)).float()
        w0 = torch.tensor(rng.uniform(0.3, 2.0, n)).float()
        cshape = tuple(nb.r2chshape(shape)) if hasattr(nb, "r2chshape") else (shape[0], shape[1], shape[2] // 2 + 1)
        ck = torch.tensor(rng.normal(size=cshape) + 1j * rng.normal(size=cshape)).to(torch.complex64).to(dev(nb))
        res = []
        for fused in (True, False):
            pos, vel, w = (t.clone().to(dev(nb)).requires_grad_() for t in (pos0, vel0, w0))
            if fused:
                out = nb.nufft(pos, shape, paint, w, 2, 2, paint_deconv=True, lattice=lattice, rsd=(vel, los, coef))
            else:
                from montecosmo_b200.model import _RsdShift
                out = nb.nufft(_RsdShift.apply(pos, vel, los, coef), shape, paint, w, 2, 2, paint_deconv=True, lattice=lattice)
            (out * ck.conj()).real.sum().backward()
            res.append((out.detach(), pos.grad, vel.grad, w.grad))
        names = ("spectrum", "posbar", "velbar", "weightsbar")
        for name, a, b, tol in zip(names, res[0], res[1], (1e-6, 2e-5, 2e-5, 2e-5)):
            assert rel(a, b) < tol, (shape, name)
        # float64 oracle of the same chain
        po, vo, wo = (t.double().requires_grad_() for t in (pos0, vel0, w0))
        xs = (q + po if lattice else po) + (vo @ torch.tensor(los, dtype=torch.float64))[:, None] * coef * torch.tensor(los, dtype=torch.float64)
        oo = O.nufft(xs, shape, paint, wo, 2, 2, paint_deconv=True)
        (oo * ck.cpu().to(torch.complex128).conj()).real.sum().backward()
        assert rel(res[0][0], oo.detach()) < 2e-5 and rel(res[0][2], vo.grad) < 1e-4 and rel(res[0][1], po.grad) < 1e-4


def test_bullfrog_vf_scan_and_host_windows(nb, golden):
    """bullfrog_vf (nbody.py:902-960) against the oracle's drift-kick-drift step; nbody_bf_scan against the same steps in
    a loop; the host-side window helpers against the golden vectors of the reference source."""
    from montecosmo_b200.cosmo import Cosmology
    rng = np.random.default_rng(31)
    shape = (8, 6, 10)
    q = O.regular_pos(shape)
    pos = (q + torch.tensor(rng.normal(scale=0.5, size=q.shape))).float()
    vel = torch.tensor(rng.normal(scale=0.3, size=q.shape), dtype=torch.float32)
    g0, dg = 0.3, 0.15
    vf = nb.bullfrog_vf(Cosmology(), dg, shape, 2)
    dp, dv = vf(g0, (pos.to(dev(nb)), vel.to(dev(nb))), None)
    po, vo = O.bullfrog_step(O.Cosmology(), (pos.double(), vel.double()), torch.tensor(g0), torch.tensor(dg), shape, 2)
    assert rel(pos.to(dev(nb)) + dg * dp, po) < 1e-6 and rel(vel.to(dev(nb)) + dg * dv, vo) < 5e-5
    # nbody_bf_scan: vel = pm_forces(pos, delta_k), then n_steps equal steps from g = 0 to g(a)
    dk0 = np.fft.rfftn(rng.normal(size=shape)) * 0.02
    a, n_steps = 0.6, 3
    xs, vs = nb.nbody_bf_scan(Cosmology(), torch.tensor(dk0, dtype=torch.complex64, device=dev(nb)), pos.to(dev(nb)), a,
                              n_steps)
    co = O.Cosmology()
    state = (pos.double(), O.pm_forces(pos.double(), torch.tensor(dk0), 2))
    dgo = O.a2g(co, a) / n_steps
    for i in range(n_steps):
        state = O.bullfrog_step(co, state, i * dgo, dgo, shape, 2)
    assert xs.shape == (1, q.shape[0], 3) and rel(xs[0], state[0]) < 1e-6 and rel(vs[0], state[1]) < 5e-5
    # host-side kernels
    g = golden("kernels")
    kvec = nb.rfftk(tuple(int(s) for s in g["shape"]))
    assert np.allclose(nb.kaiser_bessel_hat(kvec, 4, nb.optim_kcut(1.5)), g["kaiser_bessel_hat_4"], rtol=1e-12)
    kf = nb.fftk((8, 6, 10), (80.0, 60.0, 50.0))  # nbody.py:78-103
    assert [k.shape for k in kf] == [(8, 1, 1), (1, 6, 1), (1, 1, 10)]
    assert np.allclose(kf[2].ravel(), np.fft.fftfreq(10) * 2 * np.pi * 10 / 50.0, rtol=1e-15)
    assert nb.top_hat(kvec, np.inf) == 1.0  # nbody.py:191-217
    th = nb.top_hat(kvec, 2.0)
    assert th.dtype == bool and th.shape == np.broadcast_shapes(*(k.shape for k in kvec))
    assert np.array_equal(th, sum(k**2 for k in kvec) < 4.0)
    assert float(nb.a2chi(Cosmology(), 1.0)) == 0.0 and float(nb.chi2a(Cosmology(), 0.0)) == 1.0
    s = np.linspace(0, 2, 41)
    for order in (1, 2, 3, 4):
        ref = O.rectangular(torch.tensor(s), order).numpy()
        assert np.allclose(nb.rectangular(s, order) * np.ones_like(s), ref, rtol=1e-14)
    assert np.allclose(nb.kaiser_bessel(s[:21], 2, nb.optim_kcut(2.0)),
                       O.kaiser_bessel(torch.tensor(s[:21]), 2, nb.optim_kcut(2.0)).numpy(), rtol=1e-12)
    assert float(nb.alpha_bf(Cosmology(), torch.tensor(0.3), torch.tensor(0.1))) == pytest.approx(
        float(O.alpha_bf(O.Cosmology(), torch.tensor(0.3), torch.tensor(0.1))), rel=1e-12)
    # alpha_fpm (nbody.py:921-931) against the golden vector, and the FastPM-weighted loop against the oracle's steps
    gn = golden("nbody")
    g0n, dgn = gn["bf4_g0_dg"]
    assert np.allclose([float(nb.alpha_fpm(Cosmology(), g0n + n * dgn, dgn)) for n in range(4)], gn["bf4_alpha_fpm"],
                       rtol=1e-10)
    shp = tuple(int(s) for s in gn["shape"])
    dkn = torch.tensor(gn["delta_k"], dtype=torch.complex64, device=dev(nb))
    qn = O.regular_pos(shp)
    pf, vf_ = nb.nbody_bf(Cosmology(), dkn, qn.float().to(dev(nb)), 0.1, 0.8, 3, integrator="fastpm")
    co = O.Cosmology()
    dpo, vlo = O.lpt(co, torch.tensor(gn["delta_k"]), qn, 0.1, 2, 1)
    state, g_lo, g_hi = (qn + dpo, vlo), O.a2g(co, 0.1), O.a2g(co, 0.8)
    dgo = (g_hi - g_lo) / 3
    for i in range(3):
        t = g_lo + i * dgo
        x = state[0] + state[1] * (dgo / 2)
        al = O.alpha_fpm(co, t, dgo)
        v = al * state[1] + (1 - al) * O.pm_forces(x, shp, 2) / (t + dgo / 2)
        state = (x + v * (dgo / 2), v)
    assert np.abs(pf[0].cpu().numpy() - state[0].numpy()).max() < 2e-4 and rel(vf_[0], state[1]) < 2e-4
    with pytest.raises(ValueError):
        nb.nbody_bf(Cosmology(), dkn, qn.float().to(dev(nb)), 0.1, 0.8, 3, integrator="leapfrog")


def test_nbody_bf_snapshots_dense_output(nb, golden):
    """Save times inside steps, a scale-factor list and a custom save function (nbody.py:987-996; diffrax's Euler dense
    output is linear inside a step) against the golden vectors of the reference source, and the gradient through an
    interpolated snapshot against the oracle's autograd."""
    from montecosmo_b200.cosmo import Cosmology
    g = golden("nbody")
    shape = tuple(int(s) for s in g["shape"])
    dk = torch.tensor(g["delta_k"], dtype=torch.complex64, device=dev(nb))
    q = O.regular_pos(shape).float().to(dev(nb))
    p, v = nb.nbody_bf(Cosmology(), dk, q, 0.1, 0.8, 3, snapshots=3)
    assert p.shape == (3, q.shape[0], 3)
    assert np.abs(p.cpu().numpy() - g["bf3_mid_pos"]).max() < 2e-4 and rel(v, g["bf3_mid_vel"]) < 2e-4
    p, v = nb.nbody_bf(Cosmology(), dk, q, 0.1, 0.8, 3, snapshots=list(g["bf3_alist"]))
    assert np.abs(p.cpu().numpy() - g["bf3_alist_pos"]).max() < 2e-4 and rel(v, g["bf3_alist_vel"]) < 2e-4
    d = nb.nbody_bf(Cosmology(), dk, q, 0.1, 0.8, 3, snapshots=5, fn=lambda t, y, args: y[0] - q)
    assert d.shape == (5, q.shape[0], 3) and np.abs(d.cpu().numpy() - g["bf3_fn_disp"]).max() < 2e-4
    p1, _ = nb.nbody_bf(Cosmology(), dk, q, 0.1, 0.8, 3, snapshots=1, fn=lambda t, y, args: (2 * y[0], y[1]))
    assert p1.shape == (1, q.shape[0], 3) and np.abs(p1[0].cpu().numpy() / 2 - g["bf3_mid_pos"][-1]).max() < 2e-4
    # gradient of a functional of the middle (interpolated) snapshot w.r.t. delta_k
    rng = np.random.default_rng(8)
    cot = torch.tensor(rng.normal(size=(q.shape[0], 3)), dtype=torch.float32)
    dkl = leaf(dk)
    p, v = nb.nbody_bf(Cosmology(), dkl, q, 0.1, 0.8, 3, snapshots=3)
    ((p[1] * cot.to(dev(nb))).sum() + (v[1] * cot.to(dev(nb))).sum()).backward()
    dko = torch.tensor(g["delta_k"]).requires_grad_()
    po, vo = O.nbody_bf(O.Cosmology(), dko, O.regular_pos(shape), 0.1, 0.8, 3, snapshots=3)
    ((po[1] * cot.double()).sum() + (vo[1] * cot.double()).sum()).backward()
    # same regime as test_nbody_bf_matches_golden_and_oracle_grad: CIC derivatives jump at cell faces, 5e-3 in float32
    assert rel(dkl.grad, dko.grad) < 5e-3


def test_force_tape_is_optional(nb):
    """tape_forces=False (mcpm_nbody_steps_vjp with fm = NULL): the reverse sweep recomputes every step's force meshes
    from the taped kick positions -- same log-density and gradient (float-atomic summation order aside), for CIC (float4
    force meshes) and a non-CIC order (planar ones)."""
    from montecosmo_b200.model import FieldModel
    rng = np.random.default_rng(12)
    shape = (16, 12, 20)
    white = rng.normal(size=shape).astype(np.float32)
    obs = (1.0 + rng.normal(size=shape)).astype(np.float32)
    for order in (2, 3):
        out = []
        for tf in (True, False):
            m = FieldModel(shape, (200.0, 150.0, 250.0), evolution="nbody", n_steps=3, a_start=0.1, paint_order=order,
                           tape_forces=tf)
            lp, g = m.value_and_force(white, obs)
            out.append((float(lp), g.detach().cpu().numpy().astype(np.float64)))
        assert abs(out[0][0] - out[1][0]) <= 1e-6 * abs(out[0][0])
        assert np.linalg.norm(out[0][1] - out[1][1]) <= 2e-5 * np.linalg.norm(out[0][1])


def test_physics_self_checks_without_an_oracle(nb):
    """Checks that need no reference run: the 1LPT displacement of a single plane wave equals its analytic
    value D(a) A sin(k q) / k along the wave vector, and a small-amplitude field evolved by nbody_bf to a = 1 reproduces
    the linear field (normalised to a = 1, growth factor 1) on the largest scales, converging with resolution."""
    from montecosmo_b200.cosmo import Cosmology
    c = Cosmology()
    n, m, amp, a = 16, 2, 0.05, 0.6
    shape = (n, n, n)
    q = O.regular_pos(shape).float().to(dev(nb))
    k = 2 * np.pi * m / n
    x = np.arange(n)
    delta = np.broadcast_to((amp * np.cos(k * x))[:, None, None], shape).copy()
    dk = torch.tensor(np.fft.rfftn(delta), dtype=torch.complex64, device=dev(nb))
    dpos, vel = nb.lpt(c, dk, q, a, lpt_order=1, read_order=1)
    ana = np.zeros((n**3, 3))
    ana[:, 0] = (-amp * np.sin(k * x) / k)[:, None, None].repeat(n, 1).repeat(n, 2).reshape(-1)
    D = float(nb.a2g(c, a))
    assert np.abs(vel.cpu().numpy() - ana).max() < 2e-6 and np.abs(dpos.cpu().numpy() - D * ana).max() < 2e-6
    # Linear growth on the fundamental modes: the field evolved to a = 1 and painted back correlates with the linear
    # field with amplitude 1 - O((k x cell)^2) -- the particle-mesh force is softened at the mesh scale -- so the deficit
    # must be small and shrink ~4x when the same modes are resolved by twice as many cells (measured: 6.0e-2 at 16^3,
    # 1.6e-2 at 32^3).  Per-mode ratios are NOT smooth for a barely displaced lattice (CIC weights have a kink at zero
    # displacement, which feeds harmonics), hence the correlation coefficient rather than a mode-by-mode comparison.
    deficit = {}
    for n in (16, 32):
        shape = (n, n, n)
        q = O.regular_pos(shape).float().to(dev(nb))
        rng = np.random.default_rng(12)
        kk = np.sqrt(sum(np.meshgrid(np.fft.fftfreq(n) ** 2, np.fft.fftfreq(n) ** 2, np.fft.rfftfreq(n) ** 2,
                                     indexing="ij")))
        sel = (kk < 1.5 / n) & (kk > 0)
        lin = np.fft.rfftn(rng.normal(size=shape)) * sel
        pos, _ = nb.nbody_bf(c, torch.tensor(lin, dtype=torch.complex64, device=dev(nb)), q, a0=0.05, a1=1.0, n_steps=8)
        out = nb.nufft(pos[0], shape, None, 1.0, 2, 2).cpu().numpy()  # interlaced, deconvolved: 1 + delta
        r = (out[sel] * np.conj(lin[sel])).sum() / (np.abs(lin[sel]) ** 2).sum()
        assert abs(r.imag) < 5e-3
        deficit[n] = 1.0 - r.real
    assert 0 < deficit[32] < 2.5e-2 and deficit[32] < 0.4 * deficit[16], deficit


def test_legacy_scale_factor_time_pieces(nb, golden):
    """lpt_fpm (nbody.py:1030-1073) and diffrax_vf (1076-1092) against the golden vectors of the reference source: 5e-5
    like the other force-level quantities; the vector field is differentiable in the positions."""
    from montecosmo_b200.cosmo import Cosmology
    g = golden("forces_lpt")
    shape = tuple(int(s) for s in g["shape"])
    dk = torch.tensor(g["delta_k"], dtype=torch.complex64, device=dev(nb))
    pos = torch.tensor(g["pos"], dtype=torch.float32, device=dev(nb))
    for order in (1, 2):
        dq, p = nb.lpt_fpm(Cosmology(), dk, pos, 0.3, order, 2)
        assert rel(dq, g[f"lpt_fpm{order}_dq"]) < 5e-5 and rel(p, g[f"lpt_fpm{order}_p"]) < 5e-5
    vel = torch.tensor(g["vf_vel"], dtype=torch.float32, device=dev(nb))
    pl = leaf(pos)
    dp, dv = nb.diffrax_vf(Cosmology(), shape, 2)(0.5, (pl, vel), None)
    assert rel(dp, g["vf_dpos"]) < 1e-6 and rel(dv, g["vf_dvel"]) < 5e-5
    dv.sum().backward()
    assert pl.grad is not None and bool(torch.isfinite(pl.grad).all())


def test_observation_chain_golden(nb, golden):
    """The cell -> physical -> redshift-space helpers of bricks.py:628-877 (frames with a rotated box, curved and flat
    sky lines of sight, light-cone scale factors, redshift-space distortions, Alcock-Paczynski) against the golden
    vectors of the reference source: float32 tensors on the engine's device, 2e-6 of the largest value; the light-cone
    displacement is differentiable in the velocities and in the cosmology."""
    from scipy.spatial.transform import Rotation
    from montecosmo_b200 import bricks as B
    from montecosmo_b200.cosmo import Cosmology
    g = golden("observation")
    gr = golden("growth")
    shape, box, center = tuple(int(s) for s in g["shape"]), tuple(g["box_size"]), tuple(g["box_center"])
    rot = Rotation.from_matrix(g["rot_matrix"])
    d = dev(nb)
    pos = torch.tensor(g["pos"], dtype=torch.float32, device=d)
    vel = torch.tensor(g["vel"], dtype=torch.float32, device=d)
    oc, ob, h, ns, s8 = gr["other_params"]
    cosmo, fid = Cosmology(), Cosmology(Omega_c=oc, Omega_b=ob, h=h, n_s=ns, sigma8=s8)

    def near(x, ref, tol=2e-6):
        x = x.detach().cpu().numpy().astype(np.float64)
        assert np.abs(x - ref).max() <= tol * max(np.abs(ref).max(), 1e-30), np.abs(x - ref).max()

    phys = B.cell2phys_pos(pos, center, rot, box, shape)
    near(phys, g["cell2phys_pos"])
    near(B.phys2cell_pos(phys, center, rot, box, shape), g["phys2cell_roundtrip"], 1e-5)
    near(B.cell2phys_vel(vel, rot, box, shape), g["cell2phys_vel"])
    near(B.phys2cell_vel(vel, rot, box, shape), g["phys2cell_vel"])
    near(B.cell2phys_pos(pos, center, g["rot_matrix"], box, shape), g["cell2phys_pos"])  # a matrix works as well
    assert np.allclose(B.pos_mesh(center, rot, box, shape).numpy(), g["pos_mesh"], rtol=1e-12, atol=1e-9)
    pvel = B.cell2phys_vel(vel, rot, box, shape)
    for tag, curved in (("curved", True), ("flat", False)):
        assert np.allclose(B.radius_mesh(center, rot, box, shape, curved).numpy(), g[f"radius_mesh_{tag}"], rtol=1e-12)
        los, a = B.los_scalefactor_pos(pos, center, rot, box, shape, cosmo, None, curved)
        near(los * torch.ones(1, 3, device=d), g[f"los_{tag}"])
        near(a, g[f"a_{tag}"], 1e-6)
        near(B.rsd(cosmo, vel, los, a, rot, box, shape, dvel=0.01), g[f"rsd_{tag}"], 5e-6)
        near(B.ap_auto(phys, los, cosmo, fid, curved), g[f"ap_auto_{tag}"], 5e-6)
        near(B.ap_param(phys, los, dict(alpha_iso=1.03, alpha_ap=0.97), curved), g[f"ap_param_{tag}"])
        rpos = phys.norm(dim=-1, keepdim=True) if curved else (phys * los).sum(-1, keepdim=True).abs()
        near(B.rsd_ap_auto(phys, pvel, rpos, los, a, cosmo, fid, curved), g[f"rsd_ap_auto_{tag}"], 5e-6)
    near(B.scale_pos(phys, torch.tensor(g["los_flat"][:1], dtype=torch.float32, device=d), 1.1, 0.9), g["scale_pos"])
    assert np.allclose(B.isoap2parperp(1.03, 0.97), g["isoap2parperp"]) and np.allclose(B.parperp2isoap(1.05, 0.98),
                                                                                      g["parperp2isoap"])
    red, aa = B.redges_and_scalefactors(cosmo, 500.0, 2500.0, 4)
    assert np.allclose(red.numpy(), g["redges"], rtol=1e-10) and np.allclose(aa.numpy(), g["redges_a"], rtol=1e-10)
    # gradients: velocities, and Omega_c through the growth table, on the light cone
    ocl = torch.tensor(0.26447041, dtype=torch.float64, requires_grad=True)
    vl = leaf(vel)
    los, a = B.los_scalefactor_pos(pos, center, rot, box, shape, cosmo, None, True)
    B.rsd(Cosmology(Omega_c=ocl), vl, los, a, rot, box, shape).pow(2).sum().backward()
    assert bool(torch.isfinite(vl.grad).all()) and float(vl.grad.abs().sum()) > 0 and float(ocl.grad.abs()) > 0


OBS_CASES = {
    # name: (curved, a_obs, ap_auto, rotated, velocity bias, lattice-relative positions, paint shape, rsd)
    "curved_lightcone_auto": (True, None, True, True, True, False, (24, 24, 24), True),
    "flat_lightcone_auto": (False, None, True, True, True, False, None, True),
    "flat_scalar_param": (False, 0.7, False, True, False, False, (24, 24, 24), True),
    "curved_scalar_param": (True, 0.7, False, False, True, True, None, True),
    "flat_scalar_plain": (False, 0.5, None, False, False, True, None, True),
    "curved_no_rsd_auto": (True, None, True, True, False, False, None, False),
}


@pytest.mark.parametrize("case", list(OBS_CASES))
def test_fused_observation_chain(nb, case):
    """nufft_observed -- the observation chain of model.py:780-799 applied inside the paint kernels (csrc/obs.h,
    mcpm_nufft_obs) -- against the chain itself in float64: cell2phys_pos, los_scalefactor_pos, rsd, ap_auto | ap_param,
    phys2cell_pos (the mirrors pinned to the reference source by test_observation_chain_golden) followed by the
    oracle's nufft.  Half spectrum 2e-5 relative L2; cotangents of the positions, velocities, velocity bias and weights
    1e-4; of Omega_c (through the light-cone growth table, the scalar D f and the ap_auto table) and of the
    Alcock-Paczynski parameters 2e-4.  Rotated and unrotated boxes 1.5 Gpc/h from the observer, curved and flat sky,
    absolute positions and lattice-relative displacements, paint mesh 1.5x the final one or equal to it."""
    from scipy.spatial.transform import Rotation
    from montecosmo_b200 import bricks as B
    from montecosmo_b200.cosmo import Cosmology
    curved, a_obs, ap_auto, rotated, with_dvel, relative, paint, rsd_on = OBS_CASES[case]
    rng = np.random.default_rng(sum(map(ord, case)))
    shape, box, center = (16, 16, 16), (640.0, 480.0, 560.0), (300.0, -200.0, 1500.0)
    rot = Rotation.from_rotvec([0.2, -0.4, 0.6]) if rotated else None
    n = int(np.prod(shape))
    q = O.regular_pos(shape)
    disp = torch.tensor(rng.normal(scale=0.7, size=q.shape))
    vel0 = torch.tensor(rng.normal(scale=0.4, size=q.shape))
    dvel0 = torch.tensor(rng.normal(scale=3.0, size=q.shape)) if with_dvel else None
    w0 = torch.tensor(rng.uniform(0.3, 2.0, n))
    cshape = (shape[0], shape[1], shape[2] // 2 + 1)
    ck = torch.tensor(rng.normal(size=cshape) + 1j * rng.normal(size=cshape))
    fid = Cosmology(Omega_c=0.21, Omega_b=0.05, h=0.7)
    apv = dict(alpha_iso=1.03, alpha_ap=0.97)

    def cosmo_and_ap():
        oc = torch.tensor(0.2589, dtype=torch.float64, requires_grad=True)
        ap = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in apv.items()}
        return oc, Cosmology(Omega_c=oc), ap

    # the engine, transform inside the paint
    oc, cosmo, ap = cosmo_and_ap()
    pos = leaf(disp if relative else q + disp, nb, torch.float32)
    vel, w = leaf(vel0, nb, torch.float32), leaf(w0, nb, torch.float32)
    dvel = leaf(dvel0, nb, torch.float32) if with_dvel else None
    obs = B.observation(cosmo, center, rot, box, shape, a_obs, curved, rsd_on, ap_auto, fid, ap)
    out = nb.nufft_observed(pos, vel if rsd_on else None, shape, obs, paint, w, dvel, 2, 2, paint_deconv=True,
                            lattice=shape if relative else None)
    (out * ck.to(torch.complex64).to(dev(nb)).conj()).real.sum().backward()
    # the chain in float64
    oc2, cosmo2, ap2 = cosmo_and_ap()
    po, vo, wo = leaf(q + disp), leaf(vel0), leaf(w0)
    do = leaf(dvel0) if with_dvel else 0.0
    geo = (center, rot, box)
    los, a = B.los_scalefactor_pos(po, *geo, shape, cosmo2, a_obs, curved)
    phys = B.cell2phys_pos(po, *geo, shape)
    if rsd_on:
        phys = phys + B.rsd(cosmo2, vo, los, a, rot, box, shape, do)
    if ap_auto is not None:
        phys = B.ap_auto(phys, los, cosmo2, fid, curved) if ap_auto else B.ap_param(phys, los, ap2, curved)
    ref = O.nufft(B.phys2cell_pos(phys, *geo, shape), shape, paint, wo, 2, 2, paint_deconv=True)
    (ref * ck.conj()).real.sum().backward()
    print(case, "spectrum", rel(out, ref.detach()), "pos", rel(pos.grad, po.grad), "w", rel(w.grad, wo.grad),
          "Omega_c", None if oc.grad is None else (float(oc.grad), float(oc2.grad)))
    assert rel(out, ref.detach()) < 2e-5
    assert rel(pos.grad, po.grad) < 1e-4 and rel(w.grad, wo.grad) < 1e-4
    if rsd_on:
        assert rel(vel.grad, vo.grad) < 1e-4
    if with_dvel and rsd_on:
        assert rel(dvel.grad, do.grad) < 1e-4
    if oc2.grad is not None and float(oc2.grad.abs()) > 0:
        assert oc.grad is not None and abs(float(oc.grad) - float(oc2.grad)) < 1e-3 * abs(float(oc2.grad)), (oc.grad, oc2.grad)
    if ap_auto is False:
        for k in (("alpha_iso",) if curved else ("alpha_iso", "alpha_ap")):
            assert abs(float(ap[k].grad) - float(ap2[k].grad)) < 2e-4 * abs(float(ap2[k].grad)), k


def test_observed_nufft_golden(nb, golden):
    """nufft_observed (the observation chain inside the paint kernels, mcpm_nufft_obs) against golden vectors of the
    REFERENCE SOURCE: bricks.los_scalefactor_pos, cell2phys_pos, rsd with a velocity bias, ap_auto | ap_param,
    phys2cell_pos composed as model.py:780-805 does, followed by nbody.nufft on a 1.5x paint mesh
    (tests/golden/make_golden.py: observed_nufft).  Curved and flat sky, light cone and scalar a_obs, both
    Alcock-Paczynski forms and none: half spectrum 2e-5 relative L2."""
    from scipy.spatial.transform import Rotation
    from montecosmo_b200 import bricks as B
    from montecosmo_b200.cosmo import Cosmology
    g, gr = golden("observed_nufft"), golden("growth")
    shape, paint = tuple(int(s) for s in g["shape"]), tuple(int(s) for s in g["paint_shape"])
    box, center, rot = tuple(g["box_size"]), tuple(g["box_center"]), Rotation.from_matrix(g["rot_matrix"])
    oc, ob, h, ns, s8 = gr["other_params"]
    cosmo, fid = Cosmology(), Cosmology(Omega_c=oc, Omega_b=ob, h=h, n_s=ns, sigma8=s8)
    d = dev(nb)
    pos, vel, dvel, w = (torch.tensor(g[k], dtype=torch.float32, device=d) for k in ("pos", "vel", "dvel", "weights"))
    ap = dict(alpha_iso=float(g["alpha_iso"]), alpha_ap=float(g["alpha_ap"]))
    for tag, curved, a_obs, ap_auto in (("curved_lightcone_auto", True, None, True), ("flat_lightcone_auto", False, None, True),
                                        ("curved_scalar_param", True, 0.7, False), ("flat_scalar_param", False, 0.7, False),
                                        ("flat_scalar_plain", False, 0.7, None)):
        obs = B.observation(cosmo, center, rot, box, shape, a_obs, curved, True, ap_auto, fid, ap)
        out = nb.nufft_observed(pos, vel, shape, obs, paint, w, dvel, 2, 2, "rectangular", True)
        assert rel(out, g[f"nufft_{tag}"]) < 2e-5, tag


@pytest.mark.parametrize("order", [3, 4])
def test_fused_observation_chain_other_windows(order):
    """The observed paint with the assignment windows the benchmark does not use -- TSC / PCS, and the Kaiser-Bessel
    family on a 1.5x paint mesh -- against the float64 chain + the oracle's nufft: the window only changes how the
    observed position is deposited, so the same transform and transpose serve every instantiation.  CPU port only (the
    B200 runs of this transform use the CIC instantiation; these share its source).  Spectrum 2e-5 (Kaiser-Bessel 1e-4, see
    below), cotangents 1e-4."""
    import montecosmo_b200.nbody as nbody
    from montecosmo_b200 import bricks as B
    from montecosmo_b200.cosmo import Cosmology
    from oracle import cpu_port
    old = nbody._OPS
    nbody._OPS = cpu_port.cpu_ops()
    try:
        rng = np.random.default_rng(order)
        shape, paint, box, center = (12, 12, 12), (18, 18, 18), (480.0, 480.0, 480.0), (-200.0, 300.0, 1200.0)
        q = O.regular_pos(shape)
        disp, vel0 = torch.tensor(rng.normal(scale=0.7, size=q.shape)), torch.tensor(rng.normal(scale=0.4, size=q.shape))
        w0 = torch.tensor(rng.uniform(0.3, 2.0, q.shape[0]))
        cshape = (shape[0], shape[1], shape[2] // 2 + 1)
        ck = torch.tensor(rng.normal(size=cshape) + 1j * rng.normal(size=cshape))
        fid = Cosmology(Omega_c=0.21, Omega_b=0.05, h=0.7)
        for kernel_type in ("rectangular", "kaiser_bessel"):
            cosmo = Cosmology()
            pos, vel, w = leaf(q + disp, None, torch.float32), leaf(vel0, None, torch.float32), leaf(w0, None, torch.float32)
            obs = B.observation(cosmo, center, None, box, shape, None, True, True, True, fid)
            out = nbody.nufft_observed(pos, vel, shape, obs, paint, w, None, order, 2, kernel_type, True)
            (out * ck.to(torch.complex64).conj()).real.sum().backward()
            po, vo, wo = leaf(q + disp), leaf(vel0), leaf(w0)
            los, a = B.los_scalefactor_pos(po, center, None, box, shape, cosmo, None, True)
            phys = B.cell2phys_pos(po, center, None, box, shape) + B.rsd(cosmo, vo, los, a, None, box, shape)
            phys = B.ap_auto(phys, los, cosmo, fid, True)
            ref = O.nufft(B.phys2cell_pos(phys, center, None, box, shape), shape, paint, wo, order, 2, kernel_type, True)
            (ref * ck.conj()).real.sum().backward()
            # Kaiser-Bessel: the deconvolution divides by a window transform that is small at high k for these orders
            # and amplifies the float32 rounding of the deposits, summed by atomics in whatever order (run-to-run spread
            # of the engine's own result 1-2e-5 at order 4, exactly 0 with one thread)
            assert rel(out, ref.detach()) < (2e-5 if kernel_type == "rectangular" else 1e-4), kernel_type
            assert rel(pos.grad, po.grad) < 1e-4 and rel(vel.grad, vo.grad) < 1e-4 and rel(w.grad, wo.grad) < 1e-4, kernel_type
    finally:
        nbody._OPS = old


@pytest.mark.parametrize("curved", [True, False])
def test_lightcone_functions_on_the_device(nb, curved):
    """bricks.lightcone_functions (one engine pass, mcpm_radial_tables: growth lookups of the light cone at every
    particle's comoving distance) against the chain it replaces -- los_scalefactor_pos with a_obs None followed by a2g,
    a2g2, a2dg2dg, a2f on the host's float64 tables: values 2e-6 of the largest; cotangents of Omega_c (through every
    table node) 2e-4 and of the positions 2e-3 (the slope of a re-tabulated piecewise-linear function)."""
    from scipy.spatial.transform import Rotation
    from montecosmo_b200 import bricks as B
    from montecosmo_b200 import cosmo as CO
    rng = np.random.default_rng(3 + curved)
    shape, box, center = (12, 10, 14), (600.0, 500.0, 700.0), (250.0, -300.0, 1400.0)
    rot = Rotation.from_rotvec([0.3, 0.2, -0.5])
    pos0 = torch.tensor(rng.uniform(-1.0, 1.0, size=(500, 3)) + rng.uniform(0, 1, size=(500, 3)) * np.asarray(shape))
    cot = torch.tensor(rng.normal(size=(500, 4)))
    fns = (CO.a2g, CO.a2g2, CO.a2dg2dg, CO.a2f)
    oc = torch.tensor(0.2589, dtype=torch.float64, requires_grad=True)
    pos = leaf(pos0, nb, torch.float32)
    out = torch.stack(B.lightcone_functions(CO.Cosmology(Omega_c=oc), pos, center, rot, box, shape, curved, fns), dim=1)
    (out * cot.float().to(dev(nb))).sum().backward()
    oc2 = torch.tensor(0.2589, dtype=torch.float64, requires_grad=True)
    po = leaf(pos0)
    c2 = CO.Cosmology(Omega_c=oc2)
    _, a = B.los_scalefactor_pos(po, center, rot, box, shape, c2, None, curved)
    ref = torch.stack([fn(c2, a.reshape(-1)) for fn in fns], dim=1)
    (ref * cot).sum().backward()
    err = (out.detach().cpu().double() - ref.detach()).abs().max(0).values / ref.detach().abs().max(0).values
    assert float(err.max()) < 2e-6, err
    assert abs(float(oc.grad) - float(oc2.grad)) < 2e-4 * abs(float(oc2.grad)), (oc.grad, oc2.grad)
    assert rel(pos.grad, po.grad) < 2e-3


@pytest.mark.parametrize("case", ["lightcone_lpt_curved", "nbody_flat", "nbody_curved_lattice"])
def test_general_evolve_against_oracle(nb, case):
    """FieldLevelModel.evolve -- the general 'lpt' / 'nbody' branch of model.py:683-837 -- against the oracle's float64
    restatement of the same chain (every callee of which is pinned to golden vectors of the reference source):
    predicted mesh 1e-4 relative L2, gradient of a linear functional w.r.t. the white field 1e-3.
      lightcone_lpt_curved: rotated box 1500 Mpc/h from the observer, curved sky, per-particle scale factors from the
        comoving distance, full bias expansion with the velocity term, RSD, automatic Alcock-Paczynski, 1.5x paint mesh;
      nbody_flat: 3 BullFrog steps to a scalar a_obs, flat sky along the box centre, 8^3 particles in a 16^3 mesh;
      nbody_curved_lattice: 2 steps, one particle per cell (displacements carried from the loop into the observed paint),
        rotated box, curved sky, velocity bias, automatic Alcock-Paczynski."""
    from scipy.spatial.transform import Rotation
    from montecosmo_b200.cosmo import Cosmology
    from montecosmo_b200.model import FieldLevelModel
    rng = np.random.default_rng(17)
    shape, box = (16, 16, 16), (640.0, 640.0, 640.0)
    if case == "lightcone_lpt_curved":
        rot = Rotation.from_rotvec([0.2, -0.4, 0.6])
        bias = dict(b1=0.9, b2=0.3, bs2=-0.2, bn2=5.0, bnpar=8.0)
        cfg = dict(evolution="lpt", a_obs=None, box_center=(300.0, -200.0, 1500.0), box_rot=rot, curved_sky=True,
                   bias=bias, paint_oversamp=1.5, ap_auto=True, cosmo_fid=Cosmology(Omega_c=0.21, Omega_b=0.05, h=0.7))
        okw = dict(evolution="lpt", a_obs=None, box_center=cfg["box_center"], box_rot=rot, curved_sky=True, bias=bias,
                   paint_shape=(24, 24, 24), ap_fid=O.Cosmology(Omega_c=0.21, Omega_b=0.05, h=0.7))
    elif case == "nbody_curved_lattice":
        # one particle per cell: the loop and the observed paint carry lattice-relative displacements (model.py mirror)
        rot = Rotation.from_rotvec([-0.3, 0.1, 0.2])
        bias = dict(b1=0.8, b2=-0.1, bnpar=3.0)
        cfg = dict(evolution="nbody", n_steps=2, a_start=0.1, a_obs=0.7, box_center=(-400.0, 250.0, 1300.0), box_rot=rot,
                   curved_sky=True, bias=bias, ap_auto=True, cosmo_fid=Cosmology(Omega_c=0.21, Omega_b=0.05, h=0.7))
        okw = dict(evolution="nbody", n_steps=2, a_start=0.1, a_obs=0.7, box_center=cfg["box_center"], box_rot=rot,
                   curved_sky=True, bias=bias, ap_fid=O.Cosmology(Omega_c=0.21, Omega_b=0.05, h=0.7))
    else:
        bias = dict(b1=0.7, b2=0.2)
        cfg = dict(evolution="nbody", n_steps=3, a_start=0.1, a_obs=0.8, box_center=(0.0, 0.0, 2000.0), curved_sky=False,
                   bias=bias, ptcl_oversamp=0.5)
        okw = dict(evolution="nbody", n_steps=3, a_start=0.1, a_obs=0.8, box_center=cfg["box_center"], curved_sky=False,
                   bias=bias, ptcl_shape=(8, 8, 8))
    m = FieldLevelModel(shape, box, **cfg)
    transfer = m.transfer.cpu().numpy().astype(np.float64)
    white = rng.normal(size=shape).astype(np.float32)
    w = leaf(torch.tensor(white), nb)
    out = m.evolve(w)
    wo = torch.tensor(white, dtype=torch.float64, requires_grad=True)
    ref = MO.evolve_general(wo, transfer, O.Cosmology(), shape, box, **okw)
    assert tuple(out.shape) == tuple(ref.shape)
    assert rel(out, ref) < 1e-4
    cot = torch.tensor(rng.normal(size=tuple(ref.shape)))
    (out * cot.float().to(dev(nb))).sum().backward()
    (ref * cot).sum().backward()
    assert rel(w.grad, wo.grad) < 1e-3


def test_general_evolve_fused_observation_matches_elementwise_chain(nb):
    """FieldLevelModel.evolve with the observation chain inside the paint (the default) against the same model with the
    chain as elementwise passes over the particle arrays (fused_observation=False), where the evolution mesh is 1.5x the
    initial one (positions in evol_shape cells painted onto init_shape: model.py:796) and the Alcock-Paczynski factors
    are parameters: predicted mesh 2e-5, gradient w.r.t. the white field 2e-4."""
    from montecosmo_b200.model import FieldLevelModel
    rng = np.random.default_rng(5)
    shape, box = (8, 8, 8), (400.0, 400.0, 400.0)
    white = rng.normal(size=shape).astype(np.float32)
    cot = torch.tensor(rng.normal(size=shape)).float().to(dev(nb))
    res = []
    for fused in (True, False):
        m = FieldLevelModel(shape, box, evolution="lpt", a_obs=0.6, box_center=(100.0, 50.0, 1200.0), curved_sky=False,
                            bias=dict(b1=0.8, b2=0.1, bnpar=4.0), evol_oversamp=1.5, ap_auto=False, fused_observation=fused)
        w = leaf(torch.tensor(white), nb)
        out = m.evolve(w, ap=dict(alpha_iso=1.02, alpha_ap=0.98))
        (out * cot).sum().backward()
        res.append((out.detach(), w.grad))
    assert tuple(res[0][0].shape) == shape
    assert rel(res[0][0], res[1][0]) < 2e-5 and rel(res[0][1], res[1][1]) < 2e-4


def test_kaiser_model_golden(nb, golden):
    """kaiser_boost / kaiser_model (bricks.py:170-232) in its three configurations -- flat sky (one Fourier multiply, with
    the scale-dependent PNG bias), flat sky on the light cone (a mesh of scale factors), curved sky (a mesh of lines of
    sight; six second-derivative transforms here, a spherical-harmonic sum in the reference) -- and los_scalefactor_mesh
    (bricks.py:768-786), against the golden vectors of the reference source."""
    from scipy.spatial.transform import Rotation
    from montecosmo_b200 import bricks as B
    from montecosmo_b200.cosmo import Cosmology
    g = golden("observation")
    shape, box = tuple(int(s) for s in g["kaiser_shape"]), tuple(g["kaiser_box"])
    center, rot = tuple(g["box_center"]), Rotation.from_matrix(g["rot_matrix"])
    kpow, los = (g["kaiser_kpow_k"], g["kaiser_kpow_p"]), g["kaiser_los"]
    dk = torch.tensor(g["kaiser_delta_k"], dtype=torch.complex64, device=dev(nb))
    c = Cosmology()
    assert np.allclose(B.kaiser_boost(c, 0.7, shape, box, 1.8, 0.5, "fNL", los, kpow), g["kaiser_boost"], rtol=1e-10)
    assert rel(B.kaiser_model(c, 0.7, dk, box, 1.8, 0.5, "fNL", los, kpow) - 1, g["kaiser_flat"] - 1) < 1e-5
    los_m, a_m = B.los_scalefactor_mesh(center, rot, box, shape, c, None, False)
    assert np.allclose(a_m.numpy(), g["kaiser_a_mesh_flat"], rtol=1e-10) and np.allclose(los_m.numpy(), los, rtol=1e-12)
    assert rel(B.kaiser_model(c, a_m, dk, box, 1.8, los=los_m) - 1, g["kaiser_lightcone"] - 1) < 1e-5
    los_c, a_c = B.los_scalefactor_mesh(center, rot, box, shape, c, None, True)
    assert np.allclose(a_c.numpy(), g["kaiser_a_mesh_curved"], rtol=1e-10)
    cell_los = torch.as_tensor(rot.apply(los_c.numpy().reshape(-1, 3), inverse=True).reshape(shape + (3,)))
    assert np.allclose(cell_los.numpy(), g["kaiser_cell_los"], rtol=1e-10, atol=1e-12)
    assert rel(B.kaiser_model(c, a_c, dk, box, 1.8, los=cell_los) - 1, g["kaiser_curved"] - 1) < 1e-5
    # the posterior mean / std are the Wiener filter of the same boost (bricks.py:234-247)
    mean, std = B.kaiser_posterior(dk, c, 0.7, box, 0.3, 1.8, los, kpow)
    boost = B.kaiser_boost(c, 0.7, shape, box, 1.8, los=los)
    kvec = nb.rfftk(shape, box)
    pm = np.interp(np.sqrt(sum(k**2 for k in kvec)).ravel(), kpow[0], kpow[1] * float(c.sigma8) ** 2, left=0.0,
                   right=0.0).reshape(boost.shape) * np.divide(shape, box).prod()
    s2 = pm / (1 + boost**2 / 0.3 * pm)
    assert rel(std, np.sqrt(s2)) < 1e-6 and rel(mean, s2 * boost / 0.3 * g["kaiser_delta_k"]) < 1e-6


def test_kaiser_evolution_of_the_general_model(nb):
    """evolution='kaiser' (model.py:690-699, 733-736): flat sky without light cone is one Fourier multiply of the linear
    field -- checked against numpy float64 from the model's own transfer mesh -- and differentiable in the white field."""
    from montecosmo_b200 import bricks as B
    from montecosmo_b200.model import FieldLevelModel
    rng = np.random.default_rng(19)
    shape, box, center = (16, 16, 16), (640.0,) * 3, (0.0, 300.0, 2000.0)
    m = FieldLevelModel(shape, box, evolution="kaiser", a_obs=0.6, box_center=center, bias=dict(b1=0.8))
    white = rng.normal(size=shape).astype(np.float32)
    w = leaf(torch.tensor(white), nb)
    out = m.evolve(w)
    los = np.array(center) / np.linalg.norm(center)
    boost = B.kaiser_boost(m.cosmology, 0.6, shape, box, 1.8, los=los)
    ref = 1 + np.fft.irfftn(np.fft.rfftn(white.astype(np.float64)) * m.transfer.cpu().numpy() * boost, s=shape, axes=(0, 1, 2))
    assert rel(out - 1, ref - 1) < 1e-5
    out.sum().backward()
    assert bool(torch.isfinite(w.grad).all())
    # a finer evolution mesh comes back to the initial shape; the two differ only on the Nyquist planes, whose modes the
    # padding splits into +-k_N halves that see different mu^2 before the crop merges them again (measured 7.7e-3)
    m2 = FieldLevelModel(shape, box, evolution="kaiser", a_obs=0.6, box_center=center, bias=dict(b1=0.8), evol_oversamp=1.5)
    out2 = m2.evolve(torch.tensor(white))
    assert tuple(out2.shape) == shape and rel(out2 - 1, ref - 1) < 2e-2
    k1, k2 = np.fft.rfftn(out2.cpu().numpy().astype(np.float64)), np.fft.rfftn(ref)
    inner = np.s_[1:7, 1:7, 1:7]  # away from every Nyquist plane the two agree to float32 rounding
    assert np.abs(k1[inner] - k2[inner]).max() < 1e-4 * np.abs(k2[inner]).max()


def test_power_bias_and_eulerian_expansion_golden(nb, golden):
    """lin_power_mesh / trans_phi2delta_interp / white2lin / lin2white (bricks.py:67-166), the bias parametrisations and
    fNL_bias (454-508), eulerian_bias (513-585) and count2delta (927-937) against the reference source."""
    from montecosmo_b200 import bricks as B
    from montecosmo_b200.cosmo import Cosmology
    g = golden("lagrangian_bias")
    shape, box = tuple(int(s) for s in g["shape"]), tuple(g["box_size"])
    kpow = (g["kpow_k"], g["kpow_p"])
    c = Cosmology()
    d = dev(nb)
    dk = torch.tensor(g["delta_k"], dtype=torch.complex64, device=d)
    assert np.allclose(B.lin_power_mesh(c, shape, box, kpow=kpow), g["lin_power_mesh"], rtol=1e-12)
    assert np.allclose(B.trans_phi2delta_interp(c, kpow=kpow)(np.array([1e-3, 0.05, 0.7])), g["trans_phi2delta"], rtol=1e-10)
    lin = B.white2lin(c, dk, shape, box, kpow)
    assert rel(lin, g["white2lin"]) < 1e-6
    assert rel(B.lin2white(c, lin, shape, box, kpow), g["lin2white"]) < 1e-6
    bias = {k[5:]: float(v) for k, v in g.items() if k.startswith("bias_")}
    png = B.fNL_bias(dict(fNL=20.0, fNL_bp=0.0, fNL_bpd=0.0), bias, p=1.0, png_type="fNL")
    assert np.allclose([png["fNL_bp"], png["fNL_bpd"]], g["fNL_bias"], rtol=1e-12)
    assert B.b1_E2L(B.b1_L2E(0.3)) == pytest.approx(0.3) and B.b2_E2L(B.b2_L2E(0.2, 0.5), 0.5) == pytest.approx(0.2)
    assert B.bpd_E2L(B.bpd_L2E(0.4, 0.6), 0.6) == pytest.approx(0.4)
    phik = torch.tensor(g["eul_phik"], dtype=torch.complex64, device=d)
    dkl = leaf(dk)
    we, dvel = B.eulerian_bias(dkl, phik, box, bias, png, png_type="fNL")
    assert dvel == 0.0 and rel(we, g["eulerian_weights_png"]) < 2e-5
    we.sum().backward()
    assert bool(torch.isfinite(torch.view_as_real(dkl.grad)).all())
    assert rel(B.eulerian_bias(dk, phik, box, bias, png, None)[0], g["eulerian_weights"]) < 2e-5
    cm, sm = (torch.tensor(g[k], dtype=torch.float32, device=d) for k in ("count_mesh", "selec_mesh"))
    assert rel(B.count2delta(cm, sm), g["count2delta"]) < 1e-5
    with pytest.raises(NotImplementedError):
        B.lin_power(c)  # kpow=None is jax_cosmo's Eisenstein-Hu power


def test_catalogue_registration_golden(nb, golden):
    """The callers of nufft / paint outside the model (bricks.py:879-897, 1026-1100): sky <-> cartesian
    coordinates, cutsky2count, cutsky2selection, fullsky2count (streamed chunks with redshift-space shift), against the
    golden vectors of the reference source.  Catalogue particles are unordered: the generic scatter kernels."""
    from montecosmo_b200 import bricks as B
    from montecosmo_b200.cosmo import Cosmology
    g = golden("observation")
    c = Cosmology()
    cat = {k: g[f"cat_{k}"] for k in ("RA", "DEC", "Z", "WEIGHT")}
    size, center, rotvec = tuple(g["cat_box"]), tuple(g["cat_center"]), tuple(g["cat_rotvec"])
    cart = B.radecz2cart(c, cat)
    assert np.allclose(cart.numpy(), g["cat_cart"], rtol=1e-10)
    back = B.cart2radecz(c, cart)
    assert np.allclose(np.stack([back["RA"].numpy(), back["DEC"].numpy(), back["Z"].numpy()]), g["cat_back"], rtol=1e-9)
    cnt = B.cutsky2count(cat, c, (12, 14, 12), 1.5, size, center, rotvec)
    assert rel(cnt, g["cutsky2count"]) < 5e-5 and abs(float(cnt.sum()) - cat["WEIGHT"].sum()) < 1e-3 * cat["WEIGHT"].sum()
    sel, msk = B.cutsky2selection(cat, c, (6, 8, 6), (12, 14, 12), None, size, center, rotvec)
    assert rel(sel, g["cutsky_selection"]) < 5e-5 and np.array_equal(msk.cpu().numpy(), g["cutsky_mask"])
    flos = np.array(center) / np.linalg.norm(center)
    w = cat["WEIGHT"]
    chunks = [{"pos": g["full_pos"][:100], "vel": g["full_vel"][:100], "WEIGHT": w[:100]},
              {"pos": g["full_pos"][100:], "vel": g["full_vel"][100:], "WEIGHT": w[100:]}]
    assert rel(B.fullsky2count(chunks, c, 0.65, flos, size, center, rotvec, (12, 14, 12), None), g["fullsky2count"]) < 5e-5


def test_baseline_config_c1_full_size(nb):
    """BASELINE.json configs[0] at its full size -- 64^3 mesh / 64^3 particles, 640 Mpc/h box, 2LPT + 5 BullFrog steps
    to a = 1, linear bias + flat-sky RSD, interlaced deconvolved paint, Gaussian likelihood -- against the float64 oracle:
    log-density 1e-5, grad(log-density) cosine >= 0.9999 and relative L2 <= 1e-3.  Round 1 needed 3e-3 here:
    absolute float32 positions put particles on the other side of a cell face, where the CIC derivative jumps (9.6e-4,
    1.3e-3, 8.2e-4, 1.3e-3 for white-noise seeds 0-3 on the CPU port).  FieldModel now carries displacements from the
    lattice sites (mcpm_engine_set_relative): 2.9e-5, 2.2e-5, 5.5e-5, 2.8e-4 for the same seeds."""
    from montecosmo_b200.model import FieldModel
    rng = np.random.default_rng(0)
    shape = (64, 64, 64)
    m = FieldModel(shape, (640.0,) * 3, evolution="nbody", n_steps=5, a_obs=1.0, b1=1.0, sigma_obs=1.0)
    kw = dict(evolution="nbody", n_steps=5, a_obs=1.0, b1=1.0)
    white = rng.normal(size=shape).astype(np.float32)
    transfer = m.transfer.cpu().numpy().astype(np.float64)
    with torch.no_grad():
        truth = MO.evolve(torch.tensor(rng.normal(size=shape)), transfer, O.Cosmology(), shape, **kw)
    obs = (truth.numpy() + rng.normal(size=shape)).astype(np.float32)
    lp, g = m.value_and_force(white, obs)
    lpo, go = MO.value_and_force(white.astype(np.float64), obs.astype(np.float64), transfer, O.Cosmology(), shape,
                                 sigma_obs=1.0, **kw)
    assert abs(float(lp) - float(lpo)) < 1e-5 * abs(float(lpo))
    gn, gon = g.detach().cpu().numpy().ravel().astype(np.float64), go.numpy().ravel()
    assert np.linalg.norm(gn - gon) <= 1e-3 * np.linalg.norm(gon)
    assert gn @ gon / np.linalg.norm(gn) / np.linalg.norm(gon) >= 0.9999
