"""
Parity of the C ABI (include/mcpm.h) with the float64 oracle and the golden vectors of the reference source.

Every test runs on two backends through the SAME ctypes prototypes:
  * "hostemu"  -- the engine sources compiled as serial C++ (tests/hostemu.py), CPU, always run;
  * "cuda"     -- libmcpm.so on a B200 (marked gpu).
Engine arithmetic is float32; the oracle is float64.  Tolerances (stated per test) are relative L2 unless noted.
"""
import numpy as np
import pytest
import torch

from oracle import pm_oracle as O
from tests.backends import NumpyAdapter, to_numpy

INF = float("inf")


def _make_ops(kind):
    from montecosmo_b200.ops import Ops
    if kind == "hostemu":
        from tests import hostemu
        return Ops(hostemu.load(), NumpyAdapter())
    from montecosmo_b200 import _lib
    from montecosmo_b200.ops import TorchCudaAdapter
    return Ops(_lib.load(), TorchCudaAdapter())


@pytest.fixture(scope="module", params=["hostemu", pytest.param("cuda", marks=pytest.mark.gpu)])
def ops(request):
    return _make_ops(request.param)


def rel(a, b):
    a, b = np.asarray(to_numpy(a), dtype=np.float64 if not np.iscomplexobj(to_numpy(a)) else np.complex128), \
        np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def f32(x):
    return np.asarray(x, dtype=np.float32)


def c64(x):
    return np.asarray(x, dtype=np.complex64)


def T(x, dtype=torch.float64):
    return torch.as_tensor(np.asarray(x), dtype=dtype)


# ------------------------------------------------------------------------------------------------ paint / read
@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_paint_read_golden(ops, golden, order):
    """vs the reference source (float64 inputs): 2e-5 covers float32 rounding of positions up to +-40 cells."""
    g = golden("paint_read")
    shape = tuple(int(s) for s in g["shape"])
    assert rel(ops.paint(f32(g["pos"]), shape, f32(g["weights"]), order=order), g[f"paint_w_{order}"]) < 2e-5
    assert rel(ops.paint(f32(g["pos"]), shape, None, order=order), g[f"paint_1_{order}"]) < 2e-5
    assert rel(ops.read(f32(g["pos"]), f32(g["mesh"]), order=order), g[f"read_{order}"]) < 2e-5


@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_paint_read_oracle_same_inputs(ops, order):
    """vs the oracle fed the SAME float32-rounded inputs: 2e-6 (pure float32 arithmetic error)."""
    rng = np.random.default_rng(10 + order)
    shape = (10, 12, 8)
    pos = f32(rng.uniform(-15, 30, (2000, 3)))
    pos[:100] = np.round(pos[:100])
    pos[100:200] = np.floor(pos[100:200]) + 0.5
    w = f32(rng.normal(size=2000))
    mesh = f32(rng.normal(size=shape))
    scale, shift = (1.25, 0.75, 1.5), 0.5
    assert rel(ops.paint(pos, shape, w, order=order), O.paint(T(pos), shape, T(w), order).numpy()) < 2e-6
    assert rel(ops.read(pos, mesh, order=order), O.read(T(pos), T(mesh), order).numpy()) < 2e-6
    # in-kernel transform x' = x*scale + shift (nufft / interlace)
    xs = T(pos) * T(f32(scale)) + shift
    assert rel(ops.paint(pos, shape, w, 0.5, order, scale, shift), O.paint(xs, shape, T(w) * 0.5, order).numpy()) < 5e-6
    assert rel(ops.read(pos, mesh, order, scale, shift), O.read(xs, T(mesh), order).numpy()) < 5e-6
    # multi-mesh read, interleaved output
    m3 = f32(rng.normal(size=(3, *shape)))
    ref = np.stack([O.read(T(pos), T(m3[i]), order).numpy() for i in range(3)], -1)
    assert rel(ops.read(pos, m3, order=order), ref) < 2e-6


@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_paint_properties(ops, order):
    """weight conservation (bricks.py:1101-1102); read = paint^T; lattice read returns mesh values (nbody.py:602)."""
    rng = np.random.default_rng(order)
    shape = (8, 6, 10)
    pos = f32(rng.uniform(-5, 15, (500, 3)))
    w = f32(rng.uniform(0.5, 1.5, 500))
    m = f32(rng.normal(size=shape))
    painted = to_numpy(ops.paint(pos, shape, w, order=order)).astype(np.float64)
    assert abs(painted.sum() - w.astype(np.float64).sum()) < 1e-4 * w.sum()
    lhs = float((to_numpy(ops.read(pos, m, order=order)).astype(np.float64) * w).sum())
    rhs = float((m.astype(np.float64) * painted).sum())
    assert abs(lhs - rhs) < 1e-4 * max(abs(lhs), 1.0)
    if order <= 2:
        q = f32(O.regular_pos(shape).numpy())
        assert rel(ops.read(q, m, order=order), m.reshape(-1)) < 1e-7


def test_paint_edge_cases(ops):
    shape = (8, 8, 8)
    # empty particle set -> zero mesh; accumulate adds
    z = to_numpy(ops.paint(np.zeros((0, 3), np.float32), shape, None, order=2))
    assert z.shape == shape and not z.any()
    pos = f32([[0.0, 0.0, 0.0], [7.5, 7.5, 7.5], [-0.25, 8.0, 15.999]])
    a = ops.paint(pos, shape, None, order=2)
    b = ops.paint(pos, shape, None, order=2, out=a, accumulate=True)
    assert rel(b, 2 * O.paint(T(pos), shape, 1.0, 2).numpy()) < 1e-6
    from montecosmo_b200._capi import McpmError
    with pytest.raises(McpmError):
        ops.paint(pos, shape, None, order=5)
    with pytest.raises(McpmError):
        ops.paint(pos, (8, 8, 2), None, order=4)


def test_empty_inputs_and_rejected_arguments(ops):
    """Empty particle sets flow through every particle operator (zero-size outputs, zero meshes); arguments the
    reference would reject, or that the engine cannot serve, come back as MCPM_EINVAL with a message -- no crash, no
    partial write.  Positions far outside the box, or non-finite, never index outside the mesh."""
    from montecosmo_b200._capi import McpmError
    shape = (8, 6, 10)
    e = np.zeros((0, 3), np.float32)
    mesh = f32(np.random.default_rng(0).normal(size=shape))
    assert to_numpy(ops.read(e, mesh)).shape == (0,)
    assert to_numpy(ops.read_grad(e, mesh)).shape == (0, 3)
    pb, wb = ops.paint_vjp(e, mesh)
    assert to_numpy(pb).shape == (0, 3) and to_numpy(wb).shape == (0,)
    assert not to_numpy(ops.paint3(e, e, shape)).any()
    assert not to_numpy(ops.nufft_paint(e, shape)).any()
    assert to_numpy(ops.pm_forces(e, shape)).shape == (0, 3)
    one = np.zeros((3, 3), np.float32)
    for bad in (lambda: ops.paint(one, shape, order=5), lambda: ops.paint(one, shape, order=0),
                lambda: ops.paint(one, (2, 2, 2), order=4), lambda: ops.nufft_paint(one, shape, interlace_order=7),
                lambda: ops.deconv(c64(np.zeros((8, 6, 6))), 5, kb_kcut=1.0)):
        with pytest.raises(McpmError) as ei:
            bad()
        assert ei.value.code == 1 and str(ei.value)
    far = f32([[1e9, -1e9, 3e8]])
    out = to_numpy(ops.paint(far, shape))
    assert abs(out.sum() - 1.0) < 1e-6                      # wrapped, mass conserved
    if isinstance(ops.A, NumpyAdapter):                     # (on the GPU this runs last: tests/test_zz_cross_check_256.py)
        bad = f32([[np.nan, 0.0, 0.0], [np.inf, 1.0, 2.0], [-np.inf, 7.999999, 9.999999]])
        ops.paint(bad, shape)                               # non-finite positions: garbage in, but no fault
        ops.read(bad, mesh)


@pytest.mark.parametrize("grad_fd", [0, 2, 4])
def test_yz_gradients_of_the_two_field_slab_transform(grad_fd):
    """mcpm_yz_gradients -- the y / z halves of the force operator that a slab-decomposed caller applies on its own (y,z)
    spectra so that two fields instead of three cross the distributed x-transform (dist.py) -- against a NumPy float64
    restatement of gradient_hat (nbody.py:136-163) with the Hermitian projection on the self-conjugate planes, and as
    an adjoint pair:  <T0 Phi, C> = <Phi, conj-transpose>  i.e.  sum conj(F_y) C_y + conj(F_z) C_z = sum conj(Phi) i (g_y C_y
    + g_z C_z).  Values 2e-6.  CPU port (the B200 checks the same kernel inside tests/test_xfuse.py's two-field tests)."""
    ops = _make_ops("hostemu")
    rng = np.random.default_rng(grad_fd)
    xl, ny, nz = 5, 8, 12
    nzc = nz // 2 + 1
    A = ops.A
    cx = lambda *sh: (rng.normal(size=sh) + 1j * rng.normal(size=sh)).astype(np.complex64)
    term = {0: lambda k: k, 2: np.sin, 4: lambda k: (8 * np.sin(k) - np.sin(2 * k)) / 6}[grad_fd]
    ky = 2 * np.pi * np.fft.fftfreq(ny)
    kz = 2 * np.pi * np.fft.rfftfreq(nz)
    gy = np.broadcast_to(term(ky)[:, None], (ny, nzc)).copy()
    gz = np.broadcast_to(term(kz)[None, :], (ny, nzc)).copy()
    sc = np.zeros(nzc, bool)
    sc[0] = sc[nz // 2] = True
    gy[ny // 2, sc] = 0.0   # ky Nyquist on the self-conjugate kz planes
    gz[:, nz // 2] = 0.0    # kz Nyquist
    buf = cx(3, xl, ny, nzc)
    phi = buf[1].astype(np.complex128)
    d = A.prepare(buf.copy(), "c64")
    ops._call("mcpm_yz_gradients", A.stream(), A.ptr(d), xl, ny, nz, grad_fd, 0)
    out = to_numpy(d)
    assert np.array_equal(out[0], buf[0])
    assert rel(out[1], -1j * gy * phi) < 2e-6 and rel(out[2], -1j * gz * phi) < 2e-6
    cot = cx(3, xl, ny, nzc)
    d2 = A.prepare(cot.copy(), "c64")
    ops._call("mcpm_yz_gradients", A.stream(), A.ptr(d2), xl, ny, nz, grad_fd, 1)
    comb = to_numpy(d2)
    assert np.array_equal(comb[0], cot[0])
    assert rel(comb[1], gy * cot[1].astype(np.complex128) + gz * cot[2].astype(np.complex128)) < 2e-6
    lhs = np.vdot(out[1].astype(np.complex128), cot[1]) + np.vdot(out[2].astype(np.complex128), cot[2])
    rhs = np.vdot(phi, 1j * comb[1].astype(np.complex128))
    assert abs(lhs - rhs) < 1e-5 * abs(lhs)


def test_observed_nufft_edge_cases(ops):
    """mcpm_nufft_obs / mcpm_radial_tables at their edges: an empty particle set paints a zero mesh; a descriptor that
    does nothing (no redshift-space term, no Alcock-Paczynski) reproduces mcpm_nufft (1e-6: atomics order); a light cone or ap_auto
    without its radius table, an unknown Alcock-Paczynski mode, a redshift-space term without velocities and a radius
    grid with dr = 0 come back as MCPM_EINVAL; distances beyond either end of the radius grid clamp to the end values."""
    from montecosmo_b200._capi import McpmError
    shape = (8, 6, 10)
    rng = np.random.default_rng(4)
    geo = dict(curved=True, cell=(50.0, 40.0, 30.0), origin=(100.0, -80.0, 900.0), los=(0.0, 0.0, 1.0))
    e = np.zeros((0, 3), np.float32)
    assert not to_numpy(ops.nufft_observed(e, e, dict(geo, rsd=True, gf=0.5), shape)).any()
    assert not to_numpy(ops.nufft_observed(e, None, dict(geo, rsd=True, gf=0.5), shape)).any()  # null arrays with np = 0
    pos = f32(rng.uniform(0, 1, size=(200, 3)) * np.asarray(shape))
    vel = f32(rng.normal(size=(200, 3)))
    w = f32(rng.uniform(0.5, 1.5, 200))
    plain = to_numpy(ops.nufft_paint(pos, shape, w))
    same = to_numpy(ops.nufft_observed(pos, None, dict(geo, rsd=False, ap=0), shape, w))
    assert rel(same, plain) < 1e-6  # the same deposits, summed by atomics in whatever order
    for bad in (dict(geo, rsd=True, lightcone=True, r0=1.0, dr=1.0),                  # light cone without tab_gf
                dict(geo, rsd=False, ap=1, r0=1.0, dr=1.0),                           # ap_auto without tab_ap
                dict(geo, rsd=False, ap=3),                                           # unknown mode
                dict(geo, rsd=False, ap=1, tab_ap=f32(np.zeros(16)), r0=1.0, dr=0.0)):  # no radius grid
        with pytest.raises(McpmError) as ei:
            ops.nufft_observed(pos, vel, bad, shape, w)
        assert ei.value.code == 1 and str(ei.value)
    with pytest.raises(McpmError):
        ops.nufft_observed(pos, None, dict(geo, rsd=True, gf=0.5), shape, w)         # rsd without velocities
    with pytest.raises(McpmError):
        ops.radial_tables(pos, dict(geo, r0=1.0, dr=0.0), f32(np.zeros((2, 8))))
    tabs = f32(np.stack([np.linspace(1.0, 2.0, 8), np.linspace(-3.0, 4.0, 8)]))
    near_far = f32([[-2.0, 2.0, -30.0], [0.0, 0.0, 1e6]])                             # r ~ 0 and r ~ 3e7 Mpc/h
    out = to_numpy(ops.radial_tables(near_far, dict(geo, r0=500.0, dr=100.0), tabs))
    assert np.allclose(out, [[1.0, -3.0], [2.0, 4.0]])


@pytest.mark.parametrize("order", [2, 3, 4])
def test_paint_read_vjp(ops, order):
    """Gathers of the window gradient vs oracle autograd (1e-5)."""
    rng = np.random.default_rng(20 + order)
    shape = (8, 10, 6)
    pos = f32(rng.uniform(-3, 12, (600, 3)))
    w = f32(rng.uniform(0.5, 1.5, 600))
    mbar = f32(rng.normal(size=shape))
    scale, shift = (1.25, 1.0, 0.75), 0.5
    p = T(pos).requires_grad_()
    wt = T(w).requires_grad_()
    xs = p * T(f32(scale)) + shift
    (O.paint(xs, shape, wt * 0.7, order) * T(mbar)).sum().backward()
    pb, wb = ops.paint_vjp(pos, mbar, w, 0.7, order, scale, shift)
    assert rel(pb, p.grad.numpy()) < 1e-5
    assert rel(wb, wt.grad.numpy()) < 1e-5
    # read_grad with a cotangent over 3 meshes
    m3 = f32(rng.normal(size=(3, *shape)))
    cot = f32(rng.normal(size=(600, 3)))
    p = T(pos).requires_grad_()
    sum((O.read(p, T(m3[i]), order) * T(cot[:, i])).sum() for i in range(3)).backward()
    assert rel(ops.read_grad(pos, m3, cot, order), p.grad.numpy()) < 1e-5
    # paint3 = transpose of the 3-mesh read
    ref = np.stack([O.paint(T(pos), shape, T(cot[:, i]), order).numpy() for i in range(3)])
    assert rel(ops.paint3(pos, cot, shape, 1.0, order), ref) < 2e-6


# ------------------------------------------------------------------------------------------------ Kaiser-Bessel window
@pytest.mark.parametrize("order,oversamp", [(1, 1.0), (2, 2.0), (3, 1.25), (4, 1.5)])
def test_kaiser_bessel_paint_read_golden(ops, golden, order, oversamp):
    """kernel_type='kaiser_bessel' (nbody.py:280-290, 383-384, 415-416) vs the reference source: 2e-5 as the rectangular
    family (float32 positions up to +-40 cells; I0 in float32)."""
    g = golden("paint_read")
    shape = tuple(int(s) for s in g["shape"])
    kc = float(O.optim_kcut(oversamp))
    assert rel(ops.paint(f32(g["pos"]), shape, f32(g["weights"]), order=order, kb_kcut=kc), g[f"paint_kb_{order}"]) < 2e-5
    assert rel(ops.read(f32(g["pos"]), f32(g["mesh"]), order=order, kb_kcut=kc), g[f"read_kb_{order}"]) < 2e-5


@pytest.mark.parametrize("order", [2, 3, 4])
def test_kaiser_bessel_vjp_and_transform(ops, order):
    """Same float32 inputs on both sides: paint / read 5e-6; the window-gradient gathers vs oracle autograd (I1 through
    torch.special.i0) 2e-5; in-kernel position transform; read = paint^T."""
    rng = np.random.default_rng(40 + order)
    shape = (8, 10, 6)
    pos = f32(rng.uniform(-3, 12, (600, 3)))
    w = f32(rng.uniform(0.5, 1.5, 600))
    mesh = f32(rng.normal(size=shape))
    mbar = f32(rng.normal(size=shape))
    ov = 1.5
    kc = float(O.optim_kcut(ov))
    scale, shift = (1.25, 1.0, 0.75), 0.5
    assert rel(ops.paint(pos, shape, w, order=order, kb_kcut=kc),
               O.paint(T(pos), shape, T(w), order, "kaiser_bessel", ov).numpy()) < 5e-6
    assert rel(ops.read(pos, mesh, order=order, kb_kcut=kc),
               O.read(T(pos), T(mesh), order, "kaiser_bessel", ov).numpy()) < 5e-6
    p = T(pos).requires_grad_()
    wt = T(w).requires_grad_()
    xs = p * T(f32(scale)) + shift
    (O.paint(xs, shape, wt * 0.7, order, "kaiser_bessel", ov) * T(mbar)).sum().backward()
    pb, wb = ops.paint_vjp(pos, mbar, w, 0.7, order, scale, shift, kb_kcut=kc)
    assert rel(pb, p.grad.numpy()) < 2e-5
    assert rel(wb, wt.grad.numpy()) < 2e-5
    m3 = f32(rng.normal(size=(3, *shape)))
    cot = f32(rng.normal(size=(600, 3)))
    p = T(pos).requires_grad_()
    sum((O.read(p, T(m3[i]), order, "kaiser_bessel", ov) * T(cot[:, i])).sum() for i in range(3)).backward()
    assert rel(ops.read_grad(pos, m3, cot, order, kb_kcut=kc), p.grad.numpy()) < 2e-5
    painted = to_numpy(ops.paint(pos, shape, w, order=order, kb_kcut=kc)).astype(np.float64)
    lhs = float((to_numpy(ops.read(pos, mesh, order=order, kb_kcut=kc)).astype(np.float64) * w).sum())
    rhs = float((mesh.astype(np.float64) * painted).sum())
    assert abs(lhs - rhs) < 1e-4 * max(abs(lhs), 1.0)


def test_kaiser_bessel_deconv_and_nufft_golden(ops, golden):
    """deconv_paint / interlace / nufft with the Kaiser-Bessel window (nbody.py:293-312, 321-322, 532-577) vs the
    reference source, and the VJP of the fused nufft vs oracle autograd."""
    g = golden("nufft")
    final = tuple(int(s) for s in g["final_shape"])
    pos, w = f32(g["pos"]), f32(g["weights"])
    cm = np.fft.rfftn(g["deconv_real_in"])
    assert rel(ops.deconv(c64(cm), 2, kb_kcut=float(O.optim_kcut(2.0))), g["deconv_kb_cplx_2"]) < 5e-6
    out = ops.irfftn(ops.deconv(ops.rfftn(f32(g["deconv_real_in"])), 4, kb_kcut=float(O.optim_kcut(1.5))))
    assert rel(out, g["deconv_kb_real_4"]) < 5e-6
    kc = float(O.optim_kcut(1.5))
    assert rel(ops.nufft_paint(pos, final, w, paint_order=4, paint_deconv=False, kb_kcut=kc), g["interlace_kb_4_2"]) < 2e-5
    ps = O.scale_shape(final, 1.5)
    sc = tuple(np.divide(ps, final))
    out = ops.chreshape(ops.nufft_paint(pos, ps, w, scale=sc, paint_order=4, kb_kcut=kc), O.r2chshape(final))
    assert rel(out, g["nufft_kb_over15"]) < 2e-5
    ps = (12, 10, 14)
    ov = float(np.exp(np.log(np.divide(final, ps)).mean()))  # nbody.py:566
    kc2 = float(O.optim_kcut(ov))
    out = ops.chreshape(ops.nufft_paint(pos, ps, w, scale=tuple(np.divide(ps, final)), kb_kcut=kc2), O.r2chshape(final))
    assert rel(out, g["nufft_kb_tuple_o2"]) < 2e-5
    # VJP
    rng = np.random.default_rng(13)
    sc = tuple(np.divide(ps, final))
    cs = O.r2chshape(ps)
    obar = c64(rng.normal(size=cs) + 1j * rng.normal(size=cs))
    p, wt = T(pos).requires_grad_(), T(w).requires_grad_()
    o = O.interlace(p * T(f32(sc)), ps, wt * 0.8, 4, 2, "kaiser_bessel", 1.5) * float(np.prod(sc))
    o = O.deconv_paint(o, 4, "kaiser_bessel", 1.5)
    assert rel(ops.nufft_paint(pos, ps, w, 0.8, sc, 4, 2, True, kb_kcut=kc), o.detach().numpy()) < 2e-5
    ob = T(obar, torch.complex128)
    (o.real * ob.real + o.imag * ob.imag).sum().backward()
    pb, wb = ops.nufft_paint_vjp(pos, obar, ps, w, 0.8, sc, 4, 2, True, kb_kcut=kc)
    assert rel(pb, p.grad.numpy()) < 5e-5
    assert rel(wb, wt.grad.numpy()) < 5e-5


# ------------------------------------------------------------------------------------------------ Fourier passes
def test_fft_roundtrip(ops):
    rng = np.random.default_rng(3)
    m = f32(rng.normal(size=(3, 8, 6, 10)))
    k = ops.rfftn(m)
    assert rel(k, np.fft.rfftn(m.astype(np.float64), axes=(1, 2, 3))) < 1e-6
    assert rel(ops.irfftn(k), m) < 1e-6
    assert rel(ops.irfftn(ops.rfftn(m[0])), m[0]) < 1e-6


@pytest.mark.parametrize("shape", [(8, 6, 10), (12, 12, 4), (32, 32, 32)])
def test_irfftn_of_non_hermitian_input(ops, shape):
    """jnp.fft.irfftn accepts any complex array and returns the real part of the full inverse transform.  cuFFT's C2R
    does not (inconsistent input is algorithm dependent), so mcpm_irfftn projects first.  The projection alone equals
    rfftn(irfftn(.)).  The 256^3 case, where cuFFT takes another algorithm, is in tests/test_zz_cross_check_256.py."""
    rng = np.random.default_rng(sum(shape))
    cs = O.r2chshape(shape)
    a = c64(rng.normal(size=(2, *cs)) + 1j * rng.normal(size=(2, *cs)))
    ref = np.fft.irfftn(a.astype(np.complex128), s=shape, axes=(1, 2, 3))
    assert rel(ops.irfftn(a), ref) < 5e-6
    proj = to_numpy(ops.hermitian_project(ops.A.prepare(a.copy(), "c64")))
    assert rel(proj, np.fft.rfftn(ref, axes=(1, 2, 3))) < 5e-6
    # a Hermitian input is left alone (up to rounding)
    h = c64(np.fft.rfftn(rng.normal(size=shape)))
    assert rel(ops.hermitian_project(ops.A.prepare(h.copy(), "c64")), h.astype(np.complex128)) < 1e-7


def test_fourier_kernels_golden(ops, golden):
    """force / Hessian / deconvolution kernels vs the reference's kernel arrays on a non-cubic mesh (2e-6).

    Force and Hessian spectra are compared AFTER irfftn, as the reference uses them (nbody.py:603, 620, 627): the
    engine hands cuFFT the Hermitian projection of these spectra (fourier.cu: KVec), which differs from the raw product
    only by modes that numpy / XLA's irfftn discards anyway."""
    inv = lambda k: np.fft.irfftn(np.asarray(to_numpy(k), dtype=np.complex128), s=shape, axes=(-3, -2, -1))
    g = golden("kernels")
    shape = tuple(int(s) for s in g["shape"])
    rng = np.random.default_rng(5)
    dk = c64(np.fft.rfftn(rng.normal(size=shape)))
    for lap, grad, kcut, dec in [(INF, INF, INF, 0), (2, 4, INF, 0), (4, 2, 2.0, 0), (INF, INF, INF, 2)]:
        tag = lambda fd: "inf" if fd == INF else str(fd)
        base = g[f"invlaplace_{tag(lap)}"] * (g["gaussian_kcut2"] if kcut != INF else 1.0)
        if dec:
            base = base / g[f"rectangular_hat_{dec}"] ** 2
        ref = np.stack([-g[f"gradient{i}_{tag(grad)}"] * base * dk for i in range(3)])
        assert rel(inv(ops.force_spectra(dk, lap, grad, kcut, dec)), inv(ref)) < 2e-6
    ref = np.stack([g[f"gradient{i}_inf"] * g[f"gradient{j}_inf"] * g["invlaplace_inf"] * dk
                    for i, j in [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]])
    assert rel(inv(ops.hessian_spectra(dk)), inv(ref)) < 2e-6
    for o in (1, 2, 3, 4):
        assert rel(ops.deconv(dk, o), dk / g[f"rectangular_hat_{o}"]) < 2e-6


def test_fourier_transposes(ops):
    """<K x, y> = <x, K^T y> for the force / Hessian / interlace passes, in the real inner product Re sum conj(a) b."""
    rng = np.random.default_rng(6)
    shape = (8, 6, 10)
    cs = O.r2chshape(shape)
    cplx = lambda *s: c64(rng.normal(size=s) + 1j * rng.normal(size=s))
    dot = lambda a, b: float(np.real(np.vdot(to_numpy(a).astype(np.complex128), to_numpy(b).astype(np.complex128))))
    x, y3, y6 = cplx(*cs), cplx(3, *cs), cplx(6, *cs)
    for kw in [dict(), dict(lap_fd=2, grad_fd=4, kcut=2.0, deconv_order=2)]:
        assert abs(dot(ops.force_spectra(x, **kw), y3) - dot(x, ops.force_spectra_T(y3, **kw))) < 1e-4 * abs(
            dot(x, x))
    assert abs(dot(ops.hessian_spectra(x), y6) - dot(x, ops.hessian_spectra_T(y6))) < 1e-4 * abs(dot(x, x))
    # half_weights variant = (w'/N) * plain transpose
    wq = np.full(cs, 2.0)
    wq[..., 0] = 1.0
    wq[..., -1] = 1.0
    # (the plain transpose feeds a C2R and is Hermitian-projected on the self-conjugate planes, the half_weights variant
    #  is the cotangent of a free complex array and is not: compare away from the Nyquist modes of those planes)
    free = np.ones(cs, bool)
    for pl in (0, cs[2] - 1):
        free[shape[0] // 2, :, pl] = False
        free[:, shape[1] // 2, pl] = False
    free[:, :, cs[2] - 1] = False
    a = to_numpy(ops.force_spectra_T(y3, half_weights=True))
    b = to_numpy(ops.force_spectra_T(y3)) * wq / np.prod(shape)
    assert rel(a[free], b[free]) < 1e-6
    # interlace: K^T carries N / w'
    xm = cplx(2, *cs)
    lhs = dot(ops.interlace_combine(xm, 1.7, 2), x)
    rhs = dot(xm, to_numpy(ops.interlace_combine_T(x, 2, 1.7, 2)) * wq / np.prod(shape))
    assert abs(lhs - rhs) < 1e-4 * abs(dot(x, x))


@pytest.mark.parametrize("tag", ["down", "up", "mixed", "same"])
def test_chreshape_golden(ops, golden, tag):
    g = golden("chreshape")
    dst = tuple(int(s) for s in g[f"{tag}_dst"])
    assert rel(ops.chreshape(c64(g[f"{tag}_in"]), O.r2chshape(dst)), g[f"{tag}_out"]) < 2e-6


def test_lpt2_source(ops):
    rng = np.random.default_rng(7)
    h = f32(rng.normal(size=(6, 6, 6, 6)))
    ht = T(h).requires_grad_()
    d2 = ht[0] * ht[1] + ht[0] * ht[2] + ht[1] * ht[2] - ht[3] ** 2 - ht[4] ** 2 - ht[5] ** 2
    assert rel(ops.lpt2_source(h), d2.detach().numpy()) < 1e-6
    bar = f32(rng.normal(size=(6, 6, 6)))
    (d2 * T(bar)).sum().backward()
    assert rel(ops.lpt2_source_vjp(h, bar), ht.grad.numpy()) < 1e-6


# ------------------------------------------------------------------------------------------------ forces / LPT
@pytest.mark.parametrize("name", ["forces_lpt", "forces_noncubic"])
def test_pm_forces_golden(ops, golden, name):
    """vs the reference source; 5e-5 (float32 FFT + assignment against float64)."""
    g = golden(name)
    shape = tuple(int(s) for s in g["shape"])
    pos, dk = f32(g["pos"]), c64(g["delta_k"])
    assert rel(ops.pm_forces(pos, shape, 2), g["pm_forces_paint"]) < 5e-5
    assert rel(ops.pm_forces_mesh(pos, dk, 2), g["pm_forces_mesh"]) < 5e-5
    assert rel(ops.pm_forces2(pos, dk, 2), g["pm_forces2"]) < 5e-5
    if name == "forces_lpt":
        assert rel(ops.pm_forces(pos, shape, 2, paint_deconv=True, kcut=2.5), g["pm_forces_paint_deconv_kcut"]) < 5e-5
        assert rel(ops.pm_forces(pos, shape, 3, lap_fd=2, grad_fd=4), g["pm_forces_paint_o3_fd"]) < 5e-5
        q = f32(O.regular_pos(shape).numpy())
        assert rel(ops.pm_forces_mesh(q, dk, 1), g["pm_forces_mesh_ngp_lattice"]) < 5e-5


def test_lpt_golden(ops, golden):
    g = golden("forces_lpt")
    shape = tuple(int(s) for s in g["shape"])
    dk = c64(g["delta_k"])
    q = f32(O.regular_pos(shape).numpy())
    c = O.Cosmology()
    co = lambda a: (float(O.a2g(c, a)), float(O.a2g2(c, a)), float(O.a2dg2dg(c, a)))
    for order in (1, 2):
        dp, vl = ops.lpt(dk, q, *co(0.3), lpt_order=order, read_order=1)
        assert rel(dp, g[f"lpt{order}_a0.3_dpos"]) < 5e-5
        assert rel(vl, g[f"lpt{order}_a0.3_vel"]) < 5e-5
    dp, vl = ops.lpt(dk, f32(g["pos"]), *co(0.0), lpt_order=2, read_order=2)
    assert rel(dp, g["lpt2_a0_cic_dpos"]) < 5e-5
    assert rel(vl, g["lpt2_a0_cic_vel"]) < 5e-5


def test_pm_forces_vjp(ops):
    """VJP of pm_forces(pos, shape) w.r.t. pos vs oracle autograd (2e-4)."""
    rng = np.random.default_rng(8)
    shape = (8, 10, 12)
    pos = f32(O.regular_pos(shape).numpy() + rng.normal(scale=0.6, size=(np.prod(shape), 3)))
    fbar = f32(rng.normal(size=pos.shape))
    for kw in [dict(order=2), dict(order=3, paint_deconv=True, lap_fd=2, grad_fd=4)]:
        okw = dict(read_order=kw["order"], paint_deconv=kw.get("paint_deconv", False), lap_fd=kw.get("lap_fd", np.inf),
                   grad_fd=kw.get("grad_fd", np.inf))
        p = T(pos).requires_grad_()
        (O.pm_forces(p, shape, **okw) * T(fbar)).sum().backward()
        forces, fm = ops.pm_forces(pos, shape, want_meshes=True, **kw)
        assert rel(forces, O.pm_forces(T(pos), shape, **okw).numpy()) < 5e-5
        assert rel(ops.pm_forces_vjp(pos, fbar, fm, **kw), p.grad.numpy()) < 2e-4


@pytest.mark.parametrize("lpt_order,read_order", [(1, 1), (2, 1), (2, 2)])
def test_lpt_vjp(ops, lpt_order, read_order):
    """VJP of lpt w.r.t. delta_k (torch complex convention) and the growth coefficients vs oracle autograd (2e-4)."""
    rng = np.random.default_rng(9)
    shape = (8, 6, 10)
    dk = c64(np.fft.rfftn(rng.normal(size=shape)) * 0.02)
    q = O.regular_pos(shape).numpy()
    pos = f32(q if read_order == 1 else q + rng.normal(scale=0.4, size=q.shape))
    d1, d2, dv2 = 0.61, -0.17, -0.52
    dpb, vlb = f32(rng.normal(size=pos.shape)), f32(rng.normal(size=pos.shape))

    class Cst:  # oracle cosmology stub returning the chosen coefficients as differentiable scalars
        pass
    coef = T([d1, d2, dv2]).requires_grad_()
    dkt = T(dk, torch.complex128).requires_grad_()
    f1 = O.pm_forces(T(pos), dkt, read_order)
    dpo, vlo = coef[0] * f1, f1
    if lpt_order == 2:
        f2 = O.pm_forces2(T(pos), dkt, read_order)
        dpo, vlo = dpo - coef[1] * f2, vlo - coef[2] * f2
    ((dpo * T(dpb)).sum() + (vlo * T(vlb)).sum()).backward()
    dp, vl, tape = ops.lpt(dk, pos, d1, d2, dv2, lpt_order=lpt_order, read_order=read_order, tape=True)
    assert rel(dp, dpo.detach().numpy()) < 5e-5 and rel(vl, vlo.detach().numpy()) < 5e-5
    dkbar, cb = ops.lpt_vjp(pos, dk.shape, d1, d2, dv2, dpb, vlb, tape, lpt_order=lpt_order, read_order=read_order,
                            want_coef=True)
    assert rel(dkbar, dkt.grad.numpy()) < 2e-4
    ncoef = 1 if lpt_order == 1 else 3
    assert rel(to_numpy(cb)[:ncoef], coef.grad.numpy()[:ncoef]) < 2e-4


# ------------------------------------------------------------------------------------------------ BullFrog loop
def _bf_coeffs(c, a0, a1, n):
    g0, g1 = float(O.a2g(c, a0)), float(O.a2g(c, a1))
    dg = (g1 - g0) / n
    alpha = [float(O.alpha_bf(c, g0 + s * dg, dg)) for s in range(n)]
    beta = [(1 - al) / (g0 + s * dg + dg / 2) for s, al in enumerate(alpha)]
    return alpha, beta, [dg / 2] * n, [dg / 2] * n


def test_nbody_golden(ops, golden):
    """lpt + 4 BullFrog steps at 16^3 vs the reference source: max |dx| < 2e-4 cell, velocities 2e-4 relative."""
    g = golden("nbody")
    shape = tuple(int(s) for s in g["shape"])
    dk = c64(g["delta_k"])
    q = f32(O.regular_pos(shape).numpy())
    c = O.Cosmology()
    dp, vel = ops.lpt(dk, q, float(O.a2g(c, 0.0)), float(O.a2g2(c, 0.0)), float(O.a2dg2dg(c, 0.0)), 2, 1)
    pos = to_numpy(dp) + q
    ab = _bf_coeffs(c, 0.0, 1.0, 4)
    assert rel(np.array(ab[0]), g["bf4_alpha"]) < 1e-9
    pos, vel = ops.A.prepare(pos), ops.A.prepare(to_numpy(vel))
    ops.nbody_steps(pos, vel, shape, *ab)
    assert np.abs(to_numpy(pos) - g["bf4_pos"][0]).max() < 2e-4
    assert rel(vel, g["bf4_vel"][0]) < 2e-4


def test_nbody_steps_vjp(ops):
    """Reverse sweep of the BullFrog loop vs oracle autograd through the same coefficients (5e-4), incl. coefficients."""
    rng = np.random.default_rng(11)
    shape = (8, 8, 8)
    n = 3
    q = O.regular_pos(shape).numpy()
    pos0 = f32(q + rng.normal(scale=0.5, size=q.shape))
    vel0 = f32(rng.normal(scale=0.5, size=q.shape))
    alpha, beta = [0.6, 0.8, 0.9], [0.9, 0.5, 0.3]
    pre, post = [0.11, 0.13, 0.17], [0.12, 0.14, 0.16]
    pb, vb = f32(rng.normal(size=q.shape)), f32(rng.normal(size=q.shape))
    # oracle
    co = T(np.array([alpha, beta, pre, post]).T.copy()).requires_grad_()
    x, v = T(pos0).requires_grad_(), T(vel0).requires_grad_()
    xs, vs = x, v
    for s in range(n):
        xs = xs + vs * co[s, 2]
        vs = co[s, 0] * vs + co[s, 1] * O.pm_forces(xs, shape, 2)
        xs = xs + vs * co[s, 3]
    ((xs * T(pb)).sum() + (vs * T(vb)).sum()).backward()
    # engine
    A = ops.A
    pos, vel = A.prepare(pos0.copy()), A.prepare(vel0.copy())
    tape = ops.nbody_steps(pos, vel, shape, alpha, beta, pre, post, tape=True, tape_vel=True)
    assert np.abs(to_numpy(pos) - xs.detach().numpy()).max() < 1e-4
    assert rel(vel, vs.detach().numpy()) < 1e-4
    posbar, velbar = A.prepare(pb.copy()), A.prepare(vb.copy())
    coef = ops.nbody_steps_vjp(posbar, velbar, shape, alpha, beta, pre, post, tape, v0=A.prepare(vel0),
                               want_coef=True)
    assert rel(posbar, x.grad.numpy()) < 5e-4
    assert rel(velbar, v.grad.numpy()) < 5e-4
    assert rel(coef, co.grad.numpy()) < 5e-4
    # without the velocity tape (no coefficient cotangents): same state cotangents
    pos, vel = A.prepare(pos0.copy()), A.prepare(vel0.copy())
    tape = ops.nbody_steps(pos, vel, shape, alpha, beta, pre, post, tape=True)
    posbar, velbar = A.prepare(pb.copy()), A.prepare(vb.copy())
    ops.nbody_steps_vjp(posbar, velbar, shape, alpha, beta, pre, post, tape)
    assert rel(posbar, x.grad.numpy()) < 5e-4


# ------------------------------------------------------------------------------------------------ NUFFT
def test_nufft_golden(ops, golden):
    g = golden("nufft")
    final = tuple(int(s) for s in g["final_shape"])
    pos, w = f32(g["pos"]), f32(g["weights"])
    assert rel(ops.nufft_paint(pos, final, w, paint_deconv=False), g["interlace_2_2"]) < 2e-5
    assert rel(ops.nufft_paint(pos, final, w, paint_order=4, interlace_order=3, paint_deconv=False),
               g["interlace_4_3"]) < 2e-5
    assert rel(ops.nufft_paint(pos, final, w), g["nufft_same"]) < 2e-5
    ps = O.scale_shape(final, 1.5)
    sc = tuple(np.divide(ps, final))
    out = ops.chreshape(ops.nufft_paint(pos, ps, w, scale=sc), O.r2chshape(final))
    assert rel(out, g["nufft_over15"]) < 2e-5
    out = ops.chreshape(ops.nufft_paint(pos, ps, None, scale=sc, paint_order=3, paint_deconv=False), O.r2chshape(final))
    assert rel(out, g["nufft_over15_nodeconv_o3"]) < 2e-5
    ps = (12, 10, 14)
    out = ops.chreshape(ops.nufft_paint(pos, ps, w, scale=tuple(np.divide(ps, final))), O.r2chshape(final))
    assert rel(out, g["nufft_tuple"]) < 2e-5


def test_nufft_vjp(ops):
    rng = np.random.default_rng(12)
    final, ps = (8, 8, 8), (12, 10, 12)
    sc = tuple(np.divide(ps, final))
    pos = f32(rng.uniform(-2, 10, (500, 3)))
    w = f32(rng.uniform(0.5, 1.5, 500))
    cs = O.r2chshape(ps)
    obar = c64(rng.normal(size=cs) + 1j * rng.normal(size=cs))
    p, wt = T(pos).requires_grad_(), T(w).requires_grad_()
    xs = p * T(f32(sc))
    out = O.interlace(xs, ps, wt * 0.8, 2, 2) * float(np.prod(sc))
    out = O.deconv_paint(out, 2)
    assert rel(ops.nufft_paint(pos, ps, w, 0.8, sc), out.detach().numpy()) < 2e-5
    ob = T(obar, torch.complex128)
    (out.real * ob.real + out.imag * ob.imag).sum().backward()
    pb, wb = ops.nufft_paint_vjp(pos, obar, ps, w, 0.8, sc)
    assert rel(pb, p.grad.numpy()) < 5e-5
    assert rel(wb, wt.grad.numpy()) < 5e-5


# ------------------------------------------------------------------------------------------------ brick-tiled scatter
@pytest.mark.parametrize("mesh", [(32, 24, 96), (36, 22, 100)])
def test_brick_scatter_matches_generic(ops, mesh):
    """With the lattice hint (mcpm_engine_set_lattice) the CIC density paint of pm_forces and the reverse-step scatter
    of nbody_steps_vjp run brick-tiled (shared-memory tile, fixed-point accumulation, red.v4 flush).  Results must agree
    with the generic kernels to the fixed-point quantum (2e-5 rel L2; 5e-5 vs the float64 oracle), including stray
    particles far from their brick and lattices that do not divide into bricks.  On the CPU port the hint is ignored."""
    rng = np.random.default_rng(21)
    n = int(np.prod(mesh))
    q = O.regular_pos(mesh).numpy()
    k = [2 * np.pi / m for m in mesh]
    disp = np.stack([1.5 * np.sin(k[1] * q[:, 1]) + 1.0 * np.cos(2 * k[2] * q[:, 2]),
                     1.2 * np.sin(2 * k[2] * q[:, 2]) + 0.8 * np.cos(k[0] * q[:, 0]),
                     2.0 * np.sin(k[0] * q[:, 0]) + 1.5 * np.cos(3 * k[1] * q[:, 1])], -1)
    pos = q + disp + rng.normal(scale=0.4, size=q.shape)
    stray = rng.choice(n, n // 40, replace=False)
    pos[stray] += rng.uniform(-30, 30, (len(stray), 3))
    pos[::7] -= np.array(mesh) * 2.0  # some positions wrapped by the caller
    pos = f32(pos)
    vel = f32(rng.normal(size=q.shape) * np.exp(rng.normal(size=(n, 1))))  # heavy-tailed cotangent
    alpha, beta, pre, post = [0.7, 0.9], [0.6, 0.3], [0.05, 0.05], [0.05, 0.05]
    res = {}
    A = ops.A
    for hint in (None, mesh):
        ops.set_lattice(mesh, hint)
        f, fm = ops.pm_forces(pos, mesh, 2, want_meshes=True)
        p2, v2 = A.prepare(pos.copy()), A.prepare(vel.copy())
        tape = ops.nbody_steps(p2, v2, mesh, alpha, beta, pre, post, 2, tape=True)
        pb, vb = A.prepare(vel.copy()), A.prepare(pos.copy() * 0.01)
        ops.nbody_steps_vjp(pb, vb, mesh, alpha, beta, pre, post, tape, 2)
        res[hint is None] = [to_numpy(x).copy() for x in (f, fm, p2, v2, pb, vb)]
    ops.set_lattice(mesh, None)
    for name, a, b in zip(["forces", "force meshes", "pos", "vel", "posbar", "velbar"], res[False], res[True]):
        assert rel(a, b) < 2e-5, name
    assert rel(res[False][0], O.pm_forces(T(pos), mesh, 2).numpy()) < 5e-5
    # the interlaced, weighted paints of nufft take the same kernel (pos + shift, per-particle weights)
    w = f32(rng.uniform(0.2, 3.0, n))
    nu = {}
    for hint in (None, mesh):
        ops.set_lattice(mesh, hint)
        nu[hint is None] = to_numpy(ops.nufft_paint(pos, mesh, w, 0.7, None, 2, 2, True)).copy()
    ops.set_lattice(mesh, None)
    assert rel(nu[True], nu[False]) < 2e-5
    if not isinstance(A, NumpyAdapter):
        # CUDA build: these shapes must really be handled by the brick kernels (MCPM_EUNSUP raises otherwise)
        ops.set_lattice(mesh, mesh)
        eng = ops.engine(mesh)
        pd, wd = A.prepare(pos), A.prepare(w)
        out = A.zeros(mesh)
        ops._call("mcpm_paint_lattice", eng.handle, A.stream(), A.ptr(pd), A.ptr(wd), 0.5, n, A.ptr(out))
        assert rel(out, O.paint(T(pos), mesh, T(w) * 0.5, 2).numpy()) < 2e-5
        vb, xb = A.prepare(vel.copy()), A.prepare(f32(rng.normal(size=q.shape)))
        out3 = A.zeros((3, *mesh))
        ops._call("mcpm_paint3_lattice", eng.handle, A.stream(), A.ptr(pd), A.ptr(vb), A.ptr(xb), 0.25, 1.5, n,
                  A.ptr(out3))
        vnew = vel.astype(np.float64) + 0.25 * to_numpy(xb).astype(np.float64)
        assert rel(vb, vnew) < 1e-6
        for c in range(3):
            assert rel(out3[c], O.paint(T(pos), mesh, T(vnew[:, c]) * 1.5, 2).numpy()) < 2e-5
        ops.set_lattice(mesh, None)


def test_paint_read_random_configurations(ops):
    """Seeded sweep over mesh shapes (non-cubic, down to the smallest legal side), orders, both window families, position
    ranges and in-kernel transforms: paint and read against the oracle on identical float32 inputs (5e-6; 5e-5 where
    positions reach +-50 box lengths, whose float32 spacing after the in-kernel transform is ~3e-5 cell), and
    <read(x, m), w> = <m, paint(x, w)> for each."""
    rng = np.random.default_rng(2024)
    for case in range(40):
        order = int(rng.integers(1, 5))
        shape = tuple(int(s) for s in rng.integers(max(order, 2), 14, 3))
        n = int(rng.integers(1, 400))
        span = float(rng.choice([1.0, 3.0, 50.0]))
        pos = f32(rng.uniform(-span, span, (n, 3)) * np.array(shape))
        if case % 4 == 0:
            pos[: n // 2] = np.round(pos[: n // 2] * 2) / 2  # on cell centres / faces
        w = f32(rng.normal(size=n))
        mesh = f32(rng.normal(size=shape))
        kb = bool(rng.integers(0, 2))
        ov = float(rng.choice([1.0, 1.5, 2.0]))
        scale, shift = ((1.0, 1.0, 1.0), 0.0) if case % 3 else (tuple(rng.uniform(0.5, 2.0, 3)), float(rng.uniform(0, 1)))
        if (order == 1 or kb) and span > 3:
            # windows that jump at the edge of their support (NGP, Kaiser-Bessel): a float32-rounded transform of a far
            # position may pick the other base cell than the float64 one; without a transform both sides see the same x
            scale, shift = (1.0, 1.0, 1.0), 0.0
        kt, kc = ("kaiser_bessel", float(O.optim_kcut(ov))) if kb else ("rectangular", 0.0)
        xs = T(pos) * T(f32(scale)) + shift
        p = ops.paint(pos, shape, w, 1.0, order, scale, shift, kb_kcut=kc)
        r = ops.read(pos, mesh, order, scale, shift, kb_kcut=kc)
        tag = (case, order, shape, n, kt)
        tol = 5e-5 if span > 3 else 5e-6
        assert rel(p, O.paint(xs, shape, T(w), order, kt, ov).numpy()) < tol, tag
        assert rel(r, O.read(xs, T(mesh), order, kt, ov).numpy()) < tol, tag
        lhs = float((to_numpy(r).astype(np.float64) * w).sum())
        rhs = float((mesh.astype(np.float64) * to_numpy(p).astype(np.float64)).sum())
        assert abs(lhs - rhs) < 1e-4 * max(abs(lhs), abs(rhs), 1.0), tag


def test_absmax_strided(ops):
    """mcpm_absmax (the halo guard of a slab-decomposed caller): running maximum of |x| over a strided column, exact;
    accumulates across calls; a NaN sticks; an empty input leaves the value alone."""
    A = ops.A
    rng = np.random.default_rng(5)
    for n in (1, 31, 4096, 70001):
        pos = f32(rng.normal(scale=3.0, size=(n, 3)))
        pos[rng.integers(0, n), 0] = -17.25
        out = A.zeros((1,))
        pd = A.prepare(pos)
        ops._call("mcpm_absmax", A.stream(), A.ptr(pd), n, 3, A.ptr(out))
        assert float(to_numpy(out)[0]) == float(np.abs(pos[:, 0]).max())
        small = A.prepare(f32(np.full((5, 3), 0.5)))
        ops._call("mcpm_absmax", A.stream(), A.ptr(small), 5, 3, A.ptr(out))   # a smaller batch does not lower it
        ops._call("mcpm_absmax", A.stream(), A.ptr(small), 0, 3, A.ptr(out))   # nor does an empty one
        assert float(to_numpy(out)[0]) == float(np.abs(pos[:, 0]).max())
    bad = f32(rng.normal(size=(100, 3)))
    bad[7, 0] = np.nan
    out = A.zeros((1,))
    bd = A.prepare(bad)
    ops._call("mcpm_absmax", A.stream(), A.ptr(bd), 100, 3, A.ptr(out))
    v = float(to_numpy(out)[0])
    assert not v <= 1e30  # NaN (or above every finite bound): the caller's `not max <= limit` test raises
