"""
Fused x-transform path of pm_forces (montecosmo_b200/csrc/xfft.cu): 2-D cuFFT per x-plane + one kernel doing the x-FFT,
the force kernel of nbody.py:591-603 and the inverse x-FFTs.  Compared with the 3-D cuFFT + separate multiply path of
the same engine (float32 rounding only: 2e-5 relative L2) and with the float64 oracle (5e-5 forces, 2e-4 VJP -- the
tolerances of tests/test_abi_parity.py).  GPU only: the CPU port has no such path.
"""
import numpy as np
import pytest
import torch

from oracle import pm_oracle as O
from tests.backends import to_numpy

pytestmark = pytest.mark.gpu
INF = float("inf")


@pytest.fixture(scope="module")
def ops():
    from montecosmo_b200 import _lib
    from montecosmo_b200.ops import Ops, TorchCudaAdapter
    return Ops(_lib.load(), TorchCudaAdapter())


def rel(a, b):
    a, b = np.asarray(to_numpy(a), dtype=np.float64), np.asarray(to_numpy(b), dtype=np.float64)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def T(x):
    return torch.as_tensor(np.asarray(x), dtype=torch.float64)


# nx = 64 (8 x 8), 128 (16 x 8), 256 (16 x 16), 512 (32 x 16), 1024 (32 x 32) register transforms; ny * (nz/2+1) not a multiple of the 16-column tile
SHAPES = [(64, 12, 16), (128, 6, 10), (256, 4, 6), (64, 16, 30), (512, 4, 6), (1024, 4, 4)]
OPTS = [dict(order=2), dict(order=3, paint_deconv=True, lap_fd=2, grad_fd=4), dict(order=2, kcut=2.0, lap_fd=4)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("kw", OPTS)
def test_forces_and_vjp(ops, shape, kw):
    rng = np.random.default_rng(sum(shape))
    n = int(np.prod(shape))
    pos = (O.regular_pos(shape).numpy() + rng.normal(scale=0.7, size=(n, 3))).astype(np.float32)
    fbar = rng.normal(size=pos.shape).astype(np.float32)
    okw = dict(read_order=kw["order"], paint_deconv=kw.get("paint_deconv", False), lap_fd=kw.get("lap_fd", np.inf),
               grad_fd=kw.get("grad_fd", np.inf), kcut=kw.get("kcut", np.inf))
    res = {}
    for fused in (False, True):
        ops.set_fused_fft(shape, fused)
        f, fm = ops.pm_forces(pos, shape, want_meshes=True, **kw)
        g = ops.pm_forces_vjp(pos, fbar, fm, **kw)
        res[fused] = [to_numpy(x).copy() for x in (f, fm, g)]
    for name, a, b in zip(["forces", "force meshes", "vjp"], res[True], res[False]):
        assert rel(a, b) < 2e-5, name
    p = T(pos).requires_grad_()
    fo = O.pm_forces(p, shape, **okw)
    (fo * T(fbar)).sum().backward()
    # 1 / k^2 at the fundamental of a 1024-cell axis amplifies the float32 rounding of the painted density by 2.7e4:
    # both engine paths agree with each other (above) and sit 1.2e-4 from the float64 oracle on the longest meshes
    tol = 5e-5 if shape[0] <= 256 else 3e-4
    assert rel(res[True][0], fo.detach().numpy()) < tol
    assert rel(res[True][2], p.grad.numpy()) < 4 * tol


def test_nbody_steps_and_reverse_sweep(ops):
    shape = (64, 12, 16)
    rng = np.random.default_rng(5)
    n = int(np.prod(shape))
    pos = (O.regular_pos(shape).numpy() + rng.normal(scale=0.5, size=(n, 3))).astype(np.float32)
    vel = rng.normal(scale=0.3, size=(n, 3)).astype(np.float32)
    alpha, beta, pre, post = [0.7, 0.9, 0.95], [0.6, 0.3, 0.2], [0.05, 0.04, 0.03], [0.05, 0.04, 0.03]
    A = ops.A
    res = {}
    for fused in (False, True):
        ops.set_fused_fft(shape, fused)
        p2, v2 = A.prepare(pos.copy()), A.prepare(vel.copy())
        tape = ops.nbody_steps(p2, v2, shape, alpha, beta, pre, post, 2, tape=True)
        pb, vb = A.prepare(vel.copy()), A.prepare(pos.copy() * 0.01)
        ops.nbody_steps_vjp(pb, vb, shape, alpha, beta, pre, post, tape, 2)
        res[fused] = [to_numpy(x).copy() for x in (p2, v2, pb, vb)]
    ops.set_fused_fft(shape, True)
    for name, a, b in zip(["pos", "vel", "posbar", "velbar"], res[True], res[False]):
        assert rel(a, b) < 2e-5, name
    # and against the oracle's loop
    x, v = T(pos), T(vel)
    for s in range(3):
        x = x + v * pre[s]
        v = alpha[s] * v + beta[s] * O.pm_forces(x, shape, 2)
        x = x + v * post[s]
    assert np.abs(res[True][0] - x.numpy()).max() < 2e-4


def test_unsupported_shape_reports(ops):
    from montecosmo_b200._capi import McpmError
    ops.set_fused_fft((12, 8, 8), False)
    with pytest.raises(McpmError):
        ops.set_fused_fft((12, 8, 8), True)


@pytest.mark.parametrize("nx,ny,nz,ny_loc,y0", [(64, 12, 16, 12, 0), (128, 8, 10, 4, 4), (256, 6, 8, 3, 2), (512, 6, 4, 2, 3),
                                               (1024, 4, 6, 4, 0)])
def test_kernel_against_numpy_x_transforms(ops, nx, ny, nz, ny_loc, y0):
    """The kernel alone on a (ky-block of a) half spectrum: out = IFFT_x(kernel * FFT_x(in)) with numpy's FFT along x and
    the engine's own streaming multiply (mcpm_force_spectra[_T]_slab) in between.  float32 vs float64 FFT: 5e-6."""
    rng = np.random.default_rng(nx + ny_loc)
    nzc = nz // 2 + 1
    A = ops.A
    st = A.stream()
    shape_c = (nx, ny_loc, nzc)
    x1 = (rng.normal(size=shape_c) + 1j * rng.normal(size=shape_c)).astype(np.complex64)
    x3 = (rng.normal(size=(3, *shape_c)) + 1j * rng.normal(size=(3, *shape_c))).astype(np.complex64)
    norm = 0.37
    for lap_fd, grad_fd, kcut, dec in [(0, 0, 0.0, 0), (2, 4, 6.0, 2)]:
        # forward operator
        d_in, d_out = A.prepare(x1, "c64"), A.empty((3, *shape_c), "c64")
        ops._call("mcpm_xfuse_force_slab", st, A.ptr(d_in), A.ptr(d_out), nx, ny, nz, ny_loc, y0, lap_fd, grad_fd, kcut,
                  dec, norm)
        fk = A.prepare(np.fft.fft(x1.astype(np.complex128), axis=0).astype(np.complex64), "c64")
        mid = A.empty((3, *shape_c), "c64")
        ops._call("mcpm_force_spectra_slab", st, A.ptr(fk), A.ptr(mid), nx, ny, nz, ny_loc, y0, lap_fd, grad_fd, kcut,
                  dec, norm)
        ref = np.fft.ifft(to_numpy(mid).astype(np.complex128), axis=1) * nx
        got = to_numpy(d_out).astype(np.complex128)
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 5e-6
        # transpose operator
        d_in3, d_out1 = A.prepare(x3, "c64"), A.empty(shape_c, "c64")
        ops._call("mcpm_xfuse_force_T_slab", st, A.ptr(d_in3), A.ptr(d_out1), nx, ny, nz, ny_loc, y0, lap_fd, grad_fd,
                  kcut, dec, norm)
        fk3 = A.prepare(np.fft.fft(x3.astype(np.complex128), axis=1).astype(np.complex64), "c64")
        mid1 = A.empty(shape_c, "c64")
        ops._call("mcpm_force_spectra_T_slab", st, A.ptr(fk3), A.ptr(mid1), nx, ny, nz, ny_loc, y0, lap_fd, grad_fd, kcut,
                  dec, 0, 0, norm)
        ref1 = np.fft.ifft(to_numpy(mid1).astype(np.complex128), axis=0) * nx
        got1 = to_numpy(d_out1).astype(np.complex128)
        assert np.linalg.norm(got1 - ref1) / np.linalg.norm(ref1) < 5e-6


@pytest.mark.parametrize("shape", [(64, 12, 16), (128, 6, 10), (512, 4, 4)])
@pytest.mark.parametrize("lpt_order,read_order,fd", [(1, 1, (0, 0)), (2, 2, (0, 0)), (2, 1, (2, 4))])
def test_lpt_and_its_vjp(ops, shape, lpt_order, read_order, fd):
    """lpt runs spectrum-in (force / Hessian set -> inverse x-FFTs) and spectrum-out (x-FFTs -> transposed kernels with
    Hermitian weights, accumulated) variants of the fused kernel: fused vs 3-D cuFFT path 2e-5, forward vs oracle 5e-5."""
    rng = np.random.default_rng(11)
    n = int(np.prod(shape))
    dk = (np.fft.rfftn(rng.normal(size=shape)) * 0.02).astype(np.complex64)
    q = O.regular_pos(shape).numpy()
    pos = (q if read_order == 1 else q + rng.normal(scale=0.4, size=q.shape)).astype(np.float32)
    d1, d2, dv2 = 0.61, -0.17, -0.52
    dpb, vlb = (rng.normal(size=pos.shape).astype(np.float32) for _ in range(2))
    lap_fd, grad_fd = (np.inf if v == 0 else v for v in fd)
    res = {}
    for fused in (False, True):
        ops.set_fused_fft(shape, fused)
        dp, vl, tape = ops.lpt(dk, pos, d1, d2, dv2, lpt_order, read_order, lap_fd, grad_fd, tape=True)
        dkbar, cb = ops.lpt_vjp(pos, dk.shape, d1, d2, dv2, dpb, vlb, tape, lpt_order, read_order, lap_fd, grad_fd,
                                want_coef=True)
        res[fused] = [to_numpy(x).copy() for x in (dp, vl, dkbar, cb)]
    ops.set_fused_fft(shape, True)
    for name, a, b in zip(["dpos", "vel", "dkbar", "coefbar"], res[True], res[False]):
        err = np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel())
        assert err < 2e-5, (name, err)
    dkt = torch.as_tensor(dk, dtype=torch.complex128)
    f1 = O.pm_forces(T(pos), dkt, read_order, grad_fd=grad_fd, lap_fd=lap_fd)
    dpo, vlo = d1 * f1, f1
    if lpt_order == 2:
        f2 = O.pm_forces2(T(pos), dkt, read_order, grad_fd=grad_fd, lap_fd=lap_fd)
        dpo, vlo = dpo - d2 * f2, vlo - dv2 * f2
    tol = 5e-5 if shape[0] <= 256 else 3e-4  # float32 on a 512-cell axis (see test_forces_and_vjp)
    assert rel(res[True][0], dpo.numpy()) < tol and rel(res[True][1], vlo.numpy()) < tol


@pytest.mark.parametrize("nx,ny,nz,grad_fd", [(64, 12, 16, 0), (128, 8, 10, 4), (512, 6, 4, 2)])
def test_two_field_forms_match_the_three_field_ones(ops, nx, ny, nz, grad_fd):
    """The two-field forms of the distributed force transform (xfft_kernel.h: FORCE2 / FORCE2_T + fourier.cu: yz_gradients)
    against the three-field ones, through the peer entry point with ONE rank (its own buffers as the peer table): the y and
    z gradient factors applied on the local (y,z) spectra before / after the x-transform give the same three force spectra
    (1 -> 3) and the same density cotangent (3 -> 1), Nyquist rows and planes included.  2e-6."""
    import ctypes as C
    rng = np.random.default_rng(nx + grad_fd)
    nzc = nz // 2 + 1
    A = ops.A
    st = A.stream()
    shape_c = (nx, ny, nzc)
    x1 = (rng.normal(size=(1, *shape_c)) + 1j * rng.normal(size=(1, *shape_c))).astype(np.complex64)
    x3 = (rng.normal(size=(3, *shape_c)) + 1j * rng.normal(size=(3, *shape_c))).astype(np.complex64)
    tab = lambda t: (C.c_void_p * 1)(A.ptr(t))
    rel = lambda a, b: float(np.linalg.norm(to_numpy(a).astype(np.complex128) - to_numpy(b).astype(np.complex128))
                             / np.linalg.norm(to_numpy(b).astype(np.complex128)))

    def peer(inp, out, transpose):
        ops._call("mcpm_xfuse_force_peer", st, tab(inp), tab(out), 1, transpose, nx, ny, nz, ny, 0, 0, grad_fd, 0.0, 0, 0.37)

    # 1 -> 3
    d_in = A.prepare(x1, "c64")
    ref3, two3 = A.empty((3, *shape_c), "c64"), A.zeros((3, *shape_c), "c64")
    peer(d_in, ref3, 0)
    peer(d_in, two3, 2)
    ops._call("mcpm_yz_gradients", st, A.ptr(two3), nx, ny, nz, grad_fd, 0)
    assert rel(two3, ref3) < 2e-6
    # 3 -> 1
    d_in3, d_in3b = A.prepare(x3, "c64"), A.prepare(x3.copy(), "c64")
    ref1, two1 = A.empty((1, *shape_c), "c64"), A.empty((1, *shape_c), "c64")
    peer(d_in3, ref1, 1)
    ops._call("mcpm_yz_gradients", st, A.ptr(d_in3b), nx, ny, nz, grad_fd, 1)
    peer(d_in3b, two1, 3)
    assert rel(two1, ref1) < 2e-6


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("nx,ny,nz,grad_fd", [(64, 12, 16, 0), (256, 8, 6, 4)])
def test_peer_transform_with_two_emulated_ranks(ops, mode, nx, ny, nz, grad_fd):
    """The distributed x-transform kernel (mcpm_xfuse_force_peer) with TWO ranks' buffers on one device: each "rank" owns
    nx/2 x-planes of every component and transforms half of the ky rows, reading and writing both buffers through the
    peer table -- against the same operator on the whole spectrum with one rank.  All four modes (1 -> 3, 3 -> 1 and the
    two-field forms 1 -> 2, 2 -> 1 of the slab step loop): identical arithmetic per column, 1e-6."""
    import ctypes as C
    P = 2
    xl, kyl, nzc = nx // P, ny // P, nz // 2 + 1
    n_in, n_out = {0: (1, 3), 1: (3, 1), 2: (1, 2), 3: (2, 1)}[mode]
    rng = np.random.default_rng(10 * nx + mode)
    A = ops.A
    st = A.stream()
    x = (rng.normal(size=(n_in, nx, ny, nzc)) + 1j * rng.normal(size=(n_in, nx, ny, nzc))).astype(np.complex64)
    tab = lambda ts: (C.c_void_p * len(ts))(*[A.ptr(t) for t in ts])
    whole_in, whole_out = A.prepare(x, "c64"), A.zeros((n_out, nx, ny, nzc), "c64")
    ops._call("mcpm_xfuse_force_peer", st, tab([whole_in]), tab([whole_out]), 1, mode, nx, ny, nz, ny, 0, 0, grad_fd, 0.0,
              0, 0.25)
    parts_in = [A.prepare(np.ascontiguousarray(x[:, r * xl:(r + 1) * xl]), "c64") for r in range(P)]
    parts_out = [A.zeros((n_out, xl, ny, nzc), "c64") for _ in range(P)]
    for r in range(P):
        ops._call("mcpm_xfuse_force_peer", st, tab(parts_in), tab(parts_out), P, mode, nx, ny, nz, kyl, r * kyl, 0,
                  grad_fd, 0.0, 0, 0.25)
    got = np.concatenate([to_numpy(t) for t in parts_out], axis=1)
    ref = to_numpy(whole_out)
    assert np.abs(ref).max() > 0
    assert np.linalg.norm((got - ref).ravel()) <= 1e-6 * np.linalg.norm(ref.ravel())
