"""
The bulk-copy staged gathers (csrc/cic4_tma.cu: kick_drift4 / read_grad4v with the particle arrays moved by
cp.async.bulk + mbarrier) against the one-thread-per-particle kernels of cic4.cu they replace, through the C ABI.

Same arithmetic in the same association order: results must be BIT-IDENTICAL, for absolute and lattice-relative
positions, every segment length, in place (as the engine's step loop calls them), with wrapped / far-out particles,
and on particle counts that are not a whole number of segments (which must fall back, silently and correctly).
Both are separately checked against the float64 oracle by tests/test_abi_parity.py.
"""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    import montecosmo_b200.nbody as nbody
    return nbody.ops()


def _case(ops, mesh, lattice, rel, seed=0, big=False):
    from montecosmo_b200._capi import frame as make_frame
    A, lib = ops.A, ops.lib
    dev = A.device
    n = int(np.prod(lattice))
    g = torch.Generator(device=dev).manual_seed(seed)
    disp = torch.randn((n, 3), device=dev, generator=g) * (6.0 if big else 1.5)
    ax = [torch.arange(s, device=dev, dtype=torch.float32) * (m / s) for s, m in zip(lattice, mesh)]
    q = torch.stack(torch.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
    pos = disp if rel else (q + disp)
    if not rel:
        pos[::7] += 3.0 * mesh[0]  # far outside the box: exact modulo path
        pos[1::11] -= 2.0 * mesh[1]
    vel = torch.randn((n, 3), device=dev, generator=g)
    fm4 = torch.randn((*mesh, 4), device=dev, generator=g)
    rho = torch.randn(mesh, device=dev, generator=g)
    xbar = torch.randn((n, 3), device=dev, generator=g)
    fr = make_frame(lattice, span=mesh) if rel else None
    frp = C.byref(fr) if rel else None
    return dict(n=n, pos=pos.contiguous(), vel=vel, fm4=fm4, rho=rho, xbar=xbar, frp=frp, st=A.stream(), lib=lib,
                mesh=mesh)


def _kick(c):
    p, v = c["pos"].clone(), c["vel"].clone()
    rc = c["lib"].mcpm_kick_drift4_f(c["st"], c["frp"], p.data_ptr(), v.data_ptr(), c["fm4"].data_ptr(), c["n"], *c["mesh"],
                                     0.7, 0.4, 0.25)
    assert rc == 0
    return p, v


def _grad(c):
    vb, xb = c["vel"].clone(), c["xbar"].clone()
    rc = c["lib"].mcpm_read_grad4v_f(c["st"], c["frp"], c["pos"].data_ptr(), c["fm4"].data_ptr(), c["rho"].data_ptr(),
                                     vb.data_ptr(), 0.6, 0.8, c["n"], *c["mesh"], xb.data_ptr())
    assert rc == 0
    return vb, xb


@pytest.mark.parametrize("mesh,lattice,rel", [
    ((64, 32, 128), (64, 32, 128), True),    # lattice == mesh, relative (what FieldModel runs)
    ((64, 32, 128), (64, 32, 128), False),   # absolute positions, wrapped and far-out particles
    ((48, 40, 96), (48, 40, 96), True),      # pz = 96: whole 32-segments only -> 64 / 128 must fall back
    ((34, 6, 64), (34, 6, 64), True),        # py % 4 != 0: no 2 x 4 patches of rows -> linear segment order
    ((40, 24, 64), (20, 12, 32), True),      # lattice coarser than the mesh (spacing 2): not a unit frame -> falls back
    ((33, 17, 50), (33, 17, 50), False),     # particle count not a multiple of any segment -> falls back
])
def test_tma_gathers_are_bit_identical(mesh, lattice, rel):
    ops = _ops()
    c = _case(ops, mesh, lattice, rel, big=not rel)
    tune = lambda k, v: ops._call("mcpm_tune", k, v)
    try:
        tune(b"gather_tma", 0)
        ref_k, ref_g = _kick(c), _grad(c)
        for seg, brick in ((32, 1), (64, 1), (128, 1), (32, 0)):
            tune(b"gather_tma", 1)
            tune(b"gather_seg", seg)
            tune(b"gather_brick", brick)
            k, g = _kick(c), _grad(c)
            torch.cuda.synchronize()
            for a, b in zip(k + g, ref_k + ref_g):
                assert torch.equal(a, b), (mesh, lattice, rel, seg, brick, float((a - b).abs().max()))
    finally:
        tune(b"gather_tma", 1)
        tune(b"gather_seg", 32)
        tune(b"gather_brick", 0)


def test_tma_gathers_launch_when_applicable():
    """The staged kernels really are the ones that run on the step-loop geometry (no silent fallback): with the knob on,
    the launch counter moves exactly as with it off, and the two runs agree bit for bit at 128^3 (several waves of
    CTAs, every warp with four segments in flight)."""
    ops = _ops()
    c = _case(ops, (128, 128, 128), (128, 128, 128), True, seed=3)
    tune = lambda k, v: ops._call("mcpm_tune", k, v)
    try:
        tune(b"gather_tma", 0)
        ref = _kick(c) + _grad(c)
        tune(b"gather_tma", 1)
        out = _kick(c) + _grad(c)
        torch.cuda.synchronize()
        assert all(torch.equal(a, b) for a, b in zip(out, ref))
    finally:
        tune(b"gather_tma", 1)
