"""
The streaming brick scatter (csrc/brick.cu: brick_stream_kernel -- persistent CTAs, particle rows staged by
cp.async.bulk + mbarrier, tile flushed over the bounding box of the deposits) against the generic global-atomic CIC
kernels (paint.cu) and the per-brick kernel it replaces, through the C ABI.

The tile is fixed point with a per-brick scale, so agreement is to the fixed-point quantum (2e-5 relative L2, the bound
of tests/test_abi_parity.py::test_brick_scatter_matches_generic), not bit for bit.  Cases: lattice == mesh in relative
and absolute coordinates (wrapped and far-out particles: strays), a slab-like frame (lattice thinner than the mesh,
halo offset), per-particle weights, the interlacing shift, both tile row strides (44 | 48 words), a clustered brick
that overflows the fine fixed-point scale (the whole brick then takes float atomics), and many bricks per CTA.
"""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    import montecosmo_b200.nbody as nbody
    return nbody.ops()


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def _scatter(ops, lattice, mesh, origin, pos, w, ws, shift, vbar, scale, frame_rel):
    """density mesh and the 3 reverse-step meshes through mcpm_paint_brick_f / mcpm_paint3_brick_f"""
    from montecosmo_b200._capi import Frame
    lib, A = ops.lib, ops.A
    n = pos.shape[0]
    fr = Frame(1 if frame_rel else 0, *lattice, *origin, *lattice)
    out = torch.zeros(mesh, device=pos.device)
    rc = lib.mcpm_paint_brick_f(A.stream(), C.byref(fr), *lattice, pos.data_ptr(), w.data_ptr() if w is not None else None,
                                ws, shift, n, *mesh, out.data_ptr())
    assert rc == 0, ops.last_error() if hasattr(ops, "last_error") else rc
    out3 = torch.zeros((3, *mesh), device=pos.device)
    vb = vbar.clone()
    rc = lib.mcpm_paint3_brick_f(A.stream(), C.byref(fr), *lattice, pos.data_ptr(), vb.data_ptr(), None, 0.0, scale, n,
                                 *mesh, out3.data_ptr())
    assert rc == 0
    return out, out3


def _reference(ops, lattice, mesh, origin, pos, w, ws, shift, vbar, scale, frame_rel):
    """the same sums with the generic kernels on absolute positions (float64 site + displacement, rounded once)"""
    dev = pos.device
    ax = [torch.arange(s, device=dev, dtype=torch.float64) + o for s, o in zip(lattice, origin)]
    q = torch.stack(torch.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
    x = (q + pos.double() if frame_rel else pos.double())
    xa = x.float().contiguous()
    out = ops.paint(xa, mesh, w, ws, 2, None, shift)
    out3 = torch.stack([ops.paint(xa, mesh, vbar[:, c].contiguous(), scale, 2) for c in range(3)])
    return out, out3


CASES = [
    # lattice,        mesh,           origin,      relative, weights, shift, disp
    ((64, 32, 128), (64, 32, 128), (0, 0, 0), True, False, 0.0, 1.5),
    ((64, 32, 128), (64, 32, 128), (0, 0, 0), False, True, 0.5, 2.5),
    ((16, 40, 64), (64, 40, 64), (24, 0, 0), True, True, 0.0, 2.0),   # a slab rank: 16 planes inside a 64-plane mesh
    ((8, 8, 32), (26, 18, 44), (3, 2, 1), True, False, 0.0, 0.7),     # one brick, the smallest mesh both kernels take
    ((128, 128, 128), (128, 128, 128), (0, 0, 0), True, True, 0.0, 2.2),  # 1024 bricks: several per resident CTA
]


@pytest.mark.parametrize("lattice,mesh,origin,relative,weights,shift,disp", CASES)
def test_stream_scatter_matches_generic(lattice, mesh, origin, relative, weights, shift, disp):
    ops = _ops()
    dev = ops.A.device
    n = int(np.prod(lattice))
    g = torch.Generator(device=dev).manual_seed(n % 1000)
    # smooth displacement + jitter, a few far-out particles (strays) and, in absolute coordinates, caller-wrapped ones
    ax = [torch.arange(s, device=dev, dtype=torch.float32) for s in lattice]
    q = torch.stack(torch.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
    k = [2 * np.pi / s for s in lattice]
    d = torch.stack([disp * torch.sin(k[1] * q[:, 1]) + torch.cos(2 * k[2] * q[:, 2]),
                     disp * torch.sin(2 * k[2] * q[:, 2]) + 0.8 * torch.cos(k[0] * q[:, 0]),
                     disp * torch.sin(k[0] * q[:, 0]) + 1.5 * torch.cos(3 * k[1] * q[:, 1])], -1)
    d = d + 0.4 * torch.randn((n, 3), device=dev, generator=g)
    far = torch.randperm(n, device=dev, generator=g)[: max(n // 50, 1)]
    d[far] += (torch.rand((len(far), 3), device=dev, generator=g) - 0.5) * 40
    if relative:
        pos = d.contiguous()
    else:
        pos = q + torch.tensor(origin, device=dev, dtype=torch.float32) + d
        pos[::7] -= 2.0 * torch.tensor(mesh, device=dev, dtype=torch.float32)
        pos = pos.contiguous()
    w = (torch.rand(n, device=dev, generator=g) * 2.8 + 0.2) if weights else None
    vbar = torch.randn((n, 3), device=dev, generator=g) * torch.exp(torch.randn((n, 1), device=dev, generator=g))
    args = (ops, lattice, mesh, origin, pos, w, 0.7, shift, vbar, 1.3, relative)
    tune = lambda key, v: ops._call("mcpm_tune", key, v)
    ref, ref3 = _reference(*args)
    try:
        tune(b"brick_stream1", 1)  # the density paint takes the streaming kernel only on request
        res = {}
        for knob in (0, 44, 48):
            tune(b"brick_stream", knob)
            out, out3 = _scatter(*args)
            torch.cuda.synchronize()
            res[knob] = (out, out3)
            assert _rel(out, ref) < 2e-5, (knob, "density")
            for c in range(3):
                assert _rel(out3[c], ref3[c]) < 2e-5, (knob, "channel", c)
            # conservation: the mesh total is the sum of the weights (bricks.py:1101-1102)
            tot = float((w.double().sum() if w is not None else n) * 0.7)
            assert abs(float(out.double().sum()) - tot) < 2e-6 * abs(tot)
        # the two strides differ only in where the tile sits in shared memory: same integers, same flush order per cell
        assert _rel(res[44][0], res[48][0]) < 1e-6
    finally:
        tune(b"brick_stream", 44)
        tune(b"brick_stream1", 0)


def test_stream_scatter_overflowing_brick_falls_back():
    """A brick whose particles all sit in ONE cell overflows the fine fixed-point scale of the density paint (a cell holding
    more than 1/16 of the brick's weight): the tile is dropped and the brick deposited with float atomics -- the result
    must still be right, and the next bricks of the same resident CTA must find a clean tile."""
    ops = _ops()
    dev = ops.A.device
    lattice = mesh = (32, 32, 64)
    n = int(np.prod(lattice))
    g = torch.Generator(device=dev).manual_seed(5)
    ax = [torch.arange(s, device=dev, dtype=torch.float32) for s in lattice]
    q = torch.stack(torch.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
    d = 0.3 * torch.randn((n, 3), device=dev, generator=g)
    # collapse the 8 x 8 x 32 bricks number 0 and 5 (in x-major brick order) onto single points
    qb = q.reshape(4, 8, 4, 8, 2, 32, 3)
    db = d.reshape(4, 8, 4, 8, 2, 32, 3)
    for (bi, bj, bk), target in (((0, 0, 0), (4.3, 3.6, 17.2)), ((0, 2, 1), (5.5, 20.1, 40.9))):
        db[bi, :, bj, :, bk] = torch.tensor(target, device=dev) - qb[bi, :, bj, :, bk]
    pos = db.reshape(n, 3).contiguous()
    vbar = torch.randn((n, 3), device=dev, generator=g)
    args = (ops, lattice, mesh, (0, 0, 0), pos, None, 1.0, 0.0, vbar, 1.0, True)
    ref, ref3 = _reference(*args)
    tune = lambda key, v: ops._call("mcpm_tune", key, v)
    try:
        tune(b"brick_stream1", 1)
        for knob in (0, 44):
            tune(b"brick_stream", knob)
            out, out3 = _scatter(*args)
            torch.cuda.synchronize()
            assert _rel(out, ref) < 2e-5, knob
            assert abs(float(out.double().sum()) - n) < 2e-6 * n
            for c in range(3):
                assert _rel(out3[c], ref3[c]) < 2e-5, (knob, c)
    finally:
        tune(b"brick_stream", 44)
        tune(b"brick_stream1", 0)
