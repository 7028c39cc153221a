"""
Slab-decomposed engine (montecosmo_b200/dist.py) on 2 ranks against the single-process engine, on CPU:
torch.distributed `gloo` + the CPU port of the kernels.  The same class runs under NCCL on GPUs (tools/slab_check.py).
Checks: distributed rfftn / irfftn, halo reduce / gather (and that they are transposes), lpt + BullFrog steps forward,
and the reverse sweep (cotangent of delta_k), each rank comparing its own slab with the single-process result.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, shape, halo, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from oracle import cpu_port
        from montecosmo_b200.dist import SlabPM
        from montecosmo_b200.cosmo import Cosmology, a2g, a2g2, a2dg2dg, bullfrog_coefficients
        ops = cpu_port.cpu_ops()
        pm = SlabPM(ops, shape, halo=halo)
        nx, ny, nz = shape
        rng = np.random.default_rng(0)  # same global data on every rank
        rel = lambda a, b: float(np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel()) / np.linalg.norm(np.asarray(b).ravel()))
        res = {}

        # 1. distributed FFT
        m = rng.normal(size=(2, *shape)).astype(np.float32)
        mk = np.fft.rfftn(m.astype(np.float64), axes=(1, 2, 3))
        ck = pm.rfftn(pm.scatter_real(torch.tensor(m)))
        res["rfftn"] = rel(ck.numpy(), mk[:, :, pm.y0:pm.y0 + pm.kyl, :])
        back = pm.irfftn(ck).numpy() / pm.N
        res["irfftn"] = rel(back, m[:, pm.x0:pm.x0 + pm.xl])

        # 2. halos: reduce adds neighbours' halos; <reduce(a), b> == <a, gather(b)> on the matching subspaces
        ext = torch.tensor(rng.normal(size=(world, pm.ext, ny, nz)).astype(np.float32))  # every rank's extended mesh
        mine = ext[rank].clone()
        pm.halo_reduce(mine)
        H, xl = pm.H, pm.xl
        exp = ext[rank].clone()
        exp[xl:xl + H] += ext[(rank + 1) % world][:H]
        exp[H:2 * H] += ext[(rank - 1) % world][H + xl:]
        res["halo_reduce"] = rel(mine[H:H + xl].numpy(), exp[H:H + xl].numpy())
        g = ext[rank].clone()
        pm.halo_gather(g)
        res["halo_gather"] = max(rel(g[H + xl:].numpy(), ext[(rank + 1) % world][H:2 * H].numpy()),
                                 rel(g[:H].numpy(), ext[(rank - 1) % world][xl:xl + H].numpy()))

        # 3. lpt + steps forward / backward vs the single-process engine
        # smooth (red) spectrum with O(0.5 cell) displacements: particles hugging cell faces, where the CIC gradient is
        # discontinuous and float32 rounding of the local frame can flip a base cell, are then rare
        kk = np.sqrt(sum(np.meshgrid(np.fft.fftfreq(nx) ** 2, np.fft.fftfreq(ny) ** 2, np.fft.rfftfreq(nz) ** 2,
                                     indexing="ij")))
        kk[0, 0, 0] = 1.0
        dk = (np.fft.rfftn(rng.normal(size=shape)) * 0.08 * kk ** -1.5).astype(np.complex64)
        dk[0, 0, 0] = 0
        c = Cosmology()
        a0, a1, ns = 0.1, 0.7, 2
        pos, vel, tape = pm.nbody_forward(pm.scatter_spectrum(torch.tensor(dk)), c, a0, a1, ns)
        ax = [np.arange(s, dtype=np.float32) for s in shape]
        qfull = np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
        d1, d2, dv2 = float(a2g(c, a0)), float(a2g2(c, a0)), float(a2dg2dg(c, a0))
        al, be, pre, post, _, _ = bullfrog_coefficients(c, a0, a1, ns)
        co = [t.tolist() for t in (al, be, pre, post)]
        dp, vl, ltape = ops.lpt(dk, qfull, d1, d2, dv2, 2, 1, tape=True)
        pfull, vfull = (dp + torch.tensor(qfull)).contiguous(), vl.clone()
        stape = ops.nbody_steps(pfull, vfull, shape, *co, tape=True)
        sl = slice(rank * pm.npl, (rank + 1) * pm.npl)  # lattice order is x-major: a slab is a contiguous range
        mypos = pos.numpy().copy()
        mypos[:, 0] += pm.x0 - pm.H
        res["pos"] = float(np.abs(mypos - pfull.numpy()[sl]).max())
        res["vel"] = rel(vel.numpy(), vfull.numpy()[sl])
        res["disp_rms"] = float((pfull.numpy() - qfull).std())
        pb = rng.normal(size=qfull.shape).astype(np.float32)
        vb = rng.normal(size=qfull.shape).astype(np.float32)
        pbf, vbf = torch.tensor(pb.copy()), torch.tensor(vb.copy())
        ops.nbody_steps_vjp(pbf, vbf, shape, *co, stape)
        ref = ops.lpt_vjp(qfull, dk.shape, d1, d2, dv2, pbf, vbf, ltape, 2, 1)
        # the two halves of the reverse sweep separately, then together
        pl, vl_ = torch.tensor(pb[sl].copy()), torch.tensor(vb[sl].copy())
        pm.steps_backward(tape[1], pl, vl_, *co)
        res["steps_bwd"] = max(rel(pl.numpy(), pbf.numpy()[sl]), rel(vl_.numpy(), vbf.numpy()[sl]))
        res["lpt_bwd"] = rel(pm.lpt_backward(tape[0], pbf[sl].contiguous(), vbf[sl].contiguous()).numpy(),
                             ref.numpy()[:, pm.y0:pm.y0 + pm.kyl, :])
        dkbar = pm.nbody_backward(tape, torch.tensor(pb[sl]), torch.tensor(vb[sl]))
        res["dkbar"] = rel(dkbar.numpy(), ref.numpy()[:, pm.y0:pm.y0 + pm.kyl, :])
        q.put((rank, res, None))
        dist.destroy_process_group()
    except Exception as e:  # surface the traceback in the parent
        import traceback
        q.put((rank, None, traceback.format_exc()))


@pytest.mark.parametrize("shape,halo", [((16, 16, 16), 6), ((24, 16, 20), 8)])
def test_slab_engine_world2_gloo(shape, halo):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29541 + shape[0]
    ps = [ctx.Process(target=_worker, args=(r, 2, port, shape, halo, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = [q.get(timeout=600) for _ in ps]
    for p in ps:
        p.join(60)
    for rank, res, err in out:
        assert err is None, f"rank {rank}:\n{err}"
        assert res["rfftn"] < 1e-6 and res["irfftn"] < 1e-6, res
        assert res["halo_reduce"] < 1e-6 and res["halo_gather"] < 1e-7, res
        assert res["pos"] < 1e-4 and res["vel"] < 1e-4, res
        assert res["steps_bwd"] < 5e-4 and res["lpt_bwd"] < 5e-4 and res["dkbar"] < 5e-4, res
        assert res["disp_rms"] > 0.2, res  # the comparison above is on a genuinely displaced lattice
