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
_PORTS = iter(range(29700 + os.getpid() % 200, 29999))


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
        # both engines carry displacements from the lattice sites (mcpm_engine_set_relative / mcpm_frame)
        pfull, vfull = dp.clone().contiguous(), vl.clone()
        stape = ops.nbody_steps(pfull, vfull, shape, *co, tape=True, lattice=shape)
        sl = slice(rank * pm.npl, (rank + 1) * pm.npl)  # lattice order is x-major: a slab is a contiguous range
        res["pos"] = float(np.abs(pos.numpy() - pfull.numpy()[sl]).max())
        res["vel"] = rel(vel.numpy(), vfull.numpy()[sl])
        res["disp_rms"] = float(pfull.numpy().std())
        pb = rng.normal(size=qfull.shape).astype(np.float32)
        vb = rng.normal(size=qfull.shape).astype(np.float32)
        pbf, vbf = torch.tensor(pb.copy()), torch.tensor(vb.copy())
        ops.nbody_steps_vjp(pbf, vbf, shape, *co, stape, lattice=shape)
        ref = ops.lpt_vjp(qfull, dk.shape, d1, d2, dv2, pbf, vbf, ltape, 2, 1)
        # the two halves of the reverse sweep separately, then together
        pl, vl_ = torch.tensor(pb[sl].copy()), torch.tensor(vb[sl].copy())
        pm.steps_backward(tape[1], pl, vl_, *co)
        res["steps_bwd"] = max(rel(pl.numpy(), pbf.numpy()[sl]), rel(vl_.numpy(), vbf.numpy()[sl]))
        res["lpt_bwd"] = rel(pm.lpt_backward(tape[0], pbf[sl].contiguous(), vbf[sl].contiguous()).numpy(),
                             ref.numpy()[:, pm.y0:pm.y0 + pm.kyl, :])
        dkbar = pm.nbody_backward(tape, torch.tensor(pb[sl]), torch.tensor(vb[sl]))
        res["dkbar"] = rel(dkbar.numpy(), ref.numpy()[:, pm.y0:pm.y0 + pm.kyl, :])
        # 4. halo sized from the measurement: the same number on every rank (an all-reduce inside), at least the largest
        # x-displacement + the CIC stencil, at most what was allotted; a slab engine rebuilt with it gives the same result
        need = pm.halo_needed(factor=1.0, margin=2)
        dmax = float(np.abs(pfull.numpy()[:, 0]).max())
        res["halo_needed_ok"] = float(not (dmax + 1.0 <= need <= H + 1))
        pm2 = SlabPM(ops, shape, halo=min(need, H))
        pos2, vel2, _ = pm2.nbody_forward(pm2.scatter_spectrum(torch.tensor(dk)), c, a0, a1, ns)
        res["halo_resized"] = max(float(np.abs(pos2.numpy() - pos.numpy()).max()), rel(vel2.numpy(), vel.numpy()))
        # 5. per-step active halo planes from the measured kick positions: the exchanges of the early steps move fewer
        # planes; forward and reverse sweeps unchanged; a schedule that is too narrow trips the guard
        sched = pm.halo_schedule(ns, factor=1.0, margin=2)
        res["sched_ok"] = float(not (len(sched) == ns and all(1 <= h <= H for h in sched) and sched[0] <= sched[-1]))
        pm.set_halo_schedule(sched)
        pos3, vel3, tape3 = pm.nbody_forward(pm.scatter_spectrum(torch.tensor(dk)), c, a0, a1, ns)
        dkbar3 = pm.nbody_backward(tape3, torch.tensor(pb[sl]), torch.tensor(vb[sl]))
        res["sched_same"] = max(float(np.abs(pos3.numpy() - pos.numpy()).max()), rel(vel3.numpy(), vel.numpy()),
                                rel(dkbar3.numpy(), dkbar.numpy()))
        pm.set_halo_schedule([1] * ns)
        try:
            pm.nbody_forward(pm.scatter_spectrum(torch.tensor(dk)), c, a0, a1, ns)
            res["sched_guard"] = 1.0
        except RuntimeError:
            res["sched_guard"] = 0.0
        pm.set_halo_schedule(None)
        q.put((rank, res, None))
        dist.destroy_process_group()
    except Exception as e:  # surface the traceback in the parent
        import traceback
        q.put((rank, None, traceback.format_exc()))


@pytest.mark.parametrize("world,shape,halo", [(2, (16, 16, 16), 6), (2, (24, 16, 20), 8), (4, (16, 16, 16), 4)])
def test_slab_engine_world2_gloo(world, shape, halo):
    """2 ranks, and 4 ranks (where a rank's left and right neighbours differ, so a swapped direction cannot hide)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29541 + shape[0] + 11 * world
    ps = [ctx.Process(target=_worker, args=(r, world, port, shape, halo, q)) for r in range(world)]
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
        assert res["halo_needed_ok"] == 0.0 and res["halo_resized"] < 1e-5, res
        assert res["sched_ok"] == 0.0 and res["sched_same"] < 1e-5 and res["sched_guard"] == 0.0, res


def _model_worker(rank, world, port, shape, halo, kw, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from oracle import cpu_port
        import montecosmo_b200.nbody as nbody
        from montecosmo_b200.dist import SlabPM
        from montecosmo_b200.dist_model import SlabFieldModel
        from montecosmo_b200.model import FieldModel
        ops = cpu_port.cpu_ops()
        nbody._OPS = ops  # the single-process model of this (CPU) test runs on the same checker library
        box = tuple(40.0 * s for s in shape)
        rng = np.random.default_rng(3)  # same global fields on every rank
        white = torch.tensor(rng.normal(size=shape).astype(np.float32))
        truth = torch.tensor(rng.normal(size=shape).astype(np.float32))
        kw = dict(kw)
        evo = kw.pop("evolution", "nbody")
        ref = FieldModel(shape, box, evo, a_start=0.1, out_shape="mesh", **kw)
        obs = ref.evolve(truth).detach() + torch.tensor(rng.normal(size=shape).astype(np.float32))
        lp_ref, f_ref = ref.value_and_force(white, obs)
        pm = SlabPM(ops, shape, halo=halo)
        mdl = SlabFieldModel(pm, box, evo, a_start=0.1, **kw)
        sl = slice(pm.x0, pm.x0 + pm.xl)
        lp, f = mdl.value_and_force(white[sl].contiguous(), obs[sl].contiguous())
        rel = lambda a, b: float(np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel()) / np.linalg.norm(np.asarray(b).ravel()))
        res = {"lp": abs(float(lp) - float(lp_ref)) / abs(float(lp_ref)), "force": rel(f.numpy(), f_ref.numpy()[sl]),
               "pred": rel(mdl.predict(truth[sl].contiguous()).numpy(), ref.evolve(truth).detach().numpy()[sl]),
               "force_norm": float(np.linalg.norm(f_ref.numpy() + white.numpy()))}
        q.put((rank, res, None))
        dist.destroy_process_group()
    except Exception:
        import traceback
        q.put((rank, None, traceback.format_exc()))


@pytest.mark.parametrize("shape,halo,kw", [
    ((16, 16, 16), 6, dict(n_steps=2, b1=1.0, rsd=True)),
    ((24, 16, 20), 8, dict(n_steps=1, b1=0.0, rsd=False, paint_deconv=False, interlace_order=3, lpt_order=1)),
    # a paint mesh finer than the evolution mesh (BASELINE C5: 2x): second slab geometry + distributed Fourier crop
    ((16, 16, 16), 6, dict(n_steps=2, b1=1.0, rsd=True, paint_oversamp=2.0)),
    ((16, 16, 16), 4, dict(n_steps=1, b1=0.5, rsd=True, paint_oversamp=1.5, interlace_order=3)),
    ((16, 16, 16), 4, dict(evolution="lpt", a_obs=0.8, b1=1.0, rsd=True)),  # evolution = 'lpt' (model.py:763)
])
def test_slab_field_model_world2_gloo(shape, halo, kw):
    """grad(log-density) of the whole model chain on 2 slabs against the single-process FieldModel (same kernels, same
    white noise and observation): log-density 1e-5, force and predicted mesh 5e-4 relative L2."""
    _run_model_workers(2, shape, halo, kw)


def test_slab_field_model_world4_gloo():
    """The same on 4 ranks, where the left and the right neighbour of a rank differ (on 2 ranks they coincide, which
    would hide a swapped halo direction or row exchange), with the 2x paint mesh and its distributed Fourier crop."""
    _run_model_workers(4, (16, 16, 16), 4, dict(n_steps=1, b1=1.0, rsd=True, paint_oversamp=2.0))


def _run_model_workers(world, shape, halo, kw):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = next(_PORTS)  # a fresh rendezvous port per case: no reliance on a just-closed socket being reusable
    ps = [ctx.Process(target=_model_worker, args=(r, world, port, shape, halo, kw, q)) for r in range(world)]
    for p in ps:
        p.start()
    out = [q.get(timeout=600) for _ in ps]
    for p in ps:
        p.join(60)
    for rank, res, err in out:
        assert err is None, f"rank {rank}:\n{err}"
        assert res["lp"] < 1e-5 and res["force"] < 5e-4 and res["pred"] < 5e-4, res
        assert res["force_norm"] > 1e-3, res  # the likelihood term of the force is not trivially zero


@pytest.mark.parametrize("oversamp", [1.0, 2.0])
@pytest.mark.parametrize("backend", ["hostemu", pytest.param("cuda", marks=pytest.mark.gpu)])
def test_slab_field_model_single_rank(backend, oversamp):
    """The slab model on ONE rank (periodic halos exchanged with itself) against FieldModel: exercises every kernel and
    the hand-chained reverse sweep of dist_model.py on whichever device the backend names -- on a B200 this is the
    single-GPU check of the code path the multi-GPU runs take (tools/slab_bench.py --model)."""
    import montecosmo_b200.nbody as nbody
    from montecosmo_b200.dist import SlabPM
    from montecosmo_b200.dist_model import SlabFieldModel
    from montecosmo_b200.model import FieldModel
    from montecosmo_b200.ops import Ops
    old = nbody._OPS
    try:
        if backend == "hostemu":
            from oracle import cpu_port
            nbody._OPS = cpu_port.cpu_ops()
        else:
            from montecosmo_b200 import _lib
            from montecosmo_b200.ops import TorchCudaAdapter
            nbody._OPS = Ops(_lib.load(), TorchCudaAdapter())
        ops = nbody._OPS
        dev = ops.A.device
        # 64^3 on the GPU so that the fused x-transform and the brick-tiled scatters (both need nx >= 64) are on the path
        shape, halo = ((32, 32, 32), 8) if backend == "hostemu" else ((64, 64, 64), 16)
        box = tuple(20.0 * s for s in shape)
        rng = np.random.default_rng(5)
        white = torch.tensor(rng.normal(size=shape).astype(np.float32), device=dev)
        truth = torch.tensor(rng.normal(size=shape).astype(np.float32), device=dev)
        if backend == "hostemu" and oversamp != 1.0:
            shape, halo = (16, 16, 16), 6  # the 2-rank test covers this on the CPU; keep the suite short
            box = tuple(20.0 * s for s in shape)
            white, truth = white[:16, :16, :16].contiguous(), truth[:16, :16, :16].contiguous()
        kw = dict(n_steps=3, b1=0.7, rsd=True, paint_oversamp=oversamp)  # 2.0: finer paint mesh + Fourier crop (C5)
        ref = FieldModel(shape, box, "nbody", a_start=0.1, out_shape="mesh", **kw)
        obs = ref.evolve(truth).detach() + torch.tensor(rng.normal(size=shape).astype(np.float32), device=dev)
        lp_ref, f_ref = ref.value_and_force(white, obs)
        mdl = SlabFieldModel(SlabPM(ops, shape, halo=halo), box, a_start=0.1, **kw)
        lp, f = mdl.value_and_force(white, obs)
        rel = lambda a, b: float((a - b).norm() / b.norm())
        assert abs(float(lp) - float(lp_ref)) < 1e-5 * abs(float(lp_ref))
        # float32 on both sides, different kernels for the same operators (brick / fused paths vs the slab sequence).
        # Round 1 held this to 2e-3 (8.1e-4 measured on a B200): particle x was an absolute float32 coordinate kept in
        # two different frames (global vs halo-shifted).  Both engines now carry displacements from the lattice sites
        # (mcpm_engine_set_relative / mcpm_frame), the frame no longer enters the rounding, and the bound is back to 5e-4.
        assert rel(f, f_ref) < 5e-4
        assert rel(mdl.predict(truth), ref.evolve(truth).detach()) < 5e-4
    finally:
        nbody._OPS = old


def _resize_worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from oracle import cpu_port
        from montecosmo_b200.dist import SlabPM, SlabResize
        ops = cpu_port.cpu_ops()
        res = {}
        for sin, sout in [((16, 12, 20), (8, 8, 12)), ((12, 16, 8), (12, 8, 8)), ((8, 8, 8), (8, 8, 8)),
                          ((16, 16, 16), (8, 16, 8))]:
            a, b = SlabPM(ops, sin, halo=2), SlabPM(ops, sout, halo=2)
            rz = SlabResize(a, b)
            rng = np.random.default_rng(1)
            full = np.fft.rfftn(rng.normal(size=sin)).astype(np.complex64)
            ref = ops.chreshape(torch.tensor(full), (sout[0], sout[1], sout[2] // 2 + 1)).numpy()
            out = rz.forward(a.scatter_spectrum(torch.tensor(full))).numpy()
            res[f"fwd {sin}->{sout}"] = float(np.abs(out - ref[:, b.y0:b.y0 + b.kyl, :]).max() / np.abs(ref).max())
            ob = (rng.normal(size=ref.shape) + 1j * rng.normal(size=ref.shape)).astype(np.complex64)
            refb = ops.chreshape_vjp(torch.tensor(ob), full.shape).numpy()
            inb = rz.backward(b.scatter_spectrum(torch.tensor(ob))).numpy()
            res[f"bwd {sin}->{sout}"] = float(np.abs(inb - refb[:, a.y0:a.y0 + a.kyl, :]).max() / np.abs(refb).max())
        q.put((rank, res, None))
        dist.destroy_process_group()
    except Exception:
        import traceback
        q.put((rank, None, traceback.format_exc()))


def test_slab_resize_world2_gloo():
    """The distributed Fourier crop (dist.SlabResize: local x / kz crop kernel, gathered Nyquist plane, row all-to-all)
    and its transpose against the single-process mcpm_chreshape / mcpm_chreshape_vjp (themselves pinned to the golden
    vectors of utils.chreshape): 1e-6 of the largest element, every rank's rows."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_resize_worker, args=(r, 2, 29733, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = [q.get(timeout=600) for _ in ps]
    for p in ps:
        p.join(60)
    for rank, res, err in out:
        assert err is None, f"rank {rank}:\n{err}"
        assert max(res.values()) < 1e-6, res
