"""
Slab-decomposed nbody_bf forward + reverse sweep under torchrun / NCCL (BASELINE config C4).

  python -m torch.distributed.run --nproc-per-node P tools/slab_bench.py --mesh 512 --steps 3 [--check 128]

--check N first verifies, at mesh N, that every rank's slab of (pos, vel, cotangent of delta_k) equals the
single-GPU engine's result computed on the same GPU.  Then times K evaluations (lpt + 10 BullFrog steps + full reverse
sweep to the cotangent of delta_k) at --mesh, device time = max over ranks, and prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def local_delta_k(pm, cosmo, box, seed):
    """delta_k block of a Gaussian field: each rank draws its own white slab, distributed rfftn, times sqrt(P N / V)."""
    from montecosmo_b200.model import linear_power_table
    g = torch.Generator(device=pm.A.device).manual_seed(seed + pm.rank)
    white = torch.randn((1, pm.xl, pm.ny, pm.nz), device=pm.A.device, generator=g)
    wk = pm.rfftn(white)[0]
    ks, pows = linear_power_table(cosmo)
    kx = np.fft.fftfreq(pm.nx)[:, None, None] * pm.nx * 2 * np.pi / box
    ky = (np.fft.fftfreq(pm.ny)[pm.y0:pm.y0 + pm.kyl])[None, :, None] * pm.ny * 2 * np.pi / box
    kz = np.fft.rfftfreq(pm.nz)[None, None, :] * pm.nz * 2 * np.pi / box
    kk = np.sqrt(kx ** 2 + ky ** 2 + kz ** 2)
    t = np.sqrt(np.interp(kk.reshape(-1), ks, pows, left=0.0, right=0.0).reshape(kk.shape) * (pm.N / box ** 3))
    return pm.o.scale_spectrum(wk, torch.tensor(t.astype(np.float32), device=pm.A.device))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", type=int, default=512)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--halo", type=int, default=24)
    ap.add_argument("--nbody-steps", type=int, default=10)
    ap.add_argument("--check", type=int, default=0)
    ap.add_argument("--model", action="store_true",
                    help="time the whole grad(log-density) chain (dist_model.SlabFieldModel: prior, bias weights, RSD, "
                         "interlaced final paint, likelihood and the full reverse sweep to the white field) instead of "
                         "nbody_bf forward + reverse alone")
    ap.add_argument("--cell", type=float, default=2.5, help="cell size in Mpc/h (box = cell * mesh); 2.5 = BASELINE C3-C5")
    ap.add_argument("--oversamp", type=float, default=1.0,
                    help="with --model: paint mesh = oversamp x evolution mesh (BASELINE C5 uses 2)")
    ap.add_argument("--no-force-tape", action="store_true", help="tape kick positions only; recompute force meshes in the reverse sweep")
    ap.add_argument("--auto-halo", action="store_true",
                    help="with --model: after the warm-up evaluations, set the per-step active halo planes from the measured "
                         "x-displacements (SlabPM.halo_schedule); the allotted halo stays --halo")
    ap.add_argument("--model-check", action="store_true",
                    help="with --model: compare log-density and force with the single-GPU FieldModel (computed on every "
                         "rank's own GPU from the same global fields) before timing")
    ap.add_argument("--model-check-mesh", type=int, default=0, help="mesh of that comparison (default: --mesh)")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from montecosmo_b200 import nbody as nb
    from montecosmo_b200.cosmo import Cosmology, a2g, a2g2, a2dg2dg, bullfrog_coefficients
    from montecosmo_b200.dist import SlabPM
    ops = nb.ops()
    dev = ops.A.device
    cosmo = Cosmology()
    out = {}

    if a.check:
        n = a.check
        shape = (n, n, n)
        pm = SlabPM(ops, shape, halo=min(a.halo, n // world))
        rng = np.random.default_rng(0)
        kk = np.sqrt(sum(np.meshgrid(np.fft.fftfreq(n) ** 2, np.fft.fftfreq(n) ** 2, np.fft.rfftfreq(n) ** 2, indexing="ij")))
        kk[0, 0, 0] = 1.0
        dk = (np.fft.rfftn(rng.normal(size=shape)) * 0.02 * kk ** -1.5).astype(np.complex64)
        dk[0, 0, 0] = 0
        a0, a1, ns = 0.1, 0.8, 3
        pos, vel, tape = pm.nbody_forward(pm.scatter_spectrum(torch.tensor(dk)), cosmo, a0, a1, ns)
        ax = [np.arange(s, dtype=np.float32) for s in shape]
        q = torch.tensor(np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3), device=dev)
        d1, d2, dv2 = float(a2g(cosmo, a0)), float(a2g2(cosmo, a0)), float(a2dg2dg(cosmo, a0))
        co = [t.tolist() for t in bullfrog_coefficients(cosmo, a0, a1, ns)[:4]]
        dp, vl, lt = ops.lpt(torch.tensor(dk, device=dev), q, d1, d2, dv2, 2, 1, tape=True)
        # both engines carry displacements from the lattice sites (mcpm_engine_set_relative / mcpm_frame)
        pf, vf = dp.clone(), vl.clone()
        stp = ops.nbody_steps(pf, vf, shape, *co, tape=True, lattice=shape)
        sl = slice(rank * pm.npl, (rank + 1) * pm.npl)
        g = torch.Generator(device=dev).manual_seed(5)
        pb, vb = torch.randn(q.shape, device=dev, generator=g), torch.randn(q.shape, device=dev, generator=g)
        dkbar = pm.nbody_backward(tape, pb[sl].contiguous(), vb[sl].contiguous())
        pbf, vbf = pb.clone(), vb.clone()
        ops.nbody_steps_vjp(pbf, vbf, shape, *co, stp, lattice=shape)
        pl, vl2 = pb[sl].clone(), vb[sl].clone()
        pm.steps_backward(tape[1], pl, vl2, *tape[2])
        ref = ops.lpt_vjp(q, dk.shape, d1, d2, dv2, pbf, vbf, lt, 2, 1)[:, pm.y0:pm.y0 + pm.kyl, :]
        errs = torch.stack([(pos - pf[sl]).abs().max(), (vel - vf[sl]).norm() / vf[sl].norm(),
                            (pl - pbf[sl]).norm() / pbf[sl].norm(), (dkbar - ref).norm() / ref.norm()]).to(torch.float64)
        if world > 1:
            dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        out["check"] = {"mesh": n, "max_abs_disp_err_cells": float(errs[0]), "rel_vel_err": float(errs[1]),
                        "rel_posbar_err": float(errs[2]), "rel_dkbar_err": float(errs[3]), "disp_rms": float(pf.std()),
                        "p2p": pm.p2p_note}
        assert errs[0] < 2e-4 and errs[1] < 1e-4 and errs[2] < 2e-3 and errs[3] < 2e-3, out
        del pm, tape, stp, lt

    n = a.mesh
    pm = SlabPM(ops, (n, n, n), halo=min(a.halo, n // world))
    if a.no_force_tape:
        pm.tape_forces = False
    dk = local_delta_k(pm, cosmo, a.cell * n, 1234)
    g = torch.Generator(device=dev).manual_seed(9 + rank)
    pb, vb = torch.randn((pm.npl, 3), device=dev, generator=g), torch.randn((pm.npl, 3), device=dev, generator=g)

    if a.model:
        from montecosmo_b200.dist_model import SlabFieldModel
        mdl = SlabFieldModel(pm, (a.cell * n,) * 3, n_steps=a.nbody_steps, cosmology=cosmo, paint_oversamp=a.oversamp)
        gw = torch.Generator(device=dev).manual_seed(77 + rank)
        white = torch.randn((pm.xl, n, n), device=dev, generator=gw)
        obs = mdl.predict(torch.randn((pm.xl, n, n), device=dev, generator=gw)) \
            + torch.randn((pm.xl, n, n), device=dev, generator=gw)
        del dk
        if a.model_check:
            # Every rank computes the single-GPU FieldModel on the same GLOBAL white field / observation (same seed on
            # identical devices gives the same sequence) and compares its own slab; errors are the max over ranks.
            try:
                from montecosmo_b200.model import FieldModel
                nc = a.model_check_mesh or n
                pmc = pm if nc == n else SlabPM(ops, (nc, nc, nc), halo=min(a.halo, nc // world))
                mdc = mdl if nc == n else SlabFieldModel(pmc, (a.cell * nc,) * 3, n_steps=a.nbody_steps, cosmology=cosmo,
                                                         paint_oversamp=a.oversamp)
                gg = torch.Generator(device=dev).manual_seed(4242)
                gwhite, gtruth, gnoise = (torch.randn((nc, nc, nc), device=dev, generator=gg) for _ in range(3))
                ref = FieldModel((nc, nc, nc), (a.cell * nc,) * 3, "nbody", n_steps=a.nbody_steps, cosmology=cosmo,
                                 paint_oversamp=a.oversamp, out_shape="mesh")
                gobs = ref.evolve(gtruth).detach() + gnoise
                lp_ref, f_ref = ref.value_and_force(gwhite, gobs)
                sl = slice(pmc.x0, pmc.x0 + pmc.xl)
                lp, f = mdc.value_and_force(gwhite[sl].contiguous(), gobs[sl].contiguous())
                fr = f_ref[sl]
                sums = torch.stack([((f - fr) ** 2).sum(), (fr ** 2).sum(), (f * fr).sum(), (f ** 2).sum()]).double()
                if world > 1:
                    dist.all_reduce(sums)
                out["model_check"] = {"mesh": nc, "cell_mpc_h": a.cell, "ranks": world, "logp": float(lp),
                                      "logp_ref": float(lp_ref),
                                      "rel_logp_err": abs(float(lp) - float(lp_ref)) / abs(float(lp_ref)),
                                      "rel_force_err": float((sums[0] / sums[1]).sqrt()),
                                      "force_cosine": float(sums[2] / (sums[1] * sums[3]).sqrt())}
                del ref, lp_ref, f_ref, lp, f, gwhite, gtruth, gnoise, gobs
                if nc != n:
                    del mdc, pmc
                torch.cuda.empty_cache()
            except Exception as e:  # the timing below still runs
                out["model_check"] = {"error": f"{type(e).__name__}: {e}"}

    def step():
        if a.model:
            return mdl.value_and_force(white, obs)
        pos, vel, tape = pm.nbody_forward(dk, cosmo, 0.0, 1.0, a.nbody_steps)
        return pm.nbody_backward(tape, pb, vb)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
    barrier()
    if a.auto_halo:
        sched = pm.halo_schedule(a.nbody_steps)
        pm.set_halo_schedule(sched)
        out["halo_schedule"] = sched
        step()  # one evaluation under the schedule before timing (and its guard check)
        barrier()
    if pm.sections.on:
        pm.sections.report()  # drop the warm-up's sections
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        r = step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    mem = torch.tensor([torch.cuda.max_memory_allocated() / 2 ** 30], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(mem, op=dist.ReduceOp.MAX)
    if rank == 0 and pm.sections.on:
        out["sections_ms_total_over_timed_steps"] = pm.sections.report()
    if rank == 0:
        per = float(ms) / a.steps
        # NVLink bytes out of each GPU per evaluation: all-to-alls (8 B * N/P * (P-1)/P per transform) + halo planes
        nfft = 13 + 13 + (6 if (pm.two_field and pm.p2p) else 8) * a.nbody_steps  # two-field step-loop transforms (dist.py)
        a2a = nfft * 8 * (n ** 3 / 2) / world * (world - 1) / world * 1.004
        halo = a.nbody_steps * (2 + 2 * 4 + 2 * 4 + 2) * pm.H * n * n * 4
        if a.model:  # + white rfftn / its transpose, bias irfftn / its transpose, final paint: 2 R2C + 1 C2R and transposes
            nfft += 2 + 2 + 2 * 3 + 2
        out.update({"metric": ("slab-decomposed grad(log-density) of the field-level model, evaluations/s" if a.model else
                               "slab-decomposed nbody_bf forward + reverse sweep, evaluations/s"), "mesh": n, "n_gpus": world,
                    "value": 1e3 / per, "unit": "evals/s", "ms_per_eval": per, "nbody_steps": a.nbody_steps,
                    "halo_planes": pm.H, "force_tape": bool(pm.tape_forces), "paint_oversamp": a.oversamp if a.model else None, "max_mem_GiB": float(mem), "fused_x_transform": bool(pm.xfuse), "p2p": pm.p2p_note,
                    "nvlink_GB_out_per_gpu_per_eval": (a2a + halo) / 1e9 if world > 1 else 0.0,
                    "nvlink_GBps_per_gpu_if_all_time_were_comm": (a2a + halo) / 1e9 / (per * 1e-3) if world > 1 else 0.0})
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
