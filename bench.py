#!/usr/bin/env python
"""
bench.py -- grad(log-density) evaluations per second of the field-level PM model (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            this engine (CUDA, libmcpm.so through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  CPU implementation of the same path, rank 0 only

A "step" is one evaluation of (logpdf, d logpdf / d white) of the model in montecosmo_b200/model.py on one synthetic
white-noise field: BASELINE configs[2] ("C3", the configuration the metric is quoted on) -- 256^3 mesh / 256^3 particles,
640 Mpc/h box, 2LPT + 10 BullFrog steps, linear bias + flat-sky RSD, interlaced deconvolved CIC paint, Gaussian
likelihood.

N = 1: `FieldModel` on one GPU (the whole evaluation replayed from a CUDA graph).
N > 1: the SLAB-DECOMPOSED model (`dist_model.SlabFieldModel` on `dist.SlabPM`) with real exchanges on the data path --
halo planes after every paint and before every readout, the distributed FFT transposes (inside the fused x-transform
kernel over NVLink peer memory in the step loop, NCCL all-to-all elsewhere), an all-reduced log-density.  Weak scaling:
every rank holds 256^3 cells and 256^3 particles (mesh 512x256x256 on 2, 512x512x256 on 4, 512^3 on 8, same 2.5 Mpc/h
cells), and `value` counts 256^3-equivalents: N x evaluations/s of the N-times-larger mesh.  The replica mode of round 1
(one independent chain per GPU, the reference's own multi-device mode, script.py:13-20) is reported under "replicas".

One JSON line is printed by rank 0 (contract in the task statement): value = device-resident throughput, e2e = the same
through pinned host buffers (H2D of the white field and D2H of gradient + value inside the timed region), roofline for
the dominant kernel measured live with CUDA events, cpu_baseline = the CPU port really run at 256^3 on this box's cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "grad(logp) evals/s at 256^3 mesh"
UNIT = "evals/s"
CELL = 2.5  # Mpc/h
SLAB_SHAPES = {1: (256, 256, 256), 2: (512, 256, 256), 4: (512, 512, 256), 8: (512, 512, 512)}


def workload(shape):
    shape = (shape,) * 3 if isinstance(shape, int) else tuple(shape)
    return dict(mesh_shape=shape, box_size=tuple(CELL * s for s in shape), evolution="nbody", n_steps=10, a_start=0.0,
                a_obs=1.0, lpt_order=2, paint_order=2, interlace_order=2, paint_deconv=True, b1=1.0, rsd=True,
                sigma_obs=1.0)


def config_dict(shape, n_gpus, parallelism=None):
    shape = (shape,) * 3 if isinstance(shape, int) else tuple(shape)
    cells = int(np.prod(shape))
    par = parallelism or ("single GPU" if n_gpus == 1 else f"slab x{n_gpus}")
    per = cells // n_gpus
    side = round(per ** (1 / 3))
    return {"workload": f"C3: {side}^3 mesh / {side}^3 particles per GPU ({'x'.join(map(str, shape))} mesh on {n_gpus} GPU(s), "
                        f"{CELL:g} Mpc/h cells), 2LPT + 10-step BullFrog (DKD), linear Lagrangian bias b1=1 + flat-sky RSD, "
                        "interlaced (x2) deconvolved CIC paint, Gaussian likelihood; one step = value and gradient of "
                        "log-density w.r.t. the white field",
            "mesh": list(shape), "particles": cells, "n_body_steps": 10, "parallelism": par,
            "l2": "working set per step (>6 GB of particle/mesh tape) exceeds the 126 MB L2; no explicit flush"}


# ---------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------- CPU arm
def _host_threads():
    """All host cores for the CPU legs.  torchrun exports OMP_NUM_THREADS=1 to its workers, which throttled round 1's
    CPU arm threefold: the OpenMP runtime of the CPU port is told explicitly, before and after it loads."""
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    return cores


def _set_omp(cores):
    import ctypes
    for name in ("libgomp.so.1", "libgomp.so"):
        try:
            ctypes.CDLL(name).omp_set_num_threads(int(cores))
            return
        except OSError:
            continue


def cpu_port_arm(steps, warmup, n=256):
    """CPU implementation of the path, REALLY RUN at the benchmark size: the engine's own kernel sources compiled for the
    host (oracle/cpu_port.py: every kernel body an OpenMP loop, FFTs by pocketfft, float32), driven through the same
    `FieldModel`.  It is not the reference's JAX (not installable here) and not an independent implementation -- the
    independent float64 restatement is timed beside it at C1 (`oracle_f64_c1`) as the anchor SURVEY 8d asks for."""
    cores = _host_threads()
    import torch
    from oracle import cpu_port
    import montecosmo_b200.nbody as nb
    from montecosmo_b200.model import FieldModel
    torch.set_num_threads(cores)
    saved = nb._OPS
    nb._OPS = cpu_port.cpu_ops()  # the checker's operator table, for this leg only
    _set_omp(cores)
    try:
        model = FieldModel(**workload(n))
        gen = torch.Generator().manual_seed(0)
        obs = 1.0 + torch.randn(model.mesh_shape, generator=gen)
        times = []
        for i in range(warmup + steps):
            white = torch.randn(model.mesh_shape, generator=gen)
            t0 = time.perf_counter()
            lp, g = model.value_and_force(white, obs)
            float(lp)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    finally:
        nb._OPS = saved
    per_eval = float(np.mean(times))
    return {"value": 1.0 / per_eval, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(times)} grad evaluation(s) of the SAME {n}^3 model (no extrapolation) after {warmup} warm-up, "
                      f"{cores} host threads: the engine's kernel sources built for the host (OpenMP loops + pocketfft, "
                      f"float32), {per_eval:.2f} s each",
            "seconds_per_eval": per_eval, "mesh": n, "evaluations": len(times)}


def oracle_anchor():
    """The independent float64 restatement (oracle/pm_oracle.py, torch autograd) timed at C1 (64^3, 5 steps): the
    SURVEY 8d anchor.  One evaluation, all host threads."""
    cores = _host_threads()
    import torch
    from oracle import model_oracle as MO
    from oracle import pm_oracle as O
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_full_size_fixture as FX
    torch.set_num_threads(cores)
    shape, white, obs = FX.inputs(64, 0)
    transfer = FX.transfer_mesh(shape, (640.0,) * 3)
    kw = dict(FX.KW, n_steps=5)
    t0 = time.perf_counter()
    MO.value_and_force(white, obs, transfer, O.Cosmology(), shape, **kw)
    return {"config": "C1: 64^3, 640 Mpc/h, 2LPT + 5 BullFrog steps, bias + RSD + interlaced paint", "dtype": "f64",
            "seconds_per_eval": time.perf_counter() - t0, "cores": cores,
            "what": "oracle/pm_oracle.py + model_oracle.py (torch-CPU float64, autograd gradient): the independent "
                    "restatement of the reference algorithm"}


def max_over_ranks(ms, device, world):
    """Step time of the job = the slowest rank's (all-reduce MAX); identity for one rank."""
    if world <= 1:
        return float(ms)
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, min(args.steps, 4)), min(args.warmup, 1)
    # MCPM_BENCH_SAMPLE: test hook (tests/test_bench_contract.py runs this arm at 32^3 on a CPU-only box); the line then
    # says which mesh was really run.  The driver runs without it: 256^3, measured, no extrapolation.
    n = int(os.environ.get("MCPM_BENCH_SAMPLE", "256"))
    cb = cpu_port_arm(steps, warm, n)
    if n == 256:
        try:
            cb["oracle_f64_c1"] = oracle_anchor()
        except Exception as e:
            cb["oracle_f64_c1"] = {"error": f"{type(e).__name__}: {e}"}
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": 1e3 / cb["value"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(n, 1, "host CPU, all cores (rank 0 only)"), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference's JAX cannot be installed in this image (no jax / jax_cosmo / numpyro wheels); this arm "
                    "times the CPU port of the same path -- the engine's own kernel sources as OpenMP loops -- at the "
                    f"full 256^3 size on all {cb['cores']} host threads (requested steps {args.steps} / warmup "
                    f"{args.warmup} capped at {steps} / {warm}: one evaluation takes seconds); the independent float64 "
                    "oracle is timed at C1 under cpu_baseline.oracle_f64_c1"}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------- roofline
def _timer(flush):
    import torch

    def t(fn, reps=5):
        fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.mean(ts))
    return t


def kernel_rooflines(model, white, peaks):
    """Time the engine's particle kernels in isolation on the evolved (a = 1) state of this run, CUDA events on the
    launching stream, L2 flushed between launches; returns per-kernel rows and the dominant one (time x launches per step).
    Positions are displacements from the lattice sites, as the model carries them."""
    import ctypes as C
    import torch
    from montecosmo_b200 import nbody as nb
    from montecosmo_b200._capi import frame as make_frame
    o = nb.ops()
    shape = model.mesh_shape
    N = int(np.prod(shape))
    with torch.no_grad():
        dk = model.linear_field(nb._f32(white))
        pos, vel = nb.nbody_bf(model.cosmology, dk, model.q, model.a_start, model.a_obs, model.n_steps, ptcl_shape=shape,
                               relative=True, snapshots=3)  # start, middle (in growth time: the run's step 5) and end
        pos_mid = pos[1].contiguous()
        pos, vel = pos[-1].contiguous(), vel[-1].contiguous()
        _, fm = o.pm_forces(pos, shape, want_meshes=True, lattice=shape)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=pos.device)
    t = _timer(flush)
    mesh = torch.empty(shape, device=pos.device)
    A, lib = o.A, o.lib
    p2, v2 = pos.clone(), vel.clone()
    rho = torch.randn(shape, device=pos.device)
    steps = model.n_steps
    rows = []
    traffic, tsrc = {}, None
    for name in ("r2_traffic.json", "r1_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            traffic, tsrc = json.load(open(path)), name
            break

    def add(key, name, fn, alg_bytes, launches, note, in_step=True, fn_mid=None):
        """fn_mid: the same launch on the run's mid-point particles.  The scatters' time depends on the clustering (same-cell
        collisions of the shared-memory atomics: 0.35 ms early, 0.44 at a = 1 for paint3), so they are timed on both states
        and reported as the mean -- what the 10 steps of an evaluation see on average."""
        ms = t(fn)
        states = None
        if fn_mid is not None:
            states = {"a_obs": ms, "mid_run": t(fn_mid)}
            ms = 0.5 * (states["a_obs"] + states["mid_run"])
        rows.append({"kernel": name, "key": key, "ms": ms, "launches_per_step": launches, "alg_bytes": alg_bytes,
                     "achieved_GBps": alg_bytes / ms / 1e6, "frac": alg_bytes / ms / 1e6 / peaks,
                     "ms_per_step_total": ms * launches, "alg_bytes_note": note, "in_step": in_step,
                     "traffic": traffic.get(key, {}).get("dram_bytes"), "ms_by_state": states})

    st = A.stream()
    fr = make_frame(shape)
    frp = C.byref(fr)
    fm4 = torch.empty((*shape, 4), device=pos.device)
    lib.mcpm_interleave3(st, fm.data_ptr(), fm4.data_ptr(), N)
    mesh4 = torch.zeros((*shape, 4), device=pos.device)
    planar3 = torch.empty((3, *shape), device=pos.device)
    xbar, vbar = torch.randn_like(pos), torch.randn_like(pos)
    one = (C.c_float * 3)(1.0, 1.0, 1.0)
    add("brick_paint", "brick paint (CIC density: shared-memory tile, fixed-point ATOMS, red.v4 flush)",
        lambda: (mesh.zero_(), lib.mcpm_paint_brick_f(st, frp, *shape, pos.data_ptr(), 0, 1.0, 0.0, N, *shape,
                                                      mesh.data_ptr())),
        16 * N, steps + 2, "pos 12N + mesh 4N (includes the 4N memset)",
        fn_mid=lambda: (mesh.zero_(), lib.mcpm_paint_brick_f(st, frp, *shape, pos_mid.data_ptr(), 0, 1.0, 0.0, N, *shape,
                                                             mesh.data_ptr())))
    add("brick_paint3", "brick paint3 (reverse-step scatter of the 3 channels of beta * vbar; streaming kernel: persistent CTAs, bulk-copy staged rows)",
        lambda: (planar3.zero_(), lib.mcpm_paint3_brick_f(st, frp, *shape, pos.data_ptr(), vbar.data_ptr(), 0, 0.0, 0.5,
                                                          N, *shape, planar3.data_ptr())),
        36 * N, steps, "SURVEY 8d's paint3: pos 12N + vbar 12N + 3 meshes 12N (the 12N memset is timed too; round 1's "
                       "60N counted the vbar += xbar*drift update that now rides in read_grad4v's epilogue)",
        fn_mid=lambda: (planar3.zero_(), lib.mcpm_paint3_brick_f(st, frp, *shape, pos_mid.data_ptr(), vbar.data_ptr(), 0, 0.0,
                                                                 0.5, N, *shape, planar3.data_ptr())))
    add("kick_drift4", "kick_drift4 (float4 force readout + kick + drift)",
        lambda: lib.mcpm_kick_drift4_f(st, frp, p2.data_ptr(), v2.data_ptr(), fm4.data_ptr(), N, *shape, 1.0, 0.0, 0.0),
        64 * N, steps, "pos 12N r/w + vel 12N r/w + mesh4 16N")
    add("read_grad4v", "read_grad4v (reverse-step gradient gather, fused xbar / vbar update)",
        lambda: lib.mcpm_read_grad4v_step_f(st, frp, pos.data_ptr(), fm4.data_ptr(), rho.data_ptr(), vbar.data_ptr(), 0.5,
                                            1.0, 1e-3, N, *shape, xbar.data_ptr()), 80 * N, steps,
        "pos 12N + vbar 12N r/w + xbar 12N r/w + mesh4 16N + rhobar 4N")
    add("interleave3", "interleave3 (3 planar meshes -> float4 mesh)",
        lambda: lib.mcpm_interleave3(st, fm.data_ptr(), fm4.data_ptr(), N), 28 * N, steps, "12N r + 16N w")
    mk = o.rfftn(rho)
    mk3 = torch.empty((3, *mk.shape), dtype=mk.dtype, device=pos.device)
    fused = bool(lib.mcpm_xfuse_force_slab(st, mk.data_ptr(), mk3.data_ptr(), *shape, shape[1], 0, 0, 0, 0.0, 0, 1.0) == 0)
    if fused:
        add("xfuse_force", "xfuse_force (x-FFT + force kernel + 3 inverse x-FFTs, 1 -> 3 spectra)",
            lambda: lib.mcpm_xfuse_force_slab(st, mk.data_ptr(), mk3.data_ptr(), *shape, shape[1], 0, 0, 0, 0.0, 0, 1.0),
            16 * N, steps, "4N r + 12N w")
        add("xfuse_force_T", "xfuse_force_T (3 x-FFTs + transposed force kernel + inverse x-FFT, 3 -> 1 spectra)",
            lambda: lib.mcpm_xfuse_force_T_slab(st, mk3.data_ptr(), mk.data_ptr(), *shape, shape[1], 0, 0, 0, 0.0, 0, 1.0),
            16 * N, steps, "12N r + 4N w")
    add("force_spectra", "force_spectra (Green x gradient streaming pass, 1 -> 3 spectra; unused on the fused path)",
        lambda: o.force_spectra(mk), 16 * N, 0 if fused else 2 * (steps + 1), "4N r + 12N w", not fused)
    add("cufft_r2c", "cuFFT 3-D R2C (library; 3 passes)", lambda: o.rfftn(rho), 8 * N, 8 if fused else 116,
        "4N r + 4N w per transform; per evaluation on the fused path: 8 3-D transforms (initial field, final paint, "
        "likelihood) + 48 batched 2-D (y,z) launches covering 108 mesh transforms", False)
    # generic global-atomic kernels the step loop does not run under the lattice hint: listed for comparison only
    add("paint_generic", "paint (generic CIC scatter, 8 red.f32 / particle)",
        lambda: lib.mcpm_paint_f(st, frp, pos.data_ptr(), 0, 1.0, N, *shape, 2, one, 0.0, mesh.data_ptr(), 0), 16 * N, 0,
        "pos 12N + mesh 4N", False)
    add("paint3v4_generic", "paint3v4 (generic reverse-step scatter, 8 red.v4.f32 / particle)",
        lambda: lib.mcpm_paint3v4_f(st, frp, pos.data_ptr(), vbar.data_ptr(), xbar.data_ptr(), 1e-3, 0.5, N, *shape,
                                    mesh4.data_ptr()), 64 * N, 0, "pos 12N + vbar 12N r + 12N w + xbar 12N + mesh4 16N",
        False)
    dom = max([r for r in rows if r["in_step"]], key=lambda r: r["ms_per_step_total"])
    return rows, dom, tsrc, pos


def paint_microbench(model, disp256, peaks):
    """SURVEY 8d's paint-only microbenchmark P: CIC density paint at 256^3 and 512^3, particles = cells, on (a) the
    lattice + Gaussian jitter sigma = 1.5 cells and (b) the evolved a = 1 state (clustered: stresses the atomics), with
    the brick-tiled kernel (what a PM run takes) and the generic global-atomic kernel (what a catalogue takes).
    Gparticles/s, mesh clear included, L2 flushed between launches; 16 B/particle algorithmic."""
    import ctypes as C
    import torch
    from montecosmo_b200 import nbody as nb
    from montecosmo_b200._capi import frame as make_frame
    from montecosmo_b200.model import FieldModel
    o = nb.ops()
    A, lib = o.A, o.lib
    dev = A.device
    st = A.stream()
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    t = _timer(flush)
    one = (C.c_float * 3)(1.0, 1.0, 1.0)
    out = []
    for n in (256, 512):
        shape = (n, n, n)
        N = n ** 3
        fr = make_frame(shape)
        frp = C.byref(fr)
        mesh = torch.empty(shape, device=dev)
        gen = torch.Generator(device=dev).manual_seed(5 + n)
        cases = {"jitter_sigma1.5": torch.randn((N, 3), device=dev, generator=gen) * 1.5}
        try:
            if n == 256:
                cases["evolved_a1"] = disp256
            else:
                with torch.no_grad():
                    m = FieldModel(**workload(n))
                    dk = m.linear_field(torch.randn(shape, device=dev, generator=gen))
                    d, _ = nb.nbody_bf(m.cosmology, dk, m.q, 0.0, 1.0, 10, ptcl_shape=shape, relative=True)
                    cases["evolved_a1"] = d[-1].contiguous()
                    del m, dk, d
        except Exception as e:
            out.append({"mesh": n, "case": "evolved_a1", "error": f"{type(e).__name__}: {e}"})
        for case, disp in cases.items():
            row = {"mesh": n, "case": case, "disp_rms_cells": float(disp.std())}
            for kern, fn in (("brick", lambda: (mesh.zero_(), lib.mcpm_paint_brick_f(
                                 st, frp, *shape, disp.data_ptr(), 0, 1.0, 0.0, N, *shape, mesh.data_ptr()))),
                             ("generic", lambda: lib.mcpm_paint_f(st, frp, disp.data_ptr(), 0, 1.0, N, *shape, 2, one, 0.0,
                                                                  mesh.data_ptr(), 0))):
                ms = t(fn, reps=3)
                row[kern] = {"ms": ms, "Gparticles_per_s": N / ms / 1e6, "frac_of_hbm": 16 * N / ms / 1e6 / peaks}
            out.append(row)
        del cases, mesh
        torch.cuda.empty_cache()
    o._engines.pop((512, 512, 512), None)
    torch.cuda.empty_cache()
    return out


def small_configs(dev, gen):
    """The smaller BASELINE configs, for the record (not the metric): C1 64^3 / 5 steps (run/infer_example.py scale) and
    C2 128^3 / 10 steps, each inside a 100-iteration leapfrog loop with 2 gradient evaluations per iteration (as
    isokinetic_mclachlan, samplers.py:350-351), on the graph-replayed evaluation."""
    import torch
    from montecosmo_b200.model import FieldModel
    out = {}
    try:
        for tag, nn, nsteps in (("C1_64", 64, 5), ("C2_128_leapfrog", 128, 10)):
            wl = workload(nn)
            wl["box_size"] = (640.0,) * 3
            wl["n_steps"] = nsteps
            mm = FieldModel(**wl)
            ff = mm.graphed_value_and_force(1.0 + torch.randn(mm.mesh_shape, device=dev, generator=gen))
            q = torch.randn(mm.mesh_shape, device=dev, generator=gen)
            p, eps, iters = torch.zeros_like(q), 1e-3, 100
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                p.add_(ff(q)[1], alpha=0.5 * eps)
                q.add_(p, alpha=0.5 * eps)
                p.add_(ff(q)[1], alpha=0.5 * eps)
                q.add_(p, alpha=0.5 * eps)
            e1.record()
            torch.cuda.synchronize()
            out[tag] = {"mesh": nn, "n_body_steps": nsteps, "evals_per_s": 2 * iters / (e0.elapsed_time(e1) * 1e-3),
                        "iterations": iters}
            del ff, mm
        torch.cuda.empty_cache()
    except Exception as e:
        out["error"] = f"{type(e).__name__}: {e}"
    return out


# ---------------------------------------------------------------------------------------------------- engine arm
def run_engine(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's version banner (printed when the box sets NCCL_DEBUG) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from montecosmo_b200 import nbody as nb
    from montecosmo_b200.model import FieldModel
    lib = nb.ops().lib
    n = args.mesh
    dev = nb.ops().A.device
    K, W = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn):
        for i in range(W):
            step_fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            step_fn(i)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1), dev, world)

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    sampler = ClockSampler(local)
    slab = None
    if world > 1 and not args.replicas:
        slab = run_slab(args, world, rank, dev, timed, barrier, sampler)

    # ---- the single-GPU model: the metric at N = 1; per-rank replicas (no collective) at N > 1, for the record
    model = FieldModel(**workload(n))
    N = n ** 3
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    with torch.no_grad():  # observation: the model's own output at another seed + unit noise
        truth = model.evolve(torch.randn(model.mesh_shape, device=dev, generator=gen))
        obs = (truth + torch.randn(model.mesh_shape, device=dev, generator=gen)).contiguous()
    del truth
    whites = [torch.randn(model.mesh_shape, device=dev, generator=gen) for _ in range(2)]
    # kernels of ours per evaluation, counted on one eager evaluation (a graph replay launches the same nodes)
    model.value_and_force(whites[0], obs)
    torch.cuda.synchronize()
    lib.mcpm_launch_count(1)
    model.value_and_force(whites[0], obs)
    torch.cuda.synchronize()
    launches_per_eval = int(lib.mcpm_launch_count(0))
    fn, graph_note = None, "disabled (--no-graph)"
    if not args.no_graph:
        try:
            fn, graph_note = model.graphed_value_and_force(obs), "ok"
        except Exception as e:  # never lose the measurement to the capture: time the eager call instead
            fn, graph_note = None, f"capture failed, eager call timed instead ({type(e).__name__}: {e})"
            torch.cuda.synchronize()
    if rank == 0 and slab is None:
        sampler.start()
    keep = {}

    def step_dev(i):
        keep["out"] = fn(whites[i % 2]) if fn is not None else model.value_and_force(whites[i % 2], obs)
    ms_dev = timed(step_dev)
    h_white = [torch.randn(model.mesh_shape, generator=torch.Generator().manual_seed(7 + i)).pin_memory() for i in range(2)]
    h_grad = torch.empty(model.mesh_shape, dtype=torch.float32).pin_memory()
    h_lp = torch.empty((), dtype=torch.float64).pin_memory()
    d_white = torch.empty(model.mesh_shape, device=dev)

    def step_e2e(i):
        if fn is not None:
            fn.static_white.copy_(h_white[i % 2], non_blocking=True)
            fn.graph.replay()
            lp, g = fn.out
        else:
            d_white.copy_(h_white[i % 2], non_blocking=True)
            lp, g = model.value_and_force(d_white, obs)
        h_grad.copy_(g, non_blocking=True)
        h_lp.copy_(lp, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller needs the result before proposing the next state
    ms_e2e = timed(step_e2e)
    ms_eager = timed(lambda i: model.value_and_force(whites[i % 2], obs)) if fn is not None else ms_dev
    clocks = sampler.stop() if rank == 0 else None
    lp_val = float(h_lp)
    single = {"value": world * K / (ms_dev * 1e-3), "ms_per_step": ms_dev / K,
              "e2e": {"value": world * K / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 4 * N,
                      "d2h_bytes_per_step": 4 * N + 8, "ms_per_step": ms_e2e / K},
              "eager": {"value": world * K / (ms_eager * 1e-3), "unit": UNIT, "ms_per_step": ms_eager / K}}
    other = small_configs(dev, gen) if (rank == 0 and fn is not None and n == 256 and world == 1) else {}

    # Run-to-run spread of the result on identical inputs (the order in which float atomics land is not
    # deterministic): two eager evaluations outside the timed region, relative L2 of the gradient difference.
    spread = None
    if rank == 0:
        try:
            lp_a, g_a = model.value_and_force(whites[0], obs)
            g_a, lp_a = g_a.clone(), float(lp_a)
            lp_b, g_b = model.value_and_force(whites[0], obs)
            spread = {"rel_l2_grad": float((g_b - g_a).norm() / g_a.norm()),
                      "rel_logp": abs(float(lp_b) - lp_a) / max(abs(lp_a), 1e-300), "evaluations": 2}
        except Exception as e:  # never let a diagnostic cost the measurement
            spread = {"error": f"{type(e).__name__}: {e}"}
    line = None
    if rank == 0:
        del fn, keep
        torch.cuda.empty_cache()
        rows, dom, tsrc, disp = kernel_rooflines(model, whites[0], peak)
        pmb = None
        if world == 1 and not args.no_paint_bench:
            try:
                pmb = paint_microbench(model, disp, peak)
            except Exception as e:
                pmb = [{"error": f"{type(e).__name__}: {e}"}]
        cb = None
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_port_arm(1, 1, 256)
            try:
                cb["oracle_f64_c1"] = oracle_anchor()
            except Exception as e:
                cb["oracle_f64_c1"] = {"error": f"{type(e).__name__}: {e}"}
        main = slab if slab is not None else single
        shape = SLAB_SHAPES[world] if slab is not None else (n, n, n)
        line = {"metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": config_dict(shape, world, None if slab is not None or world == 1 else
                                      f"replicas x{world} (one chain per GPU, no collective)"),
                "clocks": clocks, "e2e": main["e2e"],
                "gpu_launches": int((slab["launches_per_eval"] if slab is not None else launches_per_eval) * K),
                "gpu_launches_note": "engine kernels per evaluation on this rank (mcpm_launch_count on an eager "
                                     "evaluation) x steps; cuFFT's and NCCL's kernels not counted",
                "graph": graph_note if slab is None else "not used (the slab model interleaves NCCL / symmetric-memory "
                                                         "barriers with the kernels; eager launches)",
                "api": ("SlabFieldModel.value_and_force (slab-decomposed, one process per GPU)" if slab is not None else
                        "FieldModel.value_and_force (eager)" if graph_note != "ok" else
                        "FieldModel.graphed_value_and_force (CUDA graph of the whole evaluation, replayed per step)"),
                "roofline": {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_GBps"], "peak": peak,
                             "peak_source": peak_src, "unit": "GB/s", "frac": dom["frac"],
                             "traffic": dom["traffic"],
                             "traffic_source": (f"dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full "
                                                f"capture (profiles/{tsrc}; a frozen capture, re-taken when the kernel "
                                                f"changes -- see profiles/README.md)") if dom["traffic"] else None,
                             "alg_bytes_per_launch": dom["alg_bytes"], "ms_per_launch": dom["ms"],
                             "share_of_step": dom["ms_per_step_total"] / single["ms_per_step"],
                             "whole_evaluation": {"alg_bytes": 55e9, "frac": 55e9 / (single["ms_per_step"] * 1e-3) / 1e9 / peak,
                                                  "note": "SURVEY 8d's 3.3 kN bytes for C3 over the single-GPU step time"}},
                "kernels": rows, "cpu_baseline": cb, "logp_last": lp_val, "other_configs": other,
                "paint_Gparticles_per_s": N / (rows[0]["ms"] * 1e-3) / 1e9,
                "paint_kernel": rows[0]["kernel"], "paint_microbench": pmb, "run_to_run": spread}
        if slab is not None:
            line["roofline"]["nvlink"] = slab["nvlink"]
            line["slab"] = {k: v for k, v in slab.items() if k not in ("e2e", "nvlink")}
            line["replicas"] = dict(single, note="one independent 256^3 chain per GPU (round 1's --gpus N mode, the "
                                                 "reference's pmap over chains, script.py:13-20): no data-path collective")
        else:
            line["eager"] = single["eager"]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def run_slab(args, world, rank, dev, timed, barrier, sampler):
    """N > 1: the slab-decomposed model, weak-scaled (256^3 cells and particles per rank)."""
    import torch
    import torch.distributed as dist
    from montecosmo_b200 import nbody as nb
    from montecosmo_b200.dist import SlabPM
    from montecosmo_b200.dist_model import SlabFieldModel
    ops = nb.ops()
    lib = ops.lib
    shape = SLAB_SHAPES[world]
    K = args.steps
    pm = SlabPM(ops, shape, halo=min(args.halo, shape[0] // world))
    wl = workload(shape)
    mdl = SlabFieldModel(pm, wl["box_size"], n_steps=wl["n_steps"], a_start=wl["a_start"], a_obs=wl["a_obs"],
                         b1=wl["b1"], rsd=wl["rsd"], sigma_obs=wl["sigma_obs"])
    local_shape = (pm.xl, shape[1], shape[2])
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    with torch.no_grad():
        obs = (mdl.predict(torch.randn(local_shape, device=dev, generator=gen))
               + torch.randn(local_shape, device=dev, generator=gen)).contiguous()
    whites = [torch.randn(local_shape, device=dev, generator=gen) for _ in range(2)]
    h_white = [torch.randn(local_shape, generator=torch.Generator().manual_seed(7 + i + 10 * rank)).pin_memory()
               for i in range(2)]
    fields = whites + [h.to(dev) for h in h_white]  # every white field the timed loops will see
    for f in fields:  # the halo below is sized from ALL of them: a field never seen could outgrow it and trip the guard
        mdl.value_and_force(f, obs)
    torch.cuda.synchronize()
    # Halo sized from a measurement: the default of 24 planes is an a = 1 worst case; the warm-up evaluation's largest
    # x-displacement (max over ranks) x 1.25 + 2 planes is what this run needs -- at 64 owned planes per rank (8 GPUs) 24
    # halo planes are 75 % more cells to clear, flush, exchange and gather.  The guard still raises if a particle ever
    # leaves the extended slab (SlabPM.check_guard): a caller then rebuilds with more planes.
    halo_default, halo_used = pm.H, pm.H
    sched = None
    if not args.fixed_halo:
        need = pm.halo_needed()
        need += need % 2  # an even number of planes
        if need < pm.H:
            del mdl, pm
            torch.cuda.empty_cache()
            pm = SlabPM(ops, shape, halo=need)
            mdl = SlabFieldModel(pm, wl["box_size"], n_steps=wl["n_steps"], a_start=wl["a_start"], a_obs=wl["a_obs"],
                                 b1=wl["b1"], rsd=wl["rsd"], sigma_obs=wl["sigma_obs"])
            for f in fields:
                mdl.value_and_force(f, obs)
            torch.cuda.synchronize()
            halo_used = need
        # ... and per step: the kick positions of step s stay within a few planes early in the run, so the exchanges of
        # that step (and of its reverse step) move only the active planes next to the owned region
        sched = pm.halo_schedule(wl["n_steps"])
        pm.set_halo_schedule(sched)
    del fields
    lib.mcpm_launch_count(1)
    mdl.value_and_force(whites[0], obs)
    torch.cuda.synchronize()
    launches = int(lib.mcpm_launch_count(0))
    # The slab model launches eagerly and draws every intermediate from torch's caching allocator: its first evaluations
    # pay cudaMalloc (which synchronises the device) and NCCL's lazy channel set-up, and on 2 B200s the first timed loop
    # measured 80 ms per evaluation where the second measured 51.  Untimed settling evaluations, the same number on
    # every rank (collectives inside), until the allocator has stopped growing -- at most 8.
    grew = torch.tensor([1.0], device=dev)
    for _ in range(8):
        before = torch.cuda.memory_reserved()
        mdl.value_and_force(whites[_ % 2], obs)
        torch.cuda.synchronize()
        grew[0] = float(torch.cuda.memory_reserved() != before)
        dist.all_reduce(grew, op=dist.ReduceOp.MAX)
        if float(grew) == 0.0:
            break
    for i in range(2):  # two more on the settled allocator (GPU call 17 still saw one 58 ms first loop after the break above)
        mdl.value_and_force(whites[i % 2], obs)
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    keep = {}

    def step_dev(i):
        keep["out"] = mdl.value_and_force(whites[i % 2], obs)
    ms_dev = timed(step_dev)
    nl = int(np.prod(local_shape))
    h_grad = torch.empty(local_shape, dtype=torch.float32).pin_memory()
    h_lp = torch.empty((), dtype=torch.float64).pin_memory()
    d_white = torch.empty(local_shape, device=dev)

    def step_e2e(i):
        d_white.copy_(h_white[i % 2], non_blocking=True)
        lp, g = mdl.value_and_force(d_white, obs)
        h_grad.copy_(g, non_blocking=True)
        h_lp.copy_(lp, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    ms_e2e = timed(step_e2e)
    mem = torch.tensor([torch.cuda.max_memory_allocated() / 2 ** 30], device=dev, dtype=torch.float64)
    dist.all_reduce(mem, op=dist.ReduceOp.MAX)
    # NVLink bytes out of each GPU per evaluation: one exchange of a half spectrum per distributed mesh
    # transform, 8 B x (cells/2) / P x (P-1)/P each, + the halo planes of every paint / readout
    cells = float(np.prod(shape))
    ns = wl["n_steps"]
    # per step 1 + 3 fields forward and 3 + 1 in reverse, or 1 + 2 and 2 + 1 with the two-field transform (dist.py)
    per_step = 6 if (getattr(pm, "two_field", False) and getattr(pm, "p2p", False)) else 8
    nfft = 13 + 13 + per_step * ns + (2 + 2 + 2 * 3 + 2)
    a2a = nfft * 8 * (cells / 2) / world * (world - 1) / world * (1 + 2.0 / shape[2])
    plane = shape[1] * shape[2] * 4.0
    h_steps = float(np.mean(sched)) if sched else float(halo_used)  # mean active planes of the step-loop exchanges
    halo = (ns * (1 + 4 + 3 + 1) * h_steps + 2 * 2 * halo_used) * 2 * plane
    per = ms_dev / K
    nv = (a2a + halo) / 1e9
    del keep, mdl, pm
    torch.cuda.empty_cache()
    barrier()
    return {"value": world * K / (ms_dev * 1e-3), "ms_per_step": per, "mesh": list(shape), "halo_planes": halo_used,
            "halo_schedule": sched,
            "halo_note": (f"sized from the warm-up evaluation's largest x-displacement (x 1.25 + 2 planes; default "
                          f"{halo_default}); the guard raises if a particle leaves its extended slab"
                          if halo_used != halo_default else "default"),
            "launches_per_eval": launches, "max_mem_GiB_per_gpu": float(mem),
            "e2e": {"value": world * K / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 4 * nl * world,
                    "d2h_bytes_per_step": (4 * nl + 8) * world, "ms_per_step": ms_e2e / K,
                    "note": "every rank copies its own planes of the white field in and of the gradient out"},
            "nvlink": {"bound": "nvlink", "GB_out_per_gpu_per_eval": nv, "achieved_GBps_per_gpu": nv / (per * 1e-3),
                       "peak_GBps_per_direction": 900.0, "frac": nv / (per * 1e-3) / 900.0,
                       "note": "bytes out per GPU per evaluation / whole step time: the fraction of NVLink's 900 GB/s the "
                               "exchanges would need if they were spread over the step; transposes "
                               f"{a2a / 1e9:.2f} GB + halos {halo / 1e9:.2f} GB"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--mesh", type=int, default=256, help="mesh side (development only; the contract runs 256)")
    ap.add_argument("--halo", type=int, default=24, help="halo planes of the slab decomposition (N > 1)")
    ap.add_argument("--fixed-halo", action="store_true", help="N > 1: keep --halo planes instead of sizing them from a warm-up evaluation")
    ap.add_argument("--replicas", action="store_true", help="N > 1: time independent replicas only (round 1's mode)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-paint-bench", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time the eager call instead of the CUDA-graph replay")
    args = ap.parse_args()
    _host_threads()  # before anything loads an OpenMP runtime
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
