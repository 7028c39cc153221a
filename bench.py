#!/usr/bin/env python
"""
bench.py -- grad(log-density) evaluations per second of the field-level PM model (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            this engine (CUDA, libmcpm.so through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  CPU restatement of the reference algorithm (oracle/), rank 0 only

A "step" is one evaluation of (logpdf, d logpdf / d white) of the model in montecosmo_b200/model.py on one synthetic
white-noise field.  Workload at every N: BASELINE configs[2] ("C3", the configuration the metric is quoted on): 256^3
mesh / 256^3 particles, 640 Mpc/h box, 2LPT + 10 BullFrog steps, linear bias + flat-sky RSD, interlaced deconvolved
CIC paint, Gaussian likelihood.  With N > 1 every rank runs an independent replica (one chain per GPU, the reference's own
multi-device mode, script.py:13-20): weak scaling, no data-path collective.

One JSON line is printed by rank 0 (contract in the task statement): value = device-resident throughput, e2e = the same
through host buffers (H2D of the white field and D2H of gradient + value inside the timed region), roofline for the
dominant kernel measured live with CUDA events, cpu_baseline = the oracle timed on this box's host cores.
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


def workload(n):
    return dict(mesh_shape=(n, n, n), box_size=(2.5 * n,) * 3, evolution="nbody", n_steps=10, a_start=0.0, a_obs=1.0,
                lpt_order=2, paint_order=2, interlace_order=2, paint_deconv=True, b1=1.0, rsd=True, sigma_obs=1.0)


def config_dict(n, n_gpus):
    return {"workload": f"C3: {n}^3 mesh / {n}^3 particles, {2.5 * n:g} Mpc/h box, 2LPT + 10-step BullFrog (DKD), "
                        "linear Lagrangian bias b1=1 + flat-sky RSD, interlaced (x2) deconvolved CIC paint, Gaussian "
                        "likelihood; one step = value and gradient of log-density w.r.t. the white field",
            "mesh": n, "particles": n ** 3, "n_body_steps": 10, "parallelism": f"replicas x{n_gpus} (one chain per GPU)",
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


# ---------------------------------------------------------------------------------------------------- CPU baseline
def cpu_reference_arm(steps, warmup, n_sample=128):
    """The CPU port of the path (oracle/cpu_port.py: the same kernels as OpenMP loops + pocketfft, float32) on all host
    cores: a bounded sample (n_sample^3) of the 256^3 workload, scaled by the particle-count ratio to the metric's unit."""
    import torch
    from oracle import cpu_port
    import montecosmo_b200.nbody as nb
    from montecosmo_b200.model import FieldModel
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    saved = nb._OPS
    nb._OPS = cpu_port.cpu_ops()  # the checker's operator table, for this leg only
    try:
        model = FieldModel(**workload(n_sample))
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
    scale = (256 / n_sample) ** 3  # particle-count ratio; the FFT's log factor is ignored (favours the CPU)
    return {"value": 1.0 / (per_eval * scale), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(times)} grad evaluation(s) of the same model at {n_sample}^3 on {cores} host threads "
                      f"(OpenMP + pocketfft float32 port, {per_eval:.2f} s each), scaled by ({n_sample}/256)^3 to 256^3",
            "seconds_per_sample_eval": per_eval}


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
    cb = cpu_reference_arm(max(1, min(args.steps, 5)), min(args.warmup, 1), int(os.environ.get("MCPM_BENCH_SAMPLE", "128")))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / cb["value"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(256, args.gpus), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference's JAX cannot be installed in this image; this arm times the CPU restatement of its "
                    "algorithm (oracle/), all host threads, on a bounded sample"}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------- roofline
def kernel_rooflines(model, white, peaks):
    """Time the engine's particle kernels in isolation on the evolved (a = 1) state of this run, CUDA events on the
    launching stream, L2 flushed between launches; returns per-kernel rows and the dominant one (time x launches per step)."""
    import torch
    from montecosmo_b200 import nbody as nb
    o = nb.ops()
    shape = model.mesh_shape
    N = int(np.prod(shape))
    with torch.no_grad():
        dk = model.linear_field(nb._f32(white))
        pos, vel = nb.nbody_bf(model.cosmology, dk, model.q, model.a_start, model.a_obs, model.n_steps)
        pos, vel = pos[-1].contiguous(), vel[-1].contiguous()
        _, fm = o.pm_forces(pos, shape, want_meshes=True)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=pos.device)
    mesh = torch.empty(shape, device=pos.device)
    A, lib = o.A, o.lib
    p2, v2 = pos.clone(), vel.clone()
    rho = torch.randn(shape, device=pos.device)

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

    steps = model.n_steps
    rows = []
    traffic_path = os.path.join(ROOT, "profiles", "r1_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}

    def add(key, name, fn, alg_bytes, launches, note, in_step=True):
        ms = t(fn)
        rows.append({"kernel": name, "key": key, "ms": ms, "launches_per_step": launches, "alg_bytes": alg_bytes,
                     "achieved_GBps": alg_bytes / ms / 1e6, "frac": alg_bytes / ms / 1e6 / peaks,
                     "ms_per_step_total": ms * launches, "alg_bytes_note": note, "in_step": in_step,
                     "traffic": traffic.get(key, {}).get("dram_bytes")})

    st = A.stream()
    nzc = shape[2] // 2 + 1
    fm4 = torch.empty((*shape, 4), device=pos.device)
    lib.mcpm_interleave3(st, fm.data_ptr(), fm4.data_ptr(), N)
    mesh4 = torch.zeros((*shape, 4), device=pos.device)
    planar3 = torch.empty((3, *shape), device=pos.device)
    xbar, vbar = torch.randn_like(pos), torch.randn_like(pos)
    eng = o.engine(shape)
    o.set_lattice(shape, shape)
    add("brick_paint", "brick paint (CIC density: shared-memory tile, fixed-point ATOMS, red.v4 flush)",
        lambda: (mesh.zero_(), lib.mcpm_paint_lattice(eng.handle, st, pos.data_ptr(), 0, 1.0, N, mesh.data_ptr())),
        16 * N, steps + 2, "pos 12N + mesh 4N (includes the 4N memset)")
    add("brick_paint3", "brick paint3 (reverse-step scatter of 3 channels, fused vbar += xbar*drift)",
        lambda: (planar3.zero_(), lib.mcpm_paint3_lattice(eng.handle, st, pos.data_ptr(), vbar.data_ptr(),
                                                           xbar.data_ptr(), 1e-3, 0.5, N, planar3.data_ptr())),
        60 * N, steps, "pos 12N + vbar 12N r + 12N w + xbar 12N + 3 meshes 12N (includes the 12N memset)")
    add("kick_drift4", "kick_drift4 (float4 force readout + kick + drift)",
        lambda: lib.mcpm_kick_drift4(st, p2.data_ptr(), v2.data_ptr(), fm4.data_ptr(), N, *shape, 1.0, 0.0, 0.0),
        64 * N, steps, "pos 12N r/w + vel 12N r/w + mesh4 16N")
    add("read_grad4v", "read_grad4v (reverse-step gradient gather, fused xbar / vbar update)",
        lambda: lib.mcpm_read_grad4v(st, pos.data_ptr(), fm4.data_ptr(), rho.data_ptr(), vbar.data_ptr(), 0.5, 1.0, N,
                                     *shape, xbar.data_ptr()), 80 * N, steps,
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
        lambda: o.paint(pos, shape, None, order=2, out=mesh), 16 * N, 0, "pos 12N + mesh 4N", False)
    add("paint3v4_generic", "paint3v4 (generic reverse-step scatter, 8 red.v4.f32 / particle)",
        lambda: lib.mcpm_paint3v4(st, pos.data_ptr(), vbar.data_ptr(), xbar.data_ptr(), 1e-3, 0.5, N, *shape,
                                  mesh4.data_ptr()), 64 * N, 0, "pos 12N + vbar 12N r + 12N w + xbar 12N + mesh4 16N",
        False)
    dom = max([r for r in rows if r["in_step"]], key=lambda r: r["ms_per_step_total"])
    return rows, dom


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
    model = FieldModel(**workload(n))
    dev = nb.ops().A.device
    N = n ** 3
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    # observation: the model's own output at another seed + unit noise
    with torch.no_grad():
        truth = model.evolve(torch.randn(model.mesh_shape, device=dev, generator=gen))
        obs = (truth + torch.randn(model.mesh_shape, device=dev, generator=gen)).contiguous()
    del truth
    K, W = args.steps, args.warmup
    whites = [torch.randn(model.mesh_shape, device=dev, generator=gen) for _ in range(2)]

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

    # kernels of ours per evaluation, counted on one eager evaluation (a graph replay launches the same nodes)
    model.value_and_force(whites[0], obs)
    torch.cuda.synchronize()
    lib.mcpm_launch_count(1)
    model.value_and_force(whites[0], obs)
    torch.cuda.synchronize()
    launches_per_eval = int(lib.mcpm_launch_count(0))
    # The public call a sampler makes: FieldModel.graphed_value_and_force -- the whole evaluation captured once in a CUDA
    # graph, replayed per step on a new white field (--no-graph: the eager FieldModel.value_and_force).
    fn, graph_note = None, "disabled (--no-graph)"
    if not args.no_graph:
        try:
            fn, graph_note = model.graphed_value_and_force(obs), "ok"
        except Exception as e:  # never lose the measurement to the capture: time the eager call instead
            fn, graph_note = None, f"capture failed, eager call timed instead ({type(e).__name__}: {e})"
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # (1) device-resident
    keep = {}

    def step_dev(i):
        keep["out"] = fn(whites[i % 2]) if fn is not None else model.value_and_force(whites[i % 2], obs)
    ms_dev = timed(step_dev)
    # (2) end to end through host buffers: pinned white in, gradient + value out
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
    # (3) the eager call, for the record
    ms_eager = timed(lambda i: model.value_and_force(whites[i % 2], obs)) if fn is not None else ms_dev
    clocks = sampler.stop() if rank == 0 else None
    lp_val = float(h_lp)
    launches = launches_per_eval * K
    other = small_configs(dev, gen) if (rank == 0 and fn is not None and n == 256) else {}

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
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
        rows, dom = kernel_rooflines(model, whites[0], peak)
        cb = None if args.no_cpu_baseline else cpu_reference_arm(2, 1, 128)
        value = world * K / (ms_dev * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config_dict(n, world), "clocks": clocks,
                "e2e": {"value": world * K / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 4 * N,
                        "d2h_bytes_per_step": 4 * N + 8, "ms_per_step": ms_e2e / K},
                "gpu_launches": int(launches),
                "gpu_launches_note": f"{launches_per_eval} engine kernels per evaluation (mcpm_launch_count on an eager "
                                     "evaluation) x steps; cuFFT's kernels not counted",
                "graph": graph_note,
                "api": "FieldModel.value_and_force (eager)" if fn is None else
                       "FieldModel.graphed_value_and_force (CUDA graph of the whole evaluation, replayed per step)",
                "eager": {"value": world * K / (ms_eager * 1e-3), "unit": UNIT, "ms_per_step": ms_eager / K},
                "roofline": {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_GBps"], "peak": peak,
                             "peak_source": peak_src, "unit": "GB/s", "frac": dom["frac"],
                             "traffic": dom["traffic"],
                             "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full "
                                               "capture (profiles/r1_traffic.json)" if dom["traffic"] else None,
                             "alg_bytes_per_launch": dom["alg_bytes"], "ms_per_launch": dom["ms"],
                             "share_of_step": dom["ms_per_step_total"] / (ms_dev / K)},
                "kernels": rows, "cpu_baseline": cb, "logp_last": lp_val, "other_configs": other,
                "paint_Gparticles_per_s": N / (rows[0]["ms"] * 1e-3) / 1e9,
                "paint_kernel": rows[0]["kernel"], "run_to_run": spread}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--mesh", type=int, default=256, help="mesh side (development only; the contract runs 256)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time the eager call instead of the CUDA-graph replay")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
