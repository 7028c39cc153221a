#!/usr/bin/env python
"""
Full-size oracle fixtures for the benchmarked configurations (TEST INFRASTRUCTURE; run on a CPU box, results committed).

  python tests/golden/make_full_size_fixture.py 128     ->  tests/golden/c2_128.npz   (BASELINE configs[1], C2)
  python tests/golden/make_full_size_fixture.py 256     ->  tests/golden/c3_256.npz   (BASELINE configs[2], C3: the
                                                            configuration bench.py's metric is quoted on)

Runs `oracle/model_oracle.py` in float64 on the bench workload -- n^3 mesh / n^3 particles, 640 Mpc/h box, 2LPT + 10
BullFrog steps from a = 0 to 1, linear Lagrangian bias b1 = 1 + flat-sky RSD along z, interlaced (x2) deconvolved CIC
paint, Gaussian likelihood with unit noise, N(0,1) prior (model.py:683-837, nbody.py:967-1002) -- forward AND autograd
gradient (the BullFrog steps are recomputed in the backward pass, so 256^3 fits in RAM), and stores a SMALL fixture:

  logp                      float64 log-density
  mesh_block                the 1 + delta_obs mesh on [0:32]^3
  mesh_coarse               the same mesh block-averaged to 32^3 (every cell enters)
  pk_count, pk_kmean, pk    power spectrum of the mesh in the reference's own binning (metrics.py:121-182)
  disp_sub, vel_sub         final displacement (pos - lattice site, before RSD) and velocity of every `stride`-th particle
  grad_sub                  d logp / d white on the strided subsample [::4, ::4, ::4]  (n/4)^3 values
  grad_block                the same gradient on [0:32]^3
  grad_norm, grad_dot       |grad| and <grad, v_i> for three fixed unit-variance directions v_i = default_rng(200 + i)
  cd_dot                    central differences (logp(w + eps v_i) - logp(w - eps v_i)) / (2 eps) for the same directions
                            (only with --cd; they check the oracle's own autograd.  logp is piecewise smooth -- every
                            particle that crosses a cell face puts a kink in its derivative -- so the differences
                            converge to the autograd value only for eps <~ 1e-7 at 64^3: 1e-6 relative there)

Inputs are regenerated from seeds by both sides: white = default_rng(seed).standard_normal(shape),
obs = 1 + default_rng(100).standard_normal(shape); the transfer mesh is montecosmo_b200.model's tabulated no-wiggle power
(float32-rounded, as the engine holds it).
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import metrics_oracle as MM  # noqa: E402
from oracle import model_oracle as MO  # noqa: E402
from oracle import pm_oracle as O  # noqa: E402

BOX = 640.0
N_STEPS = 10
KW = dict(evolution="nbody", n_steps=N_STEPS, a_start=0.0, a_obs=1.0, lpt_order=2, paint_order=2, interlace_order=2,
          paint_deconv=True, b1=1.0, rsd=True)


def transfer_mesh(shape, box):
    """FieldModel.transfer_mesh without an engine: sqrt(P(k) N / V), float32-rounded."""
    from montecosmo_b200 import cosmo as _cosmo
    from montecosmo_b200.model import linear_power_table
    ks, pows = linear_power_table(_cosmo.Cosmology())
    kvec = O.rfftk(shape, box)
    kmesh = np.sqrt(sum(k ** 2 for k in kvec))
    pmesh = np.interp(kmesh.reshape(-1), ks, pows, left=0.0, right=0.0).reshape(kmesh.shape)
    return np.sqrt(pmesh * (np.prod(shape) / np.prod(box))).astype(np.float32).astype(np.float64)


def inputs(n, seed):
    shape = (n, n, n)
    white = np.random.default_rng(seed).standard_normal(shape)
    obs = 1.0 + np.random.default_rng(100).standard_normal(shape)
    # both sides see float32-representable inputs
    return shape, white.astype(np.float32).astype(np.float64), obs.astype(np.float32).astype(np.float64)


def directions(shape, k=3):
    return [np.random.default_rng(200 + i).standard_normal(shape).astype(np.float32).astype(np.float64) for i in range(k)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("n", type=int)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--threads", type=int, default=6)
    ap.add_argument("--cd", action="store_true", help="also central-difference directional derivatives (6 more forwards)")
    ap.add_argument("--eps", type=float, default=1e-4)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    torch.set_num_threads(args.threads)
    O.LEAN_ASSIGNMENT = True  # same arithmetic, the neighbour loop's autograd graph rebuilt in the backward pass
    n = args.n
    shape, white, obs = inputs(n, args.seed)
    box = (BOX,) * 3
    transfer = transfer_mesh(shape, box)
    cosmo = O.Cosmology()
    t0 = time.time()
    w = torch.tensor(white, requires_grad=True)
    state = {}
    gxy = MO.evolve(w, transfer, cosmo, shape, checkpoint=True, state_out=state, **KW)
    lp = -0.5 * ((gxy - O._t(obs)) ** 2).sum() - 0.5 * (w ** 2).sum()
    print(f"forward {time.time() - t0:.1f} s, logp = {float(lp):.10e}", flush=True)
    (g,) = torch.autograd.grad(lp, w)
    print(f"forward + gradient {time.time() - t0:.1f} s", flush=True)
    mesh = gxy.detach().numpy()
    g = g.numpy()
    stride = max(n ** 3 // 65536, 1)
    disp = (state["pos"] - state["q"]).numpy()[::stride]
    vel = state["vel"].numpy()[::stride]
    kcount, kmean, pk = MM.spectrum(mesh, box_size=box)
    dirs = directions(shape)
    c = n // 32
    out = dict(n=n, seed=args.seed, box=BOX, n_steps=N_STEPS, logp=float(lp), stride=stride,
               mesh_block=mesh[:32, :32, :32].copy(),
               mesh_coarse=mesh.reshape(32, c, 32, c, 32, c).mean(axis=(1, 3, 5)),
               pk_count=kcount, pk_kmean=kmean, pk=pk, disp_sub=disp, vel_sub=vel,
               grad_sub=g[::4, ::4, ::4].copy(), grad_block=g[:32, :32, :32].copy(),
               grad_norm=float(np.linalg.norm(g)), grad_dot=np.array([float((g * v).sum()) for v in dirs]),
               white_sum=float(white.sum()), white_sq=float((white ** 2).sum()), obs_sum=float(obs.sum()))
    if args.cd:
        cd = []
        with torch.no_grad():
            for v in dirs:
                vals = []
                for sgn in (1.0, -1.0):
                    ww = torch.tensor(white + sgn * args.eps * v)
                    gg = MO.evolve(ww, transfer, cosmo, shape, **KW)
                    vals.append(float(-0.5 * ((gg - O._t(obs)) ** 2).sum() - 0.5 * (ww ** 2).sum()))
                cd.append((vals[0] - vals[1]) / (2 * args.eps))
                print(f"cd {cd[-1]:.8e} vs autograd {out['grad_dot'][len(cd) - 1]:.8e}  ({time.time() - t0:.0f} s)", flush=True)
        out["cd_dot"], out["cd_eps"] = np.array(cd), args.eps
    path = args.out or os.path.join(ROOT, "tests", "golden", f"c{2 if n == 128 else 3 if n == 256 else 0}_{n}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB) in {time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    main()
