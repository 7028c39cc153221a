"""Time the two fusions of SURVEY 8f rows 1 and 3 against the compositions they replace, at n^3 on one GPU:
  * bricks.lagrangian_bias (fused passes of csrc/bias.cu) vs bricks.lagrangian_bias_composed (round 1's pointwise torch
    composition), value + gradient w.r.t. the linear mesh, with the full 8-coefficient expansion;
  * nufft with the redshift-space shift inside the paint kernels (mcpm_nufft_rsd) vs mcpm_rsd_shift followed by nufft,
    value + gradient w.r.t. positions, velocities and weights."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from montecosmo_b200 import bricks as B  # noqa: E402
from montecosmo_b200 import nbody as nb  # noqa: E402
from montecosmo_b200.cosmo import Cosmology  # noqa: E402
from montecosmo_b200.model import _RsdShift  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = nb.ops().A.device
shape, box = (n, n, n), (2.5 * n,) * 3
g = torch.Generator(device=dev).manual_seed(0)
N = n ** 3


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


dk0 = nb.rfftn(torch.randn(shape, device=dev, generator=g)) * 0.01
pos = B.regular_pos(shape)
bias = dict(b1=0.8, b2=0.3, bs2=-0.2, b3=0.1, bds2=0.05, bs3=-0.07, bn2=0.4, bnpar=0.6)
cw = torch.randn(N, device=dev, generator=g)
cv = torch.randn((N, 3), device=dev, generator=g)
cosmo = Cosmology()  # one object: its growth table (a 128-step RK4 solve on the host) is cached on it
for name, fn in (("fused passes (bias.cu)", B.lagrangian_bias), ("pointwise composition", B.lagrangian_bias_composed)):
    def run(fn=fn):
        dk = dk0.clone().requires_grad_()
        w, dvel, _ = fn(cosmo, pos, 0.7, box, dk, bias, read_order=2)
        ((w * cw).sum() + (dvel * cv).sum()).backward()
        return dk.grad
    ms = timeit(run)
    torch.cuda.reset_peak_memory_stats()
    run()
    print(f"lagrangian_bias {n}^3, value + gradient, {name:24s} {ms:8.2f} ms   peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB",
          flush=True)
    torch.cuda.empty_cache()

disp = torch.randn((N, 3), device=dev, generator=g) * 1.5
vel0 = torch.randn((N, 3), device=dev, generator=g)
w0 = torch.rand(N, device=dev, generator=g) + 0.5
ck = nb.rfftn(torch.randn(shape, device=dev, generator=g))
los, coef = (0.0, 0.0, 1.0), 0.4
for name, fused in (("shift inside the paint (mcpm_nufft_rsd)", True), ("mcpm_rsd_shift, then mcpm_nufft", False)):
    def run(fused=fused):
        p, v, w = disp.clone().requires_grad_(), vel0.clone().requires_grad_(), w0.clone().requires_grad_()
        if fused:
            out = nb.nufft(p, shape, None, w, 2, 2, paint_deconv=True, lattice=shape, rsd=(v, los, coef))
        else:
            out = nb.nufft(_RsdShift.apply(p, v, los, coef), shape, None, w, 2, 2, paint_deconv=True, lattice=shape)
        (out * ck.conj()).real.sum().backward()
    print(f"redshift-space nufft {n}^3, value + gradients, {name:42s} {timeit(run):8.2f} ms", flush=True)
