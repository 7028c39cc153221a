"""Time FieldLevelModel.evolve + backward with the observation chain inside the paint (mcpm_nufft_obs) against the same
chain as elementwise passes, on the GPU: curved sky + light cone + ap_auto (lpt), and flat sky + scalar a_obs (nbody).
Usage: python tools/obs_probe.py [mesh side, default 128]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from montecosmo_b200.cosmo import Cosmology  # noqa: E402
from montecosmo_b200.model import FieldLevelModel  # noqa: E402


def run(n, case, fused, reps=3):
    shape, box = (n, n, n), (float(5 * n),) * 3
    if case == "curved_lightcone_lpt":
        kw = dict(evolution="lpt", a_obs=None, box_center=(300.0, -200.0, 1500.0), curved_sky=True,
                  bias=dict(b1=0.9, b2=0.3, bnpar=4.0), ap_auto=True, cosmo_fid=Cosmology(Omega_c=0.21, Omega_b=0.05, h=0.7))
    else:
        kw = dict(evolution="nbody", n_steps=5, a_obs=0.8, box_center=(0.0, 0.0, 2000.0), curved_sky=False,
                  bias=dict(b1=0.7, b2=0.2))
    m = FieldLevelModel(shape, box, fused_observation=fused, **kw)
    g = torch.Generator(device="cpu").manual_seed(0)
    white = torch.randn(shape, generator=g).cuda()
    cot = torch.randn(shape, generator=g).cuda()
    ts, out = [], None
    for _ in range(reps + 1):
        w = white.clone().requires_grad_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = m.evolve(w)
        (out * cot).sum().backward()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return dict(case=case, mesh=n, fused=fused, ms=1e3 * float(np.median(ts[1:])), out=out.detach(), grad=w.grad)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    rel = lambda a, b: float((a - b).norm() / b.norm())
    for case in ("curved_lightcone_lpt", "flat_scalar_nbody"):
        a, b = run(n, case, True), run(n, case, False)
        print(json.dumps(dict(case=case, mesh=n, fused_ms=a["ms"], elementwise_ms=b["ms"], rel_out=rel(a["out"], b["out"]),
                              rel_grad=rel(a["grad"], b["grad"]))), flush=True)
