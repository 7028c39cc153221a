"""Whole grad(log-density) evaluations of the bench workload with the assignment order of the step loop and of the final
paint set to CIC (2), TSC (3) and PCS (4): orders 3 / 4 take the generic global-atomic kernels (paint.cu) -- the brick-tiled
and bulk-copy staged kernels are CIC-specialised (missing #10: never timed)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import workload  # noqa: E402
from montecosmo_b200 import nbody as nb  # noqa: E402
from montecosmo_b200.model import FieldModel  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = nb.ops().A.device
g = torch.Generator(device=dev).manual_seed(0)
for order in (2, 3, 4):
    wl = workload(n)
    wl["paint_order"] = order
    m = FieldModel(**wl)
    obs = 1.0 + torch.randn(m.mesh_shape, device=dev, generator=g)
    w = torch.randn(m.mesh_shape, device=dev, generator=g)
    for _ in range(2):
        lp, gr = m.value_and_force(w, obs)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lp, gr = m.value_and_force(w, obs)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"paint_order {order}: {np.median(ts):8.3f} ms / eval (min {min(ts):.3f}), eager launches   logp {float(lp):.6e}", flush=True)
    del m
    torch.cuda.empty_cache()
