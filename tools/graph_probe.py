"""Eager vs CUDA-graph replay of one grad(logp) evaluation at several mesh sizes (CUDA events, median of 10)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import workload  # noqa: E402
from montecosmo_b200 import nbody as nb  # noqa: E402
from montecosmo_b200.model import FieldModel  # noqa: E402

for n in [int(a) for a in sys.argv[1:]] or [64, 128, 256]:
    m = FieldModel(**workload(n))
    dev = nb.ops().A.device
    g = torch.Generator(device=dev).manual_seed(0)
    obs = 1.0 + torch.randn(m.mesh_shape, device=dev, generator=g)
    w = torch.randn(m.mesh_shape, device=dev, generator=g)

    def timed(fn, reps=10):
        for _ in range(3):
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
    eager = timed(lambda: m.value_and_force(w, obs))
    fn = m.graphed_value_and_force(obs)
    graphed = timed(lambda: fn(w))
    print(f"mesh {n}^3: eager {eager:8.3f} ms / eval   graph replay {graphed:8.3f} ms / eval   ({eager / graphed:.2f}x)", flush=True)
    del fn, m
    torch.cuda.empty_cache()
