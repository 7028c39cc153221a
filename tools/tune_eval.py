"""Time whole grad(logp) evaluations of the bench workload under mcpm_tune settings (one process, same inputs)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import workload  # noqa: E402
from montecosmo_b200 import nbody as nb  # noqa: E402
from montecosmo_b200.model import FieldModel  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
settings = [s for s in sys.argv[2:]] or ["base", "gather_tma=0", "base", "gather_tma=0"]
m = FieldModel(**workload(n))
lib, dev = nb.ops().lib, nb.ops().A.device
g = torch.Generator(device=dev).manual_seed(0)
obs = 1.0 + torch.randn(m.mesh_shape, device=dev, generator=g)
w = torch.randn(m.mesh_shape, device=dev, generator=g)
defaults = {"gather_blocked": 0, "gather_minb": 4, "side_zero": 0, "gather_tma": 1, "gather_seg": 32, "brick": 1, "gather_brick": 0, "yzfft": 0, "brick_stream": 44, "brick_stream1": 0}
for s in settings:
    cur = dict(defaults)
    if s != "base":
        for kv in s.split(","):
            k, v = kv.split("=")
            cur[k] = int(v)
    for k, v in cur.items():
        lib.mcpm_tune(k.encode(), v)  # process defaults (stateless entry points, engines created from now on) ...
        for eng in nb.ops()._engines.values():  # ... and the engines that already exist
            lib.mcpm_engine_tune(eng.handle, k.encode(), v)
    for _ in range(2):
        lp, gr = m.value_and_force(w, obs)
    torch.cuda.synchronize()
    ts = []
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lp, gr = m.value_and_force(w, obs)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"{s:40s} {np.median(ts):8.3f} ms / eval (min {min(ts):.3f})   logp {float(lp):.6e}  |g| {float(gr.norm()):.6e}",
          flush=True)
