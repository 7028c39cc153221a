"""How well-conditioned is grad(log-density) itself?  The float64 oracle's gradient at white noise w and at w (1 + eps):
the perturbation moves every particle by ~eps x its displacement (a few 1e-6 cell for eps = 1e-6, the size of a float32
engine's position error), the smooth response of the gradient is ~eps, and whatever exceeds that is the jump of the CIC
derivative for the particles that changed cell -- the floor ANY float32 implementation sees against the float64 one."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_full_size_fixture as FX  # noqa: E402
from oracle import model_oracle as MO  # noqa: E402
from oracle import pm_oracle as O  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
eps = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-6
torch.set_num_threads(int(sys.argv[3]) if len(sys.argv) > 3 else 4)
O.LEAN_ASSIGNMENT = True
shape, white, obs = FX.inputs(n, 0)
transfer = FX.transfer_mesh(shape, (FX.BOX,) * 3)
res = []
for scale in (1.0, 1.0 + eps):
    t0 = time.time()
    w = torch.tensor(white * scale, requires_grad=True)
    st = {}
    gxy = MO.evolve(w, transfer, O.Cosmology(), shape, checkpoint=True, state_out=st, **FX.KW)
    lp = -0.5 * ((gxy - O._t(obs)) ** 2).sum() - 0.5 * (w ** 2).sum()
    (g,) = torch.autograd.grad(lp, w)
    res.append((g.numpy(), (st["pos"] - st["q"]).numpy()))
    print(f"scale {scale!r}: {time.time() - t0:.0f} s", flush=True)
g0, d0 = res[0]
g1, d1 = res[1]
print(f"mesh {n}^3, eps {eps:g}: displacement change rms {np.sqrt(((d1 - d0) ** 2).mean()):.2e} cell, "
      f"gradient change rel L2 {np.linalg.norm(g1 - g0) / np.linalg.norm(g0):.2e} "
      f"(strided [::4] subsample {np.linalg.norm((g1 - g0)[::4, ::4, ::4]) / np.linalg.norm(g0[::4, ::4, ::4]):.2e})")
