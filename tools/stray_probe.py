import sys; sys.path.insert(0,".")
import torch, numpy as np
from montecosmo_b200 import nbody as nb
from bench import workload
from montecosmo_b200.model import FieldModel
o = nb.ops(); n=256; shape=(n,n,n)
m = FieldModel(**workload(n)); dev=o.A.device
g = torch.Generator(device=dev).manual_seed(0)
dk = m.linear_field(torch.randn(shape, device=dev, generator=g))
for a1 in (0.3, 0.6, 1.0):
    pos, vel = nb.nbody_bf(m.cosmology, dk, m.q, 0.0, a1, 10, ptcl_shape=None)
    d = (pos[0] - m.q).reshape(n, n, n, 3)
    d = d - n * torch.round(d / n)
    for (BX, BY, BZ) in ((16, 8, 32), (8, 8, 32), (8, 4, 32)):
        db = d.reshape(n//BX, BX, n//BY, BY, n//BZ, BZ, 3)
        mean = db.mean(dim=(1, 3, 5), keepdim=True)
        rel = db - mean  # displacement relative to the brick mean
        out = []
        for M in (6, 8, 10, 12):
            # tile = brick + M: allowed base-cell offsets relative to centred box: floor(x) in [o, o+T-2]
            # approx: |rel| <= (M - 1)/2 - 0.5 in each dim
            lim = (M - 2) / 2.0
            frac = float(((rel.abs() > lim).any(dim=-1)).float().mean())
            out.append(f"M={M}: {100*frac:5.1f}%")
        print(f"a={a1} brick {BX}x{BY}x{BZ}: stray fraction  " + "  ".join(out), f" | rel disp rms {float(rel.std()):.2f}", flush=True)
