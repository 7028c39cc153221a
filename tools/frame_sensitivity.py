"""How sensitive are float32 cotangents to the coordinate frame?  Single-GPU engine, same problem, particle x shifted by an
integer number of cells (a periodic relabelling of the mesh: identical physics).  Yardstick for tools/slab_bench.py --check."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from montecosmo_b200 import nbody as nb
from montecosmo_b200.cosmo import Cosmology, a2g, a2g2, a2dg2dg, bullfrog_coefficients
ops = nb.ops(); dev = ops.A.device
n = 128; shape = (n, n, n); cosmo = Cosmology()
rng = np.random.default_rng(0)
kk = np.sqrt(sum(np.meshgrid(np.fft.fftfreq(n) ** 2, np.fft.fftfreq(n) ** 2, np.fft.rfftfreq(n) ** 2, indexing="ij"))); kk[0, 0, 0] = 1.0
dk = (np.fft.rfftn(rng.normal(size=shape)) * 0.02 * kk ** -1.5).astype(np.complex64); dk[0, 0, 0] = 0
a0, a1, ns = 0.1, 0.8, 3
ax = [np.arange(s, dtype=np.float32) for s in shape]
q = torch.tensor(np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3), device=dev)
d1, d2, dv2 = float(a2g(cosmo, a0)), float(a2g2(cosmo, a0)), float(a2dg2dg(cosmo, a0))
co = [t.tolist() for t in bullfrog_coefficients(cosmo, a0, a1, ns)[:4]]
dp, vl = ops.lpt(torch.tensor(dk, device=dev), q, d1, d2, dv2, 2, 1)
g = torch.Generator(device=dev).manual_seed(5)
pb, vb = torch.randn(q.shape, device=dev, generator=g), torch.randn(q.shape, device=dev, generator=g)
res = {}
for shift in (0.0, 0.0, -40.0, 24.0):
    x = (dp + q).contiguous(); x[:, 0] += shift
    v = vl.clone()
    tape = ops.nbody_steps(x, v, shape, *co, tape=True)
    a, b = pb.clone(), vb.clone()
    ops.nbody_steps_vjp(a, b, shape, *co, tape)
    res.setdefault(shift, []).append((x.clone(), a.clone()))
ref_x, ref_a = res[0.0][0]
scale = ref_a.norm(dim=1).mean()
def cmp(tag, x, a, shift):
    d = (a - ref_a).norm(dim=1); o = d > 1e-3 * scale
    xs = x.clone(); xs[:, 0] -= shift
    print(f"{tag:22s} max|dx| {float((xs - ref_x).abs().max()):.2e}  outliers {int(o.sum()):6d}  bulk rel err {float(d[~o].norm() / ref_a[~o].norm()):.2e}")
cmp("same frame, 2nd run", *res[0.0][1], 0.0)
cmp("x shifted by -40", *res[-40.0][0], -40.0)
cmp("x shifted by +24", *res[24.0][0], 24.0)
