import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from montecosmo_b200 import nbody as nb
from montecosmo_b200.model import FieldModel
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m = FieldModel((n,)*3, (2.5*n,)*3, n_steps=int(sys.argv[2]) if len(sys.argv) > 2 else 10)
dev = nb.ops().A.device
def sync(tag):
    torch.cuda.synchronize(); print("ok:", tag, f"mem={torch.cuda.memory_allocated()/2**30:.2f}GiB", flush=True)
white = torch.randn(m.mesh_shape, device=dev).requires_grad_()
dk = m.linear_field(white); sync("linear_field")
d = nb.irfftn(dk); sync("irfftn")
dq = nb.read(m.q, d, order=1); sync("read ngp")
x, v = nb.lpt(m.cosmology, dk, m.q, 0.0, 2, 1, _displaced=True); sync("lpt")
print("lpt disp rms", float((x - m.q).std()))
pos, vel = nb.nbody_bf(m.cosmology, dk, m.q, 0.0, 1.0, m.n_steps); sync("nbody_bf")
print("disp rms", float((pos[0] - m.q).std()), "max", float((pos[0]-m.q).abs().max()))
gxy = m.evolve(white); sync("evolve")
print("gxy mean/std/min/max", float(gxy.mean()), float(gxy.std()), float(gxy.min()), float(gxy.max()))
obs = gxy.detach() + torch.randn_like(gxy)
r = nb.ops().axpby(gxy.detach(), 1.0, obs, -1.0); sync("axpby")
g = torch.autograd.grad(gxy, white, r); sync("backward")
print("grad norm", float(g[0].norm()))
lp, g2 = m.value_and_force(white.detach(), obs); sync("value_and_force")
print("logp", float(lp), "grad norm", float(g2.norm()))
