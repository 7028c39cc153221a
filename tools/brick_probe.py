import sys; sys.path.insert(0,".")
import torch, numpy as np
from montecosmo_b200 import nbody as nb
from bench import workload
from montecosmo_b200.model import FieldModel
o = nb.ops(); n=256; shape=(n,n,n); N=n**3
m = FieldModel(**workload(n)); dev=o.A.device
g = torch.Generator(device=dev).manual_seed(0)
white = torch.randn(shape, device=dev, generator=g)
dk = m.linear_field(white)
pos, vel = nb.nbody_bf(m.cosmology, dk, m.q, 0.0, 1.0, 10, ptcl_shape=None)
pos, vel = pos[0].contiguous(), vel[0].contiguous()
al,be,pre,post=[0.8],[0.5],[0.01],[0.01]
o.set_lattice(shape, shape)
torch.cuda.synchronize()
tape = o.nbody_steps(pos.clone(), vel.clone(), shape, al, be, pre, post, tape=True)
pb, vb = vel.clone(), pos.clone()
o.nbody_steps_vjp(pb, vb, shape, al, be, pre, post, tape)
torch.cuda.synchronize()
print("done")
