"""Run N gradient evaluations of the bench workload (for ncu launch lists)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import workload
from montecosmo_b200 import nbody as nb
from montecosmo_b200.model import FieldModel
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
m = FieldModel(**workload(n))
dev = nb.ops().A.device
g = torch.Generator(device=dev).manual_seed(0)
obs = 1.0 + torch.randn(m.mesh_shape, device=dev, generator=g)
for i in range(reps):
    w = torch.randn(m.mesh_shape, device=dev, generator=g)
    torch.cuda.synchronize()
    nb.ops().axpby(w, 1.0)  # marker kernel: start of an evaluation
    lp, gr = m.value_and_force(w, obs)
    torch.cuda.synchronize()
print("logp", float(lp))
