"""Where does a CUDA-graph replay of the evaluation depart from the eager run?  (development tool)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import workload  # noqa: E402
from montecosmo_b200 import nbody as nb  # noqa: E402
from montecosmo_b200.model import FieldModel  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = nb.ops().A.device
m = FieldModel(**workload(n))
g = torch.Generator(device=dev).manual_seed(3)
obs = 1.0 + torch.randn(m.mesh_shape, device=dev, generator=g)
whites = [torch.randn(m.mesh_shape, device=dev, generator=g) for _ in range(3)]


def rel(a, b):
    return float((a - b).double().norm() / b.double().norm())


def capture(fn, static_in):
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(2):
            fn(static_in)
    torch.cuda.current_stream(dev).wait_stream(side)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        out = fn(static_in)
    return gr, out


stages = {
    "linear_field": lambda w: torch.view_as_real(m.linear_field(w)),
    "lpt": lambda w: torch.cat(nb.lpt(m.cosmology, m.linear_field(w), m.q, 0.5, 2, 1)),
    "nbody pos": lambda w: nb.nbody_bf(m.cosmology, m.linear_field(w), m.q, 0.0, 1.0, m.n_steps)[0][-1],
    "evolve": lambda w: m.evolve(w),
}
for name, fn in stages.items():
    static = torch.zeros(m.mesh_shape, device=dev)
    with torch.no_grad():
        gr, out = capture(fn, static)
        errs = []
        for w in whites:
            static.copy_(w)
            gr.replay()
            errs.append(rel(out.clone(), fn(w)))
    print(f"{name:14s} forward only: graph vs eager rel err {errs}", flush=True)

for lattice, fused in ((True, True), (False, True), (True, False), (False, False)):
    shape = m.mesh_shape
    static = torch.zeros(shape, device=dev)

    def vf(w):
        lp, f = m.value_and_force(w, obs)
        nb.ops().set_lattice(shape, shape if lattice else None)
        return lp, f
    import montecosmo_b200.nbody as nbm
    orig = nbm.nbody_bf

    def nbody_bf_nolat(*a, **k):
        k["ptcl_shape"] = "auto" if lattice else None
        return orig(*a, **k)
    nbm.nbody_bf = nbody_bf_nolat
    nb.ops().set_fused_fft(shape, fused)
    gr, (lp, f) = capture(vf, static)
    res = []
    for w in whites:
        static.copy_(w)
        gr.replay()
        lpg, fg = float(lp), f.clone()
        gr.replay()
        lpg2, fg2 = float(lp), f.clone()
        lpe, fe = m.value_and_force(w, obs)
        res.append((abs(lpg - float(lpe)) / abs(float(lpe)), rel(fg, fe), abs(lpg2 - lpg) / abs(lpg), rel(fg2, fg)))
    nbm.nbody_bf = orig
    nb.ops().set_fused_fft(shape, True)
    print(f"value_and_force lattice={lattice} fused={fused}: (lp err, force err, lp replay-replay, force replay-replay) =",
          [tuple(f"{v:.1e}" for v in r) for r in res], flush=True)
