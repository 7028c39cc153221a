"""Time the brick-tiled CIC scatters (density paint, 3-channel reverse-step scatter) at n^3 under mcpm_tune("brick_stream")
settings: 0 = one short-lived CTA per 16 x 8 x 32 brick, 44 | 48 = persistent bulk-copy staged kernel with that tile row
stride.  Particles: the evolved a = 1 lattice of the bench workload (relative frame, as the step loop runs)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import workload  # noqa: E402
from montecosmo_b200 import nbody as nb  # noqa: E402
from montecosmo_b200._capi import Frame  # noqa: E402
from montecosmo_b200.model import FieldModel  # noqa: E402
from tools.microbench import timeit  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m = FieldModel(**workload(n))
ops = nb.ops()
lib, A, dev = ops.lib, ops.A, ops.A.device
g = torch.Generator(device=dev).manual_seed(0)
dk = m.linear_field(torch.randn(m.mesh_shape, device=dev, generator=g))
N = n ** 3
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
st = A.stream()
fr = Frame(1, n, n, n, 0, 0, 0, n, n, n)
mesh, mesh3 = torch.zeros((n, n, n), device=dev), torch.zeros((3, n, n, n), device=dev)
vbar = torch.randn((N, 3), device=dev, generator=g)
lib.mcpm_tune(b"brick_stream1", 1)
for a1 in (0.3, 1.0):
    pos, vel = nb.nbody_bf(m.cosmology, dk, m.q, 0.0, a1, 10, ptcl_shape=None)
    d = (pos[0] - m.q)
    d = (d - n * torch.round(d / n)).contiguous()
    for knob in (0, 44, 48, 0, 44):
        lib.mcpm_tune(b"brick_stream", knob)
        p1 = timeit(lambda: (mesh.zero_(), lib.mcpm_paint_brick_f(st, C.byref(fr), n, n, n, d.data_ptr(), None, 1.0, 0.0, N, n, n, n,
                                                                 mesh.data_ptr())), flush=flush)
        p3 = timeit(lambda: (mesh3.zero_(), lib.mcpm_paint3_brick_f(st, C.byref(fr), n, n, n, d.data_ptr(), vbar.data_ptr(), None, 0.0,
                                                                   0.5, N, n, n, n, mesh3.data_ptr())), flush=flush)
        print(f"a={a1} disp rms {float(d.std()):.2f}  brick_stream={knob:2d}: paint {p1[0]:.4f} ms (min {p1[1]:.4f})   "
              f"paint3 {p3[0]:.4f} ms (min {p3[1]:.4f})   [memsets included]", flush=True)
lib.mcpm_tune(b"brick_stream", 44)
lib.mcpm_tune(b"brick_stream1", 0)
