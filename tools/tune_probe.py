"""Time the two readout kernels of the BullFrog step for each `gather_minb` setting (mcpm_tune) at 256^3."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from montecosmo_b200 import _lib  # noqa: E402
from montecosmo_b200.ops import Ops, TorchCudaAdapter  # noqa: E402
from tools.microbench import timeit  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ops = Ops(_lib.load(), TorchCudaAdapter())
lib, A, dev = ops.lib, ops.A, ops.A.device
N = n ** 3
g = torch.Generator(device=dev).manual_seed(0)
ax = torch.arange(n, device=dev, dtype=torch.float32)
q = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)
k = 2 * np.pi / n
disp = 3.0 * torch.stack([torch.sin(k * 3 * q[:, 1]) + torch.cos(k * 5 * q[:, 2]),
                          torch.sin(k * 4 * q[:, 2]) + torch.cos(k * 2 * q[:, 0]),
                          torch.sin(k * 3 * q[:, 0]) + torch.cos(k * 6 * q[:, 1])], -1)
pos = (q + disp + 0.8 * torch.randn(N, 3, device=dev, generator=g)).contiguous()
vel = torch.randn(N, 3, device=dev, generator=g)
xbar = torch.randn(N, 3, device=dev, generator=g)
fm4 = torch.randn(n, n, n, 4, device=dev, generator=g)
rho = torch.randn(n, n, n, device=dev, generator=g)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
p2, v2 = pos.clone(), vel.clone()
st = A.stream()
for minb, blocked in ((4, 0), (4, 1), (5, 1), (6, 1)):
    lib.mcpm_tune(b"gather_minb", minb)
    lib.mcpm_tune(b"gather_blocked", blocked)
    kd = timeit(lambda: lib.mcpm_kick_drift4(st, p2.data_ptr(), v2.data_ptr(), fm4.data_ptr(), N, n, n, n, 1.0, 0.0, 0.0),
                flush=flush)
    rg = timeit(lambda: lib.mcpm_read_grad4v(st, pos.data_ptr(), fm4.data_ptr(), rho.data_ptr(), vel.data_ptr(), 0.5, 1.0,
                                             N, n, n, n, xbar.data_ptr()), flush=flush)
    print(f"gather_minb={minb} blocked={blocked}: kick_drift4 {kd[0]:.3f} ms (min {kd[1]:.3f})   read_grad4v {rg[0]:.3f} ms (min {rg[1]:.3f})",
          flush=True)
lib.mcpm_tune(b"gather_minb", 4)
lib.mcpm_tune(b"gather_blocked", 0)
