import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from montecosmo_b200 import nbody as nb
from montecosmo_b200.dist import SlabPM
n = int(sys.argv[1]); H = int(sys.argv[2]); ops = nb.ops(); dev = ops.A.device; lib = ops.lib; st = ops.A.stream()
pm = SlabPM(ops, (n, n, n), halo=H)
g = torch.Generator(device=dev).manual_seed(0)
rel = lambda a, b: float((a - b).norm() / b.norm())
ax = torch.arange(n, device=dev, dtype=torch.float32)
q = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)
x = (q + 0.7 * torch.randn(q.shape, device=dev, generator=g)).contiguous()
vbar, xbar = torch.randn(q.shape, device=dev, generator=g), torch.randn(q.shape, device=dev, generator=g)
fm4 = torch.randn((n, n, n, 4), device=dev, generator=g); fm4[..., 3] = 0
N = n ** 3
# ---- single GPU reference of one reverse step
vb1, xb1 = vbar.clone(), xbar.clone()
M4 = torch.zeros((n, n, n, 4), device=dev)
lib.mcpm_paint3v4(st, x.data_ptr(), vb1.data_ptr(), xb1.data_ptr(), 0.1, 0.5, N, n, n, n, M4.data_ptr())
PL = torch.empty((3, n, n, n), device=dev); lib.mcpm_deinterleave3(st, M4.data_ptr(), PL.data_ptr(), N)
PK = ops.rfftn(PL)
RK = ops.force_spectra_T(PK) / N
RHO = ops.irfftn(RK) * N          # unnormalised inverse of (T/N) == what the engine computes
lib.mcpm_read_grad4v(st, x.data_ptr(), fm4.data_ptr(), RHO.data_ptr(), vb1.data_ptr(), 0.5, 0.8, N, n, n, n, xb1.data_ptr())
# ---- slab
sl = slice(rank * pm.npl, (rank + 1) * pm.npl)
xl_ = x[sl].clone(); xl_[:, 0] += pm.H - pm.x0
vb2, xb2 = vbar[sl].clone(), xbar[sl].clone()
m4 = torch.zeros((pm.ext, n, n, 4), device=dev)
lib.mcpm_paint3v4(st, xl_.data_ptr(), vb2.data_ptr(), xb2.data_ptr(), 0.1, 0.5, pm.npl, pm.ext, n, n, m4.data_ptr())
pm.halo_reduce(m4)
e_m4 = rel(m4[pm.H:pm.H + pm.xl], M4[pm.x0:pm.x0 + pm.xl])
planar = torch.empty((3, pm.xl, n, n), device=dev)
lib.mcpm_deinterleave3(st, m4[pm.H:pm.H + pm.xl].data_ptr(), planar.data_ptr(), pm.xl * n * n)
e_pl = rel(planar, PL[:, pm.x0:pm.x0 + pm.xl])
pk = pm.rfftn(planar)
e_pk = rel(pk, PK[:, :, pm.y0:pm.y0 + pm.kyl])
rk = pm.force_spectra_T(pk)
e_rk = rel(rk, RK[:, pm.y0:pm.y0 + pm.kyl])
rho = pm.irfftn(rk.unsqueeze(0), overwrite=True)[0]
e_rho = rel(rho, RHO[pm.x0:pm.x0 + pm.xl])
rhoe = torch.empty((pm.ext, n, n), device=dev); rhoe[pm.H:pm.H + pm.xl] = rho; pm.halo_gather(rhoe)
f4e = torch.empty((pm.ext, n, n, 4), device=dev); f4e[pm.H:pm.H + pm.xl] = fm4[pm.x0:pm.x0 + pm.xl]; pm.halo_gather(f4e)
lib.mcpm_read_grad4v(st, xl_.data_ptr(), f4e.data_ptr(), rhoe.data_ptr(), vb2.data_ptr(), 0.5, 0.8, pm.npl, pm.ext, n, n, xb2.data_ptr())
print(f"rank {rank}: m4 {e_m4:.1e} planar {e_pl:.1e} pk {e_pk:.1e} rk {e_rk:.1e} rho {e_rho:.1e} xbar {rel(xb2, xb1[sl]):.1e} vbar {rel(vb2, vb1[sl]):.1e}", flush=True)
dist.destroy_process_group()
