"""Component self-test of montecosmo_b200/dist.py under torchrun/NCCL: distributed FFT (all batch sizes), halo ops."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from montecosmo_b200 import nbody as nb
from montecosmo_b200.dist import SlabPM
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
halo = int(sys.argv[2]) if len(sys.argv) > 2 else n // world
ops = nb.ops(); dev = ops.A.device
pm = SlabPM(ops, (n, n, n), halo=halo)
g = torch.Generator(device=dev).manual_seed(0)
rel = lambda a, b: float((a - b).norm() / b.norm())
out = {}
for nb_ in (1, 2, 3, 6):
    full = torch.randn((nb_, n, n, n), device=dev, generator=g)  # same on every rank (same seed)
    ref = torch.fft.rfftn(full, dim=(1, 2, 3))
    ck = pm.rfftn(full[:, pm.x0:pm.x0 + pm.xl].contiguous())
    out[f"rfftn{nb_}"] = rel(ck, ref[:, :, pm.y0:pm.y0 + pm.kyl])
    back = pm.irfftn(ref[:, :, pm.y0:pm.y0 + pm.kyl].contiguous()) / pm.N
    out[f"irfftn{nb_}"] = rel(back, full[:, pm.x0:pm.x0 + pm.xl])
for tail in ((), (4,)):
    ext = torch.randn((world, pm.ext, n, n, *tail), device=dev, generator=g)
    H, xl = pm.H, pm.xl
    mine = ext[rank].clone(); pm.halo_reduce(mine)
    exp = ext[rank].clone()
    exp[xl:xl + H] += ext[(rank + 1) % world][:H]
    exp[H:2 * H] += ext[(rank - 1) % world][H + xl:]
    out[f"halo_reduce{tail}"] = rel(mine[H:H + xl], exp[H:H + xl])
    gth = ext[rank].clone(); pm.halo_gather(gth)
    out[f"halo_gather{tail}"] = max(rel(gth[H + xl:], ext[(rank + 1) % world][H:2 * H]), rel(gth[:H], ext[(rank - 1) % world][xl:xl + H]))
t = torch.tensor(list(out.values()), device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print({k: float(f"{v:.2e}") for k, v in zip(out, t.tolist())}, flush=True)
if world > 1:
    dist.destroy_process_group()
