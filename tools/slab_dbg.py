import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from montecosmo_b200 import nbody as nb
from montecosmo_b200.dist import SlabPM
n = int(sys.argv[1]); ops = nb.ops(); dev = ops.A.device
pm = SlabPM(ops, (n, n, n), halo=min(8, n // world))
g = torch.Generator(device=dev).manual_seed(0)
for nb_ in (1, 3):
    full = torch.randn((nb_, n, n, n), device=dev, generator=g)
    ref = torch.fft.rfftn(full, dim=(1, 2, 3))
    a = full[:, pm.x0:pm.x0 + pm.xl].contiguous()
    b = ops.A.empty((nb_, pm.xl, n, pm.nzc), "c64")
    pm._call("mcpm_slabfft_r2c_yz", pm._fft, pm._st(), a.data_ptr(), b.data_ptr(), nb_)
    r2 = torch.fft.rfftn(a, dim=(2, 3))
    e1 = float((b - r2).norm() / r2.norm())
    ck = pm.rfftn(a)
    tgt = ref[:, :, pm.y0:pm.y0 + pm.kyl]
    e2 = float((ck - tgt).norm() / tgt.norm())
    print(f"rank {rank} nb {nb_}: r2c_yz err {e1:.3e} nan {bool(torch.isnan(torch.view_as_real(b)).any())} | rfftn err {e2:.3e} nan {bool(torch.isnan(torch.view_as_real(ck)).any())} | sum check {float(full.sum()):.4f}", flush=True)
dist.destroy_process_group()
