"""CPU simulation of the shared-memory wavefronts of the brick scatter's ATOMS (round 2): for evolved particles of the
bench workload (128^3 at the C3 cell size, 2.5 Mpc/h), the number of wavefronts per warp ATOMS instruction = max over the
32 banks of the number of lanes addressing that bank (same-address lanes serialise like different-address ones), for a
given tile row stride and warp -> particle mapping."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import cpu_port
import montecosmo_b200.nbody as nb
from montecosmo_b200.model import FieldModel
from bench import workload

n = 128
nb._OPS = cpu_port.cpu_ops()
wl = workload(n)
wl["box_size"] = tuple(2.5 * n for _ in range(3)) if "box_size" in wl else wl.get("box_size")
m = FieldModel(**wl)
g = torch.Generator().manual_seed(0)
dk = m.linear_field(torch.randn((n, n, n), generator=g))
rng = np.random.default_rng(0)


def wavefronts(cx, cy, cz, stride, ty, same_addr_merge=False):
    """cx, cy, cz: [W, 32] integer tile cells of the lanes of W warps.  Returns mean wavefronts per instruction."""
    addr = (cx * ty + cy) * stride + cz
    bank = addr & 31
    W = bank.shape[0]
    if same_addr_merge:
        # lanes with the same address count once
        out = np.zeros(W)
        for w in range(W):
            u = np.unique(addr[w])
            out[w] = np.bincount(u & 31, minlength=32).max()
        return out.mean()
    cnt = np.zeros((W, 32), dtype=np.int32)
    for l in range(32):
        np.add.at(cnt, (np.arange(W), bank[:, l]), 1)
    return cnt.max(axis=1).mean()


for a1 in (0.2, 0.5, 1.0):
    pos, vel = nb.nbody_bf(m.cosmology, dk, m.q, 0.0, a1, 10, ptcl_shape=None)
    d = (pos[0] - m.q).reshape(n, n, n, 3).numpy().astype(np.float64)
    d = d - n * np.round(d / n)
    x = d + np.stack(np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij"), -1)
    c = np.floor(x).astype(np.int64)  # base cells
    gz = np.diff(d[..., 2], axis=2)
    print(f"a={a1}: disp rms {d.std():.2f} cells, d(dz)/dz rms {gz.std():.2f}, P(adjacent-z particles share z cell) "
          f"{(c[:, :, 1:, 2] == c[:, :, :-1, 2]).mean():.3f}")
    # mapping A: a warp = 32 consecutive z at fixed (i, j)   [current kernel]
    sel = rng.choice(n * n * (n // 32), size=4000, replace=False)
    ii, jj, kk = np.unravel_index(sel, (n, n, n // 32))
    cw = np.stack([c[ii, jj, k0 * 32:(k0 + 1) * 32] for k0 in range(n // 32)])[kk, np.arange(len(sel))]  # [W,32,3]
    for stride in (44, 45, 48, 52, 56, 64, 33):
        r = [wavefronts(cw[..., 0] + dx, cw[..., 1] + dy, cw[..., 2] + dz, stride, 18) for dx in (0, 1) for dy in (0, 1) for dz in (0, 1)]
        rm = wavefronts(cw[..., 0], cw[..., 1], cw[..., 2], stride, 18, same_addr_merge=True)
        print(f"   z-row warp, stride {stride}: {np.mean(r):.2f} wavefronts / ATOMS   (if same-address lanes merged: {rm:.2f})")
    # mapping B: a warp = 2 y-rows x 16 z
    cwb = np.concatenate([c[ii, jj][np.arange(len(sel)), :][:, :16], c[ii, (jj + 1) % n][np.arange(len(sel)), :][:, :16]], axis=1)
    for stride in (32, 48):
        r = wavefronts(cwb[..., 0], cwb[..., 1], cwb[..., 2], stride, 18)
        print(f"   2y x 16z warp, stride {stride}: {r:.2f}")
