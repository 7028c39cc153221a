"""Which of the two float32 forms of the observation chain is closer to float64?  FieldLevelModel.evolve with the chain
inside the paint (mcpm_nufft_obs) and with the chain as elementwise passes over absolute Mpc/h coordinates, both on the
CPU port of the engine, against the oracle's float64 restatement of the same model (oracle/model_oracle.py):
predicted mesh and gradient of a linear functional w.r.t. the white field, relative L2.
Usage: python tools/obs_accuracy.py [mesh side, default 32]   (checker-side script: it imports oracle/)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import montecosmo_b200.nbody as nbody  # noqa: E402
from montecosmo_b200.cosmo import Cosmology  # noqa: E402
from montecosmo_b200.model import FieldLevelModel  # noqa: E402
from oracle import cpu_port, model_oracle as MO, pm_oracle as O  # noqa: E402

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    nbody._OPS = cpu_port.cpu_ops()
    shape, box = (n, n, n), (float(5 * n),) * 3
    center = (300.0, -200.0, 1500.0)
    bias = dict(b1=0.9, b2=0.3, bnpar=4.0)
    rng = np.random.default_rng(0)
    white = rng.normal(size=shape).astype(np.float32)
    cot = rng.normal(size=shape)
    rel = lambda a, b: float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))
    res = {}
    for fused in (True, False):
        m = FieldLevelModel(shape, box, evolution="lpt", a_obs=None, box_center=center, curved_sky=True, bias=bias,
                            ap_auto=True, cosmo_fid=Cosmology(Omega_c=0.21, Omega_b=0.05, h=0.7), fused_observation=fused)
        w = torch.tensor(white, requires_grad=True)
        out = m.evolve(w)
        (out * torch.tensor(cot, dtype=torch.float32)).sum().backward()
        res[fused] = (out.detach().numpy().astype(np.float64), w.grad.numpy().astype(np.float64))
        transfer = m.transfer.cpu().numpy().astype(np.float64)
    wo = torch.tensor(white, dtype=torch.float64, requires_grad=True)
    ref = MO.evolve_general(wo, transfer, O.Cosmology(), shape, box, evolution="lpt", a_obs=None, box_center=center,
                            curved_sky=True, bias=bias, ap_fid=O.Cosmology(Omega_c=0.21, Omega_b=0.05, h=0.7))
    (ref * torch.tensor(cot)).sum().backward()
    r, g = ref.detach().numpy(), wo.grad.numpy()
    print(json.dumps({"mesh": n, "cell_mpc_h": 5.0, "observer_distance_mpc_h": 1500.0,
                      "inside_the_paint": {"mesh": rel(res[True][0], r), "grad": rel(res[True][1], g)},
                      "elementwise_float32": {"mesh": rel(res[False][0], r), "grad": rel(res[False][1], g)},
                      "between_the_two": {"mesh": rel(res[True][0], res[False][0]), "grad": rel(res[True][1], res[False][1])}}))
