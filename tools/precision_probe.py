"""Which kernel path costs accuracy?  The C2 (128^3) fixture check of tests/test_full_size_oracle.py under kernel knobs:
brick-tiled scatters (fixed-point shared-memory tile) vs generic float atomics, fused x-transform vs 3-D cuFFT."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from montecosmo_b200 import nbody as nb  # noqa: E402
from test_full_size_oracle import check_against_fixture  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
fx = dict(np.load(os.path.join(ROOT, "tests", "golden", f"c{2 if n == 128 else 3}_{n}.npz")))
lib = nb.ops().lib
for tag, brick, fused in (("default", 1, 1), ("brick=0", 0, 1), ("fused_fft=0", 1, 0), ("brick=0,fused_fft=0", 0, 0)):
    lib.mcpm_tune(b"brick", brick)
    nb.ops()._engines.clear()  # engines copy the default knobs at creation
    eng = nb.ops().engine((n, n, n))
    lib.mcpm_engine_set_fused_fft(eng.handle, fused)
    rep = {}
    try:
        check_against_fixture(nb, fx, rep)
    except AssertionError:
        pass
    print(tag, {k: f"{v:.2e}" for k, v in rep.items() if k in ("disp_rms", "mesh_block_rel", "logp_rel", "grad_sub_rel",
                                                                "grad_block_rel", "grad_dot_rel")}, flush=True)
lib.mcpm_tune(b"brick", 1)
