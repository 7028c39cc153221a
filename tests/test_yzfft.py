"""
The fused (y,z) transforms (csrc/yzfft.cu: both 1-D passes of a plane in one kernel, the plane in shared memory) through
the C ABI (mcpm_slabfft_r2c_yz / _c2r_yz, which dispatch to them on square planes of side 64, 128, 256), against
NumPy's float64 FFT over the last two axes: relative L2 3e-6, unnormalised both ways like cuFFT; and against the cuFFT
2-D plans they replace (knob "yzfft" = 0).  The C2R input is Hermitian-consistent (it is the R2C of a real field), as
every spectrum the engine hands to a C2R is.
"""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    import montecosmo_b200.nbody as nbody
    return nbody.ops()


@pytest.mark.parametrize("n,nx,nb", [(64, 5, 1), (128, 3, 2), (256, 2, 1), (256, 7, 3), (256, 150, 1)])
def test_fused_yz_transforms_match_numpy(n, nx, nb):
    from montecosmo_b200._capi import check
    ops = _ops()
    A, lib = ops.A, ops.lib
    dev = A.device
    rng = np.random.default_rng(n + nx + nb)
    h = C.c_void_p()
    check(lib, lib.mcpm_slabfft_create(nx, n, n, 1, C.byref(h)))
    try:
        a = rng.normal(size=(nb, nx, n, n)).astype(np.float32)
        ref = np.fft.rfft2(a.astype(np.float64), axes=(2, 3))
        ad = torch.tensor(a, device=dev)
        outs = {}
        for knob in (1, 0):
            check(lib, lib.mcpm_tune(b"yzfft", knob))
            k = torch.empty((nb, nx, n, n // 2 + 1), dtype=torch.complex64, device=dev)
            check(lib, lib.mcpm_slabfft_r2c_yz(h, A.stream(), ad.data_ptr(), k.data_ptr(), nb))
            back = torch.empty((nb, nx, n, n), dtype=torch.float32, device=dev)
            kin = torch.tensor(ref.astype(np.complex64), device=dev)  # C2R may overwrite its input: a fresh copy
            check(lib, lib.mcpm_slabfft_c2r_yz(h, A.stream(), kin.data_ptr(), back.data_ptr(), nb))
            torch.cuda.synchronize()
            outs[knob] = (k.cpu().numpy().astype(np.complex128), back.cpu().numpy().astype(np.float64))
        rel = lambda x, y: float(np.linalg.norm((x - y).ravel()) / np.linalg.norm(y.ravel()))
        for knob in (1, 0):
            assert rel(outs[knob][0], ref) < 3e-6, (knob, rel(outs[knob][0], ref))
            assert rel(outs[knob][1], a.astype(np.float64) * n * n) < 3e-6, (knob, rel(outs[knob][1], a * n * n))
        assert rel(outs[1][0], outs[0][0]) < 3e-6 and rel(outs[1][1], outs[0][1]) < 3e-6
    finally:
        lib.mcpm_tune(b"yzfft", 0)
        lib.mcpm_slabfft_destroy(h)
