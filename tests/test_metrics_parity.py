"""
Power-spectrum estimator (montecosmo/metrics.py:121-210) and the parity report BASELINE.json asks for: on identical
white noise and cosmology, the engine's density field, displacements and power spectrum against the float64 oracle,
within the float32 tolerances of SURVEY.md 8c (density 1e-4 relative L2 here; displacement max |dx| < 2e-4 cell;
P(k) ratio within 1 +- 1e-4 below half the Nyquist wavenumber and 1 +- 1e-3 up to it; cross-correlation > 1 - 1e-6).
"""
import numpy as np
import pytest
import torch

from oracle import metrics_oracle as MX
from oracle import model_oracle as MO
from oracle import pm_oracle as O


@pytest.fixture(scope="module", params=["hostemu", pytest.param("cuda", marks=pytest.mark.gpu)])
def nb(request):
    import montecosmo_b200.nbody as nbody
    from montecosmo_b200.ops import Ops
    old = nbody._OPS
    if request.param == "hostemu":
        from tests import hostemu
        from tests.backends import torch_cpu_adapter
        nbody._OPS = Ops(hostemu.load(), torch_cpu_adapter())
    else:
        nbody._OPS = None
        nbody.ops()
    yield nbody
    nbody._OPS = old


def test_spectrum_estimator_matches_oracle(nb, golden):
    from montecosmo_b200 import metrics as M
    rng = np.random.default_rng(0)
    dev = nb.ops().A.device
    # golden vectors of montecosmo/metrics.py itself (tests/golden/make_golden.py)
    g = golden("spectrum")
    a, b = (torch.tensor(g[k], dtype=torch.float32, device=dev) for k in ("mesh0", "mesh1"))
    box = tuple(g["box_size"])
    for tag, kw in [("default", dict(kedges=None, include_corners=True, deconv=2)),
                    ("n5_nocorners", dict(kedges=5, include_corners=False, deconv=0)),
                    ("dk02", dict(kedges=0.2, include_corners=True, deconv=(1, 2)))]:
        kc, km, p = M._spectrum(a, box_size=box, **kw)
        assert np.array_equal(kc, g[f"auto_{tag}_kcount"]) and np.allclose(km, g[f"auto_{tag}_kmean"], rtol=1e-12)
        assert np.allclose(p, g[f"auto_{tag}_pow"], rtol=2e-5)
        assert np.allclose(M._spectrum(a, b, box_size=box, **kw)[2], g[f"cross_{tag}_pow"], rtol=2e-5)
    ks, p1, tr, coh = M.powtranscoh(a, b, box)
    assert np.allclose(p1, g["ptc_pow1"], rtol=2e-5) and np.allclose(tr, g["ptc_trans"], rtol=2e-5)
    assert np.allclose(coh, g["ptc_coh"], rtol=2e-5)
    for shape, box in [((12, 8, 10), (100.0, 80.0, 120.0)), ((16, 16, 16), None)]:
        a = rng.normal(size=shape).astype(np.float32)
        b = (0.7 * a + 0.5 * rng.normal(size=shape)).astype(np.float32)
        for kedges, corners in [(None, True), (5, False), (0.2 if box else 0.8, True)]:
            kc, km, p = M._spectrum(torch.tensor(a, device=dev), box_size=box, kedges=kedges, include_corners=corners,
                                    deconv=2)
            kco, kmo, po = MX.spectrum(a.astype(np.float64), box_size=box, kedges=kedges, include_corners=corners,
                                       deconv=(2, 2))
            assert np.array_equal(kc, kco)          # bin membership is exact (float64 wavenumbers on the device)
            assert np.allclose(km, kmo, rtol=1e-12)
            assert np.allclose(p, po, rtol=2e-5)    # float32 FFT
            _, px = M.spectrum(torch.tensor(a, device=dev), torch.tensor(b, device=dev), box_size=box, kedges=kedges,
                               include_corners=corners)
            assert np.allclose(px, MX.spectrum(a.astype(np.float64), b.astype(np.float64), box_size=box, kedges=kedges,
                                               include_corners=corners)[2], rtol=2e-5)
        ks, p1, tr, coh = M.powtranscoh(torch.tensor(a, device=dev), torch.tensor(b, device=dev), box or shape)
        assert np.all(coh <= 1 + 1e-6) and np.all(tr > 0)
    # multipoles (metrics.py:165-166) against the golden vectors; quadrupoles change sign, so the bound is absolute,
    # 2e-5 of the largest monopole power
    a, b = (torch.tensor(g[k], dtype=torch.float32, device=dev) for k in ("mesh0", "mesh1"))
    box, center = tuple(g["box_size"]), tuple(g["box_center"])
    top = np.abs(g["auto_ell0_pow"]).max()
    p = M._spectrum(a, box_size=box, box_center=center, ells=[0, 2, 4], deconv=2)[2]
    for ell in (0, 2, 4):
        assert np.abs(p[ell] - g[f"auto_ell{ell}_pow"]).max() < 2e-5 * top
    p = M._spectrum(a, b, box_size=box, box_center=center, ells=[1, 2], kedges=6)[2]
    assert np.abs(p[1] - g["cross_ell1_pow"]).max() < 2e-5 * top
    assert np.abs(p[2] - g["cross_ell2_pow"]).max() < 2e-5 * top
    km, p2 = M.spectrum(a, box_size=box, ells=2)
    assert np.abs(p2 - g["auto_ell2_centred_pow"]).max() < 2e-5 * top


@pytest.mark.parametrize("n", [32, 64])
def test_parity_report_density_displacement_power(nb, n):
    """BASELINE.json north_star: results match the reference on identical white noise and cosmology within a stated
    float32 tolerance on the density field, displacements, power spectrum (grad(log-density): tests/test_api_model.py).
    n = 64 with a 640 Mpc/h box and 5 steps is BASELINE configs[0] at its full size."""
    from montecosmo_b200 import metrics as M
    from montecosmo_b200.cosmo import Cosmology
    from montecosmo_b200.model import FieldModel
    rng = np.random.default_rng(12)
    shape, box = (n, n, n), (10.0 * n,) * 3
    kw = dict(evolution="nbody", n_steps=5, a_obs=1.0, b1=0.5)
    m = FieldModel(shape, box, sigma_obs=1.0, **kw)
    white = rng.normal(size=shape).astype(np.float32)
    transfer = m.transfer.cpu().numpy().astype(np.float64)
    with torch.no_grad():
        field = m.evolve(torch.tensor(white, device=nb.ops().A.device))
        field_o = MO.evolve(torch.tensor(white.astype(np.float64)), transfer, O.Cosmology(), shape, **kw)
        dk = m.linear_field(torch.tensor(white, device=nb.ops().A.device))
        pos, _ = nb.nbody_bf(Cosmology(), dk, m.q, 0.0, 1.0, 5)
        pos_o, _ = O.nbody_bf(O.Cosmology(), torch.fft.rfftn(torch.tensor(white.astype(np.float64))) * O._t(transfer),
                              O.regular_pos(shape), 0.0, 1.0, 5)
    f, fo = field.cpu().numpy().astype(np.float64), field_o.numpy()
    assert np.linalg.norm(f - fo) / np.linalg.norm(fo) < 1e-4
    assert np.abs(pos[-1].cpu().numpy() - pos_o[-1].numpy()).max() < 2e-4
    kc, km, p = M._spectrum(field - 1.0, box_size=box)
    _, _, po = MX.spectrum(fo - 1.0, box_size=box)
    knyq = np.pi * shape[0] / box[0]
    ratio = p / po
    assert np.all(np.abs(ratio[km < knyq / 2] - 1) < 1e-4), ratio
    assert np.all(np.abs(ratio[km <= knyq] - 1) < 1e-3), ratio
    _, _, _, coh = M.powtranscoh(torch.tensor(fo - 1.0, dtype=torch.float32, device=field.device), field - 1.0, box)
    assert np.all(coh[km <= knyq] > 1 - 1e-6)
