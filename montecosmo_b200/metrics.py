"""
Mirror of the power-spectrum estimators of `montecosmo/metrics.py` (spectrum 184-187, transfer 190-194, coherence
196-201, powtranscoh 203-210; binning of _waves 60-118 and _spectrum 121-182), every multipole.  The FFT and the binned
reduction over the half spectrum run on the engine (mcpm_rfftn, mcpm_spectrum_bins); the bin edges and the final
normalisation are a few host-side float64 operations, as in the reference.
"""
import numpy as np
import torch

from . import nbody as _nb
from .ops import ch2rshape


def _kedges(mesh_shape, box_size, kedges, include_corners):
    """Bin edges of metrics._waves (60-108): dk = sqrt(dim) x fundamental by default, edges at kmin + dk/2 ..."""
    mesh_shape, box_size = np.asarray(mesh_shape, dtype=float), np.asarray(box_size, dtype=float)
    if isinstance(kedges, (type(None), int, float)):
        dim = len(mesh_shape)
        kmin = 0.0
        kmax = np.pi * (mesh_shape / box_size).min()  # Nyquist
        if include_corners:  # = kmesh.max(): every axis at its Nyquist frequency (even sides)
            freq = np.array([s // 2 for s in mesh_shape.astype(int)], dtype=float)
            kmax = float(np.sqrt(((2 * np.pi * freq / box_size) ** 2).sum()))
        if kedges is None:
            dk = dim ** 0.5 * 2 * np.pi / box_size.min()
            n_kedges = max(int((kmax - kmin) / dk), 1)
        elif isinstance(kedges, int):
            n_kedges = kedges
        else:
            n_kedges = max(int((kmax - kmin) / kedges), 1)
        dk = (kmax - kmin) / n_kedges
        kedges = np.linspace(kmin, kmax, n_kedges, endpoint=False) + dk / 2
    return np.asarray(kedges, dtype=np.float64)


def _to_spectrum(mesh):
    mesh = torch.as_tensor(mesh)
    if torch.is_complex(mesh):
        return _nb._c64(mesh).detach()
    return _nb.ops().rfftn(_nb._f32(mesh).detach())


def _spectrum(mesh0, mesh1=None, box_size=None, box_center=(0.0, 0.0, 0.0), ells=0, kedges=None, include_corners=True,
              deconv=(0, 0)):
    """metrics.py:121-182: (kcount, kmean, P_ell) with P_ell an array for an int `ells`, a dict {ell: array} for a list;
    the line of sight of the multipoles is box_center / |box_center| (zero for a centred box, metrics.py:127-128)."""
    if isinstance(deconv, int):
        deconv = (deconv, deconv)
    center = np.asarray(box_center, dtype=float)
    nrm = np.linalg.norm(center)
    los = center / nrm if nrm != 0 else np.zeros_like(center)  # safe_div
    m0 = _to_spectrum(mesh0)
    m1 = None if mesh1 is None else _to_spectrum(mesh1)
    mesh_shape = np.array(ch2rshape(tuple(m0.shape)))
    box_size = mesh_shape.astype(float) if box_size is None else np.asarray(box_size, dtype=float)
    edges = _kedges(mesh_shape, box_size, kedges, include_corners)
    pow, kcount, kmean = {}, None, None
    for ell in np.atleast_1d(ells):
        sums = _nb.ops().spectrum_bins(m0, m1, box_size, edges, deconv, int(ell), los).cpu().numpy()[:, 1:-1]
        kcount = sums[0]
        kmean = sums[1] / kcount
        pmean = sums[2] if m1 is None else np.sqrt(sums[2] ** 2 + sums[3] ** 2)
        pow[int(ell)] = pmean * (box_size / mesh_shape ** 2).prod() / kcount  # from cell units to [Mpc/h]^3
    if isinstance(ells, (int, np.integer)):
        return kcount, kmean, pow[int(ells)]
    return kcount, kmean, pow


def spectrum(mesh0, mesh1=None, box_size=None, box_center=(0.0, 0.0, 0.0), ells=0, kedges=None, include_corners=True):
    """(kmean, P(k)) of a real mesh or half spectrum; cross spectrum amplitude with mesh1 (metrics.py:184-187)."""
    _, kmean, pmean = _spectrum(mesh0, mesh1, box_size, box_center, ells, kedges, include_corners)
    return kmean, pmean


def transfer(mesh0, mesh1, box_size, kedges=None, include_corners=True):
    ks, pow0 = spectrum(mesh0, box_size=box_size, kedges=kedges, include_corners=include_corners)
    ks, pow1 = spectrum(mesh1, box_size=box_size, kedges=kedges, include_corners=include_corners)
    return ks, (pow1 / pow0) ** 0.5


def coherence(mesh0, mesh1, box_size, kedges=None, include_corners=True):
    ks, pow01 = spectrum(mesh0, mesh1, box_size=box_size, kedges=kedges, include_corners=include_corners)
    ks, pow0 = spectrum(mesh0, box_size=box_size, kedges=kedges, include_corners=include_corners)
    ks, pow1 = spectrum(mesh1, box_size=box_size, kedges=kedges, include_corners=include_corners)
    return ks, pow01 / (pow0 * pow1) ** 0.5


def powtranscoh(mesh0, mesh1, box_size, kedges=None, include_corners=True):
    ks, pow01 = spectrum(mesh0, mesh1, box_size=box_size, kedges=kedges, include_corners=include_corners)
    ks, pow0 = spectrum(mesh0, box_size=box_size, kedges=kedges, include_corners=include_corners)
    ks, pow1 = spectrum(mesh1, box_size=box_size, kedges=kedges, include_corners=include_corners)
    return ks, pow1, (pow1 / pow0) ** 0.5, pow01 / (pow0 * pow1) ** 0.5
