"""
CPU restatement of montecosmo/metrics.py's power-spectrum estimator (TEST INFRASTRUCTURE ONLY; same rules as
pm_oracle.py): _waves (metrics.py:60-118) and _spectrum (121-182) with multipoles, NumPy float64.  Pinned against the
reference's own source executed under the NumPy stand-in for JAX (tests/golden/spectrum.npz, bin counts identical,
powers to 1e-11; unpinned against real JAX/XLA like the rest of the oracle).  It is the yardstick for the engine's
binned reduction and for the power-spectrum tolerance of the parity report.
"""
import numpy as np

from . import pm_oracle as O


def waves(mesh_shape, box_size, kedges=None, include_corners=True):
    mesh_shape, box_size = np.asarray(mesh_shape), np.asarray(box_size, dtype=float)
    kvec = O.rfftk(tuple(int(s) for s in mesh_shape), tuple(box_size))  # h/Mpc
    kmesh = sum(k ** 2 for k in kvec) ** 0.5
    if isinstance(kedges, (type(None), int, float)):
        kmin, kmax = 0.0, np.pi * (mesh_shape / box_size).min()
        if include_corners:
            kmax = kmesh.max()
        if kedges is None:
            n = max(int((kmax - kmin) / (len(mesh_shape) ** 0.5 * 2 * np.pi / box_size.min())), 1)
        elif isinstance(kedges, int):
            n = kedges
        else:
            n = max(int((kmax - kmin) / kedges), 1)
        dk = (kmax - kmin) / n
        kedges = np.linspace(kmin, kmax, n, endpoint=False) + dk / 2
    rfftw = np.full(kmesh.shape, 2.0)
    rfftw[..., 0] = 1.0
    if mesh_shape[-1] % 2 == 0:
        rfftw[..., -1] = 1.0
    return np.asarray(kedges), kmesh, rfftw


def spectrum(mesh0, mesh1=None, box_size=None, kedges=None, include_corners=True, deconv=(0, 0), ells=0,
             box_center=(0.0, 0.0, 0.0)):
    """(kcount, kmean, P) per bin; cross spectra return |<m0 conj m1>| as the reference does (metrics.py:172-175).
    `ells` an int gives an array, a list a dict {ell: array}; multipoles carry (2 ell + 1) L_ell(mu) with mu along
    box_center / |box_center| (metrics.py:127-128, 91-92, 165-166)."""
    from numpy.polynomial import legendre as npleg

    def to_k(m):
        m = np.asarray(m)
        return (np.fft.rfftn(m), m.shape) if np.isrealobj(m) else (m.astype(np.complex128), O.ch2rshape(m.shape))
    m0, shape = to_k(mesh0)
    kcell = O.rfftk(tuple(shape))
    m0 = m0 / O.rectangular_hat(kcell, deconv[0])
    if mesh1 is None:
        mmk = m0.real ** 2 + m0.imag ** 2
    else:
        m1 = to_k(mesh1)[0] / O.rectangular_hat(kcell, deconv[1])
        mmk = m0 * m1.conj()
    shape = np.asarray(shape)
    box_size = shape.astype(float) if box_size is None else np.asarray(box_size, dtype=float)
    kedges, kmesh, rfftw = waves(shape, box_size, kedges, include_corners)
    center = np.asarray(box_center, dtype=float)
    nrm = np.linalg.norm(center)
    los = center / nrm if nrm != 0 else np.zeros(3)
    kvec = O.rfftk(tuple(int(s) for s in shape), tuple(box_size))
    kdot = sum(k * l for k, l in zip(kvec, los))
    mumesh = np.where(kmesh == 0, 0.0, kdot / np.where(kmesh == 0, 1.0, kmesh))
    nb = len(kedges) + 1
    dig = np.digitize(kmesh.reshape(-1), kedges)
    kcount = np.bincount(dig, weights=rfftw.reshape(-1), minlength=nb)[1:-1]
    kmean = np.bincount(dig, weights=(kmesh * rfftw).reshape(-1), minlength=nb)[1:-1] / kcount
    pows = {}
    for ell in np.atleast_1d(ells):
        leg = (2 * ell + 1) * npleg.legval(mumesh, [0.0] * int(ell) + [1.0])
        w = (mmk * leg * rfftw).reshape(-1)
        if mesh1 is None:
            p = np.bincount(dig, weights=w, minlength=nb)[1:-1]
        else:
            p = np.hypot(np.bincount(dig, weights=w.real, minlength=nb)[1:-1],
                         np.bincount(dig, weights=w.imag, minlength=nb)[1:-1])
        pows[int(ell)] = p * (box_size / shape ** 2).prod() / kcount
    return kcount, kmean, (pows[int(ells)] if isinstance(ells, (int, np.integer)) else pows)
