"""jax_cosmo.background (0.1.0), the functions montecosmo calls: nbody.py:705-707,848,929-930."""
import numpy as np
from . import constants as const


def w(cosmo, a):
    return cosmo.w0 + (1.0 - a) * cosmo.wa


def f_de(cosmo, a):
    epsilon = np.finfo(np.float32).eps
    return -3.0 * (1.0 + cosmo.w0) + 3.0 * cosmo.wa * ((a - 1.0) / np.log(a - epsilon) - 1.0)


def Esqr(cosmo, a):
    return (cosmo.Omega_m * np.power(a, -3) + cosmo.Omega_k * np.power(a, -2)
            + cosmo.Omega_de * np.power(a, f_de(cosmo, a)))


def H(cosmo, a):
    return const.H0 * np.sqrt(Esqr(cosmo, a))


def Omega_m_a(cosmo, a):
    return cosmo.Omega_m * np.power(a, -3) / Esqr(cosmo, a)


def Omega_de_a(cosmo, a):
    return cosmo.Omega_de * np.power(a, f_de(cosmo, a)) / Esqr(cosmo, a)


def dchioverda(cosmo, a):
    return const.rh / (a**2 * np.sqrt(Esqr(cosmo, a)))
