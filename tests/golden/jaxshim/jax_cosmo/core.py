class Cosmology:
    """jax_cosmo.core.Cosmology: container + derived Omega_m, Omega_de, Omega; `_workspace` cache dict."""

    def __init__(self, Omega_c, Omega_b, h, n_s, sigma8, Omega_k, w0, wa, gamma=None):
        self._Omega_c, self._Omega_b, self._h, self._n_s = Omega_c, Omega_b, h, n_s
        self._sigma8, self._Omega_k, self._w0, self._wa = sigma8, Omega_k, w0, wa
        self._workspace = {}

    Omega_c = property(lambda s: s._Omega_c)
    Omega_b = property(lambda s: s._Omega_b)
    h = property(lambda s: s._h)
    n_s = property(lambda s: s._n_s)
    sigma8 = property(lambda s: s._sigma8)
    Omega_k = property(lambda s: s._Omega_k)
    w0 = property(lambda s: s._w0)
    wa = property(lambda s: s._wa)
    Omega = property(lambda s: 1.0 - s._Omega_k)
    Omega_m = property(lambda s: s._Omega_b + s._Omega_c)
    Omega_de = property(lambda s: s.Omega - s.Omega_m)
