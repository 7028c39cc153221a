"""
Host-side cosmology scalars for the engine: background, growth tables and BullFrog coefficients.

These are a handful of float64 scalars per evaluation (the reference keeps them in JAX so that d/dOmega_m flows,
SURVEY section 8a "growth helpers").  Here they are torch-CPU float64 tensors: differentiable on the host, and the
engine returns cotangents for every coefficient it consumes (mcpm_lpt_vjp / mcpm_nbody_steps_vjp `coefbar`).

Follows montecosmo/nbody.py:675-808 (growth table, interpolators), 907-919 (alpha_bf) and jax_cosmo 0.1.0
background.{w, f_de, Esqr, Omega_m_a, Omega_de_a}, scipy.ode.odeint (fixed-step RK4 over the table nodes).
"""
import numpy as np
import torch

F64 = torch.float64


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, dtype=np.float64), dtype=F64)


class Cosmology:
    """Fields of jax_cosmo.Cosmology; defaults are AbacusSummit0 (bricks.py:40-50).  Fields may be torch scalars."""

    def __init__(self, Omega_c=0.26447041, Omega_b=0.04930169, h=0.6736, n_s=0.9649, sigma8=0.8076353990239834,
                 Omega_k=0.0, w0=-1.0, wa=0.0):
        self.Omega_c, self.Omega_b, self.h, self.n_s = Omega_c, Omega_b, h, n_s
        self.sigma8, self.Omega_k, self.w0, self.wa = sigma8, Omega_k, w0, wa
        self._workspace = {}

    @property
    def Omega_m(self):
        return self.Omega_b + self.Omega_c

    @property
    def Omega_de(self):
        return 1.0 - self.Omega_k - self.Omega_m


def w(c, a):
    return c.w0 + (1.0 - a) * c.wa


def f_de(c, a):
    eps = float(np.finfo(np.float32).eps)
    return -3.0 * (1.0 + c.w0) + 3.0 * c.wa * ((a - 1.0) / torch.log(a - eps) - 1.0)


def Esqr(c, a):
    a = _t(a)
    return c.Omega_m * a**-3 + c.Omega_k * a**-2 + c.Omega_de * a ** f_de(c, a)


def Omega_m_a(c, a):
    a = _t(a)
    return c.Omega_m * a**-3 / Esqr(c, a)


def Omega_de_a(c, a):
    a = _t(a)
    return c.Omega_de * a ** f_de(c, a) / Esqr(c, a)


GROWTH_LOG10_AMIN = -3.0  # nbody.py:675
GROWTH_STEPS = 128  # nbody.py:676


def _param_key(c):
    """Hashable snapshot of the cosmology, or None when a parameter carries a gradient (then nothing is cached)."""
    vals = (c.Omega_c, c.Omega_b, c.h, c.n_s, c.sigma8, c.Omega_k, c.w0, c.wa)
    if any(isinstance(v, torch.Tensor) and v.requires_grad for v in vals):
        return None
    return tuple(float(v) for v in vals)


def growth_table(c):
    """RK4 table of first/second-order growth and their log-derivatives, normalised at a = 1 (nbody.py:679-745).

    Cached on the cosmology object and keyed by its parameter values (the reference resets `_workspace` before every
    use, model.py:762,769, which under jit costs nothing; eagerly it would be ~2e4 tiny host ops per evaluation).
    """
    key = _param_key(c)
    hit = c._workspace.get("growth")
    if hit is not None and key is not None and hit[0] == key:
        return hit[1]
    atab = _t(np.logspace(GROWTH_LOG10_AMIN, 0.0, GROWTH_STEPS))

    def rhs(y, x):
        q = (2.0 - (Omega_m_a(c, x) + (1.0 + 3.0 * w(c, x)) * Omega_de_a(c, x)) / 2) / x
        r = 1.5 * Omega_m_a(c, x) / x**2
        g1, g2, f1, f2 = y[0, 0], y[0, 1], y[1, 0], y[1, 1]
        return torch.stack([torch.stack([f1, f2]), torch.stack([-q * f1 + r * g1, -q * f2 + r * g2 - r * g1**2])])

    a0 = atab[0]
    y = torch.stack([torch.stack([a0, -3.0 / 7 * a0**2]), torch.stack([torch.ones_like(a0), -6.0 / 7 * a0])])
    ys, prev = [], atab[0]
    for ti in atab:  # one classical RK4 step per table interval; the first has h = 0
        h = ti - prev
        k1 = rhs(y, prev)
        k2 = rhs(y + h * k1 / 2, prev + h / 2)
        k3 = rhs(y + h * k2 / 2, prev + h / 2)
        k4 = rhs(y + k3 * h, ti)
        y = y + h / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
        prev = ti
        ys.append(y)
    ys = torch.stack(ys)
    y1, y2 = ys[:, 0, 0], ys[:, 0, 1]
    g, g2 = y1 / y1[-1], y2 / y2[-1]
    tab = {"a": atab, "g": g, "g2": g2, "f": ys[:, 1, 0] / y1[-1] * atab / g, "f2": ys[:, 1, 1] / y2[-1] * atab / g2}
    c._workspace["growth"] = (key, tab)
    return tab


def interp(x, xp, fp):
    """jnp.interp: linear, clamped to the end values."""
    x = _t(x)
    xs = x.reshape(-1)
    i = torch.clamp(torch.searchsorted(xp.detach().contiguous(), xs.detach().contiguous(), right=True) - 1, 0,
                    len(xp) - 2)
    t = (xs - xp[i]) / (xp[i + 1] - xp[i])
    out = fp[i] + t * (fp[i + 1] - fp[i])
    out = torch.where(xs <= xp[0], fp[0].expand_as(out), out)
    out = torch.where(xs >= xp[-1], fp[-1].expand_as(out), out)
    return out.reshape(x.shape)


def safe_div(x, y):
    """utils.py:21-29."""
    x, y = _t(x), _t(y)
    ynz = torch.where(y == 0, torch.ones_like(y), y)
    return torch.where(y == 0, torch.zeros_like(x / ynz), x / ynz)


def a2g(c, a):
    t = growth_table(c)
    return interp(a, t["a"], t["g"])


def a2g2(c, a):
    t = growth_table(c)
    return interp(a, t["a"], t["g2"]) * -3 / 7


def a2f(c, a):
    t = growth_table(c)
    return interp(a, t["a"], t["f"])


def a2f2(c, a):
    t = growth_table(c)
    return interp(a, t["a"], t["f2"])


def a2dg2dg(c, a):
    return safe_div(a2g2(c, a) * a2f2(c, a), a2g(c, a) * a2f(c, a))


def g2a(c, g):
    t = growth_table(c)
    return interp(g, t["g"], t["a"])


def g2g2(c, g):
    t = growth_table(c)
    return interp(g, t["g"], t["g2"]) * -3 / 7


def g2f(c, g):
    t = growth_table(c)
    return interp(g, t["g"], t["f"])


def g2f2(c, g):
    t = growth_table(c)
    return interp(g, t["g"], t["f2"])


def g2dg2dg(c, g):
    return safe_div(g2g2(c, g) * g2f2(c, g), _t(g) * g2f(c, g))


RH = 2997.92458  # jax_cosmo.constants.rh, h^-1 Mpc
DIST_LOG10_AMIN = -3.0  # nbody.py:813
DIST_STEPS = 256  # nbody.py:814


def distance_table(c, log10_amin=DIST_LOG10_AMIN, steps=DIST_STEPS):
    """RK4 table of the radial comoving distance chi(a) = R_H int_a^1 da' / (a'^2 E(a')) in Mpc/h on `steps` log-spaced
    scale factors (nbody.py:816-857; jax_cosmo's dchioverda = R_H / (a^2 E) and its fixed-step RK4 odeint, one step per
    table interval in ln a).  Cached like growth_table."""
    key = _param_key(c)
    ckey = None if key is None else (key, float(log10_amin), int(steps))
    hit = c._workspace.get("distance")
    if hit is not None and ckey is not None and hit[0] == ckey:
        return hit[1]
    atab = _t(np.logspace(log10_amin, 0.0, steps))
    xs = torch.log(atab)

    def rhs(x):  # d chi / d ln a
        a = torch.exp(x)
        return RH / (a * Esqr(c, a) ** 0.5)

    y, prev, out = torch.zeros((), dtype=F64), xs[0], []
    for xi in xs:
        h = xi - prev
        k1 = rhs(prev)
        k2 = rhs(prev + h / 2)
        k4 = rhs(xi)
        y = y + h / 6.0 * (k1 + 4 * k2 + k4)  # the right-hand side does not depend on chi: k2 = k3
        prev = xi
        out.append(y)
    chi = torch.stack(out)
    tab = {"a": atab, "chi": chi[-1] - chi}
    c._workspace["distance"] = (ckey, tab)
    return tab


def a2chi(c, a, log10_amin=DIST_LOG10_AMIN, steps=DIST_STEPS):
    """Radial comoving distance in Mpc/h at scale factor a (nbody.py:816-857), clipped at 0."""
    t = distance_table(c, log10_amin, steps)
    return torch.clamp(interp(a, t["a"], t["chi"]), min=0.0)


def chi2a(c, chi, log10_amin=DIST_LOG10_AMIN, steps=DIST_STEPS):
    """Scale factor at radial comoving distance chi, by reverse linear interpolation (nbody.py:860-884)."""
    t = distance_table(c, log10_amin, steps)
    return interp(chi, torch.flip(t["chi"], dims=(0,)), torch.flip(t["a"], dims=(0,)))


def k2ell(c, a, k):
    """Comoving wavenumber -> multipole, Limber approximation (nbody.py:886-890)."""
    return a2chi(c, a) * _t(k) - 0.5


def ell2k(c, a, ell):
    """Multipole -> comoving wavenumber, Limber approximation (nbody.py:892-896)."""
    return (_t(ell) + 0.5) / a2chi(c, a)


def alpha_bf(c, g0, dg):
    """BullFrog kick coefficient, nbody.py:907-919."""
    g1, g2 = g0 + dg / 2, g0 + dg
    d0, d2 = g2dg2dg(c, g0), g2dg2dg(c, g2)
    lin_ratio = (g2g2(c, g0) + d0 * dg / 2) / g1 - g1
    return (d2 - lin_ratio) / (d0 - lin_ratio)


def alpha_fpm(c, g0, dg):
    """FastPM growth-time kick coefficient, nbody.py:921-931 (the reference defines it next to alpha_bf and leaves the
    switch commented out): E(a0) g0 f(g0) a0^2 / (E(a2) g2 f(g2) a2^2) with g2 = g0 + dg."""
    g0 = _t(g0)
    g2 = g0 + dg
    a0, a2 = g2a(c, g0), g2a(c, g2)
    coeff0 = Esqr(c, a0) ** 0.5 * g0 * g2f(c, g0) * a0**2
    coeff2 = Esqr(c, a2) ** 0.5 * g2 * g2f(c, g2) * a2**2
    return coeff0 / coeff2


def bullfrog_coefficients(c, a0, a1, n_steps, integrator="bullfrog"):
    """Per-step (alpha, beta = (1 - alpha) / g1, drift_pre, drift_post) of the DKD loop (nbody.py:933-951, 974-976).

    integrator = 'fastpm' uses alpha_fpm (nbody.py:921-931) in place of alpha_bf: the same loop, another kick coefficient.
    Times advance as diffrax's constant-step controller does: t_{s+1} = t_s + dg, the last step clipped onto g1.
    Returns four float64 tensors of length n_steps (differentiable w.r.t. the cosmology) and (g0, dg).
    """
    key = _param_key(c)
    if integrator not in ("bullfrog", "fastpm"):
        raise ValueError("integrator must be 'bullfrog' or 'fastpm'")
    kick = alpha_bf if integrator == "bullfrog" else alpha_fpm
    ckey = None if key is None or isinstance(a0, torch.Tensor) or isinstance(a1, torch.Tensor) else \
        (integrator, key, float(a0), float(a1), int(n_steps))
    if ckey is not None and c._workspace.get("bf_coef", (None,))[0] == ckey:
        return c._workspace["bf_coef"][1]
    g0, g1 = a2g(c, a0), a2g(c, a1)
    dg = (g1 - g0) / n_steps
    alphas, betas = [], []
    t = g0
    for _ in range(n_steps):
        al = kick(c, t, dg)
        alphas.append(al)
        betas.append((1 - al) / (t + dg / 2))
        t = t + dg
    half = dg / 2
    out = (torch.stack(alphas).reshape(-1), torch.stack(betas).reshape(-1), half.expand(n_steps).clone(),
           half.expand(n_steps).clone(), g0, dg)
    if ckey is not None:
        c._workspace["bf_coef"] = (ckey, out)
    return out
