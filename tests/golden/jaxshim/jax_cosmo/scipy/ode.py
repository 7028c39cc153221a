"""jax_cosmo.scipy.ode.odeint (0.1.0): classical RK4, one step per interval of the given nodes."""
import numpy as np


def odeint(fn, y0, t):
    y = np.asarray(y0, dtype=float)
    t_prev = np.asarray(t)[0]
    out = []
    for ti in np.asarray(t):
        h = ti - t_prev
        k1 = fn(y, t_prev)
        k2 = fn(y + h * k1 / 2, t_prev + h / 2)
        k3 = fn(y + h * k2 / 2, t_prev + h / 2)
        k4 = fn(y + k3 * h, ti)
        y = y + 1.0 / 6.0 * h * (k1 + 2 * k2 + 2 * k3 + k4)
        t_prev = ti
        out.append(np.asarray(y))
    from jax.numpy import JArray
    return np.stack(out).view(JArray)
