"""Stand-in for diffrax 0.5.0 (montenv.yml:352) as used at nbody.py:963-1000: Euler with a constant step."""
import numpy as np
from jax import tree


class ODETerm:
    def __init__(self, vector_field):
        self.vector_field = vector_field


class Euler:
    pass


class SaveAt:
    def __init__(self, t0=False, t1=False, ts=None, fn=None):
        self.t0, self.t1, self.ts = t0, t1, ts
        self.fn = fn if fn is not None else (lambda t, y, args: y)


class _Solution:
    pass


def diffeqsolve(terms, solver, t0, t1, dt0, y0, args=None, max_steps=None, saveat=None, **kw):
    saveat = saveat or SaveAt(t1=True)
    t, y, n = t0, y0, 0
    traj_t, traj_y = [t0], [y0]
    while t < t1:
        assert max_steps is None or n < max_steps, "max_steps reached"
        tn = min(t + dt0, t1)
        if max_steps is not None and n == max_steps - 1:
            tn = t1  # constant-step controller clips the last step onto t1
        dt = tn - t
        vf = terms.vector_field(t, y, args)
        y = tree.map(lambda yi, fi: yi + fi * dt, y, vf)
        t = tn
        n += 1
        traj_t.append(t)
        traj_y.append(y)
    sol = _Solution()
    sol.stats = {"num_steps": n}
    if saveat.ts is not None:
        ts = np.atleast_1d(np.asarray(saveat.ts, dtype=float))
        tt = np.asarray(traj_t, dtype=float)
        outs = []
        for s in ts:  # Euler dense output = linear interpolation inside the step
            i = int(np.clip(np.searchsorted(tt, s, side="right") - 1, 0, len(tt) - 2))
            th = (s - tt[i]) / (tt[i + 1] - tt[i])
            ys = tree.map(lambda a, b: a + th * (b - a), traj_y[i], traj_y[i + 1])
            outs.append(saveat.fn(s, ys, args))
        sol.ts = ts
        sol.ys = tree.map(lambda *a: np.stack(a), *outs)
    else:
        out = saveat.fn(t, y, args)
        sol.ts = np.asarray([t])
        sol.ys = tree.map(lambda a: np.asarray(a)[None], out)
    return sol


class _Unused:
    """Names imported by the legacy Tsit5 path (nbody.py:1125), which the model never calls."""

    def __init__(self, *a, **k):
        raise NotImplementedError("only Euler + constant step is emulated")


Heun = Dopri5 = Tsit5 = PIDController = ConstantStepSize = _Unused
