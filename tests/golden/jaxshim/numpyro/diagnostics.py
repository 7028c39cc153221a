"""Import-only stand-in: montecosmo/metrics.py imports these chain diagnostics at module level; none is on the path."""


def effective_sample_size(*a, **k):
    raise NotImplementedError("numpyro.diagnostics stand-in")


def gelman_rubin(*a, **k):
    raise NotImplementedError("numpyro.diagnostics stand-in")
