from scipy.special import logsumexp, gammaln, erf, erfc, ndtr, ndtri, i0, i0e, i1, i1e  # noqa: F401
