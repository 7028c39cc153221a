def key(seed):
    raise NotImplementedError("jax.random is not emulated; goldens use numpy.random.default_rng")
PRNGKey = split = normal = uniform = key
