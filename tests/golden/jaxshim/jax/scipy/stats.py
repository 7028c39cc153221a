from scipy.stats import norm  # noqa: F401
