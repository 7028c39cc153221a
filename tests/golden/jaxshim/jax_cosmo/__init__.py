"""Restatement of the jax_cosmo 0.1.0 pieces montecosmo's hot path calls (montenv.yml:360). NumPy, float64."""
from . import background, constants, power  # noqa: F401
from .core import Cosmology  # noqa: F401
