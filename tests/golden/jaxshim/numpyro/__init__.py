"""Import-only stand-in: montecosmo/utils.py defines distributions at import time; none is on the hot path."""
from . import distributions  # noqa: F401
