from . import special, stats, spatial  # noqa: F401
