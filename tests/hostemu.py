"""CPU build of the engine for GPU-less test runs: see oracle/cpu_port.py (checker only, never a product path)."""
from oracle.cpu_port import build, load  # noqa: F401
