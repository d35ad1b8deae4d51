"""TEST-ONLY (see package docstring)."""
from . import registration  # noqa: F401
