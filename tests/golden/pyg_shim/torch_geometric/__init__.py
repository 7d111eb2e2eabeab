"""Minimal restatement of the torch_geometric 2.0.x symbols the reference imports (test-only)."""
__version__ = "2.0.3-shim"
from . import typing, utils, data, nn  # noqa: F401
