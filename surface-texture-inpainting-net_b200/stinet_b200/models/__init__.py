"""Host-side mirror of the reference's `models` package for the hot path (same module and symbol names)."""
from . import modules, surfacetextureinpaintingnet  # noqa: F401
from .surfacetextureinpaintingnet import GraphResnetBlock, SurfaceTextureInpaintingNet, define_G  # noqa: F401
