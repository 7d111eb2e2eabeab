"""stinet_b200 -- B200-native (sm_100a) implementation of STINet's hot path: the multi-level mesh U-Net forward and
backward over batched mesh graphs, behind the reference's own module API.

    from stinet_b200.models import surfacetextureinpaintingnet
    net = surfacetextureinpaintingnet.define_G(**config['archs']['SurfaceTextureInpaintingNet']['args'], gpu_ids=[dev])

All compute goes through libstinet_b200.so (C ABI in include/stinet_b200.h); there is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"
