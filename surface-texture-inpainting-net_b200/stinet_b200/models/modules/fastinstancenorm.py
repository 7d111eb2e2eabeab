"""Mirror of reference models/modules/fastinstancenorm.py:11-110 (affine=False, track_running_stats=False, the
only configuration the reference instantiates, surfacetextureinpaintingnet.py:246-249).

forward(x, batch=None): batch None -> one instance over all rows (:44-49); otherwise per-graph statistics with the
reference's exact row partition (linspace slices for the sums, true counts for the divisor, `batch` for the lookup).
`batch` may also be a prebuilt stinet_b200.graph.Segments (no host sync).  Runs on stinet_segnorm_*.
"""
from __future__ import annotations

import torch

from ... import ops
from ..._abi import ACT_NONE
from ._structure import as_segments


class FastInstanceNorm(torch.nn.Module):
    def __init__(self, in_channels, eps=1e-5, momentum=0.1, affine=False, track_running_stats=False):
        super().__init__()
        if affine or track_running_stats:
            raise NotImplementedError("FastInstanceNorm: the reference only uses affine=False, track_running_stats=False")
        self.num_features = in_channels
        self.eps = eps
        self.momentum = momentum
        self.affine = affine
        self.track_running_stats = track_running_stats

    def forward(self, x, batch=None, residual=None, act=ACT_NONE):
        """`residual` / `act` expose the fused tail of GraphResnetBlock (out = residual + act(norm(x)))."""
        seg = as_segments(batch, x.shape[0], x.device)
        return ops.norm_act_res(x, residual, seg, True, act, self.eps)

    def __repr__(self):
        return f"{self.__class__.__name__}({self.num_features})"
