"""Mirror of reference models/modules/singlebatchgroupnorm.py:10-74 (norm_type='graph', not used by the shipped
configs).  Round-1 status: expressed with ATen tensor ops on the device (segment sums via index_add_ over the
reference's linspace slices), NOT with hand-written kernels -- SURVEY 8a row a10, lower priority.

Reproduces the reference quirk that the variance is E[x^2] of the UN-shifted x (:66-68)."""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from ...graph import Segments
from ._structure import as_segments


class SingleBatchGraphNorm(torch.nn.Module):
    def __init__(self, in_channels: int, eps: float = 1e-5):
        super().__init__()
        self.in_channels = in_channels
        self.eps = eps
        self.weight = torch.nn.Parameter(torch.ones(in_channels))
        self.bias = torch.nn.Parameter(torch.zeros(in_channels))
        self.mean_scale = torch.nn.Parameter(torch.ones(in_channels))

    def forward(self, x: Tensor, batch=None) -> Tensor:
        seg: Segments = as_segments(batch, x.shape[0], x.device)
        n = x.shape[0]
        ptr = seg.slice_ptr.long()
        lens = (ptr[1:] - ptr[:-1])
        sid = torch.repeat_interleave(torch.arange(seg.n_seg, device=x.device), lens, output_size=n)
        gid = seg.gid.long() if seg.gid is not None else sid
        denom = lens.to(x.dtype).unsqueeze(1)           # .mean() over the slice (:60,:67): divisor = slice length
        mean = torch.zeros((seg.n_seg, x.shape[1]), dtype=x.dtype, device=x.device).index_add_(0, sid, x) / denom
        out = x - mean.index_select(0, gid) * self.mean_scale
        var = torch.zeros((seg.n_seg, x.shape[1]), dtype=x.dtype, device=x.device).index_add_(0, sid, x * x) / denom
        std = (var + self.eps).sqrt().index_select(0, gid)
        return self.weight * out / std + self.bias

    def __repr__(self):
        return f"{self.__class__.__name__}({self.in_channels})"
