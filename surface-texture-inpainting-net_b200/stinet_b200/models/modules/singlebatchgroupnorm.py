"""Mirror of reference models/modules/singlebatchgroupnorm.py:10-74 (norm_type='graph', not used by the shipped
configs): y = weight * (x - mean_scale * E[x]) / sqrt(E[x^2] + eps) + bias per graph, where -- a quirk of the reference
that is reproduced -- the "variance" is the second moment of the UN-shifted x (:66-68) and both moments are taken over the
reference's `linspace` slices of the batch (:52-60), divided by the slice length.

Evaluated by the affine segmented-norm kernels (stinet_affnorm_*, kind 1): deterministic two-stage column reductions, one
elementwise apply, explicit backward incl. the three parameter gradients.  Only a ragged batch whose linspace slices cut
across graphs (rows of one slice normalised with another slice's statistics; the reference never trains there) keeps the
composite of device tensor ops, whose gradient autograd derives."""
from __future__ import annotations

import torch
from torch import Tensor

from ... import ops
from ...graph import Segments, _segment_tables
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
        if not (x.is_cuda and seg.consistent):
            return self._forward_composite(x, seg)
        sp = seg.slice_ptr_host
        lens = tuple(max(b - a, 1) for a, b in zip(sp[:-1], sp[1:]))
        slice_ptr, cnt = _segment_tables(tuple(sp), lens, x.device)            # divisor = slice length (.mean(), :60,:67)
        out, _, _ = ops.AffNormFn.apply(x, self.mean_scale, self.weight, self.bias, slice_ptr, cnt, seg.gid, seg.n_seg,
                                        seg.max_seg_rows, 1, self.eps)
        return out

    def _forward_composite(self, x: Tensor, seg: Segments) -> Tensor:
        n = x.shape[0]
        ptr = seg.slice_ptr.long()
        lens = (ptr[1:] - ptr[:-1])
        sid = torch.repeat_interleave(torch.arange(seg.n_seg, device=x.device), lens, output_size=n)
        gid = seg.gid.long() if seg.gid is not None else sid
        denom = lens.to(x.dtype).unsqueeze(1)
        mean = torch.zeros((seg.n_seg, x.shape[1]), dtype=x.dtype, device=x.device).index_add_(0, sid, x) / denom
        out = x - mean.index_select(0, gid) * self.mean_scale
        var = torch.zeros((seg.n_seg, x.shape[1]), dtype=x.dtype, device=x.device).index_add_(0, sid, x * x) / denom
        std = (var + self.eps).sqrt().index_select(0, gid)
        return self.weight * out / std + self.bias

    def __repr__(self):
        return f"{self.__class__.__name__}({self.in_channels})"
