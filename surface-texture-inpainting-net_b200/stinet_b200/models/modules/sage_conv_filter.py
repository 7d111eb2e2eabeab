"""Mirror of reference models/modules/sage_conv_filter.py: SAGEConv / SAGEConvTransInv wrapped in SumSAGEConv
(state_dict keys `sage1.lin_l.{weight,bias}`, `sage1.lin_r.weight`).

    out_i = lin_l(mean_{j->i} msg_j) + lin_r(x_i);   TransInv: msg_j = x_j with columns 3:9 made relative to x_i
The mean over in-edges runs on stinet_aggregate_* (CSR, no atomics).  mean_j(x_j[3:9] - x_i[3:9]) is evaluated as
mean_j x_j[3:9] - x_i[3:9]*[deg_i>0] (sage_conv_filter.py:87-90).
"""
from __future__ import annotations

import torch
from torch.nn import Linear

from ... import ops
from ._structure import as_edge_csr


class SAGEConv(torch.nn.Module):
    trans_inv = False

    def __init__(self, in_channels, out_channels, normalize: bool = False, root_weight: bool = True,
                 bias: bool = True, **kwargs):
        super().__init__()
        if kwargs.get("aggr", "mean") != "mean":
            raise NotImplementedError("SAGEConv: only mean aggregation (the reference's default) is implemented")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.normalize, self.root_weight = normalize, root_weight
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        self.lin_l = Linear(in_channels[0], out_channels, bias=bias)
        if self.root_weight:
            self.lin_r = Linear(in_channels[1], out_channels, bias=False)
        self.precision = "fp32"

    def forward(self, x, edge_index, size=None):
        csr = as_edge_csr(edge_index, x.shape[0])
        agg = ops.aggregate(x, csr, "mean")
        if self.trans_inv:
            has_nbr = (csr.degree > 0).to(x.dtype).unsqueeze(1)
            agg = torch.cat([agg[:, :3], agg[:, 3:9] - x[:, 3:9] * has_nbr, agg[:, 9:]], dim=1)
        out = ops.linear(agg, self.lin_l.weight, self.lin_l.bias, None, self.precision)
        if self.root_weight:
            out = out + ops.linear(x, self.lin_r.weight, None, None, self.precision)
        if self.normalize:
            out = torch.nn.functional.normalize(out, p=2.0, dim=-1)
        return out

    def __repr__(self):
        return "{}({}, {})".format(self.__class__.__name__, self.in_channels, self.out_channels)


class SAGEConvTransInv(SAGEConv):
    trans_inv = True


def get_gcn_filter(input_size: int, output_size, activation: torch.nn.Module = None,
                   inplace: bool = False, aggregation: str = "mean", bias: bool = True,
                   module=None, double_input=False):
    """Same signature as reference sage_conv_filter.py:102-104."""
    assert input_size >= 0
    assert output_size >= 0
    if module is None:
        module = SAGEConv

    class SumSAGEConv(torch.nn.Module):
        def __init__(self, *args, **kwargs):
            super().__init__()
            self.sage1 = module(*args, **kwargs)

        def forward(self, x, edge_index, size=None):
            return self.sage1(x, edge_index, size)

    return SumSAGEConv(input_size, output_size, normalize=False, root_weight=True, bias=bias)
