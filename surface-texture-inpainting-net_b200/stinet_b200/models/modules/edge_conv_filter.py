"""Mirror of reference models/modules/edge_conv_filter.py: `get_gcn_filter` returns an EdgeConv module whose
`.nn` is Sequential(Linear(k, 2*out), ReLU, Linear(2*out, out)) -- same parameters, same state_dict keys.

The arithmetic is evaluated in the hoisted, segmented form on sm_100a kernels (SURVEY 7 "where the FLOPs are"):
    nn.0([x_i || x_j - x_i]) = P_i + Q_j          P = X (Wa - Wb)^T + b0,  Q = X Wb^T        (vertex GEMM)
    hid_i = mean_{j->i} relu(P_i + Q_j)                                                      (stinet_edge_message_*)
    out_i = hid_i W2^T + b2 * [deg_i > 0]                                                    (vertex GEMM)
which equals mean_{j->i} nn([x_i || x_j - x_i]) of PyG's EdgeConv(aggr='mean'), including 0 for isolated vertices.
"""
from __future__ import annotations

import torch
from torch.nn import Linear as Lin, Sequential as Seq

from ... import ops
from ._structure import as_edge_csr


class EdgeConv(torch.nn.Module):
    """x_i' = mean_{j->i} nn([x_i || x_j - x_i])   (PyG EdgeConv with the reference's aggr='mean')."""

    trans_inv = False

    def __init__(self, nn: torch.nn.Module, aggr: str = "mean"):
        super().__init__()
        if aggr != "mean":
            raise NotImplementedError(
                f"EdgeConv(aggr={aggr!r}): only the reference's 'mean' aggregation (edge_conv_filter.py:11) is fused; "
                "use stinet_b200.ops.aggregate for add/max over explicit messages")
        if not (isinstance(nn, Seq) and len(nn) == 3 and isinstance(nn[0], Lin) and isinstance(nn[2], Lin)
                and isinstance(nn[1], torch.nn.ReLU)):
            raise NotImplementedError("EdgeConv expects nn = Sequential(Linear, ReLU, Linear) (edge_conv_filter.py:46-55); "
                                      "the with_norm=True variant (:34-44) is not part of the STINet path")
        self.nn = nn
        self.aggr = self._aggr = aggr
        self.precision = "fp32"

    def hoisted_first_layer(self):
        """[P | Q] weights/bias of the first Linear: W [2H, din], b [2H]."""
        # trans_inv: nn.0(x_j - x_i) = (-W) x_i + W x_j;  else nn.0([x_i || x_j - x_i]) = (Wa - Wb) x_i + Wb x_j
        return ops.edgeconv_hoist(self.nn[0].weight, self.nn[0].bias, self.trans_inv)

    def forward(self, x, edge_index):
        csr = as_edge_csr(edge_index, x.shape[0])
        wcat, bcat = self.hoisted_first_layer()
        pq = ops.linear(x, wcat, bcat, None, self.precision)
        hid = ops.edge_message(pq, csr)
        return ops.linear(hid, self.nn[2].weight, self.nn[2].bias, csr.degree, self.precision)

    def __repr__(self):
        return "{}(nn={}, aggr={})".format(self.__class__.__name__, self.nn, self.aggr)


def get_gcn_filter(input_size: int, output_size, activation: torch.nn.Module = torch.nn.ReLU,
                   inplace: bool = False, aggregation: str = "mean", bias: bool = True,
                   module=None, double_input=True, with_norm=False):
    """Same signature and defaults as reference edge_conv_filter.py:10-12."""
    assert input_size >= 0
    assert output_size >= 0
    double_input_size = 2 * input_size if double_input else input_size
    if module is None:
        module = EdgeConv
    if with_norm:
        raise NotImplementedError("with_norm=True (BatchNorm1d over edges, edge_conv_filter.py:34-44) is only used by "
                                  "SingleConvMeshNet and is out of scope for the STINet hot path")
    if activation is not torch.nn.ReLU:
        raise NotImplementedError("the fused message kernel implements the reference's ReLU (edge_conv_filter.py:10)")
    inner_module = Seq(
        Lin(double_input_size, 2 * output_size, bias=bias),
        activation(inplace=inplace),
        Lin(2 * output_size, output_size, bias=bias),
    )
    return module(inner_module, aggr=aggregation)
