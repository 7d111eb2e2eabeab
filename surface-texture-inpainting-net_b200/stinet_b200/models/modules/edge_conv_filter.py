"""Mirror of reference models/modules/edge_conv_filter.py: `get_gcn_filter` returns an EdgeConv module whose
`.nn` is Sequential(Linear(k, 2*out), ReLU, Linear(2*out, out)) -- same parameters, same state_dict keys.

The arithmetic is evaluated in the hoisted, segmented form on sm_100a kernels (SURVEY 7 "where the FLOPs are"):
    nn.0([x_i || x_j - x_i]) = P_i + Q_j          P = X (Wa - Wb)^T + b0,  Q = X Wb^T        (vertex GEMM)
    hid_i = mean_{j->i} relu(P_i + Q_j)                                                      (stinet_edge_message_*)
    out_i = hid_i W2^T + b2 * [deg_i > 0]                                                    (vertex GEMM)
which equals mean_{j->i} nn([x_i || x_j - x_i]) of PyG's EdgeConv(aggr='mean'), including 0 for isolated vertices.

`with_norm=True` (reference :34-44, used by SingleConvMeshNet: BatchNorm1d over the EDGES between the two Linears and
after the second one) cannot be hoisted -- the statistics are taken over per-edge activations -- so that variant is
evaluated literally on [E, .] matrices: x_i / x_j row gathers (stinet_unpool_*, the edges being the "fine" side of a
cluster map onto their end points), the Linears as tcgen05 GEMMs over E rows, BatchNorm1d / ReLU as device tensor ops,
and the mean over in-edges as stinet_pool_mean_* (deterministic segmented sums both ways, no atomics).
"""
from __future__ import annotations

import torch
from torch.nn import Linear as Lin, Sequential as Seq

from ... import ops
from ._structure import as_edge_csr


class EdgeConv(torch.nn.Module):
    """x_i' = aggr_{j->i} nn([x_i || x_j - x_i])   (PyG EdgeConv; the reference always passes aggr='mean')."""

    trans_inv = False

    def __init__(self, nn: torch.nn.Module, aggr: str = "mean"):
        super().__init__()
        if aggr not in ("mean", "add", "max"):
            raise ValueError(f"EdgeConv(aggr={aggr!r}): expected 'mean' (the reference's choice, edge_conv_filter.py:11), "
                             "'add' or 'max'")
        if not isinstance(nn, Seq):
            raise NotImplementedError("EdgeConv expects nn = torch.nn.Sequential (edge_conv_filter.py:34-55)")
        # Sequential(Linear, ReLU, Linear) under a mean is evaluated in the hoisted form; anything else (the with_norm
        # variant; 'add' / 'max', through which the second Linear and its bias do not commute) literally, per edge
        self.hoistable = (aggr == "mean" and len(nn) == 3 and isinstance(nn[0], Lin) and isinstance(nn[2], Lin)
                          and isinstance(nn[1], torch.nn.ReLU))
        self.nn = nn
        self.aggr = self._aggr = aggr
        self.precision = "fp32"

    def hoisted_first_layer(self):
        """[P | Q] weights/bias of the first Linear: W [2H, din], b [2H]."""
        # trans_inv: nn.0(x_j - x_i) = (-W) x_i + W x_j;  else nn.0([x_i || x_j - x_i]) = (Wa - Wb) x_i + Wb x_j
        return ops.edgeconv_hoist(self.nn[0].weight, self.nn[0].bias, self.trans_inv)

    def literal_forward(self, x, csr):
        """mean_{j->i} nn([x_i || x_j - x_i]) with nn applied to the [E, .] message matrix (PyG's own order of
        evaluation); rows of the message matrix are in ORIGINAL edge order."""
        by_target, by_source = csr.edge_clusters()
        x_i = ops.unpool(x, by_target)
        x_j = ops.unpool(x, by_source)
        m = (x_j - x_i) if self.trans_inv else torch.cat([x_i, x_j - x_i], dim=-1)
        for layer in self.nn:
            if isinstance(layer, Lin):
                m = ops.linear(m, layer.weight, layer.bias, None, self.precision)
            else:
                m = layer(m)
        if self.aggr == "add":
            return ops.pool_sum(m, by_target)
        if self.aggr == "max":                       # first maximum wins, vertices without in-edges get 0 (torch_scatter)
            return ops.pool_max(m, by_target)[0]
        return ops.pool_mean(m, by_target)

    def _fused_forward(self, x, csr):
        """The whole conv as one autograd node on operand planes (ops.EdgeConvFn); input widths that are not a multiple
        of 4 (the 10-channel first block) are zero-padded -- the extra products are exact zeros."""
        lin0, lin2 = self.nn[0], self.nn[2]
        if self.precision not in ("fp32", "f16") or lin0.out_features % 4 or lin2.out_features % 4:
            return None
        w0 = lin0.weight
        pad = (-x.shape[1]) % 4
        if pad:
            x = torch.nn.functional.pad(x, (0, pad))
            if self.trans_inv:
                w0 = torch.nn.functional.pad(w0, (0, pad))
            else:
                h, k2 = w0.shape
                w0 = torch.nn.functional.pad(w0.view(h, 2, k2 // 2), (0, pad)).reshape(h, -1)
        return ops.edge_conv(x, w0, lin0.bias, lin2.weight, lin2.bias, csr, self.trans_inv, self.precision)

    def forward(self, x, edge_index):
        csr = as_edge_csr(edge_index, x.shape[0])
        if not self.hoistable:
            return self.literal_forward(x, csr)
        fused = self._fused_forward(x, csr)
        if fused is not None:
            return fused
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
        # reference :34-44 -- no biases, BatchNorm1d over the edge dimension after each Linear
        inner_module = Seq(
            Lin(double_input_size, 2 * output_size, bias=False),
            ops.BatchNorm1d(2 * output_size),
            activation(inplace=inplace),
            Lin(2 * output_size, output_size, bias=False),
            ops.BatchNorm1d(output_size),
        )
        return module(inner_module, aggr=aggregation)
    if activation is not torch.nn.ReLU:
        raise NotImplementedError("the fused message kernel implements the reference's ReLU (edge_conv_filter.py:10)")
    inner_module = Seq(
        Lin(double_input_size, 2 * output_size, bias=bias),
        activation(inplace=inplace),
        Lin(2 * output_size, output_size, bias=bias),
    )
    return module(inner_module, aggr=aggregation)
