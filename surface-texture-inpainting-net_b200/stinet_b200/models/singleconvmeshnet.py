"""Mirror of reference models/singleconvmeshnet.py:10-156: the segmentation U-Net that runs
EdgeConv with BatchNorm1d over the EDGES inside the message MLP (edge_conv_filter.py:34-44), mean / max trace
pooling and unpooling with skip concatenation.  Same constructor, same sub-module names and creation order (identical
state_dict keys incl. the BatchNorm buffers), same forward schedule -- including torch.utils.checkpoint around every
block that the reference checkpoints (:129-131, :146-148), because with BatchNorm the recomputation is observable:
those blocks' running statistics receive two momentum updates per training step.

The per-edge MLP cannot be hoisted to the vertices here (its statistics are over edges), so this network exercises
the literal path of stinet_b200.models.modules.edge_conv_filter.EdgeConv: row gathers and segmented sums on the
pooling kernels, Linears as tcgen05 GEMMs over E rows, BatchNorm1d on the segmented-reduction kernels (ops.BatchNorm1d),
the skip concatenation written in place by the unpool kernel (ops.unpool_concat); ReLU is a device tensor op.

Reference quirk kept: ResBlock adds the residual in place onto a ReLU output (:107), which makes the reference's own
backward raise for num_propagation_steps > 1; the sum is written out of place here (same values), so more than one
propagation step trains.
"""
from __future__ import annotations

import torch
from torch.nn import BatchNorm1d, Linear as Lin, ReLU, Sequential as Seq, functional as F
from torch.utils import checkpoint

from .. import ops
from ..graph import GraphCache
from .modules.edge_conv_filter import get_gcn_filter
from .modules.edge_conv_translation_invariance import EdgeConvTransInv


class SingleConvMeshNet(torch.nn.Module):
    def __init__(self, feature_number, num_propagation_steps, filter_sizes, num_classes=3, pooling_method='mean',
                 aggr='mean', precision='fp32'):
        super().__init__()
        self._pooling_method = pooling_method
        self._activation, self._act = ReLU, F.relu                  # the reference hard-wires 'ReLU' (:17, :24-26)
        self._graph_levels = len(filter_sizes)
        widths = list(filter_sizes)

        def conv(din, dout, first=False):
            # every conv carries BatchNorm over edges; only the very first one is translation invariant (:44-51)
            extra = dict(module=EdgeConvTransInv, double_input=False) if first else {}
            return get_gcn_filter(din, dout, self._activation, aggregation=aggr, with_norm=True, **extra)

        def stack(din, dout, first=False):
            return [conv(din, dout, first)] + [conv(dout, dout) for _ in range(num_propagation_steps - 1)]

        # Modules are created level by level, encoder stack before decoder stack, exactly as the reference does: a seeded
        # construction then draws the same initial weights (tests/test_singleconv.py checks it bit for bit).
        encoders, decoders = [], []
        din = feature_number
        for level, width in enumerate(widths):
            enc = stack(din, width, first=(level == 0))
            if level + 1 < len(widths):
                decoders.append(self.ResBlock(stack(width + widths[level + 1], width), self._act))   # skip || unpooled (:58)
            encoders.append(self.ResBlock(enc, self._act))
            din = width
        head = Seq(Lin(widths[0], widths[0] // 2), ops.BatchNorm1d(widths[0] // 2), self._activation(inplace=False),
                   Lin(widths[0] // 2, num_classes))
        self.left_geo_cnns = torch.nn.ModuleList(encoders)
        self.right_geo_cnns = torch.nn.ModuleList(decoders)
        self.final_convs = torch.nn.ModuleList([head])
        self.set_precision(precision)

    def set_precision(self, precision: str):
        assert precision in ('fp32', 'f16', 'bf16', 'bf16x3', 'fp32_simt', 'tf32', 'fp32_tf32x3')
        self.precision = precision
        for m in self.modules():
            if m is not self and hasattr(m, 'precision'):
                m.precision = precision
        return self

    class ResBlock(torch.nn.Module):
        def __init__(self, filters, act):
            super().__init__()
            self.filters = torch.nn.ModuleList(filters)
            self._act = act

        def forward(self, vertex_features, geo_edges, inplace=False):
            residual_geo = self.filters[0](vertex_features, geo_edges)
            vertex_features = self._act(residual_geo)
            for step in range(1, len(self.filters)):
                residual_geo = self.filters[step](vertex_features, geo_edges)
                vertex_features = self._act(vertex_features + residual_geo)     # reference :107 in place, see module doc
            return vertex_features

    def _pooling(self, vertex_features, cluster):
        if self._pooling_method == 'mean':
            return ops.pool_mean(vertex_features, cluster)
        if self._pooling_method == 'max':
            return ops.pool_max(vertex_features, cluster)[0]
        raise ValueError(f"Unkown pooling type {self._pooling_method}")

    def _dense(self, seq, x):
        for layer in seq:
            x = ops.linear(x, layer.weight, layer.bias, None, self.precision) if isinstance(layer, Lin) else layer(x)
        return x

    def forward(self, sample):
        G = self._graph_levels
        cache = GraphCache.for_sample(sample, G - 1)

        def edges(level):
            return cache.edges('edge_index' if level == 0 else f"hierarchy_edge_index_{level}", level)

        def run(block, x, e):                       # reference :129-131, :146-148
            if torch.is_grad_enabled() and x.requires_grad:
                return checkpoint.checkpoint(block, x, e, use_reentrant=True, preserve_rng_state=False)
            return block(x, e)

        levels = [self.left_geo_cnns[0](sample.x, edges(0))]
        for level in range(1, G):                   # encoder
            curr = self._pooling(levels[-1], cache.cluster(level))
            levels.append(run(self.left_geo_cnns[level], curr, edges(level)))
        current = levels[-1]
        for level in range(1, G):                   # decoder
            # [skip || unpooled] (:140-141): the gather kernel writes its rows straight into the right-hand columns
            fused = ops.unpool_concat(levels[-(level + 1)], current, cache.cluster(G - level))
            if level == G - 1:
                fused = self.right_geo_cnns[-level](fused, edges(0))
            else:
                fused = run(self.right_geo_cnns[-level], fused, edges(G - level - 1))
            current = fused
        result = current
        for conv in self.final_convs:
            result = self._dense(conv, result)
        return result
