import torch
from .message_passing import MessagePassing


class EdgeConv(MessagePassing):
    """x_i' = aggr_j nn([x_i || x_j - x_i])  (PyG 2.0.x edge_conv.py; default aggr='max', the reference passes 'mean')."""

    def __init__(self, nn, aggr="max", **kwargs):
        super().__init__(aggr=aggr, **kwargs)
        self.nn = nn

    def forward(self, x, edge_index):
        if isinstance(x, torch.Tensor):
            x = (x, x)
        return self.propagate(edge_index, x=x, size=None)

    def message(self, x_i, x_j):
        return self.nn(torch.cat([x_i, x_j - x_i], dim=-1))
