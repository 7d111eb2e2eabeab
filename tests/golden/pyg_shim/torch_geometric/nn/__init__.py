import torch
from .conv import MessagePassing, EdgeConv, SAGEConv  # noqa: F401
from . import conv, inits  # noqa: F401


class BatchNorm(torch.nn.Module):
    """torch_geometric.nn.BatchNorm: BatchNorm1d over node rows."""

    def __init__(self, in_channels, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.module = torch.nn.BatchNorm1d(in_channels, eps, momentum, affine, track_running_stats)

    def forward(self, x):
        return self.module(x)


class InstanceNorm(torch.nn.Module):  # imported by the reference, never instantiated on the STINet path
    def __init__(self, *a, **k):
        raise NotImplementedError("shim: torch_geometric.nn.InstanceNorm is not used by the hot path")


class GraphNorm(torch.nn.Module):  # idem
    def __init__(self, *a, **k):
        raise NotImplementedError("shim: torch_geometric.nn.GraphNorm is not used by the hot path")
