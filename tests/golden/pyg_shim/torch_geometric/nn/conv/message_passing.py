"""MessagePassing.propagate, flow='source_to_target' (PyG 2.0.x): j = edge_index[0] (source),
i = edge_index[1] (target); message() is evaluated per edge and reduced over the TARGET index
with torch_scatter.scatter(..., dim=0, dim_size=N, reduce=aggr)."""
import inspect
import torch
from torch_scatter import scatter


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2, **kwargs):
        super().__init__()
        assert flow == "source_to_target"
        self.aggr = aggr
        self.flow = flow
        self.node_dim = node_dim

    def propagate(self, edge_index, size=None, **kwargs):
        x = kwargs["x"]
        if isinstance(x, torch.Tensor):
            x = (x, x)
        x_src, x_dst = x[0], (x[1] if x[1] is not None else x[0])
        n_dst = x_dst.size(0) if size is None or size[1] is None else size[1]
        want = inspect.signature(self.message).parameters
        args = {}
        if "x_j" in want:
            args["x_j"] = x_src.index_select(0, edge_index[0])
        if "x_i" in want:
            args["x_i"] = x_dst.index_select(0, edge_index[1])
        msg = self.message(**args)
        return scatter(msg, edge_index[1], dim=0, dim_size=n_dst, reduce=self.aggr)

    def message(self, x_j):
        return x_j
