import torch


def degree(index, num_nodes=None, dtype=None):
    """torch_geometric.utils.degree: out[i] = #occurrences of i in index."""
    n = int(index.max()) + 1 if num_nodes is None else int(num_nodes)
    out = torch.zeros((n,), dtype=dtype, device=index.device)
    one = torch.ones((index.size(0),), dtype=out.dtype, device=out.device)
    return out.scatter_add_(0, index, one)
