"""Helpers that let the mirrored modules accept either the reference's raw tensors (edge_index [2,E] int64,
batch [N] int64) or the prebuilt device structures of stinet_b200.graph (EdgeCSR, Segments)."""
from __future__ import annotations

import torch

from ...graph import EdgeCSR, Segments


def as_edge_csr(edges, n: int) -> EdgeCSR:
    if isinstance(edges, EdgeCSR):
        return edges
    if torch.is_tensor(edges):
        # reference-style call with a COO tensor: build the CSR on the fly (stinet_csr_build, no torch sort)
        return EdgeCSR(edges, n)
    raise TypeError(f"edges must be an EdgeCSR or an edge_index tensor, got {type(edges)}")


def as_segments(batch, n: int, device) -> Segments:
    """batch: None (one instance), a Segments object, or the reference's int64 graph-id vector."""
    if batch is None:
        return Segments(n, None, None, device)
    if isinstance(batch, Segments):
        return batch
    # reference path (fastinstancenorm.py:51,60): needs B = batch.max()+1 and the per-graph counts -> one host sync
    counts = torch.bincount(batch).tolist()
    return Segments(n, counts, batch.to(torch.int32).contiguous(), device)
