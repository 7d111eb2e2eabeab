"""Hierarchy construction by vertex clustering on the device -- first version.

Replaces reference preprocessing/graph_level_generation.py:194-244 (`vertex_clustering`: Python loops over bins, points
and neighbour sets; ~30 min per ScanNet scene for all levels) by a handful of device-wide sort / unique / segmented
reductions over tensors that are already in HBM:

    voxel bin of every vertex  ->  packed int64 key  ->  sort-unique + inverse  = trace map          (integer-exact)
    (trace[src], trace[dst]) of every fine edge -> drop self loops -> sort-unique                     = coarse edge set
    per-cluster mean of the member coordinates (arithmetic in the input dtype, result float32)       = coarse vertices

Round-1 status: the steps are ATen device ops (floor_divide, unique, index_add_) -- the same code runs on any device,
which is how it is pinned here against golden vectors of the reference function (tests/test_hierarchy.py); the
hand-written hash / sort kernels and the on-disk `.pt` layout (:492-536) are the next step.  Integer outputs (trace,
coarse edges) are bit-exact; coordinates agree to float32 rounding (the reference sums a cluster's members
sequentially in ascending order, index_add_ on a GPU does not fix the order).

Input contract = the reference's arrays: `coords [N,3]` float64 (input mesh) or float32 (a level produced by a
previous call), `edge_index [2,E]` int64 rows (vertex, neighbour) -- directed, symmetric for meshes.  Every vertex is
expected to have at least one edge (true for face-derived edges; the reference mis-numbers its adjacency otherwise).
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from ._abi import StinetError


def vertex_clustering(coords: torch.Tensor, edge_index: torch.Tensor, voxel_size: float
                      ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (new_coords float32 [Nc,3], trace int64 [N], coarse edge_index int64 [2,Ec] sorted by (vertex, neighbour)).
    Like every other entry of the package this runs on the device only; the tensor program underneath is
    device-agnostic, which is what tests/test_hierarchy.py uses to pin it on host tensors."""
    if not (coords.is_cuda and edge_index.is_cuda):
        raise StinetError("stinet_b200.hierarchy works on CUDA tensors (there is no CPU path in this package)")
    return _vertex_clustering_program(coords, edge_index, voxel_size)


def _vertex_clustering_program(coords: torch.Tensor, edge_index: torch.Tensor, voxel_size: float
                               ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    assert coords.dim() == 2 and coords.size(1) == 3 and coords.dtype in (torch.float32, torch.float64)
    assert edge_index.dim() == 2 and edge_index.size(0) == 2 and edge_index.dtype == torch.int64
    n = coords.size(0)
    dev = coords.device
    if n == 0:
        return (torch.empty((0, 3), dtype=torch.float32, device=dev), torch.empty(0, dtype=torch.int64, device=dev),
                torch.empty((2, 0), dtype=torch.int64, device=dev))
    # :207  numpy's `//` on floats is Python floor division; torch.floor_divide implements the same algorithm
    bins = torch.floor_divide(coords, torch.tensor(voxel_size, dtype=coords.dtype, device=dev)).to(torch.int64)
    lo = bins.min(dim=0).values
    span = bins.max(dim=0).values - lo + 1                        # device scalars; the assert below is the one host read
    assert float(span.double().prod()) < 2.0 ** 62, "voxel grid too large for a packed 64-bit key"
    rel = bins - lo
    key = (rel[:, 0] * span[1] + rel[:, 1]) * span[2] + rel[:, 2]   # lexicographic (x, y, z) order = np.unique(axis=0)
    uniq, trace = torch.unique(key, sorted=True, return_inverse=True)                             # :208-209
    n_coarse = uniq.numel()
    # :215-228  coarse edge set
    a, b = trace[edge_index[0]], trace[edge_index[1]]
    keep = a != b
    pair = torch.unique(a[keep] * n_coarse + b[keep], sorted=True)
    coarse = torch.stack([torch.div(pair, n_coarse, rounding_mode="floor"), pair % n_coarse], 0)
    # :238-242  centre of gravity per cell, arithmetic in the input dtype, stored as float32
    sums = torch.zeros((n_coarse, 3), dtype=coords.dtype, device=dev).index_add_(0, trace, coords)
    cnt = torch.bincount(trace, minlength=n_coarse).to(coords.dtype).unsqueeze(1)
    return (sums / cnt).to(torch.float32), trace, coarse


def build_hierarchy(coords: torch.Tensor, edge_index: torch.Tensor, voxel_sizes: List[float], _step=None):
    """Chains the levels as reference process_frame does (:404-420): level l+1 is clustered from level l's float32
    coordinates and coarse edges.  -> list of dicts {coords, edge_index, trace (absent at level 0)}; `trace` maps
    level l-1 vertices to level l (the `hierarchy_trace_index_l` the model consumes)."""
    levels = [{"coords": coords, "edge_index": edge_index}]
    step = _step or vertex_clustering
    for voxel in voxel_sizes:
        c, t, e = step(levels[-1]["coords"], levels[-1]["edge_index"], float(voxel))
        levels.append({"coords": c, "edge_index": e, "trace": t})
    return levels
