"""CPU oracle for hierarchy construction by vertex clustering -- TEST INFRASTRUCTURE ONLY (see oracle/stinet_oracle.py
for the rules: only tests/, smoke() and bench.py's CPU legs may import this package).

A vectorised numpy restatement of reference preprocessing/graph_level_generation.py:194-244 (`vertex_clustering`),
pinned bit-for-bit (trace, coarse edge set) and to the last float32 bit (coordinates) against golden vectors minted
from that function itself (tests/golden/make_golden_hierarchy.py, tests/test_hierarchy.py).
"""
from __future__ import annotations

import numpy as np


def vertex_clustering(coords: np.ndarray, edges: np.ndarray, voxel_size: float):
    """coords [N,3] (float64 for the input mesh, float32 for every later level -- the arithmetic stays in that dtype,
    as in the reference); edges [E,2] rows (vertex, neighbour).
    -> new_coords float32 [Nc,3], trace int64 [N], coarse edges int64 [Ec,2] sorted by (key, neighbour).

    :207       bins = coords // voxel_size                      (numpy floor division in coords' dtype)
    :208-209   unique_bins, trace = np.unique(bins, axis=0, return_inverse=True)   (lexicographic row order)
    :215-226   coarse adjacency: for every vertex p and neighbour q: trace[p] -> trace[q], self loops discarded,
               duplicates collapsed (sets); emitted grouped by ascending key
    :238-242   new_coords[c] = coords[members of c, ascending ids].mean(axis=0) cast to float32"""
    bins = coords // voxel_size
    _, trace = np.unique(bins, axis=0, return_inverse=True)
    trace = trace.reshape(-1).astype(np.int64)
    n_coarse = int(trace.max()) + 1 if trace.size else 0
    pairs = np.stack([trace[edges[:, 0]], trace[edges[:, 1]]], 1)
    pairs = pairs[pairs[:, 0] != pairs[:, 1]]
    coarse_edges = np.unique(pairs, axis=0) if len(pairs) else pairs.reshape(0, 2)
    new_coords = np.empty((n_coarse, 3), dtype=np.float32)
    order = np.argsort(trace, kind="stable")
    starts = np.searchsorted(trace[order], np.arange(n_coarse + 1))
    for c in range(n_coarse):                                  # same call as the reference: mean over the gathered rows
        new_coords[c] = coords[order[starts[c]:starts[c + 1]]].mean(axis=0)
    return new_coords, trace, coarse_edges.astype(np.int64)
