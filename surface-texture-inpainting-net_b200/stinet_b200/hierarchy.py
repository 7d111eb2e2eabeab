"""Hierarchy construction by vertex clustering on the device.

Replaces reference preprocessing/graph_level_generation.py:194-244 (`vertex_clustering`: Python loops over bins, points
and neighbour sets; ~30 min per ScanNet scene for all levels) by hand-written integer kernels (csrc/hierarchy.cu on the
stable radix sort of csrc/sort.cu) over tensors that are already in HBM:

    voxel bin of every vertex (numpy floor division)  ->  packed 64-bit cell key  ->  stable radix sort by key with the
    vertex id as payload  ->  distinct keys + inverse                                   = trace map           (bit-exact)
    (trace[v], trace[w]) of every fine edge as key, self loops dropped -> sort -> distinct keys              = coarse edge set
    per-cluster mean of the member coordinates, members in ascending vertex id, sums in the input dtype      = coarse vertices

Every output is bit-identical to the reference function's (tests/test_hierarchy.py: golden vectors minted from it), the
float32 coordinates included.  Three host reads per level size the outputs (bounding box, cluster count, edge count), as
np.unique / torch.unique need them too.  `to_pt_data` / `save_pt` write the reference's on-disk layout (:492-536).

The same tensor program on ATen ops (`_vertex_clustering_program`, round 1) is kept as the host-tensor cross-check of the
tests; the public entry points run on CUDA tensors only.

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
    return _vertex_clustering_kernels(coords, edge_index, voxel_size)


def _sort_u64(keys: torch.Tensor, key_bits: int, with_index: bool):
    """Stable radix sort of 64-bit keys (held in int64 tensors); with_index: also the sorting permutation (int32)."""
    from . import _abi
    from .graph import _ptr, _stream
    n = keys.numel()
    out = torch.empty_like(keys)
    idx = torch.empty(n, dtype=torch.int32, device=keys.device) if with_index else None
    nb = _abi.query("stinet_sort_workspace_bytes", n)
    ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=keys.device)
    _abi.call("stinet_sort_pairs_u64", keys.data_ptr(), None, out.data_ptr(), _ptr(idx), n, int(key_bits), ws.data_ptr(), nb,
              _stream(), cost=(n * 24 * ((key_bits + 7) // 8), 0, ""))
    return out, idx


def _unique_sorted(keys_sorted: torch.Tensor, limit: int):
    """ids (int32 per item) and the number of distinct keys below `limit` (one host read)."""
    from . import _abi
    from .graph import _stream
    n = keys_sorted.numel()
    ids = torch.empty(max(n, 1), dtype=torch.int32, device=keys_sorted.device)
    count = torch.empty(1, dtype=torch.int32, device=keys_sorted.device)
    nb = _abi.query("stinet_unique_workspace_bytes", n)
    ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=keys_sorted.device)
    _abi.call("stinet_unique_sorted_u64", keys_sorted.data_ptr(), n, int(limit), ids.data_ptr(), count.data_ptr(), ws.data_ptr(),
              nb, _stream(), cost=(n * 12, 0, ""))
    return ids, int(count.item())


def _vertex_clustering_kernels(coords: torch.Tensor, edge_index: torch.Tensor, voxel_size: float
                               ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    from . import _abi
    from .graph import _stream
    assert coords.dim() == 2 and coords.size(1) == 3 and coords.dtype in (torch.float32, torch.float64)
    assert edge_index.dim() == 2 and edge_index.size(0) == 2 and edge_index.dtype == torch.int64
    coords, edge_index = coords.contiguous(), edge_index.contiguous()
    n, dev = coords.size(0), coords.device
    if n == 0:
        return (torch.empty((0, 3), dtype=torch.float32, device=dev), torch.empty(0, dtype=torch.int64, device=dev),
                torch.empty((2, 0), dtype=torch.int64, device=dev))
    s = _stream()
    f64 = int(coords.dtype == torch.float64)
    # :207-209  bins -> cell keys -> sort -> distinct cells + inverse
    bins = torch.empty((n, 3), dtype=torch.int64, device=dev)
    mm = torch.empty(6, dtype=torch.int64, device=dev)
    _abi.call("stinet_voxel_bins", coords.data_ptr(), f64, n, float(voxel_size), bins.data_ptr(), mm.data_ptr(), s,
              cost=(n * (3 * coords.element_size() + 24), 0, ""))
    lo_hi = mm.tolist()                                            # host read 1: the bounding box sizes the key
    span = [lo_hi[3 + a] - lo_hi[a] + 1 for a in range(3)]
    cells = span[0] * span[1] * span[2]
    if cells >= 2 ** 62:
        raise StinetError("voxel grid too large for a packed 64-bit key")
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    _abi.call("stinet_voxel_keys", bins.data_ptr(), mm.data_ptr(), n, keys.data_ptr(), s, cost=(n * 32, 0, ""))
    keys_sorted, idx_sorted = _sort_u64(keys, max(1, (cells - 1).bit_length()), True)
    ids, n_coarse = _unique_sorted(keys_sorted, cells)             # host read 2: the number of clusters
    trace = torch.empty(n, dtype=torch.int64, device=dev)
    start = torch.empty(n_coarse + 1, dtype=torch.int32, device=dev)
    _abi.call("stinet_cluster_finish", idx_sorted.data_ptr(), ids.data_ptr(), n, n_coarse, trace.data_ptr(), start.data_ptr(), s,
              cost=(n * 20, 0, ""))
    # :238-242  centres of gravity
    new_coords = torch.empty((n_coarse, 3), dtype=torch.float32, device=dev)
    _abi.call("stinet_cluster_centroids", coords.data_ptr(), f64, idx_sorted.data_ptr(), start.data_ptr(), n_coarse,
              new_coords.data_ptr(), s, cost=(n * (3 * coords.element_size() + 4) + n_coarse * 16, 0, ""))
    # :215-228  coarse edge set
    e = edge_index.size(1)
    if e == 0:
        return new_coords, trace, torch.empty((2, 0), dtype=torch.int64, device=dev)
    ekeys = torch.empty(e, dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    _abi.call("stinet_coarse_edge_keys", edge_index[0].data_ptr(), edge_index[1].data_ptr(), e, trace.data_ptr(), n, n_coarse,
              ekeys.data_ptr(), status.data_ptr(), s, cost=(e * 40, 0, ""))
    limit = n_coarse * n_coarse
    ekeys_sorted, _ = _sort_u64(ekeys, max(1, limit.bit_length()), False)
    eids, n_e = _unique_sorted(ekeys_sorted, limit)                # host read 3: the number of coarse edges
    if int(status.item()) != 0:
        raise StinetError("edge_index refers to vertices outside [0, N)")
    coarse = torch.empty((2, n_e), dtype=torch.int64, device=dev)
    _abi.call("stinet_coarse_edges_emit", ekeys_sorted.data_ptr(), eids.data_ptr(), e, n_coarse, n_e, coarse.data_ptr(), s,
              cost=(e * 12 + n_e * 16, 0, ""))
    return new_coords, trace, coarse


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


def to_pt_data(levels, features0: torch.Tensor = None, labels: torch.Tensor = None, dilated_edges=None, dilation_dists=None):
    """The reference's on-disk layout for one scene (preprocessing/graph_level_generation.py:492-536, read back by
    datasets/scannetcolorgraph_dataloader.py): `vertices` = per-level float32 tensors (level 0 carries all per-vertex
    features -- pos, colour, normals ... -- the coarser levels positions only), `edges` = per-level int64 [E, 2] rows
    (vertex, neighbour), `traces` = per-level int64 maps fine -> coarse, `dilated_edges` / `dilation_dists` (filled by the
    reference's graph_dilation.py, outside this path: None / []), `labels` when given."""
    v0 = features0 if features0 is not None else levels[0]["coords"]
    data = {
        "vertices": [v0.float().cpu()] + [l["coords"].float().cpu() for l in levels[1:]],
        "edges": [l["edge_index"].t().contiguous().long().cpu() for l in levels],
        "traces": [l["trace"].long().cpu() for l in levels[1:]],
        "dilated_edges": dilated_edges if dilated_edges is not None else [None] * len(levels),
        "dilation_dists": dilation_dists if dilation_dists is not None else [],
    }
    if labels is not None:
        data["labels"] = labels.long().cpu()
    return data


def save_pt(path: str, levels, **kw) -> None:
    torch.save(to_pt_data(levels, **kw), path)
