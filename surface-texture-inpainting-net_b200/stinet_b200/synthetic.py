"""Seeded synthetic meshes / image-grid graphs with multi-level hierarchies (host side, numpy).

These stand in for the reference's datasets, which need ScanNet / a texture set:
  * grid_sample      -- reference datasets/imagegraph_dataloader.py:46-160 (4-neighbour pixel graph, 2x
                        decimation traces, four radius-r disc masks at the quadrant centres, x = [rgb*~mask, mask])
  * icosphere_sample -- a closed triangle mesh with the ScanNet sample layout of
                        datasets/scannetcolorgraph_dataloader.py:113-151: x = [rgb*known, normal, pos, known] (10 ch),
                        `mask` = hop distance into the hole (0 = observed), symmetric face-derived edges grouped by
                        source (preprocessing/graph_level_generation.py:119-132, 395-400), surjective trace maps
  * plane_sample     -- triangulated height-field "scene" (configs 3 and 5 of BASELINE.json)
All index tensors are int64 and all features fp32, as the reference's loaders deliver them.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .data import GraphBatch, collate

# --------------------------------------------------------------------------------------------
# topology helpers


def edges_from_faces(faces: np.ndarray, rng: Optional[np.random.Generator] = None) -> np.ndarray:
    """Directed, symmetric, duplicate-free edge list [E,2] = (source, neighbour), grouped by source.
    Neighbour order inside a group is arbitrary in the reference (python sets); `rng` shuffles it."""
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], 0)
    e = np.concatenate([e, e[:, ::-1]], 0)
    n = int(e.max()) + 1 if len(e) else 1
    key = np.unique(e[:, 0].astype(np.int64) * n + e[:, 1])      # same lexicographic order as np.unique(e, axis=0), 10x faster
    e = np.stack([key // n, key % n], 1)
    if rng is not None:
        e = e[rng.permutation(len(e))]
        e = e[np.argsort(e[:, 0], kind="stable")]
    return e.astype(np.int64)


def _icosahedron() -> Tuple[np.ndarray, np.ndarray]:
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
                  [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    return v, f


def _subdivide(v: np.ndarray, f: np.ndarray):
    """Loop-style 1->4 split; new vertices are appended, so coarse vertex k is fine vertex k.
    Returns (verts, faces, trace) with trace[fine] = coarse parent (a midpoint joins its lower-indexed endpoint)."""
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0)
    e.sort(axis=1)
    uniq, inv = np.unique(e, axis=0, return_inverse=True)
    mid = v[uniq].mean(1)
    mid /= np.linalg.norm(mid, axis=1, keepdims=True)
    n = len(v)
    m = inv.reshape(3, -1).T + n
    m01, m12, m20 = m[:, 0], m[:, 1], m[:, 2]
    nf = np.concatenate([np.stack([f[:, 0], m01, m20], 1), np.stack([f[:, 1], m12, m01], 1),
                         np.stack([f[:, 2], m20, m12], 1), np.stack([m01, m12, m20], 1)], 0)
    trace = np.concatenate([np.arange(n), uniq[:, 0]]).astype(np.int64)
    return np.concatenate([v, mid], 0), nf, trace


def icosphere_levels(subdiv: int, n_levels: int):
    """levels[l] = (verts, faces) for l = 0 (finest, `subdiv` splits) .. n_levels; traces[l-1]: level l-1 -> l."""
    assert subdiv >= n_levels >= 0
    v, f = _icosahedron()
    meshes, traces = [(v, f)], []
    for _ in range(subdiv):
        v, f, t = _subdivide(v, f)
        meshes.append((v, f))
        traces.append(t)
    meshes = meshes[::-1][: n_levels + 1]
    traces = traces[::-1][:n_levels]
    return meshes, traces


def plane_levels(rows: int, cols: int, n_levels: int, rng: np.random.Generator, noise: float = 0.02):
    meshes, traces = [], []
    r, c = rows, cols
    for lvl in range(n_levels + 1):
        rr, cc = np.meshgrid(np.arange(r), np.arange(c), indexing="ij")
        scale = 2.0 ** lvl
        z = rng.normal(0.0, noise, size=(r, c)) if lvl == 0 else np.zeros((r, c))
        v = np.stack([cc * scale, rr * scale, z], -1).reshape(-1, 3) / max(rows, cols)
        idx = (rr * c + cc)
        a, b, d, e_ = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel()
        f = np.concatenate([np.stack([a, b, e_], 1), np.stack([a, e_, d], 1)], 0).astype(np.int64)
        meshes.append((v, f))
        if lvl < n_levels:
            r2, c2 = (r + 1) // 2, (c + 1) // 2
            traces.append(((rr // 2) * c2 + (cc // 2)).ravel().astype(np.int64))
            r, c = r2, c2
    return meshes, traces


def vertex_normals(v: np.ndarray, f: np.ndarray) -> np.ndarray:
    fn = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    n = np.zeros_like(v)
    for k in range(3):
        np.add.at(n, f[:, k], fn)
    n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-12)
    return n


def hop_distance_mask(edges: np.ndarray, n: int, rng: np.random.Generator, radius: int, cover: float) -> np.ndarray:
    """BFS disc holes (reference preprocessing/observed_texture_map_generation.py:530-603 semantics):
    returns int mask, 0 = observed, k>0 = hop distance of a hole vertex from the nearest observed vertex."""
    import scipy.sparse as sp
    adj = sp.csr_matrix((np.ones(len(edges), dtype=np.int8), (edges[:, 1], edges[:, 0])), shape=(n, n))
    hole = np.zeros(n, dtype=bool)
    guard = 0
    while hole.mean() < cover and guard < 64:
        guard += 1
        front = np.zeros(n, dtype=bool)
        front[rng.integers(n)] = True
        disc = front.copy()
        for _ in range(radius):
            front = (adj @ front.astype(np.int8) > 0) & ~disc
            if not front.any():
                break
            disc |= front
        hole |= disc
    if hole.all():
        hole[rng.integers(n)] = False
    dist = np.zeros(n, dtype=np.int64)
    known = ~hole
    front, k = known.copy(), 0
    while True:
        k += 1
        nxt = (adj @ front.astype(np.int8) > 0) & ~known
        if not nxt.any():
            break
        dist[nxt] = k
        known |= nxt
        front = nxt
    return dist


# --------------------------------------------------------------------------------------------
# samples


def _mesh_sample(meshes, traces, rng: np.random.Generator, mask_radius: int, mask_cover: float,
                 dilations: Sequence[int] = (), name: str = "synthetic") -> GraphBatch:
    v0, f0 = meshes[0]
    n0 = len(v0)
    e0 = edges_from_faces(f0, rng)
    color = rng.uniform(-1.0, 1.0, size=(n0, 3))
    normal = vertex_normals(v0, f0)
    pos = v0 / 1.5                                   # CoordsNormalization(max_sizes=1.5), reference 3D config :56-61
    mask = hop_distance_mask(e0, n0, rng, mask_radius, mask_cover)
    known = (mask == 0)[:, None]
    x = np.concatenate([color * known, normal, pos, known.astype(np.float64)], 1)
    s = GraphBatch(
        x=torch.from_numpy(x).float(), color=torch.from_numpy(color).float(),
        mask=torch.from_numpy(mask).unsqueeze(1), edge_index=torch.from_numpy(e0).t().contiguous(), name=name)
    nv = [n0]
    for lvl in range(1, len(meshes)):
        vl, fl = meshes[lvl]
        el = edges_from_faces(fl, rng)
        s[f"hierarchy_edge_index_{lvl}"] = torch.from_numpy(el).t().contiguous()
        s[f"hierarchy_trace_index_{lvl}"] = torch.from_numpy(traces[lvl - 1])
        nv.append(int(traces[lvl - 1].max()) + 1)
        if lvl == len(meshes) - 1:
            for d in dilations:
                if d > 1:
                    s[f"hierarchy_dil_{d}_edge_index_{lvl}"] = torch.from_numpy(
                        dilated_edges(el, len(vl), d, rng)).t().contiguous()
    s.num_vertices = torch.tensor(nv, dtype=torch.int)
    return s


def dilated_edges(edges: np.ndarray, n: int, dist: int, rng: np.random.Generator) -> np.ndarray:
    """Synthetic stand-in for preprocessing/graph_dilation.py:83-137: a directed, ASYMMETRIC edge set
    [dilated_vertex -> center], coalesced (sorted by source,target), in which some vertices have no in-edge."""
    import scipy.sparse as sp
    adj = sp.csr_matrix((np.ones(len(edges), dtype=np.int64), (edges[:, 1], edges[:, 0])), shape=(n, n))
    out = []
    indptr, indices = adj.indptr, adj.indices
    for center in range(n):
        if rng.random() < 0.1:                   # some centres get no dilated neighbour (reference: walk fell off the mesh)
            continue
        nbrs = indices[indptr[center]:indptr[center + 1]]
        for start in nbrs:
            prev, cur = center, int(start)
            for _ in range(dist - 1):            # keep walking "away" from where we came from
                cand = indices[indptr[cur]:indptr[cur + 1]]
                cand = cand[cand != prev]
                if len(cand) == 0:
                    cur = -1
                    break
                prev, cur = cur, int(cand[rng.integers(len(cand))])
            if cur >= 0 and cur != center and rng.random() < 0.8:
                out.append((cur, center))
    e = np.unique(np.asarray(out, dtype=np.int64).reshape(-1, 2), axis=0)
    return e


def icosphere_sample(subdiv: int, n_levels: int, seed: int = 49, mask_radius: int = 16, mask_cover: float = 0.25,
                     dilations: Sequence[int] = (), _cache={}) -> GraphBatch:
    key = (subdiv, n_levels)
    if key not in _cache:
        _cache[key] = icosphere_levels(subdiv, n_levels)
    meshes, traces = _cache[key]
    rng = np.random.default_rng(seed)
    return _mesh_sample(meshes, traces, rng, mask_radius, mask_cover, dilations, name=f"icosphere{subdiv}_s{seed}")


def plane_sample(rows: int, cols: int, n_levels: int, seed: int = 49, mask_radius: int = 16, mask_cover: float = 0.25,
                 dilations: Sequence[int] = ()) -> GraphBatch:
    rng = np.random.default_rng(seed)
    meshes, traces = plane_levels(rows, cols, n_levels, rng)
    return _mesh_sample(meshes, traces, rng, mask_radius, mask_cover, dilations, name=f"plane{rows}x{cols}_s{seed}")


def grid_edges(size: int, rng: Optional[np.random.Generator]) -> np.ndarray:
    """4-neighbour pixel graph, both directions (reference imagegraph_dataloader.py:69-108).  The reference dumps a
    python `set`, i.e. ARBITRARY edge order; `rng` reproduces that with a seeded shuffle."""
    idx = np.arange(size * size).reshape(size, size)
    h = np.stack([idx[:, :-1].ravel(), idx[:, 1:].ravel()], 1)
    v = np.stack([idx[:-1, :].ravel(), idx[1:, :].ravel()], 1)
    e = np.concatenate([h, v], 0)
    e = np.concatenate([e, e[:, ::-1]], 0).astype(np.int64)
    if rng is not None:
        e = e[rng.permutation(len(e))]
    return e


def grid_sample(size: int = 128, n_levels: int = 2, seed: int = 49, circle_radius: int = 18, _cache={}) -> GraphBatch:
    """n_levels pool levels => n_levels+1 resolutions (reference `end_level` = n_levels+1)."""
    rng = np.random.default_rng(seed)
    key = (size, n_levels)
    if key not in _cache:
        topo_rng = np.random.default_rng(1234)
        edges, traces = [], []
        for lvl in range(n_levels + 1):
            ls = size // (2 ** lvl)
            edges.append(grid_edges(ls, topo_rng))
            if lvl > 0:
                t = np.arange(ls * ls).reshape(ls, ls).repeat(2, axis=1).repeat(2, axis=0).reshape(-1)
                traces.append(t.astype(np.int64))
        _cache[key] = (edges, traces)
    edges, traces = _cache[key]
    r = min(circle_radius, max(size // 8, 1))
    rr, cc = np.meshgrid(np.arange(2 * r), np.arange(2 * r), indexing="ij")
    circle = (np.abs(rr - r) ** 2 + np.abs(cc - r) ** 2) <= r * r
    mask = np.zeros((size, size), dtype=bool)
    for i in range(4):
        xo = ((i % 2) * 2 - 1) * size // 4
        yo = ((i // 2) * 2 - 1) * size // 4
        mask[size // 2 - r + xo: size // 2 + r + xo, size // 2 - r + yo: size // 2 + r + yo] |= circle
    img = rng.uniform(-1.0, 1.0, size=(size * size, 3))
    m = mask.reshape(-1, 1)
    x = np.concatenate([img * ~m, m.astype(np.float64)], 1)
    s = GraphBatch(x=torch.from_numpy(x).float(), color=torch.from_numpy(img).float(),
                   mask=torch.from_numpy(m), edge_index=torch.from_numpy(edges[0]).t().contiguous(),
                   name=f"grid{size}_s{seed}")
    nv = [size * size]
    for lvl in range(1, n_levels + 1):
        s[f"hierarchy_edge_index_{lvl}"] = torch.from_numpy(edges[lvl]).t().contiguous()
        s[f"hierarchy_trace_index_{lvl}"] = torch.from_numpy(traces[lvl - 1])
        nv.append(int(traces[lvl - 1].max()) + 1)
    s.num_vertices = torch.tensor(nv, dtype=torch.int)
    return s


def make_samples(kind: str, batch_size: int, n_levels: int, seed: int = 49, **kw) -> List[GraphBatch]:
    """The single-graph samples `make_batch` collates.  kind: 'grid' (size=), 'icosphere' (subdiv=), 'plane' (rows=,
    cols=).  Sample b uses seed+b."""
    fn = {"grid": grid_sample, "icosphere": icosphere_sample, "plane": plane_sample}[kind]
    return [fn(n_levels=n_levels, seed=seed + b, **kw) for b in range(batch_size)]


def make_batch(kind: str, batch_size: int, n_levels: int, seed: int = 49, **kw) -> GraphBatch:
    return collate(make_samples(kind, batch_size, n_levels, seed, **kw))


def paper_graph18() -> Tuple[torch.Tensor, int]:
    """The 18-node planar fixture of reference preprocessing/graph_dilation.py:6-24 is print-only (no expected
    values); we keep a tiny irregular-degree graph of the same flavour: returns (edge_index [2,E] int64, N)."""
    und = [(0, 1), (0, 2), (1, 2), (1, 3), (2, 3), (2, 4), (3, 4), (3, 5), (4, 5), (4, 6), (5, 6), (5, 7), (6, 7),
           (6, 8), (7, 8), (7, 9), (8, 9), (9, 10), (10, 11), (10, 12), (11, 12), (12, 13), (13, 14), (13, 15),
           (14, 15), (15, 16)]                       # vertex 17 stays isolated on purpose
    e = np.asarray(und + [(b, a) for a, b in und], dtype=np.int64)
    return torch.from_numpy(e).t().contiguous(), 18
