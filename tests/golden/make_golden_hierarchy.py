"""Mint golden vectors for hierarchy construction by vertex clustering from the reference's OWN
preprocessing code (authoring container only).

    python tests/golden/make_golden_hierarchy.py        # rewrites tests/golden/hierarchy/*.pt

/root/reference/preprocessing/graph_level_generation.py (edges_from_faces :119-132, vertex_clustering :194-244) is
imported unmodified (open3d / plyfile are import-only stand-ins in tests/golden/pyg_shim).  Each fixture chains the
levels exactly as process_frame does (:404-420): level l+1 = vertex_clustering(coords_l, adjacency_l, voxel_l), where
coords_0 is float64 (open3d) and every later level is the float32 array the previous call returned.
Stored per level: coords, trace (inverse of np.unique over the voxel bins), and the directed coarse edge set.  The
reference emits a vertex's neighbours in Python-set order; the fixture stores the rows sorted (key, neighbour).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "surface-texture-inpainting-net_b200"))
sys.path.insert(0, os.path.join(HERE, "pyg_shim"))
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402
import torch  # noqa: E402
from preprocessing import graph_level_generation as ref  # noqa: E402  (REAL reference code)
from stinet_b200 import synthetic  # noqa: E402


def plane(rows, cols, seed):
    rng = np.random.default_rng(seed)
    ys, xs = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    v = np.stack([xs * 0.05 - 1.0, ys * 0.05 - 0.7, rng.normal(0, 0.02, (rows, cols))], -1).reshape(-1, 3)
    idx = np.arange(rows * cols).reshape(rows, cols)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel()
    f = np.concatenate([np.stack([a, b, c], 1), np.stack([b, d, c], 1)], 0)
    return v.astype(np.float64), f.astype(np.int64)


def icosphere(subdiv):
    meshes, _ = synthetic.icosphere_levels(subdiv, 0)
    v, f = meshes[0]
    return (np.asarray(v, dtype=np.float64) * 0.9 + 0.013), np.asarray(f, dtype=np.int64)


CASES = {
    "vc_plane_31x43": (lambda: plane(31, 43, 5), [0.11, 0.23, 0.5]),
    "vc_icosphere3": (lambda: icosphere(3), [0.2, 0.45]),
}


def main():
    out_dir = os.path.join(HERE, "hierarchy")
    os.makedirs(out_dir, exist_ok=True)
    for name, (gen, voxels) in CASES.items():
        coords, faces = gen()
        adjacency = ref.edges_from_faces(faces)
        e0 = np.array([[k, n] for k, group in enumerate(adjacency) for n in group], dtype=np.int64)
        levels = [{"coords": torch.from_numpy(coords.copy()), "edges": torch.from_numpy(np.unique(e0, axis=0))}]
        cur_coords, cur_adj = coords, adjacency
        for voxel in voxels:
            new_coords, trace, new_adj, edge_out = ref.vertex_clustering(cur_coords, cur_adj, float(voxel))
            e = np.array(edge_out, dtype=np.int64).reshape(-1, 2)
            assert (np.diff(e[:, 0]) >= 0).all()                      # grouped by key, keys ascending (:222-228)
            levels.append({"voxel": float(voxel), "coords": torch.from_numpy(new_coords.copy()),
                           "trace": torch.from_numpy(trace.astype(np.int64)),
                           "edges": torch.from_numpy(np.unique(e, axis=0))})
            cur_coords, cur_adj = new_coords, new_adj
        path = os.path.join(out_dir, f"{name}.pt")
        torch.save({"faces": torch.from_numpy(faces), "levels": levels}, path)
        print(name, [tuple(l["coords"].shape) for l in levels], [int(l["edges"].shape[0]) for l in levels],
              f"-> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
