"""Hierarchy construction by vertex clustering: the numpy oracle and the device-agnostic product
function against golden vectors minted from the reference's own preprocessing/graph_level_generation.py
(tests/golden/make_golden_hierarchy.py).  Integers (trace maps, coarse edge sets) bit-exact; coordinates to float32
rounding."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR

FIXTURES = sorted(glob.glob(os.path.join(GOLDEN_DIR, "hierarchy", "*.pt")))


def test_fixtures_present():
    assert len(FIXTURES) >= 2


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-3] for p in FIXTURES])
def test_oracle_matches_reference_golden(path):
    from oracle import hierarchy_oracle as H
    fix = torch.load(path, weights_only=False)
    levels = fix["levels"]
    coords, edges = levels[0]["coords"].numpy(), levels[0]["edges"].numpy()
    for lvl in levels[1:]:
        c, t, e = H.vertex_clustering(coords, edges, lvl["voxel"])
        assert np.array_equal(t, lvl["trace"].numpy())
        assert np.array_equal(e, lvl["edges"].numpy())
        assert c.dtype == np.float32 and np.array_equal(c, lvl["coords"].numpy())      # to the last bit
        coords, edges = c, e


def _check_product(path, device):
    from stinet_b200 import hierarchy
    fix = torch.load(path, weights_only=False)
    ref = fix["levels"]
    # the public entry points are device-only; on host tensors the test drives the tensor program underneath directly
    step = None if device == "cuda" else hierarchy._vertex_clustering_program
    got = hierarchy.build_hierarchy(ref[0]["coords"].to(device), ref[0]["edges"].t().contiguous().to(device),
                                    [l["voxel"] for l in ref[1:]], _step=step)
    assert len(got) == len(ref)
    for g, r in zip(got[1:], ref[1:]):
        assert torch.equal(g["trace"].cpu(), r["trace"])
        assert torch.equal(g["edge_index"].t().cpu(), r["edges"])
        assert g["coords"].dtype == torch.float32
        if device == "cuda":     # the kernels sum a cluster's members in ascending vertex id, as the reference does: last bit
            assert torch.equal(g["coords"].cpu(), r["coords"])
        else:                    # the ATen cross-check program (index_add_) only fixes the value, not the order of the sums
            err = (g["coords"].cpu() - r["coords"]).abs().max() / r["coords"].abs().max()
            assert float(err) <= 1e-6
        # what the model consumes: a surjective trace map onto [0, N_l)
        assert int(g["trace"].max()) + 1 == g["coords"].shape[0] and torch.unique(g["trace"]).numel() == g["coords"].shape[0]


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-3] for p in FIXTURES])
def test_product_function_matches_reference_golden_on_host_tensors(path):
    """The product function is plain device-agnostic tensor code; evaluated on host tensors it must reproduce the
    reference bit for bit in its integer outputs."""
    _check_product(path, "cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-3] for p in FIXTURES])
def test_product_function_matches_reference_golden_on_device(path):
    _check_product(path, "cuda")


def test_public_entry_refuses_host_tensors():
    from stinet_b200 import _abi, hierarchy
    with pytest.raises(_abi.StinetError):
        hierarchy.vertex_clustering(torch.zeros(4, 3), torch.zeros((2, 0), dtype=torch.int64), 0.5)


def test_floor_division_semantics_match_numpy():
    """numpy `coords // voxel` (the reference's binning) vs torch.floor_divide in both dtypes, including negative
    coordinates and values sitting exactly on cell borders."""
    g = np.random.default_rng(0)
    for dt in (np.float64, np.float32):
        x = np.concatenate([g.normal(0, 3, 5000), np.arange(-40, 40) * 0.25, [-0.0, 0.0, 1e-9, -1e-9]]).astype(dt)
        for voxel in (0.25, 0.1, 0.04, 0.3):
            want = x // dt(voxel)
            got = torch.floor_divide(torch.from_numpy(x), torch.tensor(voxel, dtype=torch.from_numpy(x).dtype)).numpy()
            assert np.array_equal(want, got), (dt, voxel)


def test_pt_layout_matches_the_reference_file_format(tmp_path):
    """to_pt_data / save_pt: the dict the reference writes per scene (preprocessing/graph_level_generation.py:492-536)."""
    from stinet_b200 import hierarchy
    fix = torch.load(FIXTURES[0], weights_only=False)
    ref = fix["levels"]
    levels = [{"coords": ref[0]["coords"], "edge_index": ref[0]["edges"].t().contiguous()}]
    for l in ref[1:]:
        levels.append({"coords": l["coords"], "edge_index": l["edges"].t().contiguous(), "trace": l["trace"]})
    feats = torch.cat([ref[0]["coords"].float(), torch.zeros(ref[0]["coords"].shape[0], 6)], 1)
    path = str(tmp_path / "scene.pt")
    hierarchy.save_pt(path, levels, features0=feats, labels=torch.zeros(feats.shape[0]))
    data = torch.load(path, weights_only=False)
    assert sorted(data) == ["dilated_edges", "dilation_dists", "edges", "labels", "traces", "vertices"]
    assert len(data["vertices"]) == len(ref) and len(data["edges"]) == len(ref) and len(data["traces"]) == len(ref) - 1
    assert data["vertices"][0].shape[1] == 9 and all(v.shape[1] == 3 and v.dtype == torch.float32 for v in data["vertices"][1:])
    for e, l in zip(data["edges"], ref):
        assert e.dtype == torch.int64 and torch.equal(e, l["edges"])
    for t, l in zip(data["traces"], ref[1:]):
        assert torch.equal(t, l["trace"])


@pytest.mark.gpu
@pytest.mark.parametrize("n,bits", [(1, 8), (1000, 5), (5000, 17), (100000, 40), (300001, 63)])
def test_radix_sort_is_stable_and_matches_torch(n, bits):
    from stinet_b200 import hierarchy
    g = torch.Generator().manual_seed(n)
    hi = (1 << bits) - 1
    keys = torch.randint(0, min(hi, 2 ** 62) + 1, (n,), generator=g, dtype=torch.int64)
    keys[::3] = keys[0]                                           # many duplicates: stability is observable
    want_k, want_i = torch.sort(keys, stable=True)
    got_k, got_i = hierarchy._sort_u64(keys.cuda(), bits, True)
    assert torch.equal(got_k.cpu(), want_k) and torch.equal(got_i.cpu().long(), want_i)
    only_k, none = hierarchy._sort_u64(keys.cuda(), bits, False)
    assert none is None and torch.equal(only_k.cpu(), want_k)


@pytest.mark.gpu
def test_degree_sorted_row_order():
    """north_star's degree sort: rows by descending degree, ties in ascending row order (stable), isolated rows last."""
    from stinet_b200 import synthetic
    from stinet_b200.graph import EdgeCSR
    s = synthetic.icosphere_sample(3, 2, seed=12, mask_radius=3, dilations=(2,))
    for ei, n in ((s.edge_index, s.num_nodes), (s["hierarchy_dil_2_edge_index_2"], int(s.num_vertices[2]))):
        csr = EdgeCSR(ei.cuda(), n)
        order = csr.degree_order().cpu().long()
        deg = torch.bincount(ei[1], minlength=n)
        want = torch.sort(-deg, stable=True)[1]
        assert torch.equal(order, want)
