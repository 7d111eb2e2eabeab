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
