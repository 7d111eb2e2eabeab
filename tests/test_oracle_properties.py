"""Property tests (hypothesis) of the CPU oracle's primitives -- the semantics SURVEY 8c fixes for the third-party
calls the reference makes (PyG propagate / torch_scatter), which no golden vector can cover exhaustively:
stable grouping, permutation invariance of the reductions, first-maximum-wins with the `src.size(0)` sentinel for empty
segments, isolated vertices -> exactly 0, batch=None == one graph, and the structural invariants of vertex clustering."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import hierarchy_oracle as H
from oracle import stinet_oracle as O

SET = settings(max_examples=40, deadline=None, derandomize=True)


@st.composite
def graphs(draw, max_n=12, max_e=40):
    n = draw(st.integers(1, max_n))
    e = draw(st.integers(0, max_e))
    src = draw(st.lists(st.integers(0, n - 1), min_size=e, max_size=e))
    dst = draw(st.lists(st.integers(0, n - 1), min_size=e, max_size=e))
    seed = draw(st.integers(0, 2 ** 16))
    return n, torch.tensor([src, dst], dtype=torch.int64).reshape(2, e), seed


@SET
@given(graphs())
def test_csr_grouping_is_stable_and_complete(g):
    n, ei, _ = g
    rowptr, perm = O.csr_by_key(ei[1], n)
    e = ei.shape[1]
    assert rowptr.shape[0] == n + 1 and int(rowptr[0]) == 0 and int(rowptr[-1]) == e
    assert bool((rowptr[1:] >= rowptr[:-1]).all())
    assert sorted(perm.tolist()) == list(range(e))                               # a permutation of the positions
    for i in range(n):
        seg = perm[int(rowptr[i]):int(rowptr[i + 1])].tolist()
        assert all(int(ei[1, k]) == i for k in seg) and seg == sorted(seg)       # grouped, original order kept


@SET
@given(graphs(), st.sampled_from(["mean", "add", "max"]))
def test_aggregation_is_invariant_under_edge_permutation(g, aggr):
    n, ei, seed = g
    gen = torch.Generator().manual_seed(seed)
    e = ei.shape[1]
    msg = torch.randn(e, 3, generator=gen, dtype=torch.float64)
    p = torch.randperm(e, generator=gen)
    a = O.aggregate(msg, ei[1], n, aggr)
    b = O.aggregate(msg[p], ei[1][p], n, aggr)
    assert torch.allclose(a, b, rtol=0, atol=1e-12)
    deg = torch.bincount(ei[1], minlength=n)
    assert bool((a[deg == 0] == 0).all())                                        # rows without in-edges are exactly 0


@SET
@given(graphs())
def test_scatter_max_first_occurrence_wins_and_empty_segments_get_the_sentinel(g):
    n, ei, seed = g
    gen = torch.Generator().manual_seed(seed)
    e = ei.shape[1]
    src = torch.randint(-2, 3, (e, 2), generator=gen).double()                   # few distinct values: ties everywhere
    out, arg = O.scatter_max(src, ei[1], n)
    for i in range(n):
        rows = [k for k in range(e) if int(ei[1, k]) == i]
        for c in range(2):
            if not rows:
                assert float(out[i, c]) == 0.0 and int(arg[i, c]) == e
            else:
                best = max(float(src[k, c]) for k in rows)
                first = min(k for k in rows if float(src[k, c]) == best)
                assert float(out[i, c]) == best and int(arg[i, c]) == first


@SET
@given(graphs(), st.booleans())
def test_edge_conv_gives_exact_zero_on_isolated_vertices(g, trans_inv):
    n, ei, seed = g
    torch.manual_seed(seed)
    din, dout = 3, 2
    mlp = torch.nn.Sequential(torch.nn.Linear(din if trans_inv else 2 * din, 2 * dout), torch.nn.ReLU(),
                              torch.nn.Linear(2 * dout, dout)).double()
    x = torch.randn(n, din, dtype=torch.float64)
    out = O.edge_conv(x, ei, mlp, "mean", trans_inv)
    deg = torch.bincount(ei[1], minlength=n)
    assert out.shape == (n, dout) and bool((out[deg == 0] == 0).all())           # not b2: PyG's mean of nothing is 0


@SET
@given(st.integers(2, 40), st.integers(1, 5), st.integers(0, 2 ** 16))
def test_instance_norm_with_one_graph_equals_batch_none(n, c, seed):
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(n, c, generator=gen, dtype=torch.float64) * 2 + 0.5
    a = O.fast_instance_norm(x, None)
    b = O.fast_instance_norm(x, torch.zeros(n, dtype=torch.long))
    assert torch.allclose(a, b, rtol=0, atol=1e-9)
    assert torch.allclose(a.mean(0), torch.zeros(c, dtype=torch.float64), atol=1e-9)


@SET
@given(st.integers(1, 4), st.integers(1, 4), st.integers(0, 2 ** 16), st.sampled_from([0.3, 0.55, 1.1]))
def test_vertex_clustering_invariants(rows, cols, seed, voxel):
    rng = np.random.default_rng(seed)
    r, c = rows + 2, cols + 2
    ys, xs = np.meshgrid(np.arange(r), np.arange(c), indexing="ij")
    coords = np.stack([xs * 0.4, ys * 0.4, rng.normal(0, 0.05, (r, c))], -1).reshape(-1, 3)
    idx = np.arange(r * c).reshape(r, c)
    und = np.concatenate([np.stack([idx[:, :-1].ravel(), idx[:, 1:].ravel()], 1),
                          np.stack([idx[:-1, :].ravel(), idx[1:, :].ravel()], 1)], 0)
    edges = np.concatenate([und, und[:, ::-1]], 0)
    new_coords, trace, coarse = H.vertex_clustering(coords, edges, voxel)
    nc = new_coords.shape[0]
    assert sorted(set(trace.tolist())) == list(range(nc))                        # surjective onto [0, Nc)
    assert (coarse[:, 0] != coarse[:, 1]).all()                                  # no self loops
    pairs = set(map(tuple, coarse.tolist()))
    assert all((b, a) in pairs for a, b in pairs)                                # symmetric input -> symmetric output
    assert len(pairs) == len(coarse)                                             # no duplicates
    bins = coords // voxel
    for k in range(nc):                                                          # one voxel per cluster, centroid inside
        members = np.where(trace == k)[0]
        assert (bins[members] == bins[members[0]]).all()
        assert np.allclose(new_coords[k], coords[members].mean(0), atol=1e-6)
