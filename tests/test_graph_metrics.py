"""Per-step graph metrics: the oracle is pinned against golden vectors minted by the reference's
own utils/metrics/graph_metrics.py (tests/golden/make_golden_metrics.py); the CUDA kernels are checked against both."""
import os

import pytest
import torch

from conftest import rel_err
from oracle import stinet_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-5


def _cases():
    return torch.load(os.path.join(HERE, "golden", "metrics", "graph_metrics.pt"), weights_only=False)


@pytest.mark.parametrize("name", ["ico3", "grid12", "graph18"])
def test_oracle_metrics_match_reference_golden(name):
    c = _cases()[name]
    ei, pred, gt, mask = c["edge_index"], c["pred"], c["gt"], c["mask"]
    assert rel_err(O.graph_laplace(pred, ei), c["laplace"]) <= TOL
    assert rel_err(O.graph_laplace_variance(pred, ei), c["lap_var"]) <= TOL
    assert rel_err(O.graph_total_variation(pred, ei), c["tv"]) <= TOL
    assert rel_err(O.psnr(pred, gt, 2.0), c["psnr"]) <= TOL
    assert rel_err(O.psnr(pred, gt, 2.0, mask), c["psnr_mask_only"]) <= TOL
    assert rel_err(O.psnr(pred, gt, 1.0), c["psnr_grey_range1"]) <= TOL


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ico3", "grid12", "graph18"])
def test_cuda_metrics_match_golden_and_oracle(name):
    from stinet_b200.graph import EdgeCSR
    from stinet_b200.utils.metrics import graph_metrics as M
    c = _cases()[name]
    dev = "cuda"
    ei, pred, gt, mask = c["edge_index"].to(dev), c["pred"].to(dev), c["gt"].to(dev), c["mask"].to(dev)
    csr = EdgeCSR(ei, c["n"])
    for edges in (ei, csr):                      # reference-style COO tensor, or the CSR the forward pass holds
        assert rel_err(M.GraphLaplaceOperator()(pred, edges), c["laplace"]) <= TOL
        lv = M.GraphLaplaceVariance()(pred, edges)
        assert lv.shape == c["lap_var"].shape and rel_err(lv, c["lap_var"]) <= TOL
        tv = M.graph_total_variation(pred, edges)
        assert tv.dim() == 0 and rel_err(tv, c["tv"]) <= TOL
    assert rel_err(M.psnr(pred, gt, data_range=2.0), c["psnr"]) <= TOL
    assert rel_err(M.psnr(pred, gt, data_range=2.0, mask=mask), c["psnr_mask_only"]) <= TOL
    assert rel_err(M.psnr(pred, gt, data_range=1.0), c["psnr_grey_range1"]) <= TOL
    # single-channel input of the Laplace operator (what GraphLaplaceVariance feeds it in the reference)
    grey = M.GraphLaplaceVariance().grayscale(pred)
    assert rel_err(M.GraphLaplaceOperator()(grey, csr), O.graph_laplace(c["pred"] @ torch.tensor([[0.299], [0.587], [0.114]]), c["edge_index"])) <= TOL


@pytest.mark.gpu
def test_cuda_metrics_full_size_properties():
    """BASELINE size (8 x 40,962 vertices): agreement with the oracle, run-to-run bit determinism, and the
    size-independent identities  laplace(const) = 0,  tv(const) = 0,  psnr(x, x) = 80 dB (the 1e-8 floor)."""
    from stinet_b200 import synthetic
    from stinet_b200.graph import GraphCache
    from stinet_b200.utils.metrics import graph_metrics as M
    b = synthetic.make_batch("icosphere", 8, 1, seed=7, subdiv=6)
    d = b.to("cuda")
    csr = GraphCache.for_sample(d, 1).edges("edge_index", 0)
    g = torch.Generator().manual_seed(3)
    pred = torch.rand(b.x.shape[0], 3, generator=g) * 2 - 1
    pd = pred.to("cuda")
    lv, tv, ps = M.GraphLaplaceVariance()(pd, csr), M.graph_total_variation(pd, csr), M.psnr(pd, d.color, 2.0, mask=d.mask)
    assert rel_err(lv, O.graph_laplace_variance(pred, b.edge_index)) <= TOL
    assert rel_err(tv, O.graph_total_variation(pred, b.edge_index)) <= TOL
    assert rel_err(ps, O.psnr(pred, b.color, 2.0, b.mask)) <= TOL
    assert torch.equal(lv, M.GraphLaplaceVariance()(pd, csr)) and torch.equal(tv, M.graph_total_variation(pd, csr))
    const = torch.full_like(pd, 0.25)
    assert float(M.GraphLaplaceOperator()(const, csr).abs().max()) == 0.0
    assert float(M.graph_total_variation(const, csr)) == 0.0
    assert abs(float(M.psnr(pd, pd, 2.0)) - 80.0) < 1e-3
