"""Pin the CPU oracle against golden vectors minted by the reference's own model code (tests/golden/make_golden.py)."""
import copy

import pytest
import torch

from conftest import GOLDEN, load_golden, rel_err
from oracle import stinet_oracle as O

FP32_TOL = 1e-5   # BASELINE.json north_star: fp32 within 1e-5 relative


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_matches_reference_golden(name):
    fix = load_golden(name)
    net = O.OracleSTINet(**{("norm_type" if k == "norm" else k): v for k, v in fix["kwargs"].items()})
    missing = net.load_state_dict(fix["state_dict"], strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    batch = copy.copy(fix["batch"])
    batch.x = fix["batch"].x.clone().requires_grad_(True)
    out, inter = net(batch, return_intermediates=True)
    loss = O.masked_l1_loss(out, batch)
    loss.backward()
    assert rel_err(out, fix["out"]) <= FP32_TOL
    assert rel_err(loss, fix["loss"]) <= FP32_TOL
    assert rel_err(batch.x.grad, fix["grad_x"]) <= FP32_TOL
    for k, p in net.named_parameters():
        assert rel_err(p.grad, fix["grads"][k]) <= FP32_TOL, k
    args = [v for k, v in sorted(inter.items()) if k.startswith("pool_arg_")]
    assert len(args) == len(fix["pool_args"])
    for a, b in zip(args, fix["pool_args"]):
        assert torch.equal(a, b)          # integer work: bit-exact


def test_param_count_matches_thesis():
    """thesis 4.3 p.20: 4.2 M parameters for the shipped 3D config (SURVEY: 4,202,051)."""
    net = O.OracleSTINet(input_nc=10, output_nc=3, filter_type="edgeconvtransinv", ngf=64, norm_type="instance",
                         n_blocks=9, n_levels=2, pooling_type="max")
    assert sum(p.numel() for p in net.parameters()) == 4202051
    net2 = O.OracleSTINet(input_nc=4, output_nc=3, filter_type="edgeconv", ngf=64, norm_type="instance",
                          n_blocks=9, n_levels=2, pooling_type="max")
    assert sum(p.numel() for p in net2.parameters()) == 4201411


def test_scatter_max_first_wins_and_empty():
    src = torch.tensor([[1., 5.], [3., 5.], [3., 2.], [0., 0.]])
    idx = torch.tensor([2, 2, 2, 0])
    out, arg = O.scatter_max(src, idx, 4)
    assert out.tolist() == [[0., 0.], [0., 0.], [3., 5.], [0., 0.]]
    assert arg.tolist() == [[3, 3], [4, 4], [1, 0], [4, 4]]


def test_decisions_record_then_replay_reproduces_the_oracle():
    """oracle.Decisions: replaying the oracle's own recorded choices gives the free-running result bit for bit, and
    replaying them in fp64 stays within fp32 rounding of it (no choice is flagged beyond rounding distance)."""
    import copy
    from stinet_b200 import synthetic
    torch.manual_seed(3)
    orc = O.OracleSTINet(input_nc=10, output_nc=3, filter_type="edgeconvtransinv", ngf=8, norm_type="instance",
                         n_blocks=2, n_levels=2, pooling_type="max")
    batch = synthetic.make_batch("icosphere", 2, 2, seed=5, subdiv=2, mask_radius=2)
    free = orc(batch)
    with O.Decisions.record() as rec:
        again = orc(batch)
    assert torch.equal(free, again)
    kinds = [k for k, _ in rec.recorded]
    assert kinds.count("pool") == 2 and kinds.count("relu") == 1 + 2 + 2 + 2 + 1
    with O.Decisions.replay(rec.recorded) as rep:
        replayed = orc(batch)
    assert torch.equal(free, replayed) and rep.n_relu_diff == 0 and rep.n_pool_diff == 0
    orc64 = copy.deepcopy(orc).double()
    b64 = copy.copy(batch)
    b64.x = batch.x.double()
    with O.Decisions.replay(rec.recorded) as rep64:
        out64 = orc64(b64)
    assert float((out64 - free.double()).abs().max()) < 1e-5
    assert rep64.max_relu_margin <= 2e-5 and rep64.max_pool_margin <= 2e-5
