"""SingleConvMeshNet on the B200 (this file sorts last, so nothing it does can disturb another GPU test): the CUDA path --
literal per-edge EdgeConv with BatchNorm1d over edges on the gather / segmented-sum / tcgen05 GEMM / affine-norm kernels,
skip concatenation written in place by the unpool kernel -- against golden vectors minted from the reference's own
models/singleconvmeshnet.py and against the fp64 oracle.

What round 1 could only record as non-strict expectations has met the hardware (scripts/diag_singleconv.py, every tensor
printed): forward and loss are within 1e-5 on every fixture; against the fp64 oracle REPLAYING the CUDA path's decisions
every gradient is within 1e-5.  The one fixture whose golden gradients differ (`singleconv_ico_max_b1`, 9e-5 .. 5e-4 on the
left-branch tensors) differs because the GOLDEN took the other side of one ReLU tie: BatchNorm centres ~10^5
pre-activations at zero, one of them lies 8.6e-8 from it, and the reference's CPU evaluation and the CUDA path round it to
different signs -- the golden itself is that far from the replaying truth, the CUDA path is at 2e-6.  So the golden
gradient test allows each tensor the golden's own distance from that truth on top of 1e-5 (exactly the slack rule of
test_model_matches_reference_golden), and the replay test is the strict one.  All tests are strict: no xfail."""
import pytest
import torch

from conftest import rel_err
from test_singleconv import (DECISION_MARGIN, FIXTURES, TOL, load, oracle_with_replayed_decisions,
                             record_product_decisions)

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _net(fix, precision="fp32"):
    from stinet_b200.models.singleconvmeshnet import SingleConvMeshNet
    net = SingleConvMeshNet(**fix["kwargs"], precision=precision)
    net.load_state_dict(fix["state_dict"], strict=True)
    return net.to(DEV).train()


@pytest.mark.parametrize("name", FIXTURES)
def test_singleconv_forward_matches_reference_golden(name):
    from stinet_b200 import _abi
    fix = load(name)
    net = _net(fix)
    before = _abi.query("stinet_launch_count")
    out = net(fix["batch"].to(DEV))
    assert rel_err(out, fix["out"]) <= TOL
    assert rel_err(out.square().mean(), fix["loss"]) <= TOL
    assert _abi.query("stinet_launch_count") > before


@pytest.mark.parametrize("name", [f for f in FIXTURES if not load(f).get("forward_only")])
def test_singleconv_gradients_buffers_and_decisions(name):
    """One training step: (1) every gradient within 1e-5 of the fp64 oracle replaying the CUDA path's ReLU signs and
    max-pool winners, every differing choice within 2e-5 of its discontinuity; (2) every gradient within 1e-5 (+ the
    golden's own distance from that truth) of the reference's golden gradients; (3) BatchNorm buffers after the step
    (checkpointed blocks: two momentum updates) equal to the reference's."""
    fix = load(name)
    net = _net(fix)
    b = fix["batch"].to(DEV)
    choices = record_product_decisions(net, b)
    b.x = b.x.detach().clone().requires_grad_(True)
    out = net(b)
    loss = out.square().mean()
    loss.backward()
    t_out, t_loss, t_grads, dec = oracle_with_replayed_decisions(fix, choices)
    assert dec.max_relu_margin <= DECISION_MARGIN and dec.max_pool_margin <= DECISION_MARGIN, \
        (dec.n_relu_diff, dec.max_relu_margin, dec.n_pool_diff, dec.max_pool_margin)
    assert rel_err(out, t_out) <= TOL and rel_err(loss, t_loss) <= TOL
    got = {k: p.grad for k, p in net.named_parameters()}
    got["__x__"] = b.x.grad
    golden = dict(fix["grads"], __x__=fix["grad_x"])
    scale = max(float(v.abs().max()) for v in t_grads.values())
    for k, t in t_grads.items():
        if float(t.abs().max()) < 1e-4 * scale:              # structurally zero (a bias in front of a BatchNorm): noise only
            assert float(got[k].abs().max()) < 1e-4 * scale, k
            continue
        assert rel_err(got[k], t) <= TOL, f"{k}: {rel_err(got[k], t):.2e} vs the replaying fp64 oracle"
        slack = rel_err(golden[k], t)
        assert rel_err(got[k], golden[k]) <= TOL + slack, f"{k}: {rel_err(got[k], golden[k]):.2e} vs golden (golden itself {slack:.2e} from the truth)"
    for k, v in net.named_buffers():
        ref = fix["buffers_after"][k]
        if ref.is_floating_point():
            assert float((v.detach().cpu() - ref).abs().max()) <= TOL * max(float(ref.abs().max()), 1e-3), k
        else:
            assert torch.equal(v.cpu(), ref), k


def test_singleconv_eval_is_deterministic_and_uses_running_stats():
    fix = load(FIXTURES[0])
    net = _net(fix).eval()
    b = fix["batch"].to(DEV)
    with torch.no_grad():
        a, c = net(b), net(b)
    assert torch.equal(a, c) and torch.isfinite(a).all()
