"""SingleConvMeshNet on the CPU: the oracle restatement against golden vectors minted from the
reference's own models/singleconvmeshnet.py (tests/golden/make_golden_singleconv.py), and the HOST LOGIC of the
product module -- schedule, checkpointing / BatchNorm double update, edge-as-cluster index plumbing, state_dict
layout -- against the same vectors with the CUDA entry points replaced by plain-torch stand-ins.  The stand-ins exist
only inside this test (the product has no CPU path); the kernels themselves are checked on the B200
(tests/test_gpu_kernels.py per kernel, tests/test_zz_gpu_singleconv.py for this network)."""
import glob
import os

import pytest
import torch

from conftest import GOLDEN_DIR, assert_grads_close, rel_err
from oracle import stinet_oracle as O

FIXTURES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "singleconv", "*.pt")))
TOL = 1e-5


def load(name):
    from stinet_b200.data import GraphBatch
    fix = torch.load(os.path.join(GOLDEN_DIR, "singleconv", f"{name}.pt"), weights_only=False)
    fix["batch"] = GraphBatch(**fix["sample"])
    return fix


def check_against_golden(net, batch, fix):
    batch.x = batch.x.clone().requires_grad_(True)
    out = net(batch)
    loss = out.square().mean()
    assert rel_err(out, fix["out"]) <= TOL and rel_err(loss, fix["loss"]) <= TOL
    if not fix.get("forward_only"):              # num_propagation_steps > 1: the reference's own backward raises (:107)
        loss.backward()
        got = {k: p.grad for k, p in net.named_parameters()}
        got["__x__"] = batch.x.grad
        assert_grads_close(got, dict(fix["grads"], __x__=fix["grad_x"]), TOL)
    # BatchNorm buffers after one training step (checkpointed blocks: two momentum updates).  The running mean of the
    # translation-invariant first layer is structurally zero (sum over a symmetric edge set of W(x_j - x_i)): noise only
    for k, v in net.named_buffers():
        ref = fix["buffers_after"][k]
        if ref.is_floating_point():
            assert float((v.detach().cpu() - ref).abs().max()) <= TOL * max(float(ref.abs().max()), 1e-3), k
        else:
            assert torch.equal(v.cpu(), ref), k


def test_fixtures_present():
    assert len(FIXTURES) >= 3


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_matches_reference_golden(name):
    fix = load(name)
    net = O.OracleSingleConvMeshNet(**fix["kwargs"])
    assert list(net.state_dict()) == list(fix["state_dict"])
    net.load_state_dict(fix["state_dict"])
    check_against_golden(net.train(), fix["batch"], fix)


@pytest.fixture
def torch_stand_ins(monkeypatch):
    """Plain-torch stand-ins for the C-ABI calls this network makes (test harness only)."""
    from stinet_b200 import graph, ops

    def build_csr(key, other, n_rows, want_key32=False, status=None):
        rowptr, perm = O.csr_by_key(key, n_rows)
        col = other[perm.long()].to(torch.int32) if other is not None else None
        return rowptr, perm, col, (key.to(torch.int32) if want_key32 else None)

    def unpool(xc, cl):
        return xc.index_select(0, cl.trace32.long())

    def unpool_concat(skip, xc, cl):
        return torch.cat((skip, xc.index_select(0, cl.trace32.long())), -1)

    def pool_mean(x, cl):
        return O.scatter_mean(x, cl.trace32.long(), cl.n_coarse)

    def pool_max(x, cl):
        return O.scatter_max(x, cl.trace32.long(), cl.n_coarse)

    def linear(x, w, b=None, rowmask=None, precision="fp32"):
        assert rowmask is None
        return torch.nn.functional.linear(x, w, b)

    monkeypatch.setattr(graph, "build_csr", build_csr)
    monkeypatch.setattr(graph, "_require_cuda", lambda *a, **k: None)
    monkeypatch.setattr(ops, "unpool", unpool)
    monkeypatch.setattr(ops, "unpool_concat", unpool_concat)
    monkeypatch.setattr(ops, "pool_mean", pool_mean)
    monkeypatch.setattr(ops, "pool_max", pool_max)
    monkeypatch.setattr(ops, "linear", linear)


@pytest.mark.parametrize("name", FIXTURES)
def test_host_logic_with_stand_in_kernels_matches_reference_golden(name, torch_stand_ins):
    from stinet_b200.models.singleconvmeshnet import SingleConvMeshNet
    fix = load(name)
    net = SingleConvMeshNet(**fix["kwargs"])
    assert list(net.state_dict()) == list(fix["state_dict"])
    net.load_state_dict(fix["state_dict"], strict=True)
    check_against_golden(net.train(), fix["batch"], fix)


@pytest.mark.parametrize("name", FIXTURES)
def test_seeded_construction_draws_the_reference_weights(name):
    """Same module creation order as the reference => torch.manual_seed(49) + construction reproduces the golden
    state_dict (minted from the reference's own constructor under the same seed) bit for bit."""
    from stinet_b200.models.singleconvmeshnet import SingleConvMeshNet
    fix = load(name)
    torch.manual_seed(49)
    sd = SingleConvMeshNet(**fix["kwargs"]).state_dict()
    assert list(sd) == list(fix["state_dict"])
    for k, v in sd.items():
        assert torch.equal(v, fix["state_dict"][k]), k


def test_edges_as_clusters_index_plumbing(torch_stand_ins):
    """x_i / x_j gathers and the mean over in-edges expressed through the pooling structures: members of a target's
    cluster are its in-edges in original order, trace32 is the end point of every original edge."""
    from stinet_b200.graph import EdgeCSR
    ei = torch.tensor([[0, 2, 1, 2, 3, 0], [1, 1, 0, 3, 3, 3]])
    csr = EdgeCSR(ei, 5)
    by_t, by_s = csr.edge_clusters()
    assert (by_t.n_fine, by_t.n_coarse) == (6, 5)
    assert by_t.trace32.tolist() == ei[1].tolist() and by_s.trace32.tolist() == ei[0].tolist()
    assert by_t.rowptr.tolist() == [0, 1, 3, 3, 6, 6] and by_t.member.tolist() == [2, 0, 1, 3, 4, 5]
    assert by_s.rowptr.tolist() == [0, 2, 3, 5, 6, 6] and by_s.member.tolist() == [0, 5, 2, 1, 3, 4]


# ------------------------------------------------------------------------------------------------------------------
# decision replay (oracle.Decisions) for SingleConvMeshNet: BatchNorm centres every pre-activation at zero, so on a real
# mesh some ReLU sign (or max-pool winner) always sits within rounding distance of its discontinuity and two correct
# fp32 evaluations differ on isolated gradient entries.  The fp64 oracle replays the choices of the implementation
# under test and is then a smooth function of the same inputs: every gradient must agree within 1e-5, and every choice
# that differs from the oracle's own must lie within DECISION_MARGIN of the discontinuity.

DECISION_MARGIN = 2e-5


def record_product_decisions(net, batch):
    """One forward of the product module under no_grad (so no checkpoint recomputation: every block runs exactly once, in
    schedule order) that records ("relu", bool mask) for every ReLU -- the nn.ReLU modules of the message MLPs and of the
    head, the functional ReLU of the ResBlocks -- and ("pool", winners) for every max-pool.  BatchNorm buffers are put
    back afterwards.  Deterministic kernels => the choices of the real training pass are the same."""
    from stinet_b200 import ops
    from stinet_b200.models.singleconvmeshnet import SingleConvMeshNet
    choices, handles, patched = [], [], []
    buffers = {k: v.detach().clone() for k, v in net.named_buffers()}

    def relu_and_record(x, *a, **k):
        choices.append(("relu", (x.detach() > 0).cpu()))
        return torch.nn.functional.relu(x)

    for m in net.modules():
        if isinstance(m, torch.nn.ReLU):
            handles.append(m.register_forward_pre_hook(lambda mod, inp: choices.append(("relu", (inp[0].detach() > 0).cpu()))))
        if isinstance(m, SingleConvMeshNet.ResBlock):
            patched.append((m, m._act))
            m._act = relu_and_record
    real_pool_max = ops.pool_max

    def pool_max(x, cl):
        out, arg = real_pool_max(x, cl)
        choices.append(("pool", arg.detach().long().cpu()))
        return out, arg

    ops.pool_max = pool_max
    try:
        with torch.no_grad():
            net(batch)
    finally:
        ops.pool_max = real_pool_max
        for h in handles:
            h.remove()
        for m, act in patched:
            m._act = act
        with torch.no_grad():
            for k, v in net.named_buffers():
                v.copy_(buffers[k])
    return choices


def oracle_with_replayed_decisions(fix, choices, dtype=torch.float64):
    import copy
    orc = O.OracleSingleConvMeshNet(**fix["kwargs"])
    orc.load_state_dict(fix["state_dict"])
    orc = orc.to(dtype).train()
    ob = copy.copy(fix["batch"])
    ob.x = fix["batch"].x.detach().cpu().to(dtype).clone().requires_grad_(True)
    with O.Decisions.replay(choices) as dec:
        out = orc(ob, double_update_checkpointed=False)      # a recomputation would consume the stream twice
        loss = out.square().mean()
        loss.backward()
    assert dec.pos == len(choices), "decision stream not fully consumed"
    grads = {k: p.grad for k, p in orc.named_parameters()}
    grads["__x__"] = ob.x.grad
    return out.detach(), loss.detach(), grads, dec


def check_against_replaying_oracle(net, batch, fix):
    choices = record_product_decisions(net, batch)
    batch.x = batch.x.detach().clone().requires_grad_(True)
    out = net(batch)
    loss = out.square().mean()
    loss.backward()
    t_out, t_loss, t_grads, dec = oracle_with_replayed_decisions(fix, choices)
    assert dec.max_relu_margin <= DECISION_MARGIN and dec.max_pool_margin <= DECISION_MARGIN, \
        (dec.n_relu_diff, dec.max_relu_margin, dec.n_pool_diff, dec.max_pool_margin)
    assert rel_err(out, t_out) <= TOL and rel_err(loss, t_loss) <= TOL
    got = {k: p.grad for k, p in net.named_parameters()}
    got["__x__"] = batch.x.grad
    assert_grads_close(got, t_grads, TOL)
    return dec


@pytest.mark.parametrize("name", FIXTURES)
def test_decision_replay_protocol_with_stand_in_kernels(name, torch_stand_ins):
    """The protocol itself, on the CPU: fp32 product module (stand-in kernels) against the fp64 oracle replaying its
    choices -- the stream lines up call for call and every gradient agrees within 1e-5."""
    from stinet_b200.models.singleconvmeshnet import SingleConvMeshNet
    fix = load(name)
    net = SingleConvMeshNet(**fix["kwargs"])
    net.load_state_dict(fix["state_dict"], strict=True)
    before = {k: v.clone() for k, v in net.named_buffers()}
    choices = record_product_decisions(net.train(), fix["batch"])
    assert all(torch.equal(v, before[k]) for k, v in net.named_buffers())        # recording leaves no trace
    n_relu = sum(1 for k, _ in choices if k == "relu")
    n_pool = sum(1 for k, _ in choices if k == "pool")
    levels, steps = len(fix["kwargs"]["filter_sizes"]), fix["kwargs"]["num_propagation_steps"]
    assert n_relu == 2 * steps * (2 * levels - 1) + 1                            # MLP + block ReLU per conv, + the head
    assert n_pool == (levels - 1 if fix["kwargs"]["pooling_method"] == "max" else 0)
    check_against_replaying_oracle(net, fix["batch"], fix)
