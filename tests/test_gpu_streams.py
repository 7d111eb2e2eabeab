"""The two concurrency devices of a step -- the batch's structure built on a side stream next to the first layers
(GraphCache.build_ahead) and the weight gradients on a second stream next to the data-gradient chain
(ops._on_wgrad_stream) -- must not change a single bit: same loss and same gradients as the one-stream schedule,
eagerly and when the step is replayed from a CUDA graph several times."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _loss(out, b):
    composed = torch.where((b.mask > 0).expand_as(b.color), out, b.color)
    return ((composed - b.color).abs() * torch.pow(0.99, b.mask.squeeze().float()).unsqueeze(1)).mean()


def _make(seed):
    from stinet_b200 import synthetic
    return synthetic.make_batch("icosphere", 3, 2, seed=seed, subdiv=4, mask_radius=3)


def _net():
    from stinet_b200.models import surfacetextureinpaintingnet as S
    torch.manual_seed(49)
    return S.define_G(input_nc=10, output_nc=3, ngf=32, filter_type="edgeconv", norm="instance", n_blocks=3,
                      n_levels=2, pooling_type="max", gpu_ids=[torch.device(DEV)]).train()


def _grads(monkeypatch, struct_side: str, wgrad_side: int, reps: int = 3):
    from stinet_b200 import ops
    monkeypatch.setenv("STINET_STRUCT_SIDE_STREAM", struct_side)
    monkeypatch.setattr(ops, "_WGRAD_SIDE", wgrad_side)
    net = _net()
    out = []
    for r in range(reps):
        b = _make(49 + r).to(DEV)
        net.zero_grad(set_to_none=True)
        loss = _loss(net(b), b)
        loss.backward()
        torch.cuda.synchronize()
        out.append((float(loss.item()), [p.grad.detach().clone() for p in net.parameters()]))
    return out


@pytest.mark.parametrize("struct_side,wgrad_side", [("1", 0), ("0", 1), ("0", 2), ("1", 2)])
def test_side_streams_do_not_change_a_bit(monkeypatch, struct_side, wgrad_side):
    ref = _grads(monkeypatch, "0", 0)
    got = _grads(monkeypatch, struct_side, wgrad_side)
    for (l0, g0), (l1, g1) in zip(ref, got):
        assert l0 == l1
        for a, b in zip(g0, g1):
            assert torch.equal(a, b)


@pytest.mark.parametrize("wgrad_side", [1, 2])
def test_side_streams_inside_a_captured_step(monkeypatch, wgrad_side):
    from stinet_b200 import ops
    from stinet_b200.engine import GraphedTrainStep
    batches = [_make(49), _make(50), _make(51), _make(49)]

    def run(struct_side, wg, graphed):
        monkeypatch.setenv("STINET_STRUCT_SIDE_STREAM", struct_side)
        monkeypatch.setattr(ops, "_WGRAD_SIDE", wg)
        net = _net()
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, amsgrad=True, fused=True, capturable=True)
        losses = []
        if graphed:
            step = GraphedTrainStep(net, _loss, opt, warmup=1)
            for b in batches:
                losses.append(float(step(b.pin_memory()).item()))
            assert step.captures == 1
        else:
            for b in batches:
                gb = b.to(DEV)
                opt.zero_grad(set_to_none=True)
                loss = _loss(net(gb), gb)
                loss.backward()
                opt.step()
                losses.append(float(loss.item()))
        torch.cuda.synchronize()
        return losses, [p.detach().clone() for p in net.parameters()]

    l0, p0 = run("0", 0, False)
    l1, p1 = run("1", wgrad_side, True)
    assert l0 == l1, (l0, l1)
    for a, b in zip(p0, p1):
        assert torch.equal(a, b)


@pytest.mark.parametrize("aware", [True, False])
def test_gradient_hooks_see_finished_gradients(monkeypatch, aware):
    """Post-accumulate hooks (what GradAllReducer hangs on every parameter) under the deferred join: a hook flagged as
    aware of the second stream joins it itself (ops.join_wgrad_stream) and then reads the deposited gradient; an unflagged
    (foreign) hook makes the producing node join before it returns.  Either way the gradient a hook reads is the
    one-stream gradient, bit for bit, and every parameter's hook runs at least once."""
    from stinet_b200 import ops
    monkeypatch.setenv("STINET_STRUCT_SIDE_STREAM", "0")
    net = _net()
    names = {p: n for n, p in net.named_parameters()}
    seen = {}

    def hook(p):
        if aware:
            ops.join_wgrad_stream()
        seen.setdefault(names[p], p.grad.detach().clone())

    def step(mode, defer):
        monkeypatch.setattr(ops, "_WGRAD_SIDE", mode)
        seen.clear()
        b = _make(49).to(DEV)
        for p in net.parameters():
            p.grad = None
        loss = _loss(net(b), b)
        if defer:
            with ops.deferred_wgrad_join():
                loss.backward()
        else:
            loss.backward()
        torch.cuda.synchronize()
        return dict(seen), {n: p.grad.detach().clone() for n, p in net.named_parameters()}

    _, ref = step(0, False)
    for p in net.parameters():
        p.register_post_accumulate_grad_hook(hook)
        if aware:
            p._stinet_wgrad_aware = True
    for mode, defer in [(1, False), (1, True), (2, False)]:
        at_hook, final = step(mode, defer)
        assert set(at_hook) == set(ref)
        for n in ref:
            assert torch.equal(at_hook[n], ref[n]), (mode, defer, n)
            assert torch.equal(final[n], ref[n]), (mode, defer, n)


def test_gradient_accumulation_under_deferred_join(monkeypatch):
    """Two backward passes without zeroing in between: the second one finds `.grad` occupied, so its weight gradients go
    through autograd's own accumulation (per-node join) instead of being deposited -- the sum must be exactly what the
    one-stream schedule accumulates."""
    from stinet_b200 import ops
    monkeypatch.setenv("STINET_STRUCT_SIDE_STREAM", "0")

    def run(mode):
        monkeypatch.setattr(ops, "_WGRAD_SIDE", mode)
        net = _net()
        for seed in (49, 50):
            b = _make(seed).to(DEV)
            loss = _loss(net(b), b)
            with ops.deferred_wgrad_join():
                loss.backward()
        torch.cuda.synchronize()
        return [p.grad.detach().clone() for p in net.parameters()]

    ref = run(0)
    for mode in (1, 2):
        for a, b in zip(ref, run(mode)):
            assert torch.equal(a, b)
