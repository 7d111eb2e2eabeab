"""GraphedTrainStep (whole-step CUDA graph) against the eager step: same losses and same updated parameters, bit for
bit, on the captured batch and on a different batch of the same shape copied into the static inputs."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _loss(out, b):
    composed = torch.where((b.mask > 0).expand_as(b.color), out, b.color)
    return ((composed - b.color).abs() * torch.pow(0.99, b.mask.squeeze().float()).unsqueeze(1)).mean()


def _make(seed):
    from stinet_b200 import synthetic
    return synthetic.make_batch("icosphere", 2, 2, seed=seed, subdiv=3, mask_radius=3)


def _net():
    from stinet_b200.models import surfacetextureinpaintingnet as S
    torch.manual_seed(49)
    return S.define_G(input_nc=10, output_nc=3, ngf=16, filter_type="edgeconvtransinv", norm="instance", n_blocks=2,
                      n_levels=2, pooling_type="max", gpu_ids=[torch.device(DEV)]).train()


@pytest.mark.parametrize("prefetch", [False, True])
def test_graphed_step_equals_eager_step(prefetch):
    from stinet_b200.engine import GraphedTrainStep, batch_signature
    batches = [_make(49), _make(50), _make(49)]
    assert batch_signature(batches[0]) == batch_signature(batches[1])

    def run(graphed: bool):
        net = _net()
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, amsgrad=True, fused=True, capturable=True)
        losses = []
        if graphed:
            # capturing is free of side effects (parameters, buffers and optimizer state are put back after the warm-up),
            # so the first batch is trained exactly once, like every other one
            step = GraphedTrainStep(net, _loss, opt, warmup=1)
            pinned = [b.pin_memory() for b in batches]
            if prefetch:                                  # loader pattern: batch k+1 moves H2D while step k runs
                assert not step.prefetch(pinned[0])       # (nothing to stage into before the shape has been captured)
            for i, b in enumerate(pinned):
                loss = step(b)
                if prefetch and i + 1 < len(pinned):
                    assert step.prefetch(pinned[i + 1])
                losses.append(float(loss.item()))
            assert step.captures == 1 and step.replayed_launches > 0
        else:
            for b in batches:
                gb = b.to(DEV)
                opt.zero_grad(set_to_none=True)
                loss = _loss(net(gb), gb)
                loss.backward()
                opt.step()
                losses.append(float(loss.item()))
        return losses, {k: v.detach().clone() for k, v in net.state_dict().items()}

    l_eager, p_eager = run(False)
    l_graph, p_graph = run(True)
    assert l_eager == l_graph, (l_eager, l_graph)
    for k in p_eager:
        assert torch.equal(p_eager[k], p_graph[k]), k


def test_graphed_step_with_changing_batch_shapes_equals_eager():
    """The 3D trainer's regime: consecutive crops of different size.  Every new shape signature is captured on first use
    (warm-up without side effects), later ones replay; losses and final parameters equal the eager loop bit for bit."""
    from stinet_b200 import synthetic
    from stinet_b200.engine import GraphedTrainStep
    shapes = {"a": dict(subdiv=3, mask_radius=3), "b": dict(subdiv=2, mask_radius=2)}
    order = ["a", "b", "a", "b", "b", "a"]
    batches = [synthetic.make_batch("icosphere", 2, 2, seed=60 + i, **shapes[k]) for i, k in enumerate(order)]

    def run(graphed: bool):
        net = _net()
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, amsgrad=True, fused=True, capturable=True)
        losses = []
        if graphed:
            step = GraphedTrainStep(net, _loss, opt, warmup=1)
            for b in batches:
                losses.append(float(step(b.pin_memory()).item()))
            assert step.captures == 2
        else:
            for b in batches:
                gb = b.to(DEV)
                opt.zero_grad(set_to_none=True)
                loss = _loss(net(gb), gb)
                loss.backward()
                opt.step()
                losses.append(float(loss.item()))
        return losses, {k: v.detach().clone() for k, v in net.state_dict().items()}

    l_eager, p_eager = run(False)
    l_graph, p_graph = run(True)
    assert l_eager == l_graph, (l_eager, l_graph)
    for k in p_eager:
        assert torch.equal(p_eager[k], p_graph[k]), k


def test_graphed_forward_equals_eager_forward():
    from stinet_b200.engine import GraphedForward
    net = _net().eval()
    fwd = GraphedForward(net)
    for seed in (49, 50, 49):
        b = _make(seed)
        with torch.no_grad():
            ref = net(b.to(DEV))
        out = fwd(b.pin_memory())
        assert torch.equal(out, ref)
    assert fwd.captures == 1 and fwd.replayed_launches > 0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_gradients_equal_single_process_gradient():
    """SURVEY 8e / T4 on hardware: under torchrun with 2 ranks (NCCL), the reducer's averaged gradients of the real network
    equal the single-process gradient of the mean loss over both ranks' batches, and the graphed 2-rank step (all-reduce
    captured inside the graph) equals the eager 2-rank step bit for bit (scripts/ddp_equivalence.py)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29541", os.path.join(root, "scripts", "ddp_equivalence.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DDP_EQUIVALENCE_OK" in r.stdout
