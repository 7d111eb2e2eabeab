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
            step = GraphedTrainStep(net, _loss, opt, warmup=1)
            # the warm-up steps inside the capture already update the weights: restore them so both runs start equal
            first = batches[0].pin_memory()
            state = copy.deepcopy(net.state_dict())
            step(first)                                   # capture (+ warm-up + first replay)
            net.load_state_dict(state)
            opt = step.opt
            for g in opt.state.values():                  # reset Adam moments / step counters in place
                for v in g.values():
                    if torch.is_tensor(v):
                        v.zero_()
            pinned = [b.pin_memory() for b in batches]
            if prefetch:                                  # loader pattern: batch k+1 moves H2D while step k runs
                assert step.prefetch(pinned[0])
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
