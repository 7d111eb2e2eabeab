"""Structure cached per sample and batched block-diagonally (stinet_b200.structure) against the
structure built from the collated COO tensors: every int32 array bit-exact, the network's outputs and gradients
bit-identical, and the whole-step CUDA graph fed with index-free batches equal to the eager step."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _samples(ragged: bool, n_levels: int = 2):
    from stinet_b200 import synthetic
    if not ragged:
        return synthetic.make_samples("icosphere", 3, n_levels, seed=49, subdiv=3, mask_radius=3)
    a = synthetic.make_samples("icosphere", 2, n_levels, seed=49, subdiv=3, mask_radius=3)
    b = synthetic.make_samples("plane", 1, n_levels, seed=7, rows=19, cols=23, mask_radius=3)
    return [a[0], b[0], a[1]]


def _structs(samples, n_levels):
    from stinet_b200.structure import SampleStructure
    return [SampleStructure.build(s, n_levels, DEV) for s in samples]


@pytest.mark.parametrize("ragged", [False, True])
def test_concatenated_structure_is_bit_identical(ragged):
    from stinet_b200.data import collate
    from stinet_b200.graph import GraphCache
    from stinet_b200.structure import attach_batch_structure
    L = 2
    samples = _samples(ragged, L)
    ref = GraphCache(collate(samples).to(DEV), L)                     # stinet_csr_build on the collated int64 tensors
    got_batch = attach_batch_structure(collate(samples, keep_index=False).to(DEV), _structs(samples, L))
    assert "edge_index" not in got_batch and "hierarchy_trace_index_1" not in got_batch
    got = GraphCache(got_batch, L)
    for key, lvl in (("edge_index", 0), ("hierarchy_edge_index_1", 1), ("hierarchy_edge_index_2", 2)):
        a, b = ref.edges(key, lvl), got.edges(key, lvl)
        assert (a.n, a.e) == (b.n, b.e)
        for name in ("rowptr_t", "col_t", "eid_t"):
            assert torch.equal(getattr(a, name), getattr(b, name)), (key, name)
        for x, y, name in zip(a.by_source(), b.by_source(), ("rowptr_s", "col_s", "eid_s")):
            assert torch.equal(x, y), (key, name)
        assert torch.equal(a.degree, b.degree)
        assert torch.equal(a.tpos_s()[:a.e], b.tpos_s()[:b.e])
    for lvl in (1, 2):
        a, b = ref.cluster(lvl), got.cluster(lvl)
        assert (a.n_fine, a.n_coarse) == (b.n_fine, b.n_coarse)
        for name in ("rowptr", "member", "trace32"):
            assert torch.equal(getattr(a, name), getattr(b, name)), (lvl, name)
        assert torch.equal(ref.graph_id(lvl), got.graph_id(lvl))
    ref.check_status()


def test_concat_many_parts_and_unaligned_offsets():
    """More parts than one launch carries (32) and odd lengths, so both the 128-bit and the scalar body run."""
    import ctypes
    from stinet_b200 import _abi
    from stinet_b200.structure import _concat
    g = torch.Generator().manual_seed(3)
    lens = [int(x) for x in torch.randint(0, 700, (70,), generator=g)]
    lens[5] = 0
    parts = [torch.randint(-1000, 1000, (n,), generator=g, dtype=torch.int32).to(DEV) for n in lens]
    adds = [int(x) for x in torch.randint(-50, 50, (70,), generator=g)]
    offs = [0]
    for n in lens:
        offs.append(offs[-1] + n)
    out = _concat(parts, lens, offs[:-1], adds, offs[-1], DEV)
    want = torch.cat([p + a for p, a in zip(parts, adds)])
    assert torch.equal(out, want)
    lib = _abi.load()
    assert lib.stinet_concat_i32(None, None, None, None, 3, None, None) == -1      # argument error, not a crash
    del ctypes


def _loss(out, b):
    composed = torch.where((b.mask > 0).expand_as(b.color), out, b.color)
    return ((composed - b.color).abs() * torch.pow(0.99, b.mask.squeeze().float()).unsqueeze(1)).mean()


def _net():
    from stinet_b200.models import surfacetextureinpaintingnet as S
    torch.manual_seed(49)
    return S.define_G(input_nc=10, output_nc=3, ngf=16, filter_type="edgeconvtransinv", norm="instance", n_blocks=2,
                      n_levels=2, pooling_type="max", gpu_ids=[torch.device(DEV)]).train()


def test_network_on_cached_structure_equals_network_on_coo():
    from stinet_b200.data import collate
    from stinet_b200.structure import attach_batch_structure
    samples = _samples(False)
    net = _net()
    full = collate(samples).to(DEV)
    lean = attach_batch_structure(collate(samples, keep_index=False).to(DEV), _structs(samples, 2))
    res = []
    for b in (full, lean):
        net.zero_grad(set_to_none=True)
        out = net(b)
        _loss(out, b).backward()
        res.append((out.detach().clone(), [p.grad.clone() for p in net.parameters()]))
    assert torch.equal(res[0][0], res[1][0])
    for g0, g1 in zip(res[0][1], res[1][1]):
        assert torch.equal(g0, g1)


def test_graphed_step_on_index_free_batches():
    """GraphedTrainStep fed with batches that carry no COO tensors at all (features from pinned host memory,
    structure arrays on the device): no structure-build kernels in the graph, same losses as the eager COO step."""
    import copy
    from stinet_b200 import synthetic
    from stinet_b200.data import collate
    from stinet_b200.engine import GraphedTrainStep
    from stinet_b200.structure import attach_batch_structure
    sets = [synthetic.make_samples("icosphere", 2, 2, seed=s, subdiv=3, mask_radius=3) for s in (49, 50, 49)]

    def run(graphed: bool):
        net = _net()
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, amsgrad=True, fused=True, capturable=True)
        losses = []
        if graphed:
            step = GraphedTrainStep(net, _loss, opt, warmup=1)
            lean = [attach_batch_structure(collate(s, keep_index=False).pin_memory(), _structs(s, 2)) for s in sets]
            state = copy.deepcopy(net.state_dict())
            step(lean[0])
            net.load_state_dict(state)
            for g in step.opt.state.values():
                for v in g.values():
                    if torch.is_tensor(v):
                        v.zero_()
            assert step.prefetch(lean[0])
            for i, b in enumerate(lean):
                loss = step(b)
                if i + 1 < len(lean):
                    assert step.prefetch(lean[i + 1])
                losses.append(float(loss.item()))
            assert step.captures == 1
        else:
            for s in sets:
                gb = collate(s).to(DEV)
                opt.zero_grad(set_to_none=True)
                loss = _loss(net(gb), gb)
                loss.backward()
                opt.step()
                losses.append(float(loss.item()))
        return losses

    assert run(False) == run(True)
