"""Host-side logic on CPU: collate offsets (bit-exact vs the reference's HierarchicalData golden batches), synthetic
generators, norm segments, module API mirror."""
import inspect

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden
from oracle import stinet_oracle as O
from stinet_b200 import data, synthetic
from stinet_b200.graph import Segments
from stinet_b200.models import modules, surfacetextureinpaintingnet as S

GENS = {"grid": synthetic.grid_sample, "icosphere": synthetic.icosphere_sample}


@pytest.mark.parametrize("name", GOLDEN)
def test_collate_matches_reference_batch_bit_exact(name):
    """Golden batches were collated by the REAL utils/data_utils.HierarchicalData rules."""
    fix = load_golden(name)
    ours = data.collate([GENS[k](**kw) for k, kw in fix["specs"]])
    ref = fix["sample"]
    for key, val in ref.items():
        if torch.is_tensor(val):
            got = ours[key]
            assert got.dtype == val.dtype and got.shape == val.shape, key
            assert torch.equal(got, val), key


def test_dilated_keys_keep_the_reference_offset_quirk():
    a = synthetic.icosphere_sample(2, 2, seed=1, mask_radius=2, dilations=(2,))
    b = synthetic.icosphere_sample(2, 2, seed=2, mask_radius=2, dilations=(2,))
    batch = data.collate([a, b])
    e = a["hierarchy_dil_2_edge_index_2"].shape[1]
    # second sample offset by N0 of the first (PyG default num_nodes), not by N2 (data_utils.py:42)
    assert torch.equal(batch["hierarchy_dil_2_edge_index_2"][:, e:], b["hierarchy_dil_2_edge_index_2"] + a.num_nodes)


def test_icosphere_hierarchy_properties():
    s = synthetic.icosphere_sample(3, 3, seed=0, mask_radius=3)
    nv = s.num_vertices.tolist()
    assert nv == [642, 162, 42, 12]
    assert s.edge_index.shape[1] == 6 * 642 - 12
    for lvl in range(1, 4):
        tr = s[f"hierarchy_trace_index_{lvl}"]
        assert tr.numel() == nv[lvl - 1]
        assert torch.equal(torch.unique(tr), torch.arange(nv[lvl]))         # surjective (graph_level_generation.py:180)
        e = s[f"hierarchy_edge_index_{lvl}"]
        assert e.shape[1] == 6 * nv[lvl] - 12 and int(e.max()) == nv[lvl] - 1
    e0 = s.edge_index
    assert set(map(tuple, e0.t().tolist())) == set(map(tuple, e0.flip(0).t().tolist()))   # symmetric
    assert s.x.shape == (642, 10) and s.mask.shape == (642, 1)
    known = s.x[:, 9] > 0
    assert torch.equal(known, s.mask.squeeze(1) == 0)
    assert float(s.x[~known][:, :3].abs().max()) == 0.0                     # colour zeroed inside the hole


def test_grid_sample_matches_reference_layout():
    s = synthetic.grid_sample(16, 2, seed=3)
    assert s.num_vertices.tolist() == [256, 64, 16]
    assert s.edge_index.shape[1] == 4 * 16 * 15
    t1 = s["hierarchy_trace_index_1"].view(16, 16)
    assert int(t1[0, 0]) == int(t1[1, 1]) == 0 and int(t1[2, 2]) == 9       # 2x decimation (imagegraph_dataloader.py:44-57)
    m = s.mask.squeeze(1)
    assert torch.equal(s.x[:, 3] > 0, m)
    assert float(s.x[m][:, :3].abs().max()) == 0.0


def test_segments_linspace_quirk():
    seg = Segments(12, [6, 6], torch.zeros(12, dtype=torch.int32), "cpu")
    assert seg.consistent and seg.slice_ptr.tolist() == [0, 6, 12] and seg.max_seg_rows == 6
    rag = Segments(10, [7, 3], torch.zeros(10, dtype=torch.int32), "cpu")
    assert not rag.consistent and rag.slice_ptr.tolist() == [0, 5, 10] and rag.cnt.tolist() == [7.0, 3.0]
    one = Segments(9, None, None, "cpu")
    assert one.n_seg == 1 and one.gid is None and one.consistent


def test_module_api_mirrors_reference_signatures():
    sig = inspect.signature(S.define_G)
    for name in ["input_nc", "output_nc", "ngf", "filter_type", "norm", "dilation_order", "use_dropout", "n_blocks",
                 "n_levels", "n_repeated_io_convs", "init_type", "pooling_type", "io_receptive_field_type",
                 "checkpoint_bottleneck", "num_blocks_per_uncheckpointed_block", "use_label_embedding", "num_classes",
                 "num_embedding", "dilations", "init_gain", "gpu_ids"]:
        assert name in sig.parameters, name
    for cls in ["EdgeConvTransInv", "SAGEConvTransInv", "FastInstanceNorm", "SingleBatchGraphNorm"]:
        assert hasattr(modules, cls)
    assert list(inspect.signature(modules.edge_conv_filter.get_gcn_filter).parameters)[:2] == ["input_size", "output_size"]
    assert "batch" in inspect.signature(modules.FastInstanceNorm.forward).parameters
    assert "batch" in inspect.signature(modules.SingleBatchGraphNorm.forward).parameters


@pytest.mark.parametrize("ft,nc", [("edgeconv", 4), ("edgeconvtransinv", 10), ("sageconv", 4), ("sageconvtransinv", 10)])
@pytest.mark.parametrize("norm", ["instance", "graph", "batch", "none"])
def test_state_dict_layout_and_seeded_init_match_oracle(ft, nc, norm):
    kw = dict(input_nc=nc, output_nc=3, ngf=8, n_blocks=2, n_levels=2, pooling_type="max")
    torch.manual_seed(7)
    ours = S.define_G(filter_type=ft, norm=norm, **kw)
    torch.manual_seed(7)
    orc = O.OracleSTINet(filter_type=ft, norm_type=norm, **kw)
    a, b = ours.state_dict(), orc.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].shape == b[k].shape and torch.equal(a[k], b[k]), k


def test_shipped_configs_param_counts():
    n3 = S.define_G(input_nc=10, output_nc=3, ngf=64, n_blocks=9, dilations=[1, 1, 1, 2, 4, 8, 16, 1, 1],
                    norm="instance", pooling_type="max", n_levels=2, filter_type="edgeconvtransinv",
                    checkpoint_bottleneck=True, num_blocks_per_uncheckpointed_block=1, use_label_embedding=False,
                    num_classes=21, num_embedding=12, use_dropout=False, init_type="normal", init_gain=0.02,
                    n_repeated_io_convs=1)
    assert sum(p.numel() for p in n3.parameters()) == 4202051           # thesis 4.3: "4.2 million"
    n2 = S.define_G(input_nc=4, output_nc=3, ngf=64, n_blocks=9, dilations=[1] * 9, norm="instance",
                    pooling_type="max", n_levels=2, filter_type="edgeconv", checkpoint_bottleneck=False)
    assert sum(p.numel() for p in n2.parameters()) == 4201411


def test_ragged_instance_norm_backward_composite_matches_autograd():
    """ops._ragged_norm_backward (plain tensor ops, device-agnostic) against autograd of the oracle's restatement of the
    reference's linspace-slice statistics (fastinstancenorm.py:53-82), fp64, with and without the fused ELU."""
    from types import SimpleNamespace
    from oracle import stinet_oracle as O
    from stinet_b200 import ops
    torch.manual_seed(0)
    counts = [500, 77, 323]
    n, c = sum(counts), 5
    x = torch.randn(n, c, dtype=torch.float64) * 3 + 1.5
    go = torch.randn(n, c, dtype=torch.float64)
    batch = torch.repeat_interleave(torch.arange(3), torch.tensor(counts))
    sp = torch.linspace(0, n, 4, dtype=torch.int).tolist()
    cnt = torch.tensor(counts, dtype=torch.float64).view(-1, 1)
    mean = torch.stack([x[sp[i]:sp[i + 1]].sum(0) for i in range(3)]) / cnt
    xc = x - mean[batch]
    rstd = (torch.stack([xc[sp[i]:sp[i + 1]].pow(2).sum(0) for i in range(3)]) / cnt + 1e-5).rsqrt()
    seg = SimpleNamespace(true_ptr=[0, 500, 577, 900], slice_ptr_host=sp, n_seg=3, gid=batch.int(), cnt=cnt.flatten())
    for act in (0, 1):
        xr = x.clone().requires_grad_(True)
        y = O.fast_instance_norm(xr, batch)
        (torch.nn.functional.elu(y) if act else y).backward(go)
        dx = ops._ragged_norm_backward(x, go, mean, rstd, seg, act)
        assert float((dx - xr.grad).abs().max() / xr.grad.abs().max()) < 1e-12


def test_segments_carry_true_and_slice_boundaries():
    from stinet_b200.graph import Segments
    rag = Segments(10, [7, 3], torch.zeros(10, dtype=torch.int32), "cpu")
    assert rag.true_ptr == [0, 7, 10] and rag.slice_ptr_host == [0, 5, 10] and not rag.consistent
    one = Segments(9, None, None, "cpu")
    assert one.true_ptr == [0, 9] and one.slice_ptr_host == [0, 9]


def test_flat_views_pack_a_batch_into_one_buffer():
    """engine._flat_views: every tensor of a batch as a 256-byte aligned view of one flat buffer (what lets a
    prefetched batch move device-to-device with a single copy)."""
    from stinet_b200.engine import _flat_views
    items = [("x", torch.randn(7, 3)), ("edge_index", torch.arange(10).view(2, 5)), ("empty", torch.zeros(0, 4)),
             ("m", torch.tensor([[1], [0], [1]], dtype=torch.int32))]
    flat, views = _flat_views(items, "cpu")
    v = views(flat)
    offs = []
    for k, t in items:
        assert v[k].shape == t.shape and v[k].dtype == t.dtype
        v[k].copy_(t)
        if t.numel():                                   # an empty view has no storage address worth checking
            offs.append(v[k].data_ptr() - flat.data_ptr())
    assert all(o % 256 == 0 for o in offs) and offs == sorted(offs)
    other = torch.empty_like(flat)
    other.copy_(flat)                                   # ONE copy moves the whole batch
    for k, t in items:
        assert torch.equal(views(other)[k], t)


def test_block_diagonal_structure_offsets_with_stand_in_kernels(monkeypatch):
    """stinet_b200.structure: the per-array (length, destination offset, added offset) plan of the block-diagonal
    batching, checked against the structure of the COLLATED batch (the reference's HierarchicalData offsets,
    utils/data_utils.py:29-42).  The two C-ABI calls involved are replaced by plain-torch stand-ins inside this test
    only; on the B200 tests/test_gpu_structure.py checks the real kernels bit for bit."""
    from oracle import stinet_oracle as O
    from stinet_b200 import graph, structure, synthetic
    from stinet_b200.data import collate

    def build_csr(key, other, n_rows, want_key32=False, status=None):
        rowptr, perm = O.csr_by_key(key, n_rows)
        col = other[perm.long()].to(torch.int32) if other is not None else None
        return rowptr, perm, col, (key.to(torch.int32) if want_key32 else None)

    def concat(parts, lens, dst_offs, adds, total, device):
        out = torch.full((total,), -12345, dtype=torch.int32)
        for p, n, o, a in zip(parts, lens, dst_offs, adds):
            out[o:o + n] = p[:n] + a
        assert not (out == -12345).any()                # every element written exactly by the plan
        return out

    def tpos(self):
        _, _, eid_s = self.by_source()
        inv = torch.empty(max(self.e, 1), dtype=torch.int32)
        inv[self.eid_t.long()] = torch.arange(self.e, dtype=torch.int32)
        return inv[eid_s.long()] if self.e else inv

    monkeypatch.setattr(graph, "build_csr", build_csr)
    monkeypatch.setattr(graph, "_require_cuda", lambda *a, **k: None)
    monkeypatch.setattr(graph.EdgeCSR, "tpos_s", tpos)
    monkeypatch.setattr(structure, "_concat", concat)
    a = synthetic.make_samples("icosphere", 2, 2, seed=49, subdiv=2, mask_radius=2)
    b = synthetic.make_samples("plane", 1, 2, seed=7, rows=9, cols=11, mask_radius=2)
    samples = [a[0], b[0], a[1]]
    L = 2
    ref = graph.GraphCache(collate(samples), L)
    structs = [structure.SampleStructure.build(s, L, "cpu") for s in samples]
    lean = structure.attach_batch_structure(collate(samples, keep_index=False), structs, "cpu")
    assert "edge_index" not in lean
    got = graph.GraphCache(lean, L)
    for key, lvl in (("edge_index", 0), ("hierarchy_edge_index_1", 1), ("hierarchy_edge_index_2", 2)):
        x, y = ref.edges(key, lvl), got.edges(key, lvl)
        for name in ("rowptr_t", "col_t", "eid_t"):
            assert torch.equal(getattr(x, name), getattr(y, name)), (key, name)
        for p, q in zip(x.by_source(), y.by_source()):
            assert torch.equal(p, q)
        assert torch.equal(x.tpos_s()[:x.e], y.tpos_s()[:y.e])
    for lvl in (1, 2):
        x, y = ref.cluster(lvl), got.cluster(lvl)
        for name in ("rowptr", "member", "trace32"):
            assert torch.equal(getattr(x, name), getattr(y, name)), (lvl, name)
