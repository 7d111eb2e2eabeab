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
