"""Per-kernel parity on the B200: CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs.
Integer / index work is bit-exact; fp32 within 1e-5 relative (||a-b||inf / ||b||inf per tensor)."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import stinet_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5
DEV = "cuda"


def _graph(kind):
    from stinet_b200 import synthetic
    if kind == "ico":
        s = synthetic.icosphere_sample(4, 2, seed=11, mask_radius=3)
        return s.edge_index, s.num_nodes
    if kind == "grid_shuffled":
        s = synthetic.grid_sample(32, 1, seed=5)
        return s.edge_index, s.num_nodes
    if kind == "graph18_isolated":
        return synthetic.paper_graph18()
    if kind == "dilated_asym":
        s = synthetic.icosphere_sample(3, 2, seed=12, mask_radius=3, dilations=(2,))
        return s["hierarchy_dil_2_edge_index_2"], int(s.num_vertices[2])
    if kind == "multi_edges":       # duplicates + self loops + a hub: nothing the builder may assume away
        g = torch.Generator().manual_seed(0)
        e = torch.randint(0, 40, (2, 600), generator=g)
        e[1, :200] = 7               # hub with in-degree > 64 exercises the long-row sort
        return e, 41
    if kind == "empty":
        return torch.zeros((2, 0), dtype=torch.int64), 5
    raise KeyError(kind)


KINDS = ["ico", "grid_shuffled", "graph18_isolated", "dilated_asym", "multi_edges", "empty"]


@pytest.mark.parametrize("kind", KINDS)
def test_csr_build_bit_exact(kind):
    from stinet_b200.graph import EdgeCSR
    ei, n = _graph(kind)
    csr = EdgeCSR(ei.to(DEV), n)
    rp, perm = O.csr_by_key(ei[1], n)
    assert torch.equal(csr.rowptr_t.cpu(), rp)
    assert torch.equal(csr.eid_t.cpu(), perm)
    assert torch.equal(csr.col_t.cpu(), ei[0][perm.long()].to(torch.int32))
    rs, cs, es = csr.by_source()
    rp2, perm2 = O.csr_by_key(ei[0], n)
    assert torch.equal(rs.cpu(), rp2) and torch.equal(es.cpu(), perm2)
    assert torch.equal(cs.cpu(), ei[1][perm2.long()].to(torch.int32))


def test_csr_flags_out_of_range_index():
    from stinet_b200.graph import EdgeCSR
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    ei = torch.tensor([[0, 1, 2], [1, 9, 0]])
    EdgeCSR(ei.to(DEV), 3, status)
    assert int(status.item()) == 1


def test_csr_out_of_range_items_are_dropped_not_dereferenced():
    """Keys outside [0, n) are counted out (status bit 0) and leave a tail in perm that every later pass must skip: the
    by-target / by-source structures and the cross positions of the VALID items equal the oracle's on the filtered edge
    list, and nothing reads or writes through a garbage index (the reference would raise an index error instead)."""
    from stinet_b200.graph import EdgeCSR
    g = torch.Generator().manual_seed(3)
    n, e = 500, 6000
    ei = torch.randint(0, n, (2, e), generator=g)
    bad = torch.randperm(e, generator=g)[:300]
    ei[1, bad[:150]] = n + torch.randint(0, 1000, (150,), generator=g)       # targets beyond the level
    ei[1, bad[150:]] = -1 - torch.randint(0, 5, (150,), generator=g)         # negative targets
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    csr = EdgeCSR(ei.to(DEV), n, status)
    tpos = csr.tpos_s()
    torch.cuda.synchronize()
    assert int(status.item()) == 1
    ok = (ei[1] >= 0) & (ei[1] < n)
    n_valid = int(ok.sum())
    rp, perm = O.csr_by_key(ei[1][ok], n)
    orig = torch.nonzero(ok).flatten()[perm.long()].to(torch.int32)           # original positions of the valid items
    assert torch.equal(csr.rowptr_t.cpu(), rp) and int(csr.rowptr_t[-1]) == n_valid
    assert torch.equal(csr.eid_t.cpu()[:n_valid], orig)
    assert bool((csr.eid_t.cpu()[n_valid:] == -1).all()) and bool((csr.col_t.cpu()[n_valid:] == 0).all())
    assert bool((tpos.cpu() >= 0).all()) and bool((tpos.cpu() < e).all())


@pytest.mark.parametrize("kind", ["ico", "graph18_isolated", "dilated_asym", "multi_edges"])
@pytest.mark.parametrize("reduce", ["mean", "add", "max"])
@pytest.mark.parametrize("c", [3, 16, 128])
def test_aggregate_fwd_bwd(kind, reduce, c):
    from stinet_b200 import ops
    from stinet_b200.graph import EdgeCSR
    ei, n = _graph(kind)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, c, generator=g)
    if reduce == "max":
        x = torch.round(x * 2) / 2          # many exact ties: the first-edge-wins rule must match bit for bit
    go = torch.randn(n, c, generator=g)
    xr = x.clone().requires_grad_(True)
    msg = xr.index_select(0, ei[0])
    if reduce == "max":
        ref, ref_arg = O.scatter_max(msg, ei[1], n)
    else:
        ref = O.aggregate(msg, ei[1], n, reduce)
    ref.backward(go)
    xd = x.to(DEV).requires_grad_(True)
    csr = EdgeCSR(ei.to(DEV), n)
    out = ops.aggregate(xd, csr, reduce)
    if reduce == "max":
        out, arg = out
        assert torch.equal(arg.cpu().long(), ref_arg)
        assert torch.equal(out.cpu(), ref.detach())
    out.backward(go.to(DEV))
    assert rel_err(out, ref) <= TOL
    assert rel_err(xd.grad, xr.grad) <= TOL


@pytest.mark.parametrize("kind", ["ico", "graph18_isolated", "dilated_asym", "multi_edges", "grid_shuffled"])
@pytest.mark.parametrize("din,dout,trans_inv", [(10, 8, True), (4, 8, False), (64, 64, False), (32, 100, False)])
def test_edge_conv_vs_literal_per_edge_mlp(kind, din, dout, trans_inv):
    """Hoisted, fused CUDA EdgeConv == the reference's per-edge MLP + scatter-mean (values and all gradients)."""
    from stinet_b200.models.modules import edge_conv_filter, edge_conv_translation_invariance
    ei, n = _graph(kind)
    module = edge_conv_translation_invariance.EdgeConvTransInv if trans_inv else None
    # ReLU'(0) is discontinuous: if a pre-activation lies within fp32 rounding of 0 the literal and the hoisted
    # evaluation order may legitimately pick different sides (scripts/diag_flip.py).  Screen the seed in fp64 so that
    # every one of the E x 2*dout decisions in this case has a margin, then demand strict 1e-5 parity.
    for seed in range(2, 40):
        torch.manual_seed(seed)
        conv = edge_conv_filter.get_gcn_filter(din, dout, module=module, double_input=not trans_inv)
        with torch.no_grad():
            for p in conv.parameters():
                if p.dim() == 1:
                    p.normal_(0, 0.5)       # non-zero biases: isolated vertices must still output exactly 0
        x = torch.randn(n, din)
        go = torch.randn(n, dout)
        xi, xj = x.double()[ei[1]], x.double()[ei[0]]
        inp = (xj - xi) if trans_inv else torch.cat([xi, xj - xi], 1)
        pre = inp @ conv.nn[0].weight.double().t() + conv.nn[0].bias.double()
        if pre.numel() == 0 or float(pre.abs().min()) > 1e-6:
            break
    else:
        pytest.skip("no seed with a ReLU margin found")
    xr = x.clone().requires_grad_(True)
    ref = O.edge_conv(xr, ei, conv.nn, "mean", trans_inv)
    ref.backward(go)
    ref_grads = {k: p.grad.clone() for k, p in conv.named_parameters()}
    conv.zero_grad()
    conv = conv.to(DEV)
    xd = x.to(DEV).requires_grad_(True)
    out = conv(xd, ei.to(DEV))
    out.backward(go.to(DEV))
    assert rel_err(out, ref) <= TOL
    assert rel_err(xd.grad, xr.grad) <= TOL
    for k, p in conv.named_parameters():
        assert rel_err(p.grad, ref_grads[k]) <= TOL, k
    deg = torch.bincount(ei[1], minlength=n)
    if (deg == 0).any():
        assert float(out[(deg == 0).to(DEV)].abs().max()) == 0.0


@pytest.mark.parametrize("kind", ["ico", "graph18_isolated", "multi_edges"])
@pytest.mark.parametrize("aggr", ["add", "max"])
@pytest.mark.parametrize("trans_inv", [False, True])
def test_edge_conv_add_and_max_aggregation(kind, aggr, trans_inv):
    """EdgeConv(aggr='add' | 'max') -- accepted by the reference's get_gcn_filter signature, never used by its configs --
    runs the literal per-edge form (the second Linear does not commute with these reductions): values and gradients
    against the oracle; vertices without in-edges output 0."""
    from stinet_b200.models.modules import edge_conv_filter, edge_conv_translation_invariance
    ei, n = _graph(kind)
    din, dout = 8, 12
    module = edge_conv_translation_invariance.EdgeConvTransInv if trans_inv else None
    torch.manual_seed(7)
    conv = edge_conv_filter.get_gcn_filter(din, dout, module=module, double_input=not trans_inv, aggregation=aggr)
    assert not conv.hoistable
    x = torch.randn(n, din)
    go = torch.randn(n, dout)
    xr = x.clone().requires_grad_(True)
    ref = O.edge_conv(xr, ei, conv.nn, aggr, trans_inv)
    ref.backward(go)
    ref_grads = {k: p.grad.clone() for k, p in conv.named_parameters()}
    conv.zero_grad()
    conv = conv.to(DEV)
    xd = x.to(DEV).requires_grad_(True)
    out = conv(xd, ei.to(DEV))
    out.backward(go.to(DEV))
    assert rel_err(out, ref) <= TOL
    assert rel_err(xd.grad, xr.grad) <= TOL
    for k, p in conv.named_parameters():
        assert rel_err(p.grad, ref_grads[k]) <= TOL, k
    deg = torch.bincount(ei[1], minlength=n)
    if (deg == 0).any():
        assert float(out[(deg == 0).to(DEV)].abs().max()) == 0.0


@pytest.mark.parametrize("c", [3, 8, 24, 64, 256, 1000])
@pytest.mark.parametrize("pool", ["max", "mean"])
def test_pool_unpool(c, pool):
    from stinet_b200 import ops, synthetic
    from stinet_b200.graph import ClusterCSR
    s = synthetic.icosphere_sample(4, 1, seed=3, mask_radius=2)
    trace = s["hierarchy_trace_index_1"]
    nf, nc = trace.numel(), int(s.num_vertices[1]) + 2          # two EMPTY clusters at the end (dim_size > max+1)
    g = torch.Generator().manual_seed(4)
    x = torch.round(torch.randn(nf, c, generator=g) * 2) / 2     # ties
    go = torch.randn(nc, c, generator=g)
    xr = x.clone().requires_grad_(True)
    if pool == "max":
        ref, ref_arg = O.scatter_max(xr, trace, nc)
    else:
        ref = O.scatter_mean(xr, trace, nc)
    ref.backward(go)
    cl = ClusterCSR(trace.to(DEV), nc)
    rp, member = O.csr_by_key(trace, nc)
    assert torch.equal(cl.rowptr.cpu(), rp) and torch.equal(cl.member.cpu(), member)
    assert torch.equal(cl.trace32.cpu().long(), trace)
    xd = x.to(DEV).requires_grad_(True)
    if pool == "max":
        out, arg = ops.pool_max(xd, cl)
        assert torch.equal(arg.cpu().long(), ref_arg)             # lowest fine id wins; empty -> n_fine sentinel
        assert torch.equal(out.cpu(), ref.detach())
    else:
        out = ops.pool_mean(xd, cl)
    out.backward(go.to(DEV))
    assert rel_err(out, ref) <= TOL and rel_err(xd.grad, xr.grad) <= TOL
    # unpool = gather; backward = segmented add
    xc = torch.randn(nc, c, generator=g)
    gf = torch.randn(nf, c, generator=g)
    xcr = xc.clone().requires_grad_(True)
    xcr[trace].backward(gf)
    xcd = xc.to(DEV).requires_grad_(True)
    up = ops.unpool(xcd, cl)
    up.backward(gf.to(DEV))
    assert torch.equal(up.cpu(), xc[trace])
    assert rel_err(xcd.grad, xcr.grad) <= TOL


@pytest.mark.parametrize("c", [4, 20, 48, 128, 160])
def test_pool_unpool_ragged_clusters(c):
    """cluster sizes 0..9 in random order: every tail length of the four-at-a-time member loop, every team width."""
    from stinet_b200 import ops
    from stinet_b200.graph import ClusterCSR
    g = torch.Generator().manual_seed(21)
    sizes = torch.randint(0, 10, (301,), generator=g)
    trace = torch.repeat_interleave(torch.arange(sizes.numel()), sizes)
    trace = trace[torch.randperm(trace.numel(), generator=g)]
    nf, nc = trace.numel(), sizes.numel()
    x = torch.round(torch.randn(nf, c, generator=g) * 2) / 2
    go = torch.randn(nc, c, generator=g)
    cl = ClusterCSR(trace.to(DEV), nc)
    for pool in ("max", "mean"):
        xr = x.clone().requires_grad_(True)
        if pool == "max":
            ref, ref_arg = O.scatter_max(xr, trace, nc)
        else:
            ref = O.scatter_mean(xr, trace, nc)
        ref.backward(go)
        xd = x.to(DEV).requires_grad_(True)
        if pool == "max":
            out, arg = ops.pool_max(xd, cl)
            assert torch.equal(arg.cpu().long(), ref_arg) and torch.equal(out.cpu(), ref.detach())
        else:
            out = ops.pool_mean(xd, cl)
        out.backward(go.to(DEV))
        assert rel_err(out, ref) <= TOL and rel_err(xd.grad, xr.grad) <= TOL
    xc = torch.randn(nc, c, generator=g)
    gf = torch.randn(nf, c, generator=g)
    xcr = xc.clone().requires_grad_(True)
    xcr[trace].backward(gf)
    xcd = xc.to(DEV).requires_grad_(True)
    up = ops.unpool(xcd, cl)
    up.backward(gf.to(DEV))
    assert torch.equal(up.cpu(), xc[trace])
    assert rel_err(xcd.grad, xcr.grad) <= TOL


def test_graph_id_pooling_int():
    from stinet_b200 import synthetic
    from stinet_b200.graph import GraphCache
    b = synthetic.make_batch("icosphere", 3, 2, subdiv=3, mask_radius=2)
    cache = GraphCache.for_sample(b.to(DEV), 2)
    gid = b.batch
    for lvl in (1, 2):
        gid = O.scatter_max(gid, b[f"hierarchy_trace_index_{lvl}"], cache.totals[lvl])[0]
        assert torch.equal(cache.graph_id(lvl).cpu().long(), gid)
    cache.check_status()


# layouts cover every instance-norm path: one cluster kernel with rows in registers (single / equal / empty_slice),
# cluster kernel re-reading rows (equal_mid: 8 CTAs x 375 rows), long slices (stats + slice apply: single_long,
# equal_long), graph-id lookups (ragged, and odd widths with several graphs)
@pytest.mark.parametrize("c", [3, 40, 64, 192])
@pytest.mark.parametrize("layout", ["single", "equal", "ragged", "equal_mid", "single_long", "equal_long"])
def test_instance_norm_elu_residual(c, layout):
    from stinet_b200.models.modules import FastInstanceNorm
    from stinet_b200._abi import ACT_ELU
    g = torch.Generator().manual_seed(5)
    counts = {"single": [700], "equal": [300, 300, 300], "ragged": [500, 77, 323], "equal_mid": [3000, 3000],
              "single_long": [17001], "equal_long": [16500, 16500]}[layout]
    n = sum(counts)
    x = torch.randn(n, c, generator=g) * 3 + 1.5
    res = torch.randn(n, c, generator=g)
    go = torch.randn(n, c, generator=g)
    batch = None if layout == "single" else torch.repeat_interleave(torch.arange(len(counts)), torch.tensor(counts))
    xr, rr = x.clone().requires_grad_(True), res.clone().requires_grad_(True)
    ref = rr + torch.nn.functional.elu(O.fast_instance_norm(xr, batch))
    ref.backward(go)
    norm = FastInstanceNorm(c)
    xd, rd = x.to(DEV).requires_grad_(True), res.to(DEV).requires_grad_(True)
    out = norm(xd, None if batch is None else batch.to(DEV), rd, ACT_ELU)
    assert rel_err(out, ref) <= TOL
    out.backward(go.to(DEV))                        # ragged: composite backward of the linspace-slice quirk
    assert rel_err(xd.grad, xr.grad) <= TOL
    assert torch.equal(rd.grad.cpu(), go)
    # no residual, no activation (the plain module call of the reference)
    x2 = x.to(DEV).requires_grad_(True)
    out2 = norm(x2, None if batch is None else batch.to(DEV))
    xr2 = x.clone().requires_grad_(True)
    ref2 = O.fast_instance_norm(xr2, batch)
    ref2.backward(go)
    out2.backward(go.to(DEV))
    assert rel_err(out2, ref2) <= TOL and rel_err(x2.grad, xr2.grad) <= TOL


@pytest.mark.parametrize("m,n,k", [(1000, 16, 10), (777, 128, 64), (130, 3, 64), (4097, 256, 20), (64, 512, 256), (5, 8, 4)])
@pytest.mark.parametrize("masked", [False, True])
def test_linear_fwd_dgrad_wgrad(m, n, k, masked):
    from stinet_b200 import ops
    g = torch.Generator().manual_seed(6)
    x = torch.randn(m, k, generator=g)
    w = torch.randn(n, k, generator=g) / k ** 0.5
    b = torch.randn(n, generator=g)
    go = torch.randn(m, n, generator=g)
    mask = (torch.rand(m, generator=g) > 0.3).to(torch.int32) if masked else None
    xr, wr, br = (t.double().requires_grad_(True) for t in (x, w, b))
    ref = xr @ wr.t() + (br * mask.double().unsqueeze(1) if masked else br)
    ref.backward(go.double())
    xd, wd, bd = (t.to(DEV).requires_grad_(True) for t in (x, w, b))
    out = ops.linear(xd, wd, bd, mask.to(DEV) if masked else None, "fp32")
    out.backward(go.to(DEV))
    assert rel_err(out, ref) <= TOL
    assert rel_err(xd.grad, xr.grad) <= TOL
    assert rel_err(wd.grad, wr.grad) <= TOL
    assert rel_err(bd.grad, br.grad) <= TOL


def test_kernels_are_bitwise_deterministic():
    """run-twice bit comparison (the reference's atomics are not deterministic; ours must be)."""
    from stinet_b200.models.modules import edge_conv_filter
    ei, n = _graph("ico")
    torch.manual_seed(8)
    conv = edge_conv_filter.get_gcn_filter(16, 32).to(DEV)
    x = torch.randn(n, 16, device=DEV)
    outs = []
    for _ in range(2):
        xd = x.clone().requires_grad_(True)
        conv.zero_grad()
        o = conv(xd, ei.to(DEV))
        o.square().sum().backward()
        outs.append((o.detach().clone(), xd.grad.clone(), [p.grad.clone() for p in conv.parameters()]))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    for a, b in zip(outs[0][2], outs[1][2]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("kind", ["ico", "graph18_isolated", "dilated_asym", "multi_edges", "empty"])
@pytest.mark.parametrize("h", [4, 64, 128, 320])
def test_edge_message_mask_kernels_equal_recomputing_kernels(kind, h):
    """Forward with saved ReLU decision masks + mask-reading backward kernels vs the kernels that re-evaluate P_i + Q_j:
    bit-identical hid / dP / dQ (same products, same summation order); widths below a warp, exactly a warp (128) and
    with a partial third chunk (320), rows that are empty, long (hub) or carry duplicate edges."""
    from stinet_b200 import _abi
    from stinet_b200.graph import EdgeCSR, _stream
    ei, n = _graph(kind)
    csr = EdgeCSR(ei.to(DEV), n)
    rs, cs, es = csr.by_source()
    g = torch.Generator().manual_seed(17 + h)
    pq = torch.randn(n, 2 * h, generator=g).to(DEV)
    pq[:, :h].mul_(0.5)
    if n > 3:
        pq[1, :h] = -pq[2, h:]                      # exact zeros of P_i + Q_j: the decision is `> 0`, ties are off
    dh = torch.randn(n, h, generator=g).to(DEV)
    P, Q, ld = pq.data_ptr(), pq.data_ptr() + 4 * h, 2 * h
    hid0, hid1 = torch.empty(n, h, device=DEV), torch.full((n, h), float("nan"), device=DEV)
    d0, d1 = torch.empty(n, 2 * h, device=DEV), torch.full((n, 2 * h), float("nan"), device=DEV)
    mask = torch.zeros(max(csr.e, 1), h // 4, dtype=torch.uint8, device=DEV)
    s = _stream()
    _abi.call("stinet_edge_message_fwd", P, ld, Q, ld, csr.rowptr_t.data_ptr(), csr.col_t.data_ptr(), n, h,
              hid0.data_ptr(), h, s)
    _abi.call("stinet_edge_message_fwd_mask", P, ld, Q, ld, csr.rowptr_t.data_ptr(), csr.col_t.data_ptr(), n, h,
              hid1.data_ptr(), h, mask.data_ptr(), s)
    assert torch.equal(hid0, hid1)
    # the masks are exactly the decisions, in by-target CSR order
    tpos = csr.tpos_s()
    if csr.e:
        perm = csr.eid_t.long()
        assert torch.equal(perm[tpos[:csr.e].long()], es.long())                   # tpos_s: by-source entry -> by-target slot
        dec = (pq[ei[1].to(DEV)[perm], :h] + pq[ei[0].to(DEV)[perm], h:]) > 0      # [E, h]
        c = torch.arange(h, device=DEV)
        bits = (mask[:csr.e, c // 4].int() >> (c % 4)) & 1
        assert torch.equal(bits.bool(), dec)
    _abi.call("stinet_edge_message_bwd_target", P, ld, Q, ld, dh.data_ptr(), h, csr.rowptr_t.data_ptr(),
              csr.col_t.data_ptr(), n, h, d0.data_ptr(), 2 * h, s)
    _abi.call("stinet_edge_message_bwd_source", P, ld, Q, ld, dh.data_ptr(), h, csr.rowptr_t.data_ptr(), rs.data_ptr(),
              cs.data_ptr(), n, h, d0.data_ptr() + 4 * h, 2 * h, s)
    _abi.call("stinet_edge_message_bwd_target_mask", dh.data_ptr(), h, csr.rowptr_t.data_ptr(), mask.data_ptr(), n, h,
              d1.data_ptr(), 2 * h, s)
    _abi.call("stinet_edge_message_bwd_source_mask", dh.data_ptr(), h, csr.rowptr_t.data_ptr(), rs.data_ptr(),
              cs.data_ptr(), tpos.data_ptr(), mask.data_ptr(), n, h, d1.data_ptr() + 4 * h, 2 * h, s)
    assert torch.equal(d0, d1)


@pytest.mark.parametrize("kind", ["ico", "graph18_isolated", "dilated_asym", "multi_edges", "empty"])
@pytest.mark.parametrize("h", [4, 64, 128, 320, 2048])
def test_edge_message_plane_kernels_equal_the_fp32_kernels(kind, h):
    """The message-stage kernels that write fp16 operand planes (hid in forward, dPQ in backward, scales from the
    bounds 2 max|PQ| and max|dhid| * dq_factor) against the fp32 mask kernels: the planes reproduce the fp32 values to
    22 significand bits of the bound, the decision masks are identical, the bounds hold, and the bias gradient that the
    target kernel accumulates on the side equals the column sums of dP."""
    from stinet_b200 import _abi, ops
    from stinet_b200.graph import EdgeCSR, _stream
    ei, n = _graph(kind)
    csr = EdgeCSR(ei.to(DEV), n)
    rs, cs, es = csr.by_source()
    g = torch.Generator().manual_seed(23 + h)
    pq = (torch.randn(n, 2 * h, generator=g) * 3).to(DEV)
    dh = (torch.randn(n, h, generator=g) * 1e-6).to(DEV)                      # gradient-sized values: the scale matters
    P, Q, ld = pq.data_ptr(), pq.data_ptr() + 4 * h, 2 * h
    s = _stream()
    hid = torch.empty(n, h, device=DEV)
    mask0 = torch.zeros(max(csr.e, 1), h // 4, dtype=torch.uint8, device=DEV)
    mask1 = torch.zeros_like(mask0)
    _abi.call("stinet_edge_message_fwd_mask", P, ld, Q, ld, csr.rowptr_t.data_ptr(), csr.col_t.data_ptr(), n, h,
              hid.data_ptr(), h, mask0.data_ptr(), s)
    pq_amax = pq.abs().max().reshape(1)
    hp = ops._new_planes(n, h, True, DEV)
    _abi.call("stinet_edge_message_fwd_planes", P, ld, Q, ld, csr.rowptr_t.data_ptr(), csr.col_t.data_ptr(), n, h,
              pq_amax.data_ptr(), hp.hi.data_ptr(), hp.lo.data_ptr(), hp.ld, hp.exp.data_ptr(), mask1.data_ptr(), s)
    assert torch.equal(mask0, mask1)

    def value(p, cols):
        return (p.hi[:, :cols].double() + p.lo[:, :cols].double() / 2048.0) * 2.0 ** int(p.exp.item())

    bound = 2.0 * float(pq_amax)
    assert float(hid.abs().max()) <= bound
    assert float((value(hp, h) - hid.double()).abs().max()) <= bound * 2.0 ** -21
    # backward
    d_ref = torch.empty(n, 2 * h, device=DEV)
    tpos = csr.tpos_s()
    _abi.call("stinet_edge_message_bwd_target_mask", dh.data_ptr(), h, csr.rowptr_t.data_ptr(), mask0.data_ptr(), n, h,
              d_ref.data_ptr(), 2 * h, s)
    _abi.call("stinet_edge_message_bwd_source_mask", dh.data_ptr(), h, csr.rowptr_t.data_ptr(), rs.data_ptr(),
              cs.data_ptr(), tpos.data_ptr(), mask0.data_ptr(), n, h, d_ref.data_ptr() + 4 * h, 2 * h, s)
    dh_amax = dh.abs().max().reshape(1)
    dp = ops._new_planes(n, 2 * h, True, DEV)
    db = torch.full((h,), float("nan"), device=DEV)
    nb = _abi.query("stinet_edge_message_bwd_workspace_bytes", n, h)
    ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=DEV)
    dh1, dh2 = dh.clone(), dh.clone()                # the call consumes dhid (rows are rescaled by 1 / deg in place)
    _abi.call("stinet_edge_message_bwd_planes", dh1.data_ptr(), h, dh_amax.data_ptr(), csr.dq_factor().data_ptr(),
              csr.rowptr_t.data_ptr(), rs.data_ptr(), cs.data_ptr(), tpos.data_ptr(), mask0.data_ptr(), n, h,
              dp.hi.data_ptr(), dp.lo.data_ptr(), dp.ld, dp.exp.data_ptr(), db.data_ptr(), ws.data_ptr(), nb, s)
    dbound = float(dh_amax) * max(1.0, float(csr.dq_factor()))
    assert float(d_ref.abs().max()) <= dbound * (1 + 1e-6)
    assert float((value(dp, 2 * h) - d_ref.double()).abs().max()) <= dbound * 2.0 ** -21
    ref_db = d_ref[:, :h].double().sum(0)
    assert float((db.double() - ref_db).abs().max()) <= 1e-5 * max(float(ref_db.abs().max()), 1e-30) + 1e-12 * dbound
    # without the bias gradient (no workspace): same planes
    dp2 = ops._new_planes(n, 2 * h, True, DEV)
    _abi.call("stinet_edge_message_bwd_planes", dh2.data_ptr(), h, dh_amax.data_ptr(), csr.dq_factor().data_ptr(),
              csr.rowptr_t.data_ptr(), rs.data_ptr(), cs.data_ptr(), tpos.data_ptr(), mask0.data_ptr(), n, h,
              dp2.hi.data_ptr(), dp2.lo.data_ptr(), dp2.ld, dp2.exp.data_ptr(), None, None, 0, s)
    assert torch.equal(dp.hi, dp2.hi) and torch.equal(dp.lo, dp2.lo) and torch.equal(dp.exp, dp2.exp)


# ------------------------------------------------------------------------------------------------------------------
# affine segmented norms and the in-place skip concatenation of the unpool kernel (row a8)


@pytest.mark.parametrize("n,c", [(1000, 16), (4097, 64), (333, 10), (20000, 128)])
def test_batch_norm_kernels_match_torch(n, c):
    """ops.BatchNorm1d (stinet_affnorm_*, kind 0) against nn.BatchNorm1d in fp64: output, gradients of x / weight / bias,
    running statistics after the step (unbiased variance), then eval mode on the updated statistics."""
    from stinet_b200 import ops
    g = torch.Generator().manual_seed(n + c)
    x = torch.randn(n, c, generator=g) * 3 + torch.randn(c, generator=g)
    go = torch.randn(n, c, generator=g)
    ref = torch.nn.BatchNorm1d(c).double()
    with torch.no_grad():
        ref.weight.copy_(torch.randn(c, generator=g))
        ref.bias.copy_(torch.randn(c, generator=g))
    bn = ops.BatchNorm1d(c)
    bn.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in ref.state_dict().items()})
    bn = bn.to(DEV)
    xr = x.double().requires_grad_(True)
    yr = ref(xr)
    yr.backward(go.double())
    xd = x.to(DEV).requires_grad_(True)
    yd = bn(xd)
    yd.backward(go.to(DEV))
    assert rel_err(yd, yr) <= TOL
    assert rel_err(xd.grad, xr.grad) <= TOL
    assert rel_err(bn.weight.grad, ref.weight.grad) <= TOL and rel_err(bn.bias.grad, ref.bias.grad) <= TOL
    assert rel_err(bn.running_mean, ref.running_mean) <= TOL and rel_err(bn.running_var, ref.running_var) <= TOL
    assert int(bn.num_batches_tracked) == int(ref.num_batches_tracked) == 1
    ref.eval(), bn.eval()
    with torch.no_grad():
        assert rel_err(bn(xd), ref(xr)) <= TOL
    # run-to-run bit determinism (no atomics anywhere in the reductions)
    bn.train()
    assert torch.equal(bn(xd), bn(xd))


@pytest.mark.parametrize("counts", [None, [500, 500, 500], [64, 64]])
@pytest.mark.parametrize("c", [8, 40, 130])
def test_graph_norm_kernels_match_the_reference_formula(counts, c):
    """SingleBatchGraphNorm on the kernels (stinet_affnorm_*, kind 1) against the reference's formula evaluated in fp64
    (models/modules/singlebatchgroupnorm.py:44-71, incl. the un-shifted second moment): output and the gradients of x,
    weight, bias and mean_scale."""
    from stinet_b200.models.modules.singlebatchgroupnorm import SingleBatchGraphNorm
    n = sum(counts) if counts else 777
    g = torch.Generator().manual_seed(n + c)
    x = torch.randn(n, c, generator=g) * 2 + 0.7
    go = torch.randn(n, c, generator=g)
    batch = None if counts is None else torch.repeat_interleave(torch.arange(len(counts)), torch.tensor(counts))
    mod = SingleBatchGraphNorm(c)
    with torch.no_grad():
        for p in mod.parameters():
            p.copy_(torch.randn(c, generator=g))
    w, b, ms = (p.detach().double().requires_grad_(True) for p in (mod.weight, mod.bias, mod.mean_scale))
    xr = x.double().requires_grad_(True)
    bsz = 1 if counts is None else len(counts)
    ptr = torch.linspace(0, n, bsz + 1, dtype=torch.int)
    bvec = torch.zeros(n, dtype=torch.long) if batch is None else batch
    mean = torch.stack([xr[ptr[i]:ptr[i + 1]].mean(0) for i in range(bsz)]).index_select(0, bvec)
    var = torch.stack([xr[ptr[i]:ptr[i + 1]].pow(2).mean(0) for i in range(bsz)])
    yr = w * (xr - mean * ms) / (var + mod.eps).sqrt().index_select(0, bvec) + b
    yr.backward(go.double())
    mod = mod.to(DEV)
    xd = x.to(DEV).requires_grad_(True)
    yd = mod(xd, None if batch is None else batch.to(DEV))
    yd.backward(go.to(DEV))
    assert rel_err(yd, yr) <= TOL and rel_err(xd.grad, xr.grad) <= TOL
    assert rel_err(mod.weight.grad, w.grad) <= TOL and rel_err(mod.bias.grad, b.grad) <= TOL
    assert rel_err(mod.mean_scale.grad, ms.grad) <= TOL


@pytest.mark.parametrize("ca,cb", [(16, 32), (64, 64), (12, 20), (3, 5)])
def test_unpool_concat_writes_the_skip_concatenation_in_place(ca, cb):
    """[skip || coarse[trace]] through the `ldo` column-slice form of the gather kernel: bit-identical to torch.cat of the
    two halves, forward and backward."""
    from stinet_b200 import ops
    from stinet_b200.graph import ClusterCSR
    g = torch.Generator().manual_seed(ca * 100 + cb)
    n_f, n_c = 5000, 1300
    trace = torch.randint(0, n_c, (n_f,), generator=g)
    skip, xc = torch.randn(n_f, ca, generator=g), torch.randn(n_c, cb, generator=g)
    go = torch.randn(n_f, ca + cb, generator=g)
    sr, xr = skip.clone().requires_grad_(True), xc.clone().requires_grad_(True)
    ref = torch.cat((sr, xr.index_select(0, trace)), -1)
    ref.backward(go)
    cl = ClusterCSR(trace.to(DEV), n_c)
    sd, xd = skip.to(DEV).requires_grad_(True), xc.to(DEV).requires_grad_(True)
    out = ops.unpool_concat(sd, xd, cl)
    out.backward(go.to(DEV))
    assert torch.equal(out.cpu(), ref)
    assert torch.equal(sd.grad.cpu(), sr.grad)
    assert rel_err(xd.grad, xr.grad) <= TOL         # segmented sum in member order vs index_add order: same values up to rounding


@pytest.mark.parametrize("n,c", [(1000, 64), (4099, 16), (77, 256), (20000, 32)])
def test_head_tanh_kernel(n, c):
    """tanh(h W^T + b) with the 3-wide Linear in registers (stinet_head_*) against torch in fp64: output, dh, dW, db."""
    from stinet_b200 import ops
    g = torch.Generator().manual_seed(n + c)
    h = torch.randn(n, c, generator=g)
    w, b = torch.randn(3, c, generator=g) / c ** 0.5, torch.randn(3, generator=g)
    go = torch.randn(n, 3, generator=g)
    hr, wr, br = (t.double().requires_grad_(True) for t in (h, w, b))
    ref = torch.tanh(hr @ wr.t() + br)
    ref.backward(go.double())
    hd, wd, bd = (t.to(DEV).requires_grad_(True) for t in (h, w, b))
    out = ops.head_tanh(hd, wd, bd)
    assert out is not None
    out.backward(go.to(DEV))
    assert rel_err(out, ref) <= TOL and rel_err(hd.grad, hr.grad) <= TOL
    assert rel_err(wd.grad, wr.grad) <= TOL and rel_err(bd.grad, br.grad) <= TOL


@pytest.mark.parametrize("n", [1000, 40962])
def test_masked_l1_loss_kernel(n):
    """The trainer's loss (trainers/inpainting3d_trainer.py:127-137) as one kernel each way against the torch chain."""
    from stinet_b200 import ops
    g = torch.Generator().manual_seed(n)
    out = torch.tanh(torch.randn(n, 3, generator=g))
    color = torch.rand(n, 3, generator=g) * 2 - 1
    mask = torch.randint(0, 12, (n, 1), generator=g).float() * (torch.rand(n, 1, generator=g) > 0.4)
    out[:5] = color[:5]                                                      # exact ties: sign(0) = 0 on both sides
    orf = out.double().requires_grad_(True)
    composed = torch.where((mask > 0).expand_as(color), orf, color.double())
    ref = ((composed - color.double()).abs() * torch.pow(0.99, mask.squeeze().double()).unsqueeze(1)).mean()
    (ref * 3.0).backward()
    od = out.to(DEV).requires_grad_(True)
    loss = ops.masked_l1_loss(od, color.to(DEV), mask.to(DEV))
    (loss * 3.0).backward()
    assert rel_err(loss, ref) <= TOL and rel_err(od.grad, orf.grad) <= TOL
    assert bool((od.grad[(mask.squeeze() <= 0).to(DEV)] == 0).all())
