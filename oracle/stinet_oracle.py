"""CPU oracle for the STINet hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file.
The product package (surface-texture-inpainting-net_b200/stinet_b200) never does, and it has no CPU fallback.

What it is: a plain-torch, fp32/fp64, CPU restatement of the reference's algorithm for the multi-level mesh U-Net
forward (+ autograd backward), written LITERALLY the way the reference evaluates it -- per-edge gathers, the message
MLP on [E, .] matrices, scatter by target -- i.e. deliberately NOT in the hoisted/segmented form the CUDA path uses,
so that the two are independent derivations of the same arithmetic.

Pinning status: the reference cannot run here as shipped (torch_geometric / torch_scatter / torch_sparse are not
installable).  The oracle is pinned instead against golden vectors minted by the reference's OWN model code
(models/surfacetextureinpaintingnet.py, models/modules/*, utils/data_utils.py imported unmodified from
/root/reference) running on tests/golden/pyg_shim, which restates only the third-party primitives
(PyG 2.0.x MessagePassing/EdgeConv/SAGEConv/Data collate, torch_scatter 2.0.x CPU scatter_{sum,mean,max}).
See tests/golden/make_golden.py and tests/test_oracle_golden.py.  The third-party primitives themselves remain
"parity unpinned" against real PyG/torch_scatter binaries (none exist in this image); their published semantics
are restated twice, independently (shim: amax + first-equal; here: sequential strict-'>' update).

Each function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------------
# discrete decisions (ReLU signs of the message MLP, max-pool winners)


class Decisions:
    """The network is smooth except for its discrete choices: the sign of every message-MLP pre-activation (ReLU) and
    the winner of every max-pool cluster.  Two correct evaluations in different precisions (fp32 vs fp64, CPU vs GPU,
    the reference's own CPU and CUDA paths) take the same choices everywhere except where a pre-activation lies within
    rounding distance of zero (or two cluster members within rounding distance of each other); there the derivative is
    discontinuous and isolated gradient entries legitimately differ.  To compare gradients at 1e-5 anyway, the oracle
    can be made to REPLAY the choices of the implementation under test:

        with Decisions.replay(choices) as d:    # choices: list of ("relu", bool [E, 2*dout]) / ("pool", long [n_c, C])
            out = oracle(batch)                 # in call order (input blocks, [pool, encoder block]*, bottleneck, ...)
        d.max_relu_margin, d.max_pool_margin    # how far from the discontinuity the differing choices were, relative
                                                # to the layer's largest pre-activation / feature (legitimacy check)

    and `Decisions.record()` returns the oracle's own choices.  Outside a `with` block the oracle is unchanged."""

    _active: Optional["Decisions"] = None

    def __init__(self, choices=None):
        self.choices = list(choices) if choices is not None else None
        self.recorded: List[Tuple[str, torch.Tensor]] = []
        self.pos = 0
        self.max_relu_margin = 0.0
        self.max_pool_margin = 0.0
        self.n_relu_diff = 0
        self.n_pool_diff = 0

    @classmethod
    def replay(cls, choices):
        return cls(choices)

    @classmethod
    def record(cls):
        return cls(None)

    def __enter__(self):
        Decisions._active = self
        return self

    def __exit__(self, *exc):
        Decisions._active = None

    def _next(self, kind):
        k, v = self.choices[self.pos]
        assert k == kind, f"decision stream out of step: expected {kind}, got {k} at {self.pos}"
        self.pos += 1
        return v

    def relu(self, pre: torch.Tensor) -> torch.Tensor:
        own = pre > 0
        if self.choices is None:
            self.recorded.append(("relu", own))
            return F.relu(pre)
        m = self._next("relu").to(pre.device)
        diff = m != own
        if bool(diff.any()):
            self.n_relu_diff += int(diff.sum())
            scale = float(pre.detach().abs().max())
            self.max_relu_margin = max(self.max_relu_margin, float(pre.detach().abs()[diff].max()) / max(scale, 1e-30))
        return pre * m.to(pre.dtype)

    def pool_max(self, src: torch.Tensor, index: torch.Tensor, n: int):
        out, arg = scatter_max(src, index, n)
        if self.choices is None:
            self.recorded.append(("pool", arg))
            return out, arg
        given = self._next("pool").to(torch.long)
        n_src = src.size(0)
        valid = given < n_src
        picked = torch.where(valid, src.gather(0, given.clamp(max=max(n_src - 1, 0))), torch.zeros_like(out))
        diff = given != arg
        if bool(diff.any()):
            self.n_pool_diff += int(diff.sum())
            scale = float(src.detach().abs().max())
            gap = (out.detach() - picked.detach()).abs()[diff].max()
            self.max_pool_margin = max(self.max_pool_margin, float(gap) / max(scale, 1e-30))
            # a replayed winner must still belong to the cluster it wins
            assert bool((index[given[diff & valid]] == torch.nonzero(diff & valid)[:, 0]).all()), "winner outside cluster"
        return picked, given


# ------------------------------------------------------------------------------------------------
# integer structure


def csr_by_key(key: torch.Tensor, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Stable counting sort of positions by `key` (SURVEY 8a 'CSR builder'):
    rowptr[n+1] (int32), perm[len(key)] (int32) with perm listing positions grouped by key, original order kept."""
    perm = torch.argsort(key, stable=True)
    rowptr = torch.zeros(n + 1, dtype=torch.int64)
    rowptr[1:] = torch.bincount(key, minlength=n).cumsum(0)
    return rowptr.to(torch.int32), perm.to(torch.int32)


def _rank_in_segment(index: torch.Tensor, n: int):
    """For each position: its rank among the positions sharing its index value (in position order)."""
    rowptr, perm = csr_by_key(index, n)
    perm = perm.long()
    rank_sorted = torch.arange(index.numel()) - rowptr.long()[index[perm]]
    rank = torch.empty_like(rank_sorted)
    rank[perm] = rank_sorted
    return rank


# ------------------------------------------------------------------------------------------------
# torch_scatter semantics (third-party, restated; call sites surfacetextureinpaintingnet.py:384,386,422)


def scatter_add(src: torch.Tensor, index: torch.Tensor, n: int) -> torch.Tensor:
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype)
    return out.index_add(0, index, src)


def scatter_mean(src: torch.Tensor, index: torch.Tensor, n: int) -> torch.Tensor:
    """sum / clamp(count, min=1); segments with no entry are 0."""
    cnt = torch.bincount(index, minlength=n).clamp(min=1).to(src.dtype)
    return scatter_add(src, index, n) / cnt.view([-1] + [1] * (src.dim() - 1))


class _ScatterMaxFirst(torch.autograd.Function):
    """torch_scatter.scatter_max, CPU tie-break: positions visited in order, update iff value > current (strict)
    => the FIRST position holding the maximum wins; empty segment => value 0, arg = src.size(0)."""

    @staticmethod
    def forward(ctx, src, index, n):
        n_src = src.size(0)
        feat = tuple(src.shape[1:])
        lowest = torch.finfo(src.dtype).min if src.is_floating_point() else torch.iinfo(src.dtype).min
        out = torch.full((n,) + feat, lowest, dtype=src.dtype)
        arg = torch.full((n,) + feat, n_src, dtype=torch.long)
        if n_src:
            rank = _rank_in_segment(index, n)
            for k in range(int(rank.max()) + 1):
                rows = torch.nonzero(rank == k).squeeze(1)          # at most one row per segment
                seg = index[rows]
                val = src[rows]
                better = val > out[seg]
                out[seg] = torch.where(better, val, out[seg])
                pos = rows.view([-1] + [1] * len(feat)).expand_as(val)
                arg[seg] = torch.where(better, pos, arg[seg])
        out = torch.where(arg == n_src, torch.zeros_like(out), out)
        ctx.save_for_backward(arg)
        ctx.n_src = n_src
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, g, _):
        (arg,) = ctx.saved_tensors
        buf = torch.zeros((ctx.n_src + 1,) + tuple(g.shape[1:]), dtype=g.dtype)
        buf.scatter_(0, arg, g)
        return buf[: ctx.n_src], None, None


def scatter_max(src: torch.Tensor, index: torch.Tensor, n: int):
    return _ScatterMaxFirst.apply(src, index, n)


# ------------------------------------------------------------------------------------------------
# message passing (PyG 2.0.x MessagePassing.propagate, flow source_to_target)


def aggregate(msg: torch.Tensor, target: torch.Tensor, n: int, aggr: str) -> torch.Tensor:
    if aggr == "mean":
        return scatter_mean(msg, target, n)
    if aggr in ("add", "sum"):
        return scatter_add(msg, target, n)
    if aggr == "max":
        return scatter_max(msg, target, n)[0]
    raise ValueError(aggr)


def edge_conv(x, edge_index, mlp: nn.Module, aggr: str = "mean", trans_inv: bool = False):
    """EdgeConv (edge_conv_filter.py:10-57 -> PyG edge_conv.py): nn([x_i || x_j - x_i]) per edge, reduced over the
    target edge_index[1].  trans_inv: nn(x_j - x_i) (edge_conv_translation_invariance.py:19-21)."""
    x_j = x.index_select(0, edge_index[0])
    x_i = x.index_select(0, edge_index[1])
    inp = (x_j - x_i) if trans_inv else torch.cat([x_i, x_j - x_i], dim=-1)
    dec = Decisions._active
    if dec is None:
        msg = mlp(inp)
    else:                                   # same arithmetic, the ReLU's sign choices recorded or replayed
        assert len(mlp) == 3 and isinstance(mlp[1], nn.ReLU)
        msg = mlp[2](dec.relu(mlp[0](inp)))
    return aggregate(msg, edge_index[1], x.size(0), aggr)


def sage_conv(x, edge_index, lin_l: nn.Linear, lin_r: nn.Linear, trans_inv: bool = False):
    """SAGEConv (sage_conv_filter.py:102-138 -> PyG sage_conv.py): lin_l(mean_j msg_j) + lin_r(x_i).
    trans_inv message: x_j[:,3:9] -= x_i[:,3:9] (sage_conv_filter.py:87-90)."""
    x_j = x.index_select(0, edge_index[0])
    if trans_inv:
        x_i = x.index_select(0, edge_index[1])
        x_j = torch.cat([x_j[:, :3], x_j[:, 3:9] - x_i[:, 3:9], x_j[:, 9:]], dim=1)
    out = lin_l(scatter_mean(x_j, edge_index[1], x.size(0)))
    return out + lin_r(x)


# ------------------------------------------------------------------------------------------------
# norms


def fast_instance_norm(x: torch.Tensor, batch: Optional[torch.Tensor], eps: float = 1e-5) -> torch.Tensor:
    """FastInstanceNorm.forward with affine=False, track_running_stats=False (fastinstancenorm.py:42-107).
    batch=None: F.instance_norm over all rows (:44-49).  Otherwise: sums over `linspace(0,N,B+1)` slices (:53,
    :68-69, :79-80) divided by the TRUE per-graph counts (:60), looked up through `batch` (:73, :99)."""
    if batch is None:
        return F.instance_norm(x.t().unsqueeze(0), None, None, None, None, True, 0.1, eps).squeeze(0).t()
    bsz = int(batch.max()) + 1
    ptr = torch.linspace(0, x.shape[0], bsz + 1, dtype=torch.int)
    norm = torch.bincount(batch, minlength=bsz).to(x.dtype).clamp(min=1).view(-1, 1)
    mean = torch.stack([x[ptr[i - 1]:ptr[i]].sum(dim=0) for i in range(1, len(ptr))]) / norm
    xc = x - mean.index_select(0, batch)
    var = torch.stack([xc[ptr[i - 1]:ptr[i]].pow(2).sum(dim=0) for i in range(1, len(ptr))]) / norm
    return xc / (var + eps).sqrt().index_select(0, batch)


def single_batch_graph_norm(x, batch, weight, bias, mean_scale, eps: float = 1e-5):
    """SingleBatchGraphNorm.forward (singlebatchgroupnorm.py:44-71) -- note var is E[x^2] of the UN-shifted x (:66-68)."""
    if batch is None:
        batch = torch.zeros(x.size(0), dtype=torch.long)
    bsz = int(batch.max()) + 1
    ptr = torch.linspace(0, x.shape[0], bsz + 1, dtype=torch.int)
    mean = torch.stack([x[ptr[i - 1]:ptr[i]].mean(dim=0) for i in range(1, len(ptr))]).index_select(0, batch)
    out = x - mean * mean_scale
    var = torch.stack([x[ptr[i - 1]:ptr[i]].pow(2).mean(dim=0) for i in range(1, len(ptr))])
    std = (var + eps).sqrt().index_select(0, batch)
    return weight * out / std + bias


class _Norm(nn.Module):
    def __init__(self, kind: str, channels: int):
        super().__init__()
        self.kind = kind
        if kind == "graph":
            self.weight = nn.Parameter(torch.ones(channels))
            self.bias = nn.Parameter(torch.zeros(channels))
            self.mean_scale = nn.Parameter(torch.ones(channels))
        elif kind == "batch":
            self.module = nn.BatchNorm1d(channels)

    def forward(self, x, batch=None):
        if self.kind == "instance":
            return fast_instance_norm(x, batch)
        if self.kind == "graph":
            return single_batch_graph_norm(x, batch, self.weight, self.bias, self.mean_scale)
        if self.kind == "batch":                                    # BatchNorm2Param ignores `batch` (:236-241)
            return self.module(x)
        return x                                                    # Identity (:257-263)


# ------------------------------------------------------------------------------------------------
# blocks and network (state_dict keys identical to the reference's)


class _EdgeFilter(nn.Module):
    def __init__(self, din, dout, double_input, trans_inv, aggr="mean"):
        super().__init__()
        k = 2 * din if double_input else din
        self.nn = nn.Sequential(nn.Linear(k, 2 * dout), nn.ReLU(), nn.Linear(2 * dout, dout))   # edge_conv_filter.py:46-55
        self.trans_inv, self.aggr = trans_inv, aggr

    def forward(self, x, edge_index):
        return edge_conv(x, edge_index, self.nn, self.aggr, self.trans_inv)


class _SageInner(nn.Module):
    def __init__(self, din, dout, trans_inv):
        super().__init__()
        self.lin_l = nn.Linear(din, dout, bias=True)
        self.lin_r = nn.Linear(din, dout, bias=False)
        self.trans_inv = trans_inv


class _SageFilter(nn.Module):
    def __init__(self, din, dout, trans_inv):
        super().__init__()
        self.sage1 = _SageInner(din, dout, trans_inv)               # SumSAGEConv.sage1 (sage_conv_filter.py:122-136)

    def forward(self, x, edge_index):
        return sage_conv(x, edge_index, self.sage1.lin_l, self.sage1.lin_r, self.sage1.trans_inv)


class OracleBlock(nn.Module):
    """GraphResnetBlock (surfacetextureinpaintingnet.py:474-521): shortcut(x) + ELU(norm(conv(x), batch))."""

    def __init__(self, din, dout, filter_type, norm, first=False):
        super().__init__()
        self.dim_in, self.dim_out = din, dout
        if filter_type.startswith("edgeconv"):
            ti = first and filter_type == "edgeconvtransinv"
            self.first_filter = _EdgeFilter(din, dout, double_input=not ti, trans_inv=ti)
        else:
            self.first_filter = _SageFilter(din, dout, trans_inv=first and filter_type == "sageconvtransinv")
        self.first_norm = _Norm(norm, dout)
        if din != dout:
            self.shortcut = nn.Linear(din, dout)

    def forward(self, x, edges, batch=None):
        out = F.elu(self.first_norm(self.first_filter(x, edges), batch))
        if self.dim_in != self.dim_out:
            x = self.shortcut(x)
        return x + out


class OracleSTINet(nn.Module):
    """SurfaceTextureInpaintingNet (surfacetextureinpaintingnet.py:202-471).  torch.utils.checkpoint is numerically a
    no-op and is omitted."""

    def __init__(self, input_nc, output_nc, filter_type, ngf=64, norm_type="instance", n_blocks=6, n_levels=2,
                 n_repeated_io_convs=1, pooling_type="mean", dilations=None, **_ignored):
        super().__init__()
        norm = norm_type if norm_type in ("batch", "instance", "graph") else "none"
        self.pooling_type = pooling_type
        self.dilations = list(dilations) if dilations is not None else [1] * n_blocks
        blk = lambda a, b, first=False: OracleBlock(a, b, filter_type, norm, first)
        self.input_blocks = nn.ModuleList(
            [blk(input_nc, ngf if i == n_repeated_io_convs - 1 else input_nc, first=(i == 0))
             for i in range(n_repeated_io_convs)])
        self.encoder_blocks = nn.ModuleList([blk(ngf * 2 ** i, ngf * 2 ** i * 2) for i in range(n_levels)])
        self.bottleneck_blocks = nn.ModuleList([blk(ngf * 2 ** n_levels, ngf * 2 ** n_levels) for _ in range(n_blocks)])
        self.decoder_blocks = nn.ModuleList(
            [blk(ngf * 2 ** (n_levels - i), ngf * 2 ** (n_levels - i) // 2) for i in range(n_levels)])
        self.output_blocks = nn.ModuleList([blk(ngf, ngf) for _ in range(n_repeated_io_convs)])
        self.final_linear1 = nn.Linear(ngf, ngf)
        self.final_norm1 = _Norm(norm, ngf)
        self.final_linear2 = nn.Linear(ngf, output_nc)
        for m in self.modules():                                    # init_weights (:360-374): Linear biases -> 0
            if isinstance(m, nn.Linear) and m.bias is not None:
                nn.init.zeros_(m.bias)

    def forward(self, sample, return_intermediates: bool = False):
        inter: Dict[str, torch.Tensor] = {}
        L = len(self.decoder_blocks)
        out = sample.x
        for blk in self.input_blocks:
            out = blk(out, sample.edge_index)                       # no batch => whole-batch norm (:406-407)
        total = sample.num_vertices.sum(dim=0) if sample.num_vertices.dim() > 1 else sample.num_vertices
        batch = sample.batch if int(sample.batch.max()) > 0 else None       # :416
        for i, blk in enumerate(self.encoder_blocks):
            lvl = i + 1
            trace = sample[f"hierarchy_trace_index_{lvl}"]
            n_l = int(total[lvl])
            if batch is not None:
                batch = scatter_max(batch, trace, n_l)[0]           # :422
            if self.pooling_type == "max":
                dec = Decisions._active
                out, arg = scatter_max(out, trace, n_l) if dec is None else dec.pool_max(out, trace, n_l)   # :386
                inter[f"pool_arg_{lvl}"] = arg
            else:
                out = scatter_mean(out, trace, n_l)                 # :384
            out = blk(out, sample[f"hierarchy_edge_index_{lvl}"], batch)
            inter[f"enc_{lvl}"] = out
        for i, blk in enumerate(self.bottleneck_blocks):
            d = self.dilations[i]
            key = f"hierarchy_dil_{d}_edge_index_{L}" if d > 1 else f"hierarchy_edge_index_{L}"
            edges = sample.edge_index if (L == 0 and d <= 1) else sample[key]
            out = blk(out, edges, batch)
        inter["bottleneck"] = out
        for i, blk in enumerate(self.decoder_blocks):
            lvl = i + 1
            trace = sample[f"hierarchy_trace_index_{L + 1 - lvl}"]
            out = out[trace]                                        # :391
            if batch is not None:
                batch = batch.index_select(0, trace)
            edges = sample.edge_index if lvl == L else sample[f"hierarchy_edge_index_{L - lvl}"]
            out = blk(out, edges, batch)
        for blk in self.output_blocks:
            out = blk(out, sample.edge_index)
        out = self.final_linear1(out)
        out = F.elu(self.final_norm1(out, batch=sample.batch))      # final norm always gets sample.batch (:465)
        out = torch.tanh(self.final_linear2(out))
        return (out, inter) if return_intermediates else out


# ------------------------------------------------------------------------------------------------
# trainer glue on the timed path (trainers/inpainting3d_trainer.py:127-137)


def masked_l1_loss(output: torch.Tensor, sample) -> torch.Tensor:
    """_graph_forward's torch.where(mask>0, out, color) (:127-129) followed by compute_loss (:132-137)."""
    color = sample.color
    mask = sample.mask
    composed = torch.where((mask > 0).expand_as(color), output, color)
    loss = (composed - color).abs()
    loss = loss * torch.pow(torch.tensor(0.99, dtype=loss.dtype), mask.squeeze().to(loss.dtype)).unsqueeze(1)
    return loss.mean()


# graph metrics ("next" row f1; utils/metrics/graph_metrics.py:6-38)


def graph_laplace_variance(x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
    gray = 0.299 * x[:, 0:1] + 0.587 * x[:, 1:2] + 0.114 * x[:, 2:3]
    xi = torch.cat([gray.new_ones(gray.shape[0], 1), gray], dim=1)
    prop = scatter_add(xi.index_select(0, edge_index[0]), edge_index[1], x.size(0))
    lap = prop[:, 1:] - prop[:, 0:1] * gray
    return torch.var(lap, dim=0, unbiased=False)


def graph_total_variation(x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
    h, w = x.shape
    return torch.abs(x[edge_index[0]] - x[edge_index[1]]).sum() / (h * w)


def graph_laplace(x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
    """GraphLaplaceOperator.forward (utils/metrics/graph_metrics.py:10-16): add-aggregate [1 | x_j], then
    sum_j x_j - deg_i x_i."""
    xi = torch.cat([x.new_ones(x.shape[0], 1), x], dim=1)
    prop = scatter_add(xi.index_select(0, edge_index[0]), edge_index[1], x.size(0))
    return prop[:, 1:] - prop[:, 0:1] * x


def psnr(x: torch.Tensor, y: torch.Tensor, data_range: float = 1.0, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """utils/metrics/graph_metrics.py:40-72 with convert_to_greyscale=False; `mask` restates the trainer's
    boolean-indexed call (trainers/inpainting3d_trainer.py:261-263)."""
    if mask is not None:
        sel = mask.reshape(-1) > 0
        x, y = x[sel], y[sel]
    x = x / data_range
    y = y / data_range
    mse = torch.mean((x - y) ** 2, dim=[0, 1])
    return -10 * torch.log10(mse + 1e-8)


# ------------------------------------------------------------------------------------------------
# SingleConvMeshNet (models/singleconvmeshnet.py:10-156): EdgeConv with BatchNorm1d over EDGES inside
# the message MLP (edge_conv_filter.py:34-44), mean / max trace pooling without dim_size (:112-118), unpool + skip
# concat (:140-141), checkpointed blocks (:129-131, :146-148 -- their BatchNorm buffers are updated twice per step).


def _mlp_with_norm(k: int, dout: int) -> nn.Sequential:
    """edge_conv_filter.py:34-44: Lin(k, 2*dout, no bias), BN1d, ReLU, Lin(2*dout, dout, no bias), BN1d"""
    return nn.Sequential(nn.Linear(k, 2 * dout, bias=False), nn.BatchNorm1d(2 * dout), nn.ReLU(),
                         nn.Linear(2 * dout, dout, bias=False), nn.BatchNorm1d(dout))


class _NormEdgeFilter(nn.Module):
    def __init__(self, din, dout, trans_inv=False, aggr="mean"):
        super().__init__()
        self.nn = _mlp_with_norm(din if trans_inv else 2 * din, dout)
        self.trans_inv, self.aggr = trans_inv, aggr

    def forward(self, x, edge_index):
        dec = Decisions._active
        if dec is None:
            return edge_conv(x, edge_index, self.nn, self.aggr, self.trans_inv)
        # same arithmetic with the ReLU's sign choices recorded or replayed (rows in original edge order)
        x_j, x_i = x.index_select(0, edge_index[0]), x.index_select(0, edge_index[1])
        m = (x_j - x_i) if self.trans_inv else torch.cat([x_i, x_j - x_i], dim=-1)
        for layer in self.nn:
            m = dec.relu(m) if isinstance(layer, nn.ReLU) else layer(m)
        return aggregate(m, edge_index[1], x.size(0), self.aggr)


def _relu(x):
    dec = Decisions._active
    return F.relu(x) if dec is None else dec.relu(x)


class _ResBlock(nn.Module):
    """singleconvmeshnet.py:95-110 (the in-place `+=` of :107 is written out of place: with more than one filter the
    reference's own backward raises, so only num_propagation_steps == 1 is reachable in training)."""

    def __init__(self, filters):
        super().__init__()
        self.filters = nn.ModuleList(filters)

    def forward(self, x, edge_index):
        x = _relu(self.filters[0](x, edge_index))
        for f in list(self.filters)[1:]:
            x = _relu(x + f(x, edge_index))
        return x


class OracleSingleConvMeshNet(nn.Module):
    def __init__(self, feature_number, num_propagation_steps, filter_sizes, num_classes=3, pooling_method="mean",
                 aggr="mean"):
        super().__init__()
        self._pooling_method = pooling_method
        self._graph_levels = len(filter_sizes)
        left, right = [], []
        curr = feature_number
        for level, size in enumerate(filter_sizes):
            filters = [_NormEdgeFilter(curr, size, trans_inv=(level == 0), aggr=aggr)]                # :44-51
            filters += [_NormEdgeFilter(size, size, aggr=aggr) for _ in range(num_propagation_steps - 1)]
            if level < len(filter_sizes) - 1:
                fused = size + filter_sizes[level + 1]                                               # :58
                rf = [_NormEdgeFilter(fused, size, aggr=aggr)]
                rf += [_NormEdgeFilter(size, size, aggr=aggr) for _ in range(num_propagation_steps - 1)]
                right.append(_ResBlock(rf))
            left.append(_ResBlock(filters))
            curr = size
        # registration order of the reference (:88-90): left, right, final
        self.left_geo_cnns = nn.ModuleList(left)
        self.right_geo_cnns = nn.ModuleList(right)
        f0 = filter_sizes[0]
        self.final_convs = nn.ModuleList([nn.Sequential(nn.Linear(f0, f0 // 2), nn.BatchNorm1d(f0 // 2), nn.ReLU(),
                                                        nn.Linear(f0 // 2, num_classes))])

    def _pooling(self, x, trace):
        n = int(trace.max()) + 1                                    # scatter_* without dim_size (:112-116)
        if self._pooling_method == "mean":
            return scatter_mean(x, trace, n)
        if self._pooling_method == "max":
            dec = Decisions._active
            return (scatter_max(x, trace, n) if dec is None else dec.pool_max(x, trace, n))[0]
        raise ValueError(self._pooling_method)

    def forward(self, sample, double_update_checkpointed: bool = True):
        """double_update_checkpointed: the reference wraps every block except the level-0 ones in
        torch.utils.checkpoint (reentrant), so in training their forward runs twice per step and their BatchNorm
        running statistics receive two momentum updates; here the block is simply evaluated a second time without
        grad when the module is training."""
        from torch.utils import checkpoint as _cp
        G = self._graph_levels

        def run(block, x, edges, checkpointed):
            if checkpointed and self.training and double_update_checkpointed and torch.is_grad_enabled():
                return _cp.checkpoint(block, x, edges, use_reentrant=True, preserve_rng_state=False)
            return block(x, edges)

        levels = [self.left_geo_cnns[0](sample.x, sample.edge_index)]                                 # :123
        for level in range(1, G):                                                                     # :128-134
            cur = self._pooling(levels[-1], sample[f"hierarchy_trace_index_{level}"])
            levels.append(run(self.left_geo_cnns[level], cur, sample[f"hierarchy_edge_index_{level}"], True))
        current = levels[-1]
        for level in range(1, G):                                                                     # :139-149
            back = current[sample[f"hierarchy_trace_index_{G - level}"]]
            fused = torch.cat((levels[-(level + 1)], back), -1)
            if level == G - 1:
                current = self.right_geo_cnns[-level](fused, sample.edge_index)
            else:
                current = run(self.right_geo_cnns[-level], fused, sample[f"hierarchy_edge_index_{G - level - 1}"], True)
        result = current
        for conv in self.final_convs:
            for layer in conv:
                result = _relu(result) if isinstance(layer, nn.ReLU) else layer(result)
        return result
