"""Device-side graph structure for one batch: CSR by target / by source for every edge set, cluster CSR for every
trace map, per-level graph ids and norm segments.  Built once per batch by `stinet_csr_build` (no torch sort, no
PyG), cached on the sample object, consumed by every kernel of the forward and backward pass.

Replaces the COO `edge_index` / `trace` tensors that the reference hands to PyG's MessagePassing.propagate and to
torch_scatter on every layer (reference models/surfacetextureinpaintingnet.py:398-471).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import os

import torch

from . import _abi


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise _abi.StinetError(f"{what} must live on a CUDA device (got {t.device}); stinet_b200 has no CPU path")


_SIDE: Dict[tuple, "torch.cuda.Stream"] = {}


def _side_stream(dev: torch.device) -> "torch.cuda.Stream":
    """The stream the structure of a batch is built on while the first layers already run (GraphCache.build_ahead)."""
    k = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if k not in _SIDE:
        _SIDE[k] = torch.cuda.Stream(device=dev)
    return _SIDE[k]


def _cuda_tensors(obj):
    for v in vars(obj).values():
        for t in (v if isinstance(v, (tuple, list)) else (v,)):
            if torch.is_tensor(t) and t.is_cuda:
                yield t


def build_csr(key: torch.Tensor, other: Optional[torch.Tensor], n_rows: int, want_key32: bool = False,
              status: Optional[torch.Tensor] = None):
    """rowptr, perm, col, key32 (all int32) for positions grouped stably by `key`."""
    _require_cuda(key, "index tensor")
    assert key.dtype == torch.int64 and key.is_contiguous()
    n_items = key.numel()
    dev = key.device
    rowptr = torch.empty(n_rows + 1, dtype=torch.int32, device=dev)
    perm = torch.empty(n_items, dtype=torch.int32, device=dev)
    col = torch.empty(n_items, dtype=torch.int32, device=dev) if other is not None else None
    key32 = torch.empty(n_items, dtype=torch.int32, device=dev) if want_key32 else None
    ws_bytes = _abi.query("stinet_csr_workspace_bytes", n_rows, n_items)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    if other is not None:
        assert other.dtype == torch.int64 and other.is_contiguous() and other.numel() == n_items
    _abi.call("stinet_csr_build", key.data_ptr(), _ptr(other), n_items, n_rows, rowptr.data_ptr(), perm.data_ptr(),
              _ptr(col), _ptr(key32), _ptr(status), ws.data_ptr(), ws_bytes, _stream(),
              cost=(n_items * (16 + 8) + n_rows * 4, 0, ""))
    return rowptr, perm, col, key32


class EdgeCSR:
    """One directed edge set (`edge_index[0]` = source j, `edge_index[1]` = target i) over n vertices."""

    def __init__(self, edge_index: torch.Tensor, n: int, status: Optional[torch.Tensor] = None):
        assert edge_index.dim() == 2 and edge_index.size(0) == 2
        self.n, self.e = int(n), int(edge_index.size(1))
        self._src = edge_index[0].contiguous()
        self._dst = edge_index[1].contiguous()
        self._status = status
        # by target: row i lists its in-edges in original order; col_t = source vertex, eid_t = original edge id
        self.rowptr_t, self.eid_t, self.col_t, _ = build_csr(self._dst, self._src, self.n, status=status)
        self._by_source = None
        self._tpos_s = None

    _bwd_event = None     # set by GraphCache.build_ahead: the backward-only arrays were built on the side stream

    def _wait_bwd(self) -> None:
        ev = self._bwd_event
        if ev is not None and torch.cuda.current_stream(self.rowptr_t.device) != _side_stream(self.rowptr_t.device):
            torch.cuda.current_stream(self.rowptr_t.device).wait_event(ev)
            self._bwd_event = None

    def build_backward_parts(self, fused: bool) -> None:
        """Everything only a backward pass reads: the by-source CSR and, for the fused EdgeConv node, the positions of
        the saved ReLU masks and the dQ bound."""
        self.by_source()
        if fused:
            self.tpos_s()
            self.dq_factor()

    @classmethod
    def from_arrays(cls, n: int, rowptr_t, col_t, eid_t, rowptr_s, col_s, eid_s, tpos_s=None) -> "EdgeCSR":
        """Wrap arrays that already exist (stinet_b200.structure: built once per sample, concatenated per batch)."""
        self = cls.__new__(cls)
        self.n, self.e = int(n), int(col_t.numel())
        self._src = self._dst = self._status = None
        self.rowptr_t, self.col_t, self.eid_t = rowptr_t, col_t, eid_t
        self._by_source = (rowptr_s, col_s, eid_s)
        self._tpos_s = tpos_s
        return self

    def edge_clusters(self):
        """(by_target, by_source): the EDGES (original order) as the fine side of two cluster maps onto their end points --
        lets the row-gather / segmented-add kernels of pooling serve the literal per-edge message path
        (x_i = unpool(x, by_target), x_j = unpool(x, by_source), mean over in-edges = pool_mean(msg, by_target))."""
        if getattr(self, "_edge_clusters", None) is None:
            if self._dst is None:
                raise _abi.StinetError("the literal per-edge path needs an EdgeCSR built from edge_index (COO), not from "
                                       "prebuilt arrays")
            rowptr_s, _, eid_s = self.by_source()
            self._edge_clusters = (
                ClusterCSR.from_arrays(self.e, self.n, self.rowptr_t, self.eid_t, self._dst.to(torch.int32)),
                ClusterCSR.from_arrays(self.e, self.n, rowptr_s, eid_s, self._src.to(torch.int32)))
        return self._edge_clusters

    def tpos_s(self) -> torch.Tensor:
        """By-target position of every by-source entry (where the saved ReLU masks of an out-edge live)."""
        self._wait_bwd()
        if getattr(self, "_tpos_s", None) is None:
            _, _, eid_s = self.by_source()
            out = torch.empty(max(self.e, 1), dtype=torch.int32, device=self.rowptr_t.device)
            scratch = torch.empty(max(self.e, 1), dtype=torch.int32, device=self.rowptr_t.device)
            _abi.call("stinet_csr_cross_positions", self.eid_t.data_ptr(), eid_s.data_ptr(), self.e, out.data_ptr(),
                      scratch.data_ptr(), _stream(), cost=(16 * self.e, 0, ""))
            self._tpos_s = out
        return self._tpos_s

    def by_source(self):
        """rowptr_s, col_s (= target vertex of each out-edge), eid_s -- only needed by backward passes."""
        self._wait_bwd()
        if self._by_source is None:
            rowptr_s, eid_s, col_s, _ = build_csr(self._src, self._dst, self.n, status=self._status)
            self._by_source = (rowptr_s, col_s, eid_s)
        return self._by_source

    def degree_order(self) -> torch.Tensor:
        """int32 [n]: the rows by DESCENDING in-degree, ties in ascending row order (stinet_csr_degree_order).  A row
        schedule for kernels that map several rows onto one warp; the warp-per-row message kernels of this package keep
        the natural order, which preserves the gather locality of neighbouring mesh vertices."""
        if getattr(self, "_degree_order", None) is None:
            out = torch.empty(max(self.n, 1), dtype=torch.int32, device=self.rowptr_t.device)
            nb = _abi.query("stinet_csr_degree_order_workspace_bytes", self.n)
            ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=out.device)
            _abi.call("stinet_csr_degree_order", self.rowptr_t.data_ptr(), self.n, out.data_ptr(), ws.data_ptr(), nb, _stream(),
                      cost=(self.n * 60, 0, ""))
            self._degree_order = out[:self.n]
        return self._degree_order

    def dq_factor(self) -> torch.Tensor:
        """Device float: max_j sum_{j->i} 1/deg_i -- |dQ| <= max|dhid| * dq_factor in the message-stage backward (the
        bound its fp16 output planes are scaled with)."""
        self._wait_bwd()
        if getattr(self, "_dq_factor", None) is None:
            rowptr_s, col_s, _ = self.by_source()
            out = torch.empty(1, dtype=torch.float32, device=self.rowptr_t.device)
            _abi.call("stinet_csr_dq_factor", self.rowptr_t.data_ptr(), rowptr_s.data_ptr(), col_s.data_ptr(), self.n,
                      out.data_ptr(), _stream(), cost=(8 * self.e + 8 * self.n, 0, ""))
            self._dq_factor = out
        return self._dq_factor

    @property
    def degree(self) -> torch.Tensor:
        """int32 in-degree per vertex (rowmask for the `isolated vertex -> 0` rule)."""
        if not hasattr(self, "_deg"):
            self._deg = (self.rowptr_t[1:] - self.rowptr_t[:-1]).contiguous()
        return self._deg


class ClusterCSR:
    """One trace map: fine vertex i belongs to coarse vertex trace[i]."""

    def __init__(self, trace: torch.Tensor, n_coarse: int, status: Optional[torch.Tensor] = None):
        self.n_fine, self.n_coarse = int(trace.numel()), int(n_coarse)
        self.rowptr, self.member, _, self.trace32 = build_csr(trace.contiguous(), None, self.n_coarse,
                                                             want_key32=True, status=status)

    @classmethod
    def from_arrays(cls, n_fine: int, n_coarse: int, rowptr, member, trace32) -> "ClusterCSR":
        self = cls.__new__(cls)
        self.n_fine, self.n_coarse = int(n_fine), int(n_coarse)
        self.rowptr, self.member, self.trace32 = rowptr, member, trace32
        return self


_SEG_TABLES: Dict[tuple, tuple] = {}


def _segment_tables(slice_ptr: tuple, cnt: tuple, device: torch.device):
    """Device copies of the (host-derived) slice boundaries and divisors, memoised per (sizes, device): batches of
    the same shape re-use them, so a step that is being captured into a CUDA graph performs no host-to-device copy."""
    key = (slice_ptr, cnt, device)
    hit = _SEG_TABLES.get(key)
    if hit is None:
        if len(_SEG_TABLES) > 4096:
            _SEG_TABLES.clear()
        hit = (torch.tensor(slice_ptr, dtype=torch.int32).to(device), torch.tensor(cnt, dtype=torch.float32).to(device))
        _SEG_TABLES[key] = hit
    return hit


class Segments:
    """Row partition used by the per-graph norm at one level (reference fastinstancenorm.py:51-60):
    slice_ptr = linspace(0, N, B+1) (int), cnt = true per-graph vertex counts clamped to >= 1, gid = graph id per row.
    `consistent` is True when the linspace slices coincide with the true graph boundaries."""

    def __init__(self, n_rows: int, counts: Optional[List[int]], gid: Optional[torch.Tensor], device):
        self.n_rows = int(n_rows)
        if counts is None or len(counts) <= 1:
            self.n_seg = 1
            slice_ptr = [0, self.n_rows]
            cnt = [max(self.n_rows, 1)]
            self.gid = None
            self.consistent = True
            self.true_ptr = [0, self.n_rows]
        else:
            self.n_seg = len(counts)
            # same call as the reference (fastinstancenorm.py:53), evaluated on the host
            slice_ptr = torch.linspace(0, self.n_rows, self.n_seg + 1, dtype=torch.int).tolist()
            cnt = [max(int(c), 1) for c in counts]
            true_ptr = [0]
            for c in counts:
                true_ptr.append(true_ptr[-1] + int(c))
            self.consistent = (true_ptr == slice_ptr)
            self.true_ptr = true_ptr
            self.gid = gid
        self.slice_ptr_host = list(slice_ptr)
        self.max_seg_rows = max(b - a for a, b in zip(slice_ptr[:-1], slice_ptr[1:]))
        self.slice_ptr, self.cnt = _segment_tables(tuple(slice_ptr), tuple(cnt), torch.device(device))


class GraphCache:
    """Everything structural the forward/backward of one batch needs.  `for_sample` memoises on the sample."""

    def __init__(self, sample, n_levels: int):
        x = sample.x
        _require_cuda(x, "sample.x")
        dev = x.device
        self.device = dev
        self.n_levels = n_levels
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        pre = getattr(sample, "_nv_host", None)    # host copy supplied by the caller (engine.GraphedTrainStep)
        if pre is not None:
            self.per_graph = [list(map(int, g)) for g in pre]
        else:
            nv_host = sample.num_vertices.detach().to("cpu")   # the one host read of the step: [B, L+1] (or [L+1]) ints
            if nv_host.dim() == 1:
                nv_host = nv_host.unsqueeze(0)
            self.per_graph = nv_host.tolist()      # [B][L+1]
        self.batch_size = len(self.per_graph)
        self.totals = [sum(g[l] for g in self.per_graph) for l in range(len(self.per_graph[0]))]
        assert self.totals[0] == x.shape[0], "num_vertices[:,0] must sum to x.shape[0]"
        self._sample = sample
        self._edges: Dict[str, EdgeCSR] = {}
        self._clusters: Dict[int, ClusterCSR] = {}
        self._segments: Dict[tuple, Segments] = {}
        self._gid: Dict[int, torch.Tensor] = {}
        self._pending: Dict[tuple, "torch.cuda.Event"] = {}      # structures being built on the side stream
        self._ahead = False
        self._forked = False

    # -- building the structure next to the first layers --------------------------------------------------------
    def build_ahead(self, plan, backward: bool, fused: bool) -> None:
        """plan: [("edges", key, level) | ("cluster", level) | ("gid", level)] in the order the forward first touches them.
        The first entry is built on the current stream (the first layer needs it now); everything else, and the
        backward-only arrays of every edge set, on a side stream that the current stream joins item by item when the item
        is first asked for -- dozens of short launches that would otherwise sit in front of the first GEMM.  Works the
        same eagerly and under CUDA-graph capture (the side stream forks from and rejoins the capturing stream).
        STINET_STRUCT_SIDE_STREAM=0 keeps the lazy single-stream construction."""
        if self._ahead:
            return
        self._ahead = True
        if not plan or os.environ.get("STINET_STRUCT_SIDE_STREAM", "1") == "0":
            return
        dev = self.device
        main = torch.cuda.current_stream(dev)
        side = _side_stream(dev)
        first = plan[0]
        built = []
        if first[0] == "edges":
            built.append(self.edges(first[1], first[2]))
        side.wait_stream(main)
        self._forked = True
        objs = []
        with torch.cuda.stream(side):
            for item in plan[1:]:
                k = tuple(item[:2])
                if k in self._pending:
                    continue
                if item[0] == "edges":
                    if item[1] in self._edges:
                        continue
                    o = self.edges(item[1], item[2])
                    built.append(o)
                elif item[0] == "cluster":
                    if item[1] in self._clusters:
                        continue
                    o = self.cluster(item[1])
                else:
                    if item[1] in self._gid or self.batch_size <= 1:
                        continue
                    o = self.graph_id(item[1])
                objs.append(o)
                ev = torch.cuda.Event()
                ev.record(side)
                self._pending[k] = ev
            if backward:
                for e in built:
                    if e._bwd_event is None and e._dst is not None:
                        e.build_backward_parts(fused)
                        ev = torch.cuda.Event()
                        ev.record(side)
                        e._bwd_event = ev
        for o in objs + built:                          # allocated on the side stream, read on this one
            for t in ([o] if torch.is_tensor(o) else _cuda_tensors(o)):
                t.record_stream(main)

    def join_ahead(self) -> None:
        """The current stream waits for whatever build_ahead still has in flight (end of the forward)."""
        if self._forked:
            torch.cuda.current_stream(self.device).wait_stream(_side_stream(self.device))
            self._forked = False
            self._pending.clear()
            for e in self._edges.values():
                e._bwd_event = None

    def _wait(self, k: tuple) -> None:
        if k in self._pending and torch.cuda.current_stream(self.device) != _side_stream(self.device):
            torch.cuda.current_stream(self.device).wait_event(self._pending.pop(k))

    @staticmethod
    def for_sample(sample, n_levels: int) -> "GraphCache":
        c = getattr(sample, "_stinet_cache", None)
        if c is None or c.n_levels != n_levels or c.device != sample.x.device:
            c = GraphCache(sample, n_levels)
            try:
                object.__setattr__(sample, "_stinet_cache", c)
            except Exception:
                pass
        return c

    # -- structure ------------------------------------------------------------------------------------------
    def edges(self, key: str, level: int) -> EdgeCSR:
        self._wait(("edges", key))
        if key not in self._edges:
            from .structure import prebuilt_edges
            pre = prebuilt_edges(self._sample, key, self.totals[level])     # cached per sample, batched by concatenation
            if pre is None:
                ei = self._sample.edge_index if key == "edge_index" else self._sample[key]
                pre = EdgeCSR(ei, self.totals[level], self.status)
            self._edges[key] = pre
        return self._edges[key]

    def cluster(self, level: int) -> ClusterCSR:
        """trace map level-1 -> level."""
        self._wait(("cluster", level))
        if level not in self._clusters:
            from .structure import prebuilt_cluster
            pre = prebuilt_cluster(self._sample, level, self.totals[level - 1], self.totals[level])
            if pre is None:
                tr = self._sample[f"hierarchy_trace_index_{level}"]
                assert tr.numel() == self.totals[level - 1]
                pre = ClusterCSR(tr, self.totals[level], self.status)
            self._clusters[level] = pre
        return self._clusters[level]

    def graph_id(self, level: int) -> Optional[torch.Tensor]:
        """int32 graph id per vertex of `level`: level 0 from sample.batch, level l by max-pooling through the trace
        (reference :422 `batch = scatter_max(batch, trace)`)."""
        if self.batch_size <= 1:
            return None
        self._wait(("gid", level))
        if level not in self._gid:
            if level == 0:
                self._gid[0] = self._sample.batch.to(torch.int32).contiguous()
            else:
                fine = self.graph_id(level - 1)
                cl = self.cluster(level)
                out = torch.empty(cl.n_coarse, dtype=torch.int32, device=self.device)
                _abi.call("stinet_pool_max_i32", fine.data_ptr(), cl.rowptr.data_ptr(), cl.member.data_ptr(),
                          cl.n_coarse, out.data_ptr(), _stream())
                self._gid[level] = out
        return self._gid[level]

    def segments(self, level: int, per_graph: bool) -> Segments:
        """per_graph=False: the whole batch is one instance (reference input/output blocks, :406-407, :459-460)."""
        k = (level, per_graph and self.batch_size > 1)
        if k not in self._segments:
            if k[1]:
                counts = [g[level] for g in self.per_graph]
                self._segments[k] = Segments(self.totals[level], counts, self.graph_id(level), self.device)
            else:
                self._segments[k] = Segments(self.totals[level], None, None, self.device)
        return self._segments[k]

    def check_status(self) -> None:
        """Synchronising validation of data-dependent errors (index out of range). Call from tests / debug."""
        if int(self.status.item()) != 0:
            raise _abi.StinetError("index tensor out of range for its level (status bit 0)")
