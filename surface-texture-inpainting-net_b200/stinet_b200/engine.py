"""Whole-step CUDA-graph execution of the STINet hot path.

One training step of the mesh U-Net is ~1000 kernel launches (graph-structure build, forward, loss, backward,
optimizer), most of them short: launched one by one from Python the step is bound by launch overhead, not by the
device.  `GraphedTrainStep` captures the whole step once per *batch shape* into a CUDA graph and replays it:

    step = GraphedTrainStep(net, loss_fn, optimizer)
    for batch in loader:                       # host (pinned) or device GraphBatch / PyG Batch
        loss = step(batch)                     # H2D copies into the static inputs + one graph launch

What is inside the graph: CSR / cluster-CSR construction from the batch's int64 `edge_index` / trace tensors (so
every replay rebuilds the structure of the *new* batch on the device), forward, `loss_fn`, backward and
`optimizer.step()`.  What must be equal between the captured batch and a replayed one is its *shape signature*:
every tensor shape plus the per-graph, per-level vertex counts `num_vertices` (they size norm segments and grids).
A batch with a new signature is captured on first use and cached; the reference's 2D trainer (equal-size image
graphs) always hits one graph, the 3D trainer (variable crops) one graph per distinct crop topology.

With more than one rank the bucketed gradient all-reduce is part of the SAME graph: NCCL collectives are capturable, the
reducer's hooks issue each bucket's all-reduce on NCCL's stream as soon as backward has produced its last gradient, and
the capture records that fork, so in a replayed step the transfers overlap the remaining backward kernels and the
optimizer waits for them inside the graph (STINET_ALLREDUCE_IN_GRAPH=0 keeps the collectives outside, between a
forward+backward graph and an optimizer graph).

Capturing is free of side effects: the eager warm-up steps (lazy initialisation of kernels, allocator and optimizer
state) run with the collectives switched off, and parameters, buffers and optimizer state are put back afterwards, so the
first batch of a new shape is trained exactly once and every rank issues the same sequence of collectives whether it hit
its graph cache or not.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import os

import torch

from . import _abi
from ._abi import StinetError
from .data import GraphBatch


def _tensor_items(batch):
    keys = batch.keys() if callable(getattr(batch, "keys", None)) else batch.keys
    for k in keys:
        v = batch[k]
        if torch.is_tensor(v):
            yield k, v


def batch_signature(batch) -> tuple:
    """Hashable shape signature of a batch; reads `num_vertices` on the host (free for host batches)."""
    shapes = tuple(sorted((k, tuple(v.shape), str(v.dtype)) for k, v in _tensor_items(batch)))
    pre = getattr(batch, "_nv_host", None)      # GraphBatch keeps a host copy across .to(device)
    if pre is not None:
        return shapes, tuple(tuple(int(x) for x in g) for g in pre)
    nv_host = batch.num_vertices.detach().to("cpu")
    if nv_host.dim() == 1:
        nv_host = nv_host.unsqueeze(0)
    return shapes, tuple(tuple(int(x) for x in g) for g in nv_host.tolist())


class _Captured:
    __slots__ = ("static", "graph_a", "graph_b", "loss", "launches", "flat", "views", "staging", "staged", "ready", "free",
                 "in_graph", "fresh")


def _flat_views(items, device):
    """One flat device buffer holding every tensor of a batch (256-byte aligned views), so that the whole batch moves
    device-to-device with a single copy."""
    layout, off = [], 0
    for k, v in items:
        nbytes = v.numel() * v.element_size()
        layout.append((k, off, nbytes, v.dtype, tuple(v.shape)))
        off += (nbytes + 255) & ~255
    flat = torch.empty(max(off, 256), dtype=torch.uint8, device=device)

    def views(buf):
        return {k: buf[o:o + n].view(dt).view(shape) for k, o, n, dt, shape in layout}
    return flat, views


class GraphedTrainStep:
    """forward + loss + backward (+ all-reduce) + optimizer step, captured per batch signature and replayed."""

    def __init__(self, net: torch.nn.Module, loss_fn: Callable, optimizer: Optional[torch.optim.Optimizer] = None,
                 reducer=None, warmup: int = 2, max_cached: int = 8):
        if optimizer is not None and not optimizer.defaults.get("capturable", False):
            raise StinetError("GraphedTrainStep needs an optimizer created with capturable=True")
        self.net, self.loss_fn, self.opt, self.reducer = net, loss_fn, optimizer, reducer
        self.warmup, self.max_cached = max(int(warmup), 1), max_cached
        self.world = reducer.world if reducer is not None else 1
        self._cache: Dict[tuple, _Captured] = {}
        self.captures = 0
        self.replayed_launches = 0      # library kernels executed through graph replays (bench.py `gpu_launches`)
        self._copy_stream = None

    # ---- pieces of one step (run eagerly for warm-up, then under capture) -------------------------------------
    def _fwd_bwd(self, static):
        static.__dict__.pop("_stinet_cache", None)      # the structure is rebuilt from the batch's index tensors
        if self.reducer is not None:
            self.reducer.zero_grad()
        elif self.opt is not None:
            self.opt.zero_grad(set_to_none=True)
        else:
            self.net.zero_grad(set_to_none=True)
        loss = self.loss_fn(self.net(static), static)
        from . import ops
        # this step owns the backward call and nothing but the reducer and the optimizer reads the gradients: the weight
        # gradients' stream is joined once, at the end (ops.deferred_wgrad_join)
        with ops.deferred_wgrad_join():
            if self.world > 1:
                (loss * self.reducer.loss_scale).backward()  # mean over ranks = sum of the ranks' scaled gradients
            else:
                loss.backward()
        return loss

    def _finish(self):
        if self.reducer is not None:
            self.reducer.finish()

    def _capture(self, batch, sig) -> _Captured:
        dev = next(self.net.parameters()).device
        c = _Captured()
        c.static = GraphBatch()
        c.flat, c.views = _flat_views(list(_tensor_items(batch)), dev)      # static inputs of the graph: views of one buffer
        for k, v in c.views(c.flat).items():
            v.copy_(batch[k], non_blocking=True)
            c.static.__dict__[k] = v
        c.static.__dict__["_nv_host"] = sig[1]
        c.staging = c.staged = c.ready = c.free = None
        torch.cuda.synchronize(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        # ---- eager warm-up without side effects: lazy inits (kernel attributes, allocator, segment tables, optimizer
        # state) happen here, then parameters / buffers / optimizer state are put back and no collective is issued
        params = [p for p in self.net.parameters()]
        buffers = [b for b in self.net.buffers()]
        with torch.no_grad():
            snap_p = [p.detach().clone() for p in params]
            snap_b = [b.detach().clone() for b in buffers]
            snap_o = {}
            if self.opt is not None:
                for p, st in self.opt.state.items():
                    snap_o[p] = {k: v.detach().clone() for k, v in st.items() if torch.is_tensor(v)}
        if self.reducer is not None:
            self.reducer.enabled = False
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self._fwd_bwd(c.static)
                if self.opt is not None:
                    self.opt.step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        with torch.no_grad():
            for p, v in zip(params, snap_p):
                p.copy_(v)
            for b, v in zip(buffers, snap_b):
                b.copy_(v)
            if self.opt is not None:
                for p, st in self.opt.state.items():          # in place: the graph must see these very tensors
                    old = snap_o.get(p)
                    for k, v in st.items():
                        if torch.is_tensor(v):
                            if old is not None and k in old:
                                v.copy_(old[k])
                            else:
                                v.zero_()                     # state created by the warm-up: back to its initial zeros
        if self.reducer is not None:
            self.reducer.enabled = True
        torch.cuda.synchronize(dev)
        from . import ops
        ops.invalidate_planes()                          # the graph must contain the split of every operand it reads
        in_graph = self.world > 1 and os.environ.get("STINET_ALLREDUCE_IN_GRAPH", "1") != "0"
        overlap = getattr(self.reducer, "overlap", False)
        if self.reducer is not None and not in_graph:
            self.reducer.overlap = False                 # hooks must not launch collectives inside the capture
        n0 = _abi.query("stinet_launch_count")
        try:
            # with a process group alive, NCCL's watchdog thread polls CUDA events while we capture: only this thread's
            # calls may be policed by the capture (the mode PyTorch documents for DDP + CUDA graphs)
            mode = "thread_local" if self.world > 1 else "global"
            c.graph_a = torch.cuda.CUDAGraph()
            with torch.cuda.graph(c.graph_a, capture_error_mode=mode):
                c.loss = self._fwd_bwd(c.static)
                if self.world == 1 or in_graph:
                    self._finish()                       # waits for the bucket all-reduces the hooks forked off
                    if self.opt is not None:
                        self.opt.step()
            c.graph_b = None
            c.fresh = None
            if self.world > 1 and not in_graph:
                c.fresh = {p: p.grad for p in self.reducer.params}     # static outputs of the backward graph
                self.reducer.repoint()
            if self.world > 1 and not in_graph and self.opt is not None:
                c.graph_b = torch.cuda.CUDAGraph()
                with torch.cuda.graph(c.graph_b, pool=c.graph_a.pool(), capture_error_mode=mode):
                    self.opt.step()
        finally:
            if self.reducer is not None:
                self.reducer.overlap = overlap
        c.in_graph = in_graph
        c.launches = _abi.query("stinet_launch_count") - n0   # kernels of this library recorded into the graphs
        self.captures += 1
        return c

    def prefetch(self, batch, structs=None) -> bool:
        """Start moving the NEXT batch (pinned host memory) to the device on a copy stream while the current step is
        still running; the following `step(batch)` call with the same object then only pays one device-to-device copy.
        What a DataLoader with pin_memory + a prefetching collate thread gives the reference trainer.
        `structs`: the samples' cached SampleStructures (stinet_b200.structure) -- their block-diagonal concatenation
        is then done here too, on the copy stream, and attached to `batch`.
        Returns False (nothing staged) for a batch shape that has not been captured yet."""
        dev = next(self.net.parameters()).device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        cur = torch.cuda.current_stream(dev)
        if any(v.is_cuda for _, v in _tensor_items(batch)):  # device-resident fields were produced on the compute stream
            self._copy_stream.wait_stream(cur)
        if structs is not None:
            from .structure import attach_batch_structure
            with torch.cuda.stream(self._copy_stream):
                attach_batch_structure(batch, structs, dev)
        c = self._cache.get(batch_signature(batch))
        if c is None:
            if structs is not None:
                cur.wait_stream(self._copy_stream)       # the attached arrays will be read by the capturing call
            return False
        if c.staging is None:
            c.staging = torch.empty_like(c.flat)
            c.ready, c.free = torch.cuda.Event(), torch.cuda.Event()
            c.free.record(cur)
        self._copy_stream.wait_event(c.free)             # the previous staged batch has been consumed
        with torch.cuda.stream(self._copy_stream):
            for k, v in c.views(c.staging).items():
                v.copy_(batch[k], non_blocking=True)
            c.ready.record(self._copy_stream)
        c.staged = batch
        return True

    def __call__(self, batch) -> torch.Tensor:
        sig = batch_signature(batch)
        c = self._cache.get(sig)
        if c is None:
            if len(self._cache) >= self.max_cached:
                self._cache.pop(next(iter(self._cache)))
            c = self._cache[sig] = self._capture(batch, sig)
        if c.staged is batch:                            # prefetched: wait for the copy stream, one D2D copy
            cur = torch.cuda.current_stream(c.flat.device)
            cur.wait_event(c.ready)
            c.flat.copy_(c.staging, non_blocking=True)
            c.free.record(cur)
            c.staged = None
        else:
            for k, v in _tensor_items(batch):
                c.static.__dict__[k].copy_(v, non_blocking=True)
        c.graph_a.replay()
        self.replayed_launches += c.launches
        if self.world > 1 and not getattr(c, "in_graph", False):
            # collectives outside the graphs: the backward graph left fresh gradients behind; pack, reduce, then step
            self.reducer.reduce_from(c.fresh)
            if c.graph_b is not None:
                c.graph_b.replay()
        return c.loss


class GraphedForward:
    """Inference counterpart of GraphedTrainStep: structure build + forward (eval, no_grad) captured once per batch shape
    signature into a CUDA graph and replayed -- full-scene inference is ~800 short kernels, launched one by one from
    Python it is bound by launch overhead, not by the device.

        fwd = GraphedForward(net.eval())
        out = fwd(batch)            # H2D / D2D copies into the static inputs + one graph launch; `out` is the graph's
                                    # static output buffer (clone it if it has to outlive the next call)
    """

    def __init__(self, net: torch.nn.Module, warmup: int = 1, max_cached: int = 8):
        self.net, self.warmup, self.max_cached = net, max(int(warmup), 1), max_cached
        self._cache: Dict[tuple, _Captured] = {}
        self.captures = 0
        self.replayed_launches = 0

    def _run(self, static):
        static.__dict__.pop("_stinet_cache", None)      # the structure is rebuilt from the batch's index tensors
        with torch.no_grad():
            return self.net(static)

    def _capture(self, batch, sig) -> _Captured:
        from . import ops
        dev = next(self.net.parameters()).device
        c = _Captured()
        c.static = GraphBatch()
        c.flat, c.views = _flat_views(list(_tensor_items(batch)), dev)
        for k, v in c.views(c.flat).items():
            v.copy_(batch[k], non_blocking=True)
            c.static.__dict__[k] = v
        c.static.__dict__["_nv_host"] = sig[1]
        torch.cuda.synchronize(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(self.warmup):                 # eager warm-up: lazy inits, allocator, segment tables (no side effects)
                self._run(c.static)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        ops.invalidate_planes()
        n0 = _abi.query("stinet_launch_count")
        c.graph_a = torch.cuda.CUDAGraph()
        with torch.cuda.graph(c.graph_a):
            c.loss = self._run(c.static)                 # the static output
        c.launches = _abi.query("stinet_launch_count") - n0
        self.captures += 1
        return c

    def __call__(self, batch) -> torch.Tensor:
        sig = batch_signature(batch)
        c = self._cache.get(sig)
        if c is None:
            if len(self._cache) >= self.max_cached:
                self._cache.pop(next(iter(self._cache)))
            c = self._cache[sig] = self._capture(batch, sig)
        for k, v in _tensor_items(batch):
            c.static.__dict__[k].copy_(v, non_blocking=True)
        c.graph_a.replay()
        self.replayed_launches += c.launches
        return c.loss
