"""torch.autograd.Function wrappers around the C ABI (one per kernel family, explicit backward kernels).

Every Function is stateless, deterministic and RNG-free, so it is re-entrant under torch.utils.checkpoint
(reference models/surfacetextureinpaintingnet.py:429,438,451-455).  Tensors are borrowed as raw device pointers
for the duration of the call; outputs and workspaces are allocated by torch's caching allocator on the caller's
current stream.  fp32 storage everywhere; `precision` selects the arithmetic of the dense layers only.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _abi
from ._abi import ACT_ELU, ACT_NONE, PREC, REDUCE
from .graph import ClusterCSR, EdgeCSR, Segments, _ptr, _stream


def _mat(t: torch.Tensor) -> torch.Tensor:
    """fp32 CUDA matrix whose rows are contiguous (column slices of a wider buffer are fine)."""
    if not t.is_cuda:
        raise _abi.StinetError(f"stinet_b200 ops need CUDA tensors (got {t.device}); there is no CPU fallback")
    if t.dtype != torch.float32:
        raise _abi.StinetError(f"expected float32 storage, got {t.dtype}")
    assert t.dim() == 2
    if t.stride(1) != 1 or (t.size(0) > 1 and t.stride(0) < t.size(1)):
        t = t.contiguous()
    return t


def _ld(t: torch.Tensor) -> int:
    return t.stride(0) if t.size(0) > 1 else max(t.stride(0), t.size(1))


def _vec_ok(c: int, *tensors) -> bool:
    """128-bit row access is possible: width, pitches and base addresses are multiples of four floats."""
    return c % 4 == 0 and all(t is None or (t.data_ptr() % 16 == 0 and _ld(t) % 4 == 0) for t in tensors)


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------------------------
# weight gradients next to the data-gradient chain
#
# In a backward pass only the data gradients are on the critical path (dgrad -> message-stage backward -> dgrad ...);
# the weight gradients (wgrad GEMM, its split reduction, the un-hoisting of dWcat) are leaves that nothing reads before
# the optimizer.  They are launched on a second stream, so the tail of a dgrad kernel that no longer fills the 148 SMs
# overlaps with the head of the wgrad next to it (and the other way round); in a captured step these are parallel branches
# of the CUDA graph.  Two ways to rejoin:
#   * per node (default, safe under any caller): the node's backward waits for the second stream before it returns, so
#     whatever autograd or a hook does with the gradient next is ordered after it;
#   * deferred (inside `deferred_wgrad_join()`, which engine.GraphedTrainStep enters around its backward): one join at
#     the end of the backward pass (autograd's final callback), and one before GradAllReducer packs a bucket.  Nothing on
#     the main stream may touch such a gradient earlier, so the node does not hand it to autograd at all (whether
#     AccumulateGrad keeps a gradient by reference or copies it is its own business -- under compute-sanitizer it
#     copies): it deposits it in `.grad` itself and calls the reducer's hook, which is what AccumulateGrad would have done
#     for a leaf parameter without a gradient.  Weights that do not qualify (`_straight_to_grad`) join per node.
# STINET_WGRAD_SIDE_STREAM = 0: one stream; 1 (default): as above; 2: deferred everywhere.
_WGRAD_SIDE = int(os.environ.get("STINET_WGRAD_SIDE_STREAM", "1"))
_WGRAD_STREAMS = {}
_wgrad_join_due = set()
_defer_depth = 0


class deferred_wgrad_join:
    """with deferred_wgrad_join(): loss.backward()   -- see above.  The caller owns the backward call and promises that
    nothing but the optimizer (or GradAllReducer) reads the parameters' gradients before backward() has returned."""

    def __enter__(self):
        global _defer_depth
        _defer_depth += 1
        return self

    def __exit__(self, *exc):
        global _defer_depth
        _defer_depth -= 1
        return False


def _deferring() -> bool:
    return _WGRAD_SIDE == 2 or (_WGRAD_SIDE == 1 and _defer_depth > 0)


def _wgrad_stream(dev: torch.device):
    k = dev.index if dev.index is not None else torch.cuda.current_device()
    if k not in _WGRAD_STREAMS:
        _WGRAD_STREAMS[k] = torch.cuda.Stream(device=dev)
    return _WGRAD_STREAMS[k]


class _on_wgrad_stream:
    """with _on_wgrad_stream(dev, reads): ...   launches of the body go to the weight-gradient stream, after everything
    issued so far on the current one.  `reads`: Planes / tensors of the current stream the body reads."""

    def __init__(self, dev, reads=()):
        self.dev, self.reads, self.ctx = dev, reads, None

    def __enter__(self):
        if not _WGRAD_SIDE:
            return self
        main, side = torch.cuda.current_stream(self.dev), _wgrad_stream(self.dev)
        side.wait_stream(main)
        if _deferring():                             # the reader may still be running when the caller drops its buffers
            for r in self.reads:
                for t in ((r.hi, r.lo, r.exp) if isinstance(r, Planes) else (r,)):
                    if t is not None:
                        t.record_stream(side)
        self.ctx = torch.cuda.stream(side)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def join_wgrad_stream(dev=None) -> None:
    """The current stream waits for the weight gradients launched so far (no-op when nothing is in flight)."""
    for k in ([dev.index if dev.index is not None else torch.cuda.current_device()] if dev is not None
              else list(_wgrad_join_due)):
        if k in _wgrad_join_due:
            _wgrad_join_due.discard(k)
            torch.cuda.current_stream(k).wait_stream(_WGRAD_STREAMS[k])


def _straight_to_grad(*weights) -> bool:
    """True when these weights are leaf parameters without a gradient yet, without tensor hooks, and with post-accumulate
    hooks only from GradAllReducer (which joins before it packs): depositing the gradient in `.grad` is then all that
    autograd would do with it.  A weight that is itself a function of a parameter (a padded or hoisted copy) has an
    autograd node downstream that reads the gradient at once."""
    for w in weights:
        if w is None:
            continue
        if not w.is_leaf or w.grad is not None or w._backward_hooks:
            return False
        if getattr(w, "_post_accumulate_grad_hooks", None) and not getattr(w, "_stinet_wgrad_aware", False):
            return False
    return True


def _wgrad_done(dev, pairs):
    """End of a backward that used _on_wgrad_stream.  pairs: [(weight, gradient produced on the second stream or None)].
    Returns the gradients to hand to autograd, in order: the tensors themselves after a join, or None for those that were
    deposited in `.grad` here (deferred join)."""
    grads = [g for _, g in pairs]
    if not _WGRAD_SIDE or all(g is None for g in grads):
        return grads
    main = torch.cuda.current_stream(dev)
    for g in grads:
        if g is not None:
            g.record_stream(main)
    k = dev.index if dev.index is not None else torch.cuda.current_device()
    _wgrad_join_due.add(k)
    if not (_deferring() and _straight_to_grad(*[w for w, g in pairs if g is not None])):
        join_wgrad_stream(dev)
        return grads
    # idempotent: the first callback to run joins, the others find nothing due
    torch.autograd.Variable._execution_engine.queue_callback(lambda: join_wgrad_stream(dev))
    for w, g in pairs:
        if g is None:
            continue
        w.grad = g
        for hook in list((getattr(w, "_post_accumulate_grad_hooks", None) or {}).values()):
            hook(w)
    return [None] * len(grads)


# ------------------------------------------------------------------------------------------------------------
# dense layers


class LinearFn(Function):
    """y = x W^T + b, bias only on rows with rowmask > 0 when a rowmask is given."""

    @staticmethod
    def forward(ctx, x, weight, bias, rowmask, precision):
        x, weight = _mat(x), _mat(weight)
        M, K = x.shape
        N = weight.shape[0]
        assert weight.shape[1] == K
        y = torch.empty((M, N), dtype=torch.float32, device=x.device)
        prec = PREC[precision]
        nb = _abi.query("stinet_gemm_workspace_bytes", M, N, K, prec)
        ws = _ws(nb, x.device)
        _abi.call("stinet_linear_fwd", x.data_ptr(), _ld(x), weight.data_ptr(), _ld(weight), _ptr(bias),
                  _ptr(rowmask), y.data_ptr(), N, M, N, K, prec, ws.data_ptr(), nb, _stream(),
                  cost=(4 * (M * K + N * K + M * N), 2 * M * N * K, f"{N}x{K}"))
        ctx.save_for_backward(x, weight, rowmask)
        ctx.has_bias = bias is not None
        ctx.prec = prec
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, weight, rowmask = ctx.saved_tensors
        dy = _mat(dy)
        M, K = x.shape
        N = weight.shape[0]
        nb = _abi.query("stinet_gemm_workspace_bytes", M, N, K, ctx.prec)
        ws = _ws(nb, x.device)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((M, K), dtype=torch.float32, device=x.device)
            _abi.call("stinet_linear_dgrad", dy.data_ptr(), _ld(dy), weight.data_ptr(), _ld(weight), dx.data_ptr(), K,
                      M, N, K, ctx.prec, ws.data_ptr(), nb, _stream(),
                      cost=(4 * (M * K + N * K + M * N), 2 * M * N * K, f"{N}x{K}"))
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw = torch.empty((N, K), dtype=torch.float32, device=x.device)
            db = torch.empty((N,), dtype=torch.float32, device=x.device) if ctx.has_bias else None
            _abi.call("stinet_linear_wgrad", dy.data_ptr(), _ld(dy), x.data_ptr(), _ld(x), _ptr(rowmask),
                      dw.data_ptr(), K, _ptr(db), M, N, K, ctx.prec, ws.data_ptr(), nb, _stream(),
                      cost=(4 * (M * K + N * K + M * N), 2 * M * N * K, f"{N}x{K}"))
        return dx, dw, db, None, None


class Planes:
    """fp16 hi / lo planes of a scaled fp32 matrix and the device scalar exp = -s (include/stinet_b200.h, "operand
    PLANES").  Produced once per matrix and shared by every dense layer that reads it (fwd + wgrad for activations,
    dgrad + wgrad for output gradients, fwd + dgrad for weights)."""
    __slots__ = ("hi", "lo", "exp", "rows", "cols", "ld")

    def __init__(self, hi, lo, exp, rows, cols, ld):
        self.hi, self.lo, self.exp, self.rows, self.cols, self.ld = hi, lo, exp, rows, cols, ld


_plane_epoch = 0


def invalidate_planes() -> None:
    """Forget every cached plane set (called before a CUDA-graph capture: a graph must contain the split kernels of
    everything it reads, it may not lean on planes computed outside of it)."""
    global _plane_epoch
    _plane_epoch += 1


def set_amax(t: torch.Tensor, amax: torch.Tensor) -> torch.Tensor:
    """A producer kernel already knows max|t| (or an upper bound): planes_of(t) then skips its reduction pass."""
    t._stinet_amax = (t._version, amax)
    return t


def planes_of(t: torch.Tensor, need_lo: bool = True) -> Planes:
    """The operand planes of fp32 matrix `t`, cached on the tensor object (keyed by its version counter)."""
    # Parameters are never cached: fused / foreach optimizers update them without touching the version counter, so a
    # split taken in one step says nothing about the next one (one split per forward, shared with backward through ctx)
    cacheable = not (t.is_leaf and t.requires_grad)
    if not cacheable:
        wp = _weight_planes(t, 0)                    # split by this forward's WeightPlanes.refresh()
        if wp is not None:
            return wp[0]
    cached = getattr(t, "_stinet_planes", None) if cacheable else None
    if cached is not None and cached[0] == t._version and cached[1] == _plane_epoch and (cached[2].lo is not None or not need_lo):
        return cached[2]
    x = _mat(t)
    rows, cols = x.shape
    dev = x.device
    s = _stream()
    known = getattr(t, "_stinet_amax", None)
    if known is not None and known[0] == t._version:
        amax = known[1]
    else:
        amax = torch.empty(1, dtype=torch.float32, device=dev)
        _abi.call("stinet_f16_amax", x.data_ptr(), _ld(x), rows, cols, amax.data_ptr(), s,
                  cost=(4 * rows * cols, 0, ""))
    ld = (cols + 7) & ~7
    hi = torch.empty((rows, ld), dtype=torch.float16, device=dev)
    lo = torch.empty((rows, ld), dtype=torch.float16, device=dev) if need_lo else None
    exp = torch.empty(1, dtype=torch.int32, device=dev)
    _abi.call("stinet_f16_split", x.data_ptr(), _ld(x), rows, cols, amax.data_ptr(), hi.data_ptr(), _ptr(lo), ld,
              exp.data_ptr(), s, cost=(rows * cols * (4 + (4 if need_lo else 2)), 0, ""))
    p = Planes(hi, lo, exp, rows, cols, ld)
    if cacheable:
        t._stinet_planes = (t._version, _plane_epoch, p)
    return p


_weight_epoch = 0


class WeightPlanes:
    """Operand planes of all dense-layer weights of a module, refreshed in two launches per forward (the weights change
    with every optimizer step; fused optimizers do not touch version counters, so nothing is cached across steps).
    entries: (weight Parameter, bias Parameter or None, kind) with kind 0 = the weight itself, 1 = hoisted EdgeConv first
    layer [Wa - Wb ; Wb], 2 = hoisted EdgeConvTransInv first layer [-W ; W] (both with bcat = [b ; 0])."""

    def __init__(self, entries):
        self.entries = [(w, b, int(k)) for w, b, k in entries]
        self._key = None

    def _build(self, dev):
        n = len(self.entries)
        self.amax = torch.zeros(max(n, 1), dtype=torch.float32, device=dev)
        self.exp = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
        esz = _abi.query("stinet_weight_entry_bytes")
        import ctypes
        host = (ctypes.c_char * (esz * max(n, 1)))()
        self.planes, chunk0 = [], 0
        lib = _abi.load()
        for i, (w, b, kind) in enumerate(self.entries):
            assert w.dim() == 2 and w.stride(1) == 1 and w.dtype == torch.float32
            rows = w.shape[0]
            cols = w.shape[1] if kind != 1 else w.shape[1] // 2
            out_rows = rows if kind == 0 else 2 * rows
            p = _new_planes(out_rows, cols, True, dev)
            p.exp = self.exp[i:i + 1]
            bcat = torch.empty(2 * rows, dtype=torch.float32, device=dev) if kind != 0 else None
            self.planes.append((p, bcat))
            chunk0 += int(lib.stinet_weight_entry_fill(
                ctypes.addressof(host) + i * esz, w.data_ptr(), w.stride(0), _ptr(b), rows, cols, kind, p.hi.data_ptr(),
                p.lo.data_ptr(), p.ld, _ptr(bcat), self.amax[i:i + 1].data_ptr(), self.exp[i:i + 1].data_ptr(), chunk0))
        self.n_chunks = chunk0
        self.table = torch.frombuffer(host, dtype=torch.uint8).clone().to(dev)

    def refresh(self):
        """Split every weight of the table (2 launches) and stamp the parameters with their fresh planes."""
        global _weight_epoch
        if not self.entries:
            return
        dev = self.entries[0][0].device
        key = tuple((w.data_ptr(), None if b is None else b.data_ptr()) for w, b, _ in self.entries)
        if key != self._key:                         # first use, or the parameters moved (.to(), load with assign=True)
            self._build(dev)
            self._key = key
        _abi.call("stinet_weight_planes_refresh", self.table.data_ptr(), len(self.entries), self.n_chunks, self.amax.data_ptr(),
                  _stream(), cost=(12 * sum(w.numel() for w, _, _ in self.entries), 0, ""))
        _weight_epoch += 1
        for (w, _, kind), (p, bcat) in zip(self.entries, self.planes):
            w._stinet_wp = (_weight_epoch, kind, p, bcat)


def _weight_planes(w: torch.Tensor, kind: int):
    """(planes, bcat) stamped by a WeightPlanes.refresh() of the CURRENT forward, else None."""
    wp = getattr(w, "_stinet_wp", None)
    if wp is not None and wp[0] == _weight_epoch and wp[1] == kind:
        return wp[2], wp[3]
    return None


def _new_planes(rows: int, cols: int, need_lo: bool, dev) -> Planes:
    ld = (cols + 7) & ~7
    hi = torch.empty((rows, ld), dtype=torch.float16, device=dev)
    lo = torch.empty((rows, ld), dtype=torch.float16, device=dev) if need_lo else None
    return Planes(hi, lo, torch.empty(1, dtype=torch.int32, device=dev), rows, cols, ld)


def _pl_fwd(xp: Planes, wp: Planes, bias, rowmask, passes: int, want_amax: bool = False):
    """y = x W^T + b on planes; optionally also max|y| (device scalar) from the GEMM's epilogue."""
    M, K, N = xp.rows, xp.cols, wp.rows
    assert wp.cols == K
    dev = xp.hi.device
    y = torch.empty((M, N), dtype=torch.float32, device=dev)
    amax = torch.empty(1, dtype=torch.float32, device=dev) if want_amax else None
    nb = _abi.query("stinet_gemm_workspace_bytes", M, N, K, 0)
    ws = _ws(nb, dev)
    _abi.call("stinet_linear_fwd_f16", xp.hi.data_ptr(), _ptr(xp.lo), xp.ld, xp.exp.data_ptr(), wp.hi.data_ptr(),
              _ptr(wp.lo), wp.ld, wp.exp.data_ptr(), _ptr(bias), _ptr(rowmask), y.data_ptr(), N, _ptr(amax), M, N, K,
              passes, ws.data_ptr(), nb, _stream(), cost=(4 * (M * K + N * K + M * N), 2 * M * N * K, f"{N}x{K}"))
    return y, amax


def _pl_dgrad(dyp: Planes, wp: Planes, passes: int, want_amax: bool = False):
    M, N, K = dyp.rows, wp.rows, wp.cols
    assert dyp.cols == N
    dev = dyp.hi.device
    dx = torch.empty((M, K), dtype=torch.float32, device=dev)
    amax = torch.empty(1, dtype=torch.float32, device=dev) if want_amax else None
    nb = _abi.query("stinet_gemm_workspace_bytes", M, N, K, 0)
    ws = _ws(nb, dev)
    _abi.call("stinet_linear_dgrad_f16", dyp.hi.data_ptr(), _ptr(dyp.lo), dyp.ld, dyp.exp.data_ptr(),
              wp.hi.data_ptr(), _ptr(wp.lo), wp.ld, wp.exp.data_ptr(), dx.data_ptr(), K, _ptr(amax), M, N, K, passes,
              ws.data_ptr(), nb, _stream(), cost=(4 * (M * K + N * K + M * N), 2 * M * N * K, f"{N}x{K}"))
    return dx, amax


def _pl_wgrad(dyp: Planes, xp: Planes, passes: int):
    M, N, K = dyp.rows, dyp.cols, xp.cols
    assert xp.rows == M
    dev = dyp.hi.device
    dw = torch.empty((N, K), dtype=torch.float32, device=dev)
    if M == 0:
        return dw.zero_()
    nb = _abi.query("stinet_gemm_workspace_bytes", M, N, K, 0)
    ws = _ws(nb, dev)
    _abi.call("stinet_linear_wgrad_f16", dyp.hi.data_ptr(), _ptr(dyp.lo), dyp.ld, dyp.exp.data_ptr(),
              xp.hi.data_ptr(), _ptr(xp.lo), xp.ld, xp.exp.data_ptr(), dw.data_ptr(), K, M, N, K, passes,
              ws.data_ptr(), nb, _stream(), cost=(4 * (M * K + N * K + M * N), 2 * M * N * K, f"{N}x{K}"))
    return dw


def _colsum(dy: torch.Tensor, rowmask) -> torch.Tensor:
    """dbias = masked column sums of dy (deterministic two-stage reduction)."""
    dym = _mat(dy)
    M, N = dym.shape
    db = torch.empty((N,), dtype=torch.float32, device=dym.device)
    nb = _abi.query("stinet_gemm_workspace_bytes", M, N, 1, 0)
    ws = _ws(nb, dym.device)
    _abi.call("stinet_colsum", dym.data_ptr(), _ld(dym), _ptr(rowmask), M, N, db.data_ptr(), ws.data_ptr(), nb,
              _stream(), cost=(4 * M * N, 0, f"N{N}"))
    return db


def planes_and_colsum(t: torch.Tensor, rowmask, need_lo: bool = True):
    """(operand planes of t, masked column sums of t) -- what the backward of a Linear with bias needs from dC -- in one
    pass over t when the layout allows the fused kernel, else as two."""
    x = _mat(t)
    rows, cols = x.shape
    cached = getattr(t, "_stinet_planes", None)
    fused = (cached is None and rows > 0 and cols % 8 == 0 and _ld(x) % 4 == 0 and x.data_ptr() % 16 == 0
             and not (t.is_leaf and t.requires_grad))
    if not fused:
        return planes_of(t, need_lo), _colsum(t, rowmask)
    dev = x.device
    s = _stream()
    known = getattr(t, "_stinet_amax", None)
    if known is not None and known[0] == t._version:
        amax = known[1]
    else:
        amax = torch.empty(1, dtype=torch.float32, device=dev)
        _abi.call("stinet_f16_amax", x.data_ptr(), _ld(x), rows, cols, amax.data_ptr(), s, cost=(4 * rows * cols, 0, ""))
    p = _new_planes(rows, cols, need_lo, dev)
    db = torch.empty((cols,), dtype=torch.float32, device=dev)
    nb = _abi.query("stinet_gemm_workspace_bytes", rows, cols, 1, 0)
    ws = _ws(nb, dev)
    _abi.call("stinet_f16_split_colsum", x.data_ptr(), _ld(x), rows, cols, amax.data_ptr(), _ptr(rowmask), p.hi.data_ptr(),
              _ptr(p.lo), p.ld, p.exp.data_ptr(), db.data_ptr(), ws.data_ptr(), nb, s,
              cost=(rows * cols * (4 + (4 if need_lo else 2)), 0, ""))
    t._stinet_planes = (t._version, _plane_epoch, p)
    return p, db


# experiment knob (scripts/diag_f16.py): tcgen05 passes of the BACKWARD GEMMs when the forward runs one pass
_BWD_PASSES = int(os.environ.get("STINET_F16_BWD_PASSES", "0"))


def _bwd_passes(passes: int) -> int:
    return _BWD_PASSES if (_BWD_PASSES and passes == 1) else passes


class LinearPlanesFn(Function):
    """y = x W^T + b on fp16 operand planes (tcgen05 kind::f16; passes = 3: fp32-class result, 1: 11-bit operands)."""

    last_amax = None      # max|y| of the most recent forward (from the GEMM's epilogue), picked up by linear(want_amax=True)

    @staticmethod
    def forward(ctx, x, weight, bias, rowmask, passes):
        need_lo = _bwd_passes(passes) == 3
        xp, wp = planes_of(x, need_lo), planes_of(weight, need_lo)
        y, LinearPlanesFn.last_amax = _pl_fwd(xp, wp, bias, rowmask, passes, want_amax=True)
        ctx.xp, ctx.wp, ctx.rowmask = xp, wp, rowmask
        ctx.has_bias, ctx.passes = bias is not None, passes
        ctx.weights = (weight,)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        xp, wp, rowmask, passes = ctx.xp, ctx.wp, ctx.rowmask, _bwd_passes(ctx.passes)
        dx = dw = db = None
        if ctx.has_bias and ctx.needs_input_grad[2]:
            dyp, db = planes_and_colsum(dy, rowmask, passes == 3)
        else:
            dyp = planes_of(dy, passes == 3)
        dev = dy.device
        if ctx.needs_input_grad[1]:
            with _on_wgrad_stream(dev, (dyp, xp)):
                dw = _pl_wgrad(dyp, xp, passes)
        if ctx.needs_input_grad[0]:
            dx, _ = _pl_dgrad(dyp, wp, passes)
        (dw,) = _wgrad_done(dev, [(ctx.weights[0], dw)])
        return dx, dw, db, None, None


# observer(pq, csr) called with the [P | Q] matrix of every fused EdgeConv message stage (tests record the ReLU decisions)
_edge_message_observer = None


class EdgeConvFn(Function):
    """out_i = mean_{j->i} nn([x_i || x_j - x_i])  (trans_inv: nn(x_j - x_i)),  nn = Lin(.,H) ReLU Lin(H,dout), evaluated in
    the hoisted form as ONE autograd node on operand planes:
        [P | Q] = x Wcat^T + [b0 ; 0]          tcgen05 GEMM, epilogue leaves max|PQ|
        hid     = mean relu(P_i + Q_j)          written as fp16 planes (scale from 2 max|PQ|), ReLU decisions saved
        out     = hid W2^T + b2 [deg > 0]       tcgen05 GEMM
    Backward: dhid = dy W2 (epilogue leaves max|dhid|), dPQ written as planes by the message-stage backward, then the
    two wgrads and the dgrad.  hid and dPQ never exist as fp32 matrices; every operand is split exactly once."""

    @staticmethod
    def forward(ctx, x, w0, b0, w2, b2, csr: EdgeCSR, trans_inv: bool, passes: int):
        need_lo = _bwd_passes(passes) == 3
        dev = x.device
        s = _stream()
        xp = planes_of(x, need_lo)
        n, din = xp.rows, xp.cols
        w0m = _mat(w0)
        h = w0m.shape[0]
        kin = w0m.shape[1]
        assert kin == (din if trans_inv else 2 * din) and n == csr.n
        dout = w2.shape[0]
        hp = _weight_planes(w0, 2 if trans_inv else 1)
        if hp is not None:                           # hoisted and split by this forward's WeightPlanes.refresh()
            wcp, bcat = hp
            if b0 is None:
                bcat = None
        else:
            wcat = torch.empty((2 * h, din), dtype=torch.float32, device=dev)
            bcat = torch.empty((2 * h,), dtype=torch.float32, device=dev) if b0 is not None else None
            _abi.call("stinet_edgeconv_hoist_fwd", w0m.data_ptr(), _ld(w0m), _ptr(b0), h, din, int(trans_inv), wcat.data_ptr(),
                      _ptr(bcat), s, cost=(4 * h * (kin + 2 * din), 0, ""))
            wcp = planes_of(wcat, need_lo)
        pq, pq_amax = _pl_fwd(xp, wcp, bcat, None, passes, want_amax=True)
        if _edge_message_observer is not None:
            _edge_message_observer(pq, csr)
        train = any(ctx.needs_input_grad[:5])
        hidp = _new_planes(n, h, need_lo, dev)
        mask = torch.empty((max(csr.e, 1), h // 4), dtype=torch.uint8, device=dev) if train else None
        _abi.call("stinet_edge_message_fwd_planes", pq.data_ptr(), 2 * h, pq.data_ptr() + 4 * h, 2 * h,
                  csr.rowptr_t.data_ptr(), csr.col_t.data_ptr(), n, h, pq_amax.data_ptr(), hidp.hi.data_ptr(),
                  _ptr(hidp.lo), hidp.ld, hidp.exp.data_ptr(), _ptr(mask), s,
                  cost=(csr.e * (4 * h + 4 + (h // 4 if train else 0)) + n * (8 * h + 4), 2 * csr.e * h, f"H{h}"))
        w2p = planes_of(w2, need_lo)
        y, _ = _pl_fwd(hidp, w2p, b2, csr.degree, passes)
        if train:
            ctx.saved = (xp, wcp, hidp, w2p, mask, csr)
        ctx.dims = (n, din, h, kin, dout, trans_inv, passes, b0 is not None, b2 is not None)
        ctx.weights = (w0, w2)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        xp, wcp, hidp, w2p, mask, csr = ctx.saved
        n, din, h, kin, dout, trans_inv, passes, has_b0, has_b2 = ctx.dims
        passes = _bwd_passes(passes)
        need_lo = passes == 3
        dev = dy.device
        s = _stream()
        db2 = None
        if has_b2 and ctx.needs_input_grad[4]:
            dyp, db2 = planes_and_colsum(dy, csr.degree, need_lo)
        else:
            dyp = planes_of(dy, need_lo)
        # second Linear
        dw2 = None
        if ctx.needs_input_grad[3]:
            with _on_wgrad_stream(dev, (dyp, hidp)):
                dw2 = _pl_wgrad(dyp, hidp, passes)
        dhid, dhid_amax = _pl_dgrad(dyp, w2p, passes, want_amax=True)
        # message stage: dPQ = [dP | dQ] as planes
        rowptr_s, col_s, _ = csr.by_source()
        dpqp = _new_planes(n, 2 * h, need_lo, dev)
        want_db0 = has_b0 and ctx.needs_input_grad[2]
        db0 = torch.empty((h,), dtype=torch.float32, device=dev) if want_db0 else None     # = sum_i dP_i, from the same kernel
        nbe = _abi.query("stinet_edge_message_bwd_workspace_bytes", n, h) if want_db0 else 0
        wse = _ws(nbe, dev) if want_db0 else None
        _abi.call("stinet_edge_message_bwd_planes", dhid.data_ptr(), h, dhid_amax.data_ptr(), csr.dq_factor().data_ptr(),
                  csr.rowptr_t.data_ptr(), rowptr_s.data_ptr(), col_s.data_ptr(), csr.tpos_s().data_ptr(), mask.data_ptr(),
                  n, h, dpqp.hi.data_ptr(), _ptr(dpqp.lo), dpqp.ld, dpqp.exp.data_ptr(), _ptr(db0), _ptr(wse), nbe, s,
                  cost=(csr.e * (4 * h + h // 2 + 12) + n * (16 * h + 8), 2 * csr.e * h, f"H{h}"))
        # first (hoisted) Linear
        dw0 = None
        if ctx.needs_input_grad[1]:
            with _on_wgrad_stream(dev, (dpqp, xp)):
                dwcat = _pl_wgrad(dpqp, xp, passes)
                dw0 = torch.empty((h, kin), dtype=torch.float32, device=dev)
                _abi.call("stinet_edgeconv_hoist_bwd", dwcat.data_ptr(), None, h, din, int(trans_inv), dw0.data_ptr(), kin,
                          None, _stream(), cost=(4 * h * (kin + 2 * din), 0, ""))
        dx = _pl_dgrad(dpqp, wcp, passes)[0] if ctx.needs_input_grad[0] else None
        dw0, dw2 = _wgrad_done(dev, [(ctx.weights[0], dw0), (ctx.weights[1], dw2)])
        return dx, dw0, db0, dw2, db2, None, None, None


def edge_conv(x, w0, b0, w2, b2, csr, trans_inv: bool, precision: str):
    """The fused EdgeConv node when the shapes allow the vector kernels (widths multiples of 4), else None."""
    passes = _abi.PLANE_PASSES.get(precision)
    h, dout, din = w0.shape[0], w2.shape[0], x.shape[1]
    if passes is None or h % 4 or dout % 4 or din % 4 or not x.is_cuda:
        return None
    return EdgeConvFn.apply(x, w0, b0, w2, b2, csr, trans_inv, passes)


def linear(x, weight, bias=None, rowmask=None, precision="fp32"):
    y = _linear(x, weight, bias, rowmask, precision)
    if LinearPlanesFn.last_amax is not None:         # the GEMM's epilogue left max|y| behind: consumers scale planes with it
        if y.dim() == 2 and y.is_contiguous():
            set_amax(y, LinearPlanesFn.last_amax)
        LinearPlanesFn.last_amax = None
    return y


def _linear(x, weight, bias=None, rowmask=None, precision="fp32"):
    """y = x W^T + b.  TMA (the tensor-core path) needs 16-byte row pitches, so a reduction width or an output width
    that is not a multiple of 4 floats (the 10-channel input, the 3-channel head) is zero-padded: the extra products
    are exact zeros and the extra output columns are sliced away again."""
    k, n = x.shape[1], weight.shape[0]
    pk, pn = (-k) % 4, (-n) % 4
    passes = _abi.PLANE_PASSES.get(precision)
    fn = LinearFn if passes is None else LinearPlanesFn
    arg = precision if passes is None else passes
    if not (pk or pn):
        return fn.apply(x, weight, bias, rowmask, arg)
    if pk:
        x = torch.nn.functional.pad(x, (0, pk))
    weight = torch.nn.functional.pad(weight, (0, pk, 0, pn))
    if bias is not None and pn:
        bias = torch.nn.functional.pad(bias, (0, pn))
    y = fn.apply(x, weight, bias, rowmask, arg)
    return y[:, :n] if pn else y


# ------------------------------------------------------------------------------------------------------------
# message passing


class EdgeMessageFn(Function):
    """hid[i] = mean_{j->i} relu(P[i] + Q[j]),  PQ = [P | Q] of shape [N, 2H].

    Main path (H % 4 == 0, aligned rows): forward saves the ReLU decisions (one byte per edge and float4 column) and
    backward reads those instead of P and Q -- dP becomes a row-local count, dQ gathers one row per
    out-edge instead of two, and [P|Q] itself does not have to be kept for this op.  Odd widths recompute P_i + Q_j."""

    @staticmethod
    def forward(ctx, pq, csr: EdgeCSR):
        pq = _mat(pq)
        n, h2 = pq.shape
        h = h2 // 2
        assert n == csr.n and h2 == 2 * h
        hid = torch.empty((n, h), dtype=torch.float32, device=pq.device)
        ld = _ld(pq)
        ctx.csr = csr
        ctx.shape = (n, h)
        # decisions are only worth storing when a backward pass will read them (not under no_grad / inference)
        ctx.masked = ctx.needs_input_grad[0] and h % 4 == 0 and ld % 4 == 0 and pq.data_ptr() % 16 == 0
        if ctx.masked:
            mask = torch.empty((max(csr.e, 1), h // 4), dtype=torch.uint8, device=pq.device)
            _abi.call("stinet_edge_message_fwd_mask", pq.data_ptr(), ld, pq.data_ptr() + 4 * h, ld,
                      csr.rowptr_t.data_ptr(), csr.col_t.data_ptr(), n, h, hid.data_ptr(), h, mask.data_ptr(), _stream(),
                      cost=(csr.e * (4 * h + 4 + h // 4) + n * (8 * h + 4), 2 * csr.e * h, f"H{h}"))
            ctx.save_for_backward(mask)
            return hid
        _abi.call("stinet_edge_message_fwd", pq.data_ptr(), ld, pq.data_ptr() + 4 * h, ld, csr.rowptr_t.data_ptr(),
                  csr.col_t.data_ptr(), n, h, hid.data_ptr(), h, _stream(),
                  cost=(csr.e * (4 * h + 4) + n * (8 * h + 4), 2 * csr.e * h, f"H{h}"))
        ctx.save_for_backward(pq)
        return hid

    @staticmethod
    @once_differentiable
    def backward(ctx, dhid):
        csr = ctx.csr
        dhid = _mat(dhid)
        n, h = ctx.shape
        h2 = 2 * h
        dpq = torch.empty((n, h2), dtype=torch.float32, device=dhid.device)
        rowptr_s, col_s, eid_s = csr.by_source()
        s = _stream()
        if ctx.masked:
            if _ld(dhid) % 4 or dhid.data_ptr() % 16:
                dhid = dhid.contiguous().clone()         # a fresh allocation is 16-byte aligned
            (mask,) = ctx.saved_tensors
            tpos_s = csr.tpos_s()
            _abi.call("stinet_edge_message_bwd_target_mask", dhid.data_ptr(), _ld(dhid), csr.rowptr_t.data_ptr(),
                      mask.data_ptr(), n, h, dpq.data_ptr(), h2, s,
                      cost=(csr.e * (h // 4) + n * (8 * h + 4), csr.e * h, f"H{h}"))
            _abi.call("stinet_edge_message_bwd_source_mask", dhid.data_ptr(), _ld(dhid), csr.rowptr_t.data_ptr(),
                      rowptr_s.data_ptr(), col_s.data_ptr(), tpos_s.data_ptr(), mask.data_ptr(), n, h,
                      dpq.data_ptr() + 4 * h, h2, s,
                      cost=(csr.e * (4 * h + h // 4 + 12) + n * (4 * h + 4), csr.e * h, f"H{h}"))
            return dpq, None
        (pq,) = ctx.saved_tensors
        ld = _ld(pq)
        _abi.call("stinet_edge_message_bwd_target", pq.data_ptr(), ld, pq.data_ptr() + 4 * h, ld, dhid.data_ptr(),
                  _ld(dhid), csr.rowptr_t.data_ptr(), csr.col_t.data_ptr(), n, h, dpq.data_ptr(), h2, s,
                  cost=(csr.e * (4 * h + 4) + n * (12 * h + 4), 2 * csr.e * h, f"H{h}"))
        _abi.call("stinet_edge_message_bwd_source", pq.data_ptr(), ld, pq.data_ptr() + 4 * h, ld, dhid.data_ptr(),
                  _ld(dhid), csr.rowptr_t.data_ptr(), rowptr_s.data_ptr(), col_s.data_ptr(), n, h,
                  dpq.data_ptr() + 4 * h, h2, s,
                  cost=(csr.e * (8 * h + 8) + n * (8 * h + 4), 2 * csr.e * h, f"H{h}"))
        return dpq, None


def edge_message(pq, csr):
    return EdgeMessageFn.apply(pq, csr)


class EdgeConvHoistFn(Function):
    """(Wcat, bcat) of the hoisted first layer from nn.0's own parameters, and the gradient fold back onto them."""

    @staticmethod
    def forward(ctx, w, b, trans_inv: bool):
        w = _mat(w)
        h, kin = w.shape
        din = kin if trans_inv else kin // 2
        assert trans_inv or kin == 2 * din
        wcat = torch.empty((2 * h, din), dtype=torch.float32, device=w.device)
        bcat = torch.empty((2 * h,), dtype=torch.float32, device=w.device) if b is not None else None
        _abi.call("stinet_edgeconv_hoist_fwd", w.data_ptr(), _ld(w), _ptr(b), h, din, int(trans_inv), wcat.data_ptr(),
                  _ptr(bcat), _stream(), cost=(4 * h * (kin + 2 * din), 0, ""))
        ctx.shape, ctx.trans_inv, ctx.has_bias = (h, kin, din), trans_inv, b is not None
        if bcat is None:
            return wcat
        return wcat, bcat

    @staticmethod
    @once_differentiable
    def backward(ctx, dwcat, dbcat=None):
        h, kin, din = ctx.shape
        dwcat = dwcat.contiguous()
        dw = torch.empty((h, kin), dtype=torch.float32, device=dwcat.device)
        db = None
        if ctx.has_bias and dbcat is not None:
            dbcat = dbcat.contiguous()
            db = torch.empty((h,), dtype=torch.float32, device=dwcat.device)
        _abi.call("stinet_edgeconv_hoist_bwd", dwcat.data_ptr(), _ptr(dbcat) if db is not None else None, h, din,
                  int(ctx.trans_inv), dw.data_ptr(), kin, _ptr(db), _stream(), cost=(4 * h * (kin + 2 * din), 0, ""))
        return dw, db, None


def edgeconv_hoist(w, b, trans_inv: bool):
    """-> (Wcat [2H, din], bcat [2H] or None)"""
    if b is None:
        return EdgeConvHoistFn.apply(w, None, trans_inv), None
    return EdgeConvHoistFn.apply(w, b, trans_inv)


class AggregateFn(Function):
    """out[i] = reduce_{j->i} x[j]  (mean / add / max over in-edges)."""

    @staticmethod
    def forward(ctx, x, csr: EdgeCSR, reduce: str):
        x = _mat(x)
        n, c = x.shape
        assert n == csr.n
        r = REDUCE[reduce]
        out = torch.empty((n, c), dtype=torch.float32, device=x.device)
        arg = torch.empty((n, c), dtype=torch.int32, device=x.device) if r == 2 else None
        _abi.call("stinet_aggregate_fwd", x.data_ptr(), _ld(x), csr.rowptr_t.data_ptr(), csr.col_t.data_ptr(),
                  csr.eid_t.data_ptr(), n, csr.e, c, r, out.data_ptr(), c, _ptr(arg), _stream(),
                  cost=(csr.e * (4 * c + 4) + n * (4 * c + 4), csr.e * c, f"C{c}"))
        ctx.csr, ctx.r, ctx.arg = csr, r, arg
        if arg is not None:
            ctx.mark_non_differentiable(arg)
            return out, arg
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g, *_):
        csr = ctx.csr
        g = _mat(g)
        n, c = g.shape
        rowptr_s, col_s, eid_s = csr.by_source()
        dx = torch.empty((n, c), dtype=torch.float32, device=g.device)
        _abi.call("stinet_aggregate_bwd", g.data_ptr(), _ld(g), rowptr_s.data_ptr(), col_s.data_ptr(),
                  eid_s.data_ptr(), csr.rowptr_t.data_ptr(), _ptr(ctx.arg), n, c, ctx.r, dx.data_ptr(), c, _stream(),
                  cost=(csr.e * (4 * c + 8) + n * (4 * c + 4), csr.e * c, f"C{c}"))
        return dx, None, None


def aggregate(x, csr, reduce="mean"):
    return AggregateFn.apply(x, csr, reduce)


# ------------------------------------------------------------------------------------------------------------
# pooling / unpooling


class PoolMaxFn(Function):
    @staticmethod
    def forward(ctx, x, cl: ClusterCSR):
        x = _mat(x)
        n, c = x.shape
        assert n == cl.n_fine
        out = torch.empty((cl.n_coarse, c), dtype=torch.float32, device=x.device)
        arg = torch.empty((cl.n_coarse, c), dtype=torch.int32, device=x.device)
        _abi.call("stinet_pool_max_fwd", x.data_ptr(), _ld(x), cl.rowptr.data_ptr(), cl.member.data_ptr(), n,
                  cl.n_coarse, c, out.data_ptr(), c, arg.data_ptr(), _stream(),
                  cost=(n * (4 * c + 4) + cl.n_coarse * 8 * c, 0, f"C{c}"))
        ctx.cl, ctx.arg = cl, arg
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    @once_differentiable
    def backward(ctx, g, _garg):
        cl = ctx.cl
        g = _mat(g)
        c = g.shape[1]
        dx = torch.empty((cl.n_fine, c), dtype=torch.float32, device=g.device)
        _abi.call("stinet_pool_max_bwd", g.data_ptr(), _ld(g), ctx.arg.data_ptr(), cl.trace32.data_ptr(), cl.n_fine,
                  c, dx.data_ptr(), c, _stream(), cost=(cl.n_coarse * 8 * c + cl.n_fine * (4 * c + 4), 0, f"C{c}"))
        return dx, None


class PoolMeanFn(Function):
    @staticmethod
    def forward(ctx, x, cl: ClusterCSR):
        x = _mat(x)
        n, c = x.shape
        assert n == cl.n_fine
        out = torch.empty((cl.n_coarse, c), dtype=torch.float32, device=x.device)
        _abi.call("stinet_pool_mean_fwd", x.data_ptr(), _ld(x), cl.rowptr.data_ptr(), cl.member.data_ptr(),
                  cl.n_coarse, c, out.data_ptr(), c, _stream(),
                  cost=(n * (4 * c + 4) + cl.n_coarse * 4 * c, 0, f"C{c}"))
        ctx.cl = cl
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        cl = ctx.cl
        g = _mat(g)
        c = g.shape[1]
        dx = torch.empty((cl.n_fine, c), dtype=torch.float32, device=g.device)
        _abi.call("stinet_pool_mean_bwd", g.data_ptr(), _ld(g), cl.rowptr.data_ptr(), cl.trace32.data_ptr(),
                  cl.n_fine, c, dx.data_ptr(), c, _stream(),
                  cost=(cl.n_coarse * 4 * c + cl.n_fine * (4 * c + 4), 0, f"C{c}"))
        return dx, None


class UnpoolFn(Function):
    @staticmethod
    def forward(ctx, xc, cl: ClusterCSR):
        xc = _mat(xc)
        nc, c = xc.shape
        assert nc == cl.n_coarse
        out = torch.empty((cl.n_fine, c), dtype=torch.float32, device=xc.device)
        _abi.call("stinet_unpool_fwd", xc.data_ptr(), _ld(xc), cl.trace32.data_ptr(), cl.n_fine, c, out.data_ptr(), c,
                  _stream(), cost=(cl.n_fine * (4 + 8 * c), 0, f"C{c}"))
        ctx.cl = cl
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        cl = ctx.cl
        g = _mat(g)
        c = g.shape[1]
        dxc = torch.empty((cl.n_coarse, c), dtype=torch.float32, device=g.device)
        _abi.call("stinet_unpool_bwd", g.data_ptr(), _ld(g), cl.rowptr.data_ptr(), cl.member.data_ptr(), cl.n_coarse,
                  c, dxc.data_ptr(), c, _stream(), cost=(cl.n_fine * (4 * c + 4) + cl.n_coarse * 4 * c, 0, f"C{c}"))
        return dxc, None


class UnpoolConcatFn(Function):
    """out = [skip || coarse[trace]]  (reference models/singleconvmeshnet.py:140-141): the row gather writes into the column
    slice [Ca, Ca+Cb) of the pre-allocated buffer (`ldo` of stinet_unpool_fwd); backward reads the same slice in place."""

    @staticmethod
    def forward(ctx, skip, xc, cl: ClusterCSR):
        skip, xc = _mat(skip), _mat(xc)
        n, ca = skip.shape
        nc, cb = xc.shape
        assert n == cl.n_fine and nc == cl.n_coarse
        out = torch.empty((n, ca + cb), dtype=torch.float32, device=xc.device)
        out[:, :ca].copy_(skip)
        _abi.call("stinet_unpool_fwd", xc.data_ptr(), _ld(xc), cl.trace32.data_ptr(), n, cb, out.data_ptr() + 4 * ca, ca + cb,
                  _stream(), cost=(n * (4 + 8 * cb), 0, f"C{cb}"))
        ctx.cl, ctx.ca, ctx.cb = cl, ca, cb
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        cl, ca, cb = ctx.cl, ctx.ca, ctx.cb
        g = _mat(g)
        dskip = g[:, :ca] if ctx.needs_input_grad[0] else None
        dxc = None
        if ctx.needs_input_grad[1]:
            dxc = torch.empty((cl.n_coarse, cb), dtype=torch.float32, device=g.device)
            _abi.call("stinet_unpool_bwd", g.data_ptr() + 4 * ca, _ld(g), cl.rowptr.data_ptr(), cl.member.data_ptr(),
                      cl.n_coarse, cb, dxc.data_ptr(), cb, _stream(),
                      cost=(cl.n_fine * (4 * cb + 4) + cl.n_coarse * 4 * cb, 0, f"C{cb}"))
        return dskip, dxc, None


def unpool_concat(skip, xc, cl):
    return UnpoolConcatFn.apply(skip, xc, cl)


def _carry_amax(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """max / mean pooling and row gathers cannot raise max|x|: the source's bound serves the result's operand planes."""
    known = getattr(src, "_stinet_amax", None)
    if known is not None and known[0] == src._version:
        set_amax(dst, known[1])
    return dst


def pool_max(x, cl):
    out, arg = PoolMaxFn.apply(x, cl)
    return _carry_amax(x, out), arg


def pool_mean(x, cl):
    return _carry_amax(x, PoolMeanFn.apply(x, cl))


class PoolSumFn(Function):
    """out[c] = sum of the rows of cluster c (in member order, deterministic): the adjoint pair of the row gather --
    forward is stinet_unpool_bwd, backward stinet_unpool_fwd.  PyG's aggr='add' over per-edge messages."""

    @staticmethod
    def forward(ctx, x, cl: ClusterCSR):
        x = _mat(x)
        n, c = x.shape
        assert n == cl.n_fine
        out = torch.empty((cl.n_coarse, c), dtype=torch.float32, device=x.device)
        _abi.call("stinet_unpool_bwd", x.data_ptr(), _ld(x), cl.rowptr.data_ptr(), cl.member.data_ptr(), cl.n_coarse, c,
                  out.data_ptr(), c, _stream(), cost=(cl.n_fine * (4 * c + 4) + cl.n_coarse * 4 * c, 0, f"C{c}"))
        ctx.cl = cl
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        cl = ctx.cl
        g = _mat(g)
        c = g.shape[1]
        dx = torch.empty((cl.n_fine, c), dtype=torch.float32, device=g.device)
        _abi.call("stinet_unpool_fwd", g.data_ptr(), _ld(g), cl.trace32.data_ptr(), cl.n_fine, c, dx.data_ptr(), c,
                  _stream(), cost=(cl.n_fine * (4 + 8 * c), 0, f"C{c}"))
        return dx, None


def pool_sum(x, cl):
    return PoolSumFn.apply(x, cl)


def unpool(xc, cl):
    return _carry_amax(xc, UnpoolFn.apply(xc, cl))


# ------------------------------------------------------------------------------------------------------------
# per-graph instance norm fused with activation + residual


class NormActResFn(Function):
    """out = residual + act(instance_norm(x; segments))   (use_norm=False: identity norm)."""
    last_amax = None      # device scalar max|out| of the most recent forward (picked up by norm_act_res)
    last_planes = None    # fp16 operand planes of the most recent forward's output, when the kernels wrote them
    emit_planes = True

    @staticmethod
    def forward(ctx, x, residual, seg: Optional[Segments], use_norm: bool, act: int, eps: float):
        x = _mat(x)
        n, c = x.shape
        dev = x.device
        mean = rstd = None
        res = _mat(residual) if residual is not None else None
        out = torch.empty((n, c), dtype=torch.float32, device=dev)
        nres = 1 if res is not None else 0
        if use_norm:
            assert seg is not None and seg.n_rows == n
            mean = torch.empty((seg.n_seg, c), dtype=torch.float32, device=dev)
            rstd = torch.empty((seg.n_seg, c), dtype=torch.float32, device=dev)
            nb = _abi.query("stinet_segnorm_workspace_bytes", seg.max_seg_rows, c, seg.n_seg)
            ws = _ws(nb, dev)
        amax = None
        if use_norm and seg.consistent and (seg.n_seg == 1 or _vec_ok(c, x, res)):
            # slices are the graphs: one entry point (a single cluster kernel when the slices are short); the kernels also
            # leave max|out| behind: the next dense layer's operand planes are scaled with it
            amax = torch.empty(1, dtype=torch.float32, device=dev)
            # ... and, when max|residual| is known, the result also as fp16 operand planes (scale from the bound
            # max|residual| + sqrt(rows)): the block that reads it next splits nothing
            planes = None
            res_amax = None
            if residual is not None:
                known = getattr(residual, "_stinet_amax", None)
                res_amax = known[1] if (known is not None and known[0] == residual._version) else None
            if _vec_ok(c, x, res) and c % 4 == 0 and (residual is None or res_amax is not None) and NormActResFn.emit_planes:
                planes = _new_planes(n, c, True, dev)
            _abi.call("stinet_segnorm_fwd", x.data_ptr(), _ld(x), n, c, seg.n_seg, seg.max_seg_rows,
                      seg.slice_ptr.data_ptr(), seg.cnt.data_ptr(), float(eps), _ptr(res),
                      _ld(res) if res is not None else 0, act, out.data_ptr(), c, mean.data_ptr(), rstd.data_ptr(),
                      amax.data_ptr(), _ptr(planes.hi) if planes else None, _ptr(planes.lo) if planes else None,
                      planes.ld if planes else 0, _ptr(res_amax), _ptr(planes.exp) if planes else None,
                      ws.data_ptr(), nb, _stream(), cost=(4 * n * c * (3 + nres + (1 if planes else 0)), 7 * n * c, f"C{c}"))
            NormActResFn.last_planes = planes
        else:
            if use_norm:
                # the reference's linspace slices cut across graphs (ragged batch), or rows of an odd width: sums by
                # slice, lookups by graph id
                _abi.call("stinet_segnorm_stats", x.data_ptr(), _ld(x), n, c, seg.n_seg, seg.max_seg_rows,
                          seg.slice_ptr.data_ptr(), seg.cnt.data_ptr(), _ptr(seg.gid), float(eps), mean.data_ptr(),
                          rstd.data_ptr(), ws.data_ptr(), nb, _stream(), cost=(8 * n * c, 3 * n * c, f"C{c}"))
            gid = seg.gid if (use_norm and seg is not None) else None
            _abi.call("stinet_segnorm_apply", x.data_ptr(), _ld(x), n, c, _ptr(gid), _ptr(mean), _ptr(rstd), _ptr(res),
                      _ld(res) if res is not None else 0, act, out.data_ptr(), c, _stream(),
                      cost=(4 * n * c * (2 + nres), 4 * n * c, f"C{c}"))
        ctx.save_for_backward(x, mean, rstd)
        ctx.seg, ctx.use_norm, ctx.act, ctx.has_res = seg, use_norm, act, residual is not None
        NormActResFn.last_amax = amax
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        x, mean, rstd = ctx.saved_tensors
        seg = ctx.seg
        dout_in = dout
        dout = _mat(dout)
        n, c = x.shape
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((n, c), dtype=torch.float32, device=x.device)
            if ctx.use_norm:
                if not seg.consistent:
                    dx = _ragged_norm_backward(x, dout, mean, rstd, seg, ctx.act)
                    dres = dout if (ctx.has_res and ctx.needs_input_grad[1]) else None
                    return dx, dres, None, None, None, None
                nb = _abi.query("stinet_segnorm_workspace_bytes", seg.max_seg_rows, c, seg.n_seg)
                ws = _ws(nb, x.device)
                amax = torch.empty(2, dtype=torch.float32, device=x.device)
                # dout is also the gradient of the shortcut branch: its Linear's backward reads max|dout| from here
                want_dmax = ctx.has_res and ctx.needs_input_grad[1] and getattr(dout, "_stinet_amax", None) is None
                _abi.call("stinet_segnorm_bwd", x.data_ptr(), _ld(x), dout.data_ptr(), _ld(dout), n, c, seg.n_seg,
                          seg.max_seg_rows, seg.slice_ptr.data_ptr(), seg.cnt.data_ptr(),
                          None if (seg.n_seg == 1 or _vec_ok(c, x, dout)) else _ptr(seg.gid),
                          mean.data_ptr(), rstd.data_ptr(), ctx.act, dx.data_ptr(), c, amax.data_ptr(),
                          amax.data_ptr() + 4 if want_dmax else None, ws.data_ptr(), nb,
                          _stream(), cost=(16 * n * c, 10 * n * c, f"C{c}"))
                set_amax(dx, amax[0:1])                  # the conv's backward splits dx into planes next
                if want_dmax:
                    set_amax(dout_in, amax[1:2])
            else:
                _abi.call("stinet_segnorm_bwd", x.data_ptr(), _ld(x), dout.data_ptr(), _ld(dout), n, c, 1, n, None,
                          None, None, None, None, ctx.act, dx.data_ptr(), c, None, None, None, 0, _stream())
        dres = dout_in if (ctx.has_res and ctx.needs_input_grad[1]) else None
        return dx, dres, None, None, None, None


def _ragged_norm_backward(x, dout, mean, rstd, seg: Segments, act: int):
    """Gradient of the reference's per-graph norm when its `linspace` slices (fastinstancenorm.py:53) cut across the
    true graph boundaries (graphs of different size in one batch).  There the statistics of graph g are sums over
    SLICE g divided by the true count of graph g, looked up through `batch` (:60-82), so the derivative couples rows
    of neighbouring graphs.  The reference never trains in this regime (3D: batch 1; 2D: equal-size images), so this is
    a short composite of device tensor ops (segment sums over a handful of contiguous row ranges) rather than a
    dedicated kernel; the equal-size case runs on stinet_segnorm_bwd.
        y_i = (x_i - m[g_i]) r[g_i],  m[g] = sum_{slice g} x / cnt_g,  v[g] = sum_{slice g} (x - m[gid])^2 / cnt_g"""
    tp, sp = seg.true_ptr, seg.slice_ptr_host
    B = seg.n_seg
    gid = seg.gid.long()
    slice_len = torch.tensor([sp[b + 1] - sp[b] for b in range(B)], device=x.device)
    m, r = mean.index_select(0, gid), rstd.index_select(0, gid)
    xc = x - m
    if act == ACT_ELU:
        v = xc * r
        dy = dout * torch.where(v > 0, torch.ones_like(v), torch.exp(v))
    else:
        dy = dout
    cnt = seg.cnt.view(-1, 1)

    def by_graph(t):          # sums over the rows whose graph id is g (contiguous: collate keeps graphs in order)
        return torch.stack([t[tp[b]:tp[b + 1]].sum(0) for b in range(B)])

    def by_slice(table):      # value of slice s(i) for every row i
        return torch.repeat_interleave(table, slice_len, dim=0)

    dr = by_graph(dy * xc)                                   # dL/dr[g]
    dv = -0.5 * dr * rstd ** 3                               # dL/dv[g]
    dxc = dy * r + 2.0 * xc * by_slice(dv / cnt)             # total derivative w.r.t. xc_i
    dm = -by_graph(dxc)                                      # dL/dm[g]
    return dxc + by_slice(dm / cnt)


def norm_act_res(x, residual, seg, use_norm=True, act=ACT_ELU, eps=1e-5):
    out = NormActResFn.apply(x, residual, seg, use_norm, act, eps)
    if NormActResFn.last_amax is not None:
        set_amax(out, NormActResFn.last_amax)        # max|out| came with the norm kernels: no reduction pass later
        NormActResFn.last_amax = None
    if NormActResFn.last_planes is not None:         # ... and so did its operand planes: no split pass either
        out._stinet_planes = (out._version, _plane_epoch, NormActResFn.last_planes)
        NormActResFn.last_planes = None
    return out


# ------------------------------------------------------------------------------------------------------------
# affine segmented norms: BatchNorm over rows, SingleBatchGraphNorm


class AffNormFn(Function):
    """y = gamma * (x - alpha * m[g]) * r[g] + beta with slice statistics (kind 0: centred variance -- BatchNorm; kind 1:
    second moment of x itself -- the reference's SingleBatchGraphNorm).  Returns (y, mean, rstd); the statistics are
    non-differentiable by-products (BatchNorm's running-stat update reads them)."""

    @staticmethod
    def forward(ctx, x, alpha, gamma, beta, slice_ptr, cnt, gid, n_seg: int, max_seg_rows: int, kind: int, eps: float):
        x = _mat(x)
        n, c = x.shape
        dev = x.device
        out = torch.empty((n, c), dtype=torch.float32, device=dev)
        mean = torch.empty((n_seg, c), dtype=torch.float32, device=dev)
        rstd = torch.empty((n_seg, c), dtype=torch.float32, device=dev)
        nb = _abi.query("stinet_affnorm_workspace_bytes", max_seg_rows, c, n_seg)
        ws = _ws(nb, dev)
        _abi.call("stinet_affnorm_fwd", x.data_ptr(), _ld(x), n, c, n_seg, max_seg_rows, slice_ptr.data_ptr(), cnt.data_ptr(),
                  _ptr(gid), kind, float(eps), _ptr(alpha), _ptr(gamma), _ptr(beta), out.data_ptr(), c, mean.data_ptr(),
                  rstd.data_ptr(), ws.data_ptr(), nb, _stream(), cost=(4 * n * c * 4, 8 * n * c, f"C{c}"))
        ctx.save_for_backward(x, mean, rstd, alpha, gamma, slice_ptr, cnt, gid)
        ctx.cfg = (n_seg, max_seg_rows, kind)
        ctx.mark_non_differentiable(mean, rstd)
        return out, mean, rstd

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, _dm, _dr):
        x, mean, rstd, alpha, gamma, slice_ptr, cnt, gid = ctx.saved_tensors
        n_seg, max_seg_rows, kind = ctx.cfg
        dy = _mat(dy)
        n, c = x.shape
        dev = x.device
        need = ctx.needs_input_grad
        dx = torch.empty((n, c), dtype=torch.float32, device=dev) if need[0] else None
        dal = torch.empty((c,), dtype=torch.float32, device=dev) if (alpha is not None and need[1]) else None
        dga = torch.empty((c,), dtype=torch.float32, device=dev) if (gamma is not None and need[2]) else None
        dbe = torch.empty((c,), dtype=torch.float32, device=dev) if need[3] else None
        nb = _abi.query("stinet_affnorm_workspace_bytes", max_seg_rows, c, n_seg)
        ws = _ws(nb, dev)
        _abi.call("stinet_affnorm_bwd", x.data_ptr(), _ld(x), dy.data_ptr(), _ld(dy), n, c, n_seg, max_seg_rows,
                  slice_ptr.data_ptr(), cnt.data_ptr(), _ptr(gid), kind, mean.data_ptr(), rstd.data_ptr(), _ptr(alpha),
                  _ptr(gamma), _ptr(dx), c, _ptr(dga), _ptr(dbe), _ptr(dal), ws.data_ptr(), nb, _stream(),
                  cost=(4 * n * c * 5, 12 * n * c, f"C{c}"))
        return dx, dal, dga, dbe, None, None, None, None, None, None, None


class AffNormEvalFn(Function):
    """y = gamma * (x - mean) * rstd + beta with GIVEN statistics (BatchNorm in eval mode)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, mean, rstd):
        x = _mat(x)
        n, c = x.shape
        out = torch.empty((n, c), dtype=torch.float32, device=x.device)
        _abi.call("stinet_affnorm_apply", x.data_ptr(), _ld(x), n, c, None, mean.data_ptr(), rstd.data_ptr(), None,
                  _ptr(gamma), _ptr(beta), out.data_ptr(), c, _stream(), cost=(8 * n * c, 3 * n * c, f"C{c}"))
        ctx.save_for_backward(x, gamma, mean, rstd)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        # statistics are constants here: three short tensor expressions (eval-mode gradients are not on the hot path)
        x, gamma, mean, rstd = ctx.saved_tensors
        xh = (x - mean) * rstd
        g = gamma if gamma is not None else 1.0
        return dy * (g * rstd), ((dy * xh).sum(0) if gamma is not None else None), dy.sum(0), None, None


def batch_norm(x, weight, bias, running_mean, running_var, training: bool, momentum: float, eps: float, updates: int = 1):
    """nn.BatchNorm1d over the rows of x [N, C] on the segmented-reduction kernels (deterministic; replaces
    F.batch_norm).  Updates the running statistics in place when training (unbiased variance, like torch); `updates` = 2
    reproduces a block the reference wraps in torch.utils.checkpoint, whose forward -- and momentum update -- runs twice."""
    n, c = x.shape
    if not training and running_mean is not None:
        rstd = torch.rsqrt(running_var + eps).reshape(1, c).contiguous()
        return AffNormEvalFn.apply(x, weight, bias, running_mean.reshape(1, c).contiguous(), rstd)
    from .graph import _segment_tables
    slice_ptr, cnt = _segment_tables((0, n), (max(n, 1),), x.device)
    out, mean, rstd = AffNormFn.apply(x, None, weight, bias, slice_ptr, cnt, None, 1, n, 0, eps)
    if training and running_mean is not None:
        for _ in range(updates):
            _abi.call("stinet_bn_running_update", mean.data_ptr(), rstd.data_ptr(), n, float(eps), float(momentum), c,
                      running_mean.data_ptr(), running_var.data_ptr(), _stream())
    return out


class BatchNorm1d(torch.nn.BatchNorm1d):
    """nn.BatchNorm1d (same parameters, buffers and state_dict keys) evaluated by ops.batch_norm on CUDA inputs."""

    updates_per_step = 1      # 2: the reference runs this module inside torch.utils.checkpoint (forward + recomputation)

    def forward(self, input):
        if not input.is_cuda or input.dim() != 2 or self.momentum is None:
            return super().forward(input)
        training = self.training or self.running_mean is None
        # the recomputation only happens when a backward pass will run
        updates = self.updates_per_step if (torch.is_grad_enabled() and input.requires_grad) else 1
        if self.training and self.track_running_stats and self.num_batches_tracked is not None:
            self.num_batches_tracked.add_(updates)
        return batch_norm(input, self.weight, self.bias, self.running_mean if self.track_running_stats else None,
                          self.running_var if self.track_running_stats else None, training, self.momentum, self.eps, updates)


# ------------------------------------------------------------------------------------------------------------
# network tail and trainer loss


class HeadTanhFn(Function):
    """out = tanh(h W^T + b) for the 3-channel output head (final_linear2 + Tanh): one kernel, the Linear in registers."""

    @staticmethod
    def forward(ctx, h, weight, bias):
        h, weight = _mat(h), weight.contiguous()
        n, c = h.shape
        out = torch.empty((n, 3), dtype=torch.float32, device=h.device)
        _abi.call("stinet_head_fwd", h.data_ptr(), _ld(h), weight.data_ptr(), _ptr(bias), n, c, out.data_ptr(), _stream(),
                  cost=(4 * n * (c + 3), 6 * n * c, f"C{c}"))
        ctx.save_for_backward(h, weight, out)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        h, weight, out = ctx.saved_tensors
        n, c = h.shape
        dev = h.device
        dout = dout.contiguous()
        dh = torch.empty((n, c), dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        dw = torch.empty((3, c), dtype=torch.float32, device=dev)
        db = torch.empty((3,), dtype=torch.float32, device=dev) if ctx.has_bias else None
        nb = _abi.query("stinet_head_workspace_bytes", n, c)
        ws = _ws(nb, dev)
        _abi.call("stinet_head_bwd", h.data_ptr(), _ld(h), weight.data_ptr(), out.data_ptr(), dout.data_ptr(), n, c, _ptr(dh), c,
                  dw.data_ptr(), _ptr(db), ws.data_ptr(), nb, _stream(), cost=(4 * n * (2 * c + 6), 12 * n * c, f"C{c}"))
        return dh, dw, db


def head_tanh(h, weight, bias):
    """The fused head when the shapes allow it (3 outputs, C a power of two in 8..256, aligned rows), else None."""
    c = h.shape[1]
    if weight.shape[0] != 3 or c not in (8, 16, 32, 64, 128, 256) or not h.is_cuda:
        return None
    h = _mat(h)
    if _ld(h) % 4 or h.data_ptr() % 16:
        return None
    return HeadTanhFn.apply(h, weight, bias)


class MaskedL1Fn(Function):
    """loss = mean(|where(mask > 0, out, color) - color| * 0.99^mask)  (trainers/inpainting3d_trainer.py:127-137)."""

    @staticmethod
    def forward(ctx, out, color, mask):
        out, color = out.contiguous(), color.contiguous()
        n, c = out.shape
        m = mask.reshape(-1).to(torch.float32).contiguous()
        assert color.shape == out.shape and m.numel() == n
        loss = torch.empty((), dtype=torch.float32, device=out.device)
        nb = _abi.query("stinet_masked_l1_workspace_bytes", n)
        ws = _ws(nb, out.device)
        _abi.call("stinet_masked_l1_fwd", out.data_ptr(), color.data_ptr(), m.data_ptr(), n, c, loss.data_ptr(), ws.data_ptr(), nb,
                  _stream(), cost=(4 * n * (2 * c + 1), 4 * n * c, ""))
        ctx.save_for_backward(out, color, m)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        out, color, m = ctx.saved_tensors
        n, c = out.shape
        g = g.reshape(1).to(torch.float32).contiguous()
        dout = torch.empty_like(out)
        _abi.call("stinet_masked_l1_bwd", out.data_ptr(), color.data_ptr(), m.data_ptr(), g.data_ptr(), n, c, dout.data_ptr(),
                  _stream(), cost=(4 * n * (3 * c + 1), 4 * n * c, ""))
        return dout, None, None


def masked_l1_loss(out, color, mask):
    """The 3D trainer's training loss as one forward and one backward kernel (drop-in for its torch.where / L1Loss / pow /
    mean chain on CUDA tensors)."""
    return MaskedL1Fn.apply(out, color, mask)
