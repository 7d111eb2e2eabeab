"""Data-parallel training plumbing: one process per GPU, each rank runs its own graph batch, gradients are
averaged with ONE collective family -- all-reduce over NCCL / NVLink 5 / NVSwitch.  The reference has
no distributed code at all (its trainers assert a single GPU, trainers/inpainting3d_trainer.py:25).

Gradients are reduced in a few large flat fp32 buckets filled in reverse registration order (decoder / output blocks
first, which is the order backward produces them).  When the last gradient of a bucket has been produced, ONE
multi-tensor copy moves the bucket's fresh gradient tensors into the flat buffer (no pre-zeroing, no accumulate pass),
the parameters' .grad are re-pointed at the buffer's views, and the bucket's all-reduce is issued at once on NCCL's
stream, so the transfer overlaps the remaining backward kernels; finish() only waits.  The same sequence is what
GraphedTrainStep captures into its CUDA graph (NCCL collectives are graph-capturable), so a replayed step has the
collectives forked off the backward kernels exactly as the eager step has.  The mean is taken by scaling the LOSS with
1 / world (`loss_scale`) before backward, so the buckets need no extra pass after the sum.
The mesh path itself has no data-path collective: graphs never span ranks.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist


def init_distributed(backend: Optional[str] = None):
    """torchrun-style initialisation (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, local, world


class GradAllReducer:
    """Bucketed, overlapped gradient averaging for a replicated module.

        reducer = GradAllReducer(net)
        reducer.zero_grad()
        (loss_fn(net(batch)) * reducer.loss_scale).backward()     # hooks issue the all-reduces bucket by bucket
        reducer.finish()                                          # waits; p.grad now holds the mean over ranks
        optimizer.step()
    """

    def __init__(self, module: torch.nn.Module, bucket_bytes: int = 32 << 20, overlap: bool = True):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.loss_scale = 1.0 / self.world
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.overlap = overlap and self.world > 1
        self.enabled = True                         # False: hooks and finish() do nothing (side-effect-free warm-up)
        self.buckets: List[torch.Tensor] = []
        self._groups: List[list] = []
        self._views: List[list] = []
        self._bucket_of = {}
        self._pending: List[int] = []
        self._issued: List[bool] = []
        self._handles = []
        self._counted = set()
        if self.world == 1:
            return
        # reverse order: the last layers' gradients are produced first
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= bucket_bytes:
                self._groups.append(cur)
                cur, cur_bytes = [], 0
        if cur:
            self._groups.append(cur)
        for b, group in enumerate(self._groups):
            flat = torch.zeros(sum(p.numel() for p in group), dtype=torch.float32, device=group[0].device)
            views, off = [], 0
            for p in group:
                views.append(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
                self._bucket_of[p] = b
            self.buckets.append(flat)
            self._views.append(views)
        self._sizes = [len(g) for g in self._groups]
        self._pending = list(self._sizes)
        self._issued = [False] * len(self._groups)
        for p in self.params:
            p.register_post_accumulate_grad_hook(self._on_grad)
            p._stinet_wgrad_aware = True             # ops._straight_to_grad: _issue() joins the weight-gradient stream

    def zero_grad(self):
        """Gradients are handed over by autograd as fresh tensors (no zero-fill, no accumulate pass)."""
        for p in self.params:
            p.grad = None
        if self.world > 1:
            self._pending = list(self._sizes)
            self._issued = [False] * len(self._groups)
            self._handles = []
            self._counted = set()

    def _issue(self, b: int):
        """Pack bucket b (one multi-tensor copy; a parameter without a gradient contributes zeros), re-point the .grad of
        its parameters at the flat buffer and start the all-reduce."""
        from . import ops
        ops.join_wgrad_stream()                      # weight gradients still in flight on their own stream
        group, views = self._groups[b], self._views[b]
        have = [(v, p.grad) for v, p in zip(views, group) if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
        for v, p in zip(views, group):
            if p.grad is None:
                v.zero_()
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        for v, p in zip(views, group):
            p.grad = v
        self._issued[b] = True
        self._handles.append(dist.all_reduce(self.buckets[b], op=dist.ReduceOp.SUM, async_op=True))

    def _on_grad(self, p):
        if not (self.enabled and self.overlap):
            return
        # once per parameter and step: a node that deposits a weight gradient itself (ops._wgrad_done) calls this hook, and
        # autograd may call it again for the same parameter although it was handed no gradient
        if id(p) in self._counted:
            return
        self._counted.add(id(p))
        b = self._bucket_of[p]
        self._pending[b] -= 1
        if self._pending[b] == 0 and not self._issued[b]:
            self._issue(b)

    def finish(self):
        """Call after backward(): issues the all-reduce of every bucket the hooks have not started (overlap off, or a
        parameter that received no gradient this step) and waits for all of them."""
        if self.world == 1 or not self.enabled:
            return
        for b in range(len(self._groups)):
            if not self._issued[b]:
                self._issue(b)
        for h in self._handles:
            h.wait()
        self._handles = []

    # ---- collectives outside a captured graph (GraphedTrainStep with STINET_ALLREDUCE_IN_GRAPH=0) ------------------
    def repoint(self):
        """Point every .grad at its bucket view without packing or reducing (what a captured optimizer step must read)."""
        for views, group in zip(self._views, self._groups):
            for v, p in zip(views, group):
                p.grad = v

    def reduce_from(self, fresh: dict):
        """Pack the given gradient tensors (parameter -> tensor, e.g. the static outputs of a backward graph) into the
        buckets and all-reduce them; returns when the reduced means are in the bucket views."""
        if self.world == 1:
            return
        handles = []
        for b, (views, group) in enumerate(zip(self._views, self._groups)):
            have = [(v, fresh[p]) for v, p in zip(views, group) if fresh.get(p) is not None]
            for v, p in zip(views, group):
                if fresh.get(p) is None:
                    v.zero_()
            if have:
                torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
            handles.append(dist.all_reduce(self.buckets[b], op=dist.ReduceOp.SUM, async_op=True))
        for h in handles:
            h.wait()
