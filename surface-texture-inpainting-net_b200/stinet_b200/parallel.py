"""Data-parallel training plumbing: one process per GPU, each rank runs its own graph batch, gradients are
averaged with ONE collective family -- all-reduce over NCCL / NVLink 5 / NVSwitch.  The reference has
no distributed code at all (its trainers assert a single GPU, trainers/inpainting3d_trainer.py:25).

Gradients live in a few large flat fp32 buckets (parameters' .grad are views into them), so a bucket is reduced
in place with a single all_reduce and no packing copies.  Buckets are filled in reverse registration order
(decoder / output blocks first, which is the order backward produces them) and each all-reduce is launched from a
post-accumulate-grad hook as soon as the bucket is complete, so the transfer overlaps the remaining backward
kernels.  The mesh path itself has no data-path collective: graphs never span ranks.
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
    """Bucketed, overlapped gradient averaging for a replicated module."""

    def __init__(self, module: torch.nn.Module, bucket_bytes: int = 64 << 20, overlap: bool = True):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.overlap = overlap and self.world > 1
        self.buckets: List[torch.Tensor] = []
        self._bucket_of = {}
        self._pending: List[int] = []
        self._handles = []
        if self.world == 1:
            # nothing to reduce: leave .grad unset so autograd hands each weight gradient over without the
            # zero-fill + accumulate pass that pre-allocated bucket views cost (two launches per parameter)
            self._sizes = []
            return
        # reverse order: the last layers' gradients are produced first
        cur, cur_bytes = [], 0
        groups = []
        for p in reversed(self.params):
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= bucket_bytes:
                groups.append(cur)
                cur, cur_bytes = [], 0
        if cur:
            groups.append(cur)
        for b, group in enumerate(groups):
            flat = torch.zeros(sum(p.numel() for p in group), dtype=torch.float32, device=group[0].device)
            off = 0
            for p in group:
                p.grad = flat[off:off + p.numel()].view_as(p)       # autograd accumulates in place into the view
                off += p.numel()
                self._bucket_of[p] = b
            self.buckets.append(flat)
            self._pending.append(len(group))
        self._sizes = list(self._pending)
        if self.world > 1:
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._on_grad)

    def zero_grad(self):
        if self.world == 1:
            for p in self.params:
                p.grad = None
            return
        for flat in self.buckets:
            flat.zero_()
        self._pending = list(self._sizes)
        self._handles = []

    def _on_grad(self, p):
        if not self.overlap:                       # e.g. while a step is being captured into a CUDA graph
            return
        b = self._bucket_of[p]
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._handles.append(dist.all_reduce(self.buckets[b], op=dist.ReduceOp.SUM, async_op=True))

    def finish(self):
        """Call after backward(): waits for (or issues) the all-reduces and turns sums into means."""
        if self.world == 1:
            return
        if not self.overlap:
            self._handles = [dist.all_reduce(f, op=dist.ReduceOp.SUM, async_op=True) for f in self.buckets]
        for h in self._handles:
            h.wait()
        inv = 1.0 / self.world
        for flat in self.buckets:
            flat.mul_(inv)
        self._handles = []
