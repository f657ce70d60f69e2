"""Gradient synchronisation of the data-parallel path: ONE flat fp32 buffer per optimiser whose slices are the parameters' .grad, summed
across ranks with a single all-reduce (NCCL over NVLink on the GPUs, gloo in the CPU tests); the 1/world_size of the mean is folded into the
SGD kernel's grad_scale.  This is the DDP semantics the reference intended but does not execute in stage 1 (it calls `.module(...)`, which
bypasses the DDP reducer: stage1_trainer.py:80-89, 339-350)."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


class FlatGradSync:
    def __init__(self, params: Sequence[torch.nn.Parameter], process_group=None, enabled: Optional[bool] = None):
        self.params = list(params)
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.views: List[torch.Tensor] = []
        off = 0
        for p in self.params:
            v = self.flat[off:off + p.numel()].view_as(p)
            p.grad = v
            self.views.append(v)
            off += p.numel()
        self.pg = process_group
        active = dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1
        import os
        if os.environ.get("AFIGAN_NO_GRAD_SYNC") == "1":      # measurement aid: independent replicas (isolates rank skew from collective cost)
            active = False
        self.enabled = active if enabled is None else enabled
        self.world = dist.get_world_size(process_group) if self.enabled else 1
        self._work = None
        if self.enabled:
            self.broadcast_parameters()

    def broadcast_parameters(self, extra: Sequence[torch.Tensor] = ()):
        """What DistributedDataParallel does at construction (reference stage1_trainer.py:80-89): rank 0's parameters -- and any `extra`
        tensors, e.g. restored momentum buffers -- replace every other rank's, so that averaged gradients are applied to IDENTICAL
        replicas (detectron2 seeds each rank differently).  BatchNorm buffers are left alone: the reference passes broadcast_buffers=False."""
        if not self.enabled:
            return
        for t in list(self.params) + list(extra):
            dist.broadcast(t.data if isinstance(t, torch.nn.Parameter) else t, src=dist.get_global_rank(self.pg, 0) if self.pg is not None else 0,
                           group=self.pg)

    @property
    def grad_scale(self) -> float:
        return 1.0 / self.world

    def all_reduce(self, async_op: bool = False):
        if self.enabled:
            return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=async_op)
        return None

    def start(self):
        """Launch the all-reduce without blocking the issuing stream (NCCL runs it on its own stream once the work queued so far is done);
        `finish()` makes the current stream wait for it.  Kernels issued in between overlap the collective."""
        self._work = self.all_reduce(async_op=True)

    def finish(self):
        if self._work is not None:
            self._work.wait()
            self._work = None
