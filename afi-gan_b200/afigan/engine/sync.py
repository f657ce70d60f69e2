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
        self.enabled = active if enabled is None else enabled
        self.world = dist.get_world_size(process_group) if self.enabled else 1

    @property
    def grad_scale(self) -> float:
        return 1.0 / self.world

    def all_reduce(self, async_op: bool = False):
        if self.enabled:
            return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=async_op)
        return None
