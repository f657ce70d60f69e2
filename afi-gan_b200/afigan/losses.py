"""Autograd wrappers of the loss kernels (nn.BCEWithLogitsLoss against a constant target, F.l1_loss), for code paths that keep torch
autograd in charge (stage 2: the generator-side losses are added to the detector's loss dict and back-propagated through the whole model,
reference stage2_trainer.py:343-384)."""
from __future__ import annotations

import torch

from . import native as N


class _BCEConstTargetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits: torch.Tensor, target: float):
        if not logits.is_cuda:
            raise RuntimeError("afigan.losses: tensors must live on an sm_100a CUDA device (no CPU fallback)")
        lg = logits.float().contiguous()
        out = torch.zeros((), dtype=torch.float32, device=lg.device)
        dl = torch.empty_like(lg) if ctx.needs_input_grad[0] else None
        N.check(N.lib().afi_bce_with_logits(lg.data_ptr(), lg.numel(), float(target), out.data_ptr(), None, 0.0, N.ptr(dl), 1.0, N.stream_ptr()))
        ctx.dl, ctx.shape = dl, logits.shape
        return out

    @staticmethod
    def backward(ctx, g):
        return (ctx.dl * g).view(ctx.shape), None


class _L1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a: torch.Tensor, b: torch.Tensor):
        if not a.is_cuda:
            raise RuntimeError("afigan.losses: tensors must live on an sm_100a CUDA device (no CPU fallback)")
        a, b = a.float(), b.float()
        n, c, h, w = a.shape
        out = torch.zeros((), dtype=torch.float32, device=a.device)
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        da = torch.empty((n, c, h, w), dtype=torch.float32, device=a.device) if need else None
        N.check(N.lib().afi_l1_loss(N.view4(a), N.view4(b), n, c, h, w, out.data_ptr(), None, 0.0, N.ptr(da), 1.0, N.stream_ptr()))
        ctx.da = da
        return out

    @staticmethod
    def backward(ctx, g):
        ga = ctx.da * g if ctx.needs_input_grad[0] else None
        gb = -ctx.da * g if ctx.needs_input_grad[1] else None
        return ga, gb


def bce_with_logits_const(logits: torch.Tensor, target: float) -> torch.Tensor:
    """nn.BCEWithLogitsLoss()(logits, full_like(logits, target))  (stage1_trainer.py:154, 355-359)."""
    return _BCEConstTargetFn.apply(logits, float(target))


def l1_loss(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """F.l1_loss(a, b) on [N,C,H,W] tensors or strided views (stage1_trainer.py:410)."""
    return _L1Fn.apply(a, b)
