"""Stage-2 (multi-scale AF extractor) loss block: reference afigan/engine/stage2_trainer.py:298-384 on pre-computed features.

  real_l = nearest-0.5x of the guide model's p_l on the HR image          (:302, F.interpolate(scale_factor=0.5) == x[:, :, ::2, ::2])
  fake_l = the AFI-FPN model's p_l on the 0.5x image                      (:303)
  D phase: d_l = BCE(D0(real_l), 1) + BCE(D0(fake_l.detach()), 0)         (:306-322)  -> backward into D only
  G phase: g_l = 1e-3 * BCE(D0(fake_l).detach(), 1) + L1(fake_l, real_l)  (:343-364)  -> added to the detector's loss dict; the adversarial
           term carries no gradient (App. D-1), D0(real_l) is evaluated for its BatchNorm side effect only.
Everything numerical runs in the library through the autograd-aware modules; torch only sums the scalars."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch

from ..losses import bce_with_logits_const, l1_loss
from .stage1 import _StepBase


def nearest_half(x: torch.Tensor) -> torch.Tensor:
    """F.interpolate(x, scale_factor=0.5) (nearest): out[i, j] = x[2i, 2j], size floor(H/2) -- a strided view, no copy."""
    h, w = x.size(2) // 2, x.size(3) // 2
    return x[:, :, : 2 * h : 2, : 2 * w : 2]


def crop_pair(a: torch.Tensor, b: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """_reshape_feature applied both ways (stage2_trainer.py:307-308, 491-496)."""
    h, w = min(a.size(2), b.size(2)), min(a.size(3), b.size(3))
    return a[:, :, :h, :w], b[:, :, :h, :w]


def stage2_discriminator_losses(D, guide_feats: Sequence[torch.Tensor], model_feats: Sequence[torch.Tensor]) -> Dict[str, torch.Tensor]:
    """d_loss_p2..: call sum(values()).backward() and step the D optimiser (stage2_trainer.py:304-341)."""
    stack = D.Discriminators[0]
    out = {}
    for lv, (hr, up) in enumerate(zip(guide_feats, model_feats), 2):
        real, fake = crop_pair(nearest_half(hr.detach()), up.detach())
        out[f"d_loss_p{lv}"] = bce_with_logits_const(stack(real), 1.0) + bce_with_logits_const(stack(fake), 0.0)
    return out


def stage2_generator_losses(D, guide_feats: Sequence[torch.Tensor], model_feats: Sequence[torch.Tensor]) -> Dict[str, torch.Tensor]:
    """g_loss_p2..: merged into the detector's loss dict (stage2_trainer.py:343-366)."""
    stack = D.Discriminators[0]
    out = {}
    for lv, (hr, up) in enumerate(zip(guide_feats, model_feats), 2):
        real, fake = crop_pair(nearest_half(hr.detach()), up)
        with torch.no_grad():
            logit_fake = stack(fake.detach())        # .detach()-ed in the reference: no gradient reaches the model through D
            stack(real)                              # BatchNorm running-stat side effect only
        adv = bce_with_logits_const(logit_fake, 1.0)
        out[f"g_loss_p{lv}"] = adv * 1e-3 + l1_loss(fake, real)
    return out


class Stage2Step(_StepBase):
    """The stage-2 discriminator / generator loss block (reference stage2_trainer.py:298-384) with GROUPED launches, the counterpart of
    `Stage1Step` for a detector that owns the AFI neck: all levels' real / fake discriminator calls of a phase run as grouped kernels through
    the C ABI (one launch per layer for the ten calls instead of ten autograd calls), D's gradients live in ONE flat buffer that is
    all-reduced once, and D's SGD runs in the library.  The generator side stays in torch autograd: `g_losses` returns `g_loss_p*` tensors
    (1e-3 * adv [no gradient, App. D-1] + L1 with gradient w.r.t. the model's features) that the caller adds to the detector's loss dict."""

    def __init__(self, D, lr: float = 1e-3, momentum: float = 0.9, weight_decay: float = 1e-4, weight_decay_norm: float = 0.0,
                 precision=None, process_group=None, distributed=None, overlap: bool = True):
        from .. import native as N
        dev = D.Discriminators[0]._params()[0].device
        self._init_common(dev, lr, momentum, weight_decay, weight_decay_norm, precision or D.Discriminators[0].precision or N.default_precision(), process_group,
                          overlap, False)
        self._init_d(D, distributed)
        self._pack_key = None
        self._refresh_packed()

    @staticmethod
    def _pairs(guide_feats, model_feats):
        """real_l = nearest-0.5x of the guide's p_l, fake_l = the model's p_l, cropped to the element-wise min size (:301-308, 491-496)."""
        return [crop_pair(nearest_half(hr.detach()), up) for hr, up in zip(guide_feats, model_feats)]

    @torch.no_grad()
    def d_phase(self, guide_feats: Sequence[torch.Tensor], model_feats: Sequence[torch.Tensor], apply_updates: bool = True) -> torch.Tensor:
        """D phase (:304-341): d_l = BCE(D0(real_l), 1) + BCE(D0(fake_l.detach()), 0) for every level in grouped launches, backward into D,
        all-reduce, SGD.  Returns the device row of d_loss_p* values (no host sync)."""
        from .. import native as N
        from ..functional import d_grad_struct
        import ctypes as C
        self._refresh_packed()
        self.losses.zero_()
        N.check(self.lib.afi_zero(self.d_acc.data_ptr(), self.d_acc.numel(), N.stream_ptr()))
        xs, tags = [], []
        for l, (real, fake) in enumerate(self._pairs(guide_feats, model_feats)):
            xs += [real, fake.detach()]                                   # D0(real) BEFORE D0(fake)  (:311-312)
            tags += [f"r{l}", f"f{l}"]
        self._d_phase(xs, tags, [1.0 if i % 2 == 0 else 0.0 for i in range(len(xs))], [0] * len(xs), True, True)
        gs = d_grad_struct(self.d_grads)
        N.check(self.lib.afi_d_unpack_grads(self.ctx, self.prec, self.d_acc.data_ptr(), C.byref(gs), 1.0, 0, N.stream_ptr()))
        self.d_sync.all_reduce()
        if apply_updates:
            self._d_update()
            self.steps_done += 1
        return self.losses[0]

    def g_losses(self, guide_feats: Sequence[torch.Tensor], model_feats: Sequence[torch.Tensor]) -> Dict[str, torch.Tensor]:
        """G phase (:343-366): g_loss_p{l} = 1e-3 * BCE(D0(fake_l).detach(), 1) + L1(fake_l, real_l); D0(fake) before D0(real), the latter
        for its BatchNorm side effect only.  The adversarial values come out of one grouped forward; L1 keeps its autograd path."""
        pairs = self._pairs(guide_feats, model_feats)
        with torch.no_grad():
            self.losses[2].zero_()
            xs, tags = [], []
            for l, (real, fake) in enumerate(pairs):
                xs += [fake.detach(), real]                               # D0(fake) BEFORE D0(real) here (:350-351)
                tags += [f"f{l}", f"r{l}"]
            self._d_phase(xs, tags, [1.0] * len(xs), [2 if i % 2 == 0 else None for i in range(len(xs))], False, False)
            adv = self.losses[2].clone()
        out = {}
        for l, (real, fake) in enumerate(pairs):
            out[f"g_loss_p{l + 2}"] = adv[l] * 1e-3 + l1_loss(fake, real)
        return out

    def metrics(self, n_levels: int = 5) -> Dict[str, float]:
        v = self.losses.cpu()
        return {f"d_loss_p{l + 2}": float(v[0, l]) for l in range(n_levels)} | {f"adv_loss_p{l + 2}": float(v[2, l]) for l in range(n_levels)}
