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
