"""Feature-patch discriminator: drop-in for reference afigan/modeling/feat_interpol/feature_patch_discriminator.py.

State dict = SURVEY.md App. B (Discriminators.0.{0,1,2}.0.{weight,bias,norm.*}, Discriminators.0.3.0.{weight,bias}),
init = c2_msra_fill on every conv in construction order (feature_patch_discriminator.py:43-46).  The trainers call
`D.Discriminators[0](x)` directly (stage1_trainer.py:349-353), so the CUDA path hangs off that Sequential.
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import nn

from .. import _holder
from ... import native
from ...functional import PatchDiscriminatorFn


class _ConvNorm(nn.Conv2d):
    """Parameter holder shaped like detectron2.layers.Conv2d(..., norm=BN): bias stays even with a norm."""

    def __init__(self, cin, cout, norm: Optional[nn.Module]):
        super().__init__(cin, cout, kernel_size=3, stride=1, padding=1)
        self.norm = norm
        self.activation = None


class PatchDiscriminatorStack(nn.Sequential):
    def __init__(self, in_filters: int = 256, precision: Optional[str] = None):
        blocks, f_mult = [], 1
        for n in range(1, 4):
            f_prev, f_mult = f_mult, min(2 ** n, 4)
            norm = nn.BatchNorm2d(in_filters * f_mult)
            blocks.append(nn.Sequential(_ConvNorm(in_filters * f_prev, in_filters * f_mult, norm), nn.LeakyReLU(0.2, True)))
        blocks.append(nn.Sequential(_ConvNorm(in_filters * f_mult, 1, None)))
        super().__init__(*blocks)
        for blk in self:
            conv = blk[0]
            nn.init.kaiming_normal_(conv.weight, mode="fan_out", nonlinearity="relu")   # fvcore c2_msra_fill
            nn.init.constant_(conv.bias, 0)
        self.precision = precision
        self._native = _holder.Holder()

    def _params(self) -> List[nn.Parameter]:
        ps = []
        for i in range(3):
            c = self[i][0]
            ps += [c.weight, c.bias, c.norm.weight, c.norm.bias]
        return ps + [self[3][0].weight, self[3][0].bias]

    def _buffers_list(self):
        bs = []
        for i in range(3):
            nm = self[i][0].norm
            bs += [nm.running_mean, nm.running_var, nm.num_batches_tracked]
        return bs

    def forward(self, feature: torch.Tensor) -> torch.Tensor:
        prec = native.PRECISIONS[self.precision or native.default_precision()]
        bn = self[0][0].norm
        momentum = 0.1 if bn.momentum is None else bn.momentum
        return PatchDiscriminatorFn.apply(feature, self._native, prec, bn.training, momentum, bn.eps, self._buffers_list(),
                                          torch.is_grad_enabled(), *self._params())


class Discriminator(nn.Module):
    def __init__(self, precision: Optional[str] = None):
        super().__init__()
        self.current_step = 0
        self.kw, self.padw, self.stw = 3, 1, 1
        self.Discriminators = nn.ModuleList([PatchDiscriminatorStack(256, precision)])

    def forward(self, feature: torch.Tensor) -> torch.Tensor:
        return self.Discriminators[self.current_step](feature)
