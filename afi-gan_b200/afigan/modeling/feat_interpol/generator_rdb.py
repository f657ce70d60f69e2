"""AF interpolator ("Generator"): drop-in for reference afigan/modeling/feat_interpol/generator_rdb.py.

Same constructor signature, parameter names / shapes / creation order and RNG consumption as the reference
(generator_rdb.py:34-62, :75-121; state-dict layout SURVEY.md App. B), so optimisers, DDP, freezing
(fpn_sr.py:67-69) and both checkpointers see the module they expect.  The modules below only HOLD parameters:
`Generator.forward` runs the whole trunk in libafigan_b200.so (afi_g_forward / afi_g_backward).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import nn

from .. import _holder
from ... import native
from ...functional import AFInterpolatorFn, PackedWeights


class ResidualDenseBlock(nn.Module):
    """Parameter holder for one dense block: conv1..conv4 (C+32i -> 32, LeakyReLU) and conv5 (C+128 -> C), all bias-free."""

    def __init__(self, in_features, growth_rate, residual_scale, kw, stw, padw):
        super().__init__()
        self.residual_scale = residual_scale
        for i in range(4):
            setattr(self, f"conv{i + 1}", nn.Sequential(
                nn.Conv2d(in_features + i * growth_rate, growth_rate, kw, stw, padw, bias=False),
                nn.LeakyReLU(negative_slope=0.2, inplace=True)))
        self.conv5 = nn.Conv2d(in_features + 4 * growth_rate, in_features, kw, stw, padw, bias=False)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
                m.weight.data *= 0.1

    def weights(self) -> List[nn.Parameter]:
        return [self.conv1[0].weight, self.conv2[0].weight, self.conv3[0].weight, self.conv4[0].weight, self.conv5.weight]

    def forward(self, x):
        raise RuntimeError("ResidualDenseBlock is evaluated inside Generator.forward by the CUDA library")


class ResidualInResidual(nn.Module):
    def __init__(self, n_residual_dense_blocks, in_features, growth_rate, residual_scale, kw, stw, padw):
        super().__init__()
        self.RDBs = nn.Sequential(*[ResidualDenseBlock(in_features, growth_rate, residual_scale, kw, stw, padw)
                                    for _ in range(n_residual_dense_blocks)])
        self.residual_scale = residual_scale

    def forward(self, x):
        raise RuntimeError("ResidualInResidual is evaluated inside Generator.forward by the CUDA library")


class Generator(nn.Module):
    def __init__(self, in_channels=256, n_residual_dense_blocks=2, growth_rate=32, residual_scale=0.2, scale=2,
                 precision: Optional[str] = None):
        super().__init__()
        if in_channels != 256 or growth_rate != 32 or residual_scale != 0.2 or scale != 2:
            raise ValueError("the sm_100a kernels are specialised for in_channels=256, growth_rate=32, residual_scale=0.2, "
                             "scale=2 (every caller in the reference uses these: fpn_sr.py:65, stage1_trainer.py:505)")
        if not 1 <= n_residual_dense_blocks <= native.MAX_RDB:
            raise ValueError(f"n_residual_dense_blocks must be in [1, {native.MAX_RDB}]")
        self.in_channels, self.n_residual_dense_blocks = in_channels, n_residual_dense_blocks
        self.growth_rate, self.residual_scale, self.scale = growth_rate, residual_scale, scale
        self.kw, self.padw, self.stw = 3, 1, 1
        self.precision = precision

        stages = [
            nn.Sequential(nn.Conv2d(in_channels, in_channels, 3, 1, 1), nn.LeakyReLU(0.2, True)),
            ResidualInResidual(n_residual_dense_blocks, in_channels, growth_rate, residual_scale, 3, 1, 1),
            nn.Sequential(nn.Conv2d(in_channels, in_channels, 3, 1, 1), nn.LeakyReLU(0.2, True)),
            nn.Sequential(nn.ConvTranspose2d(in_channels, in_channels, kernel_size=6, stride=2, padding=2), nn.LeakyReLU(0.2, True)),
            nn.Sequential(nn.Conv2d(in_channels, in_channels, 3, 1, 1)),
        ]
        for st in stages:
            if isinstance(st, ResidualInResidual):
                continue
            layer = st[0]
            nn.init.kaiming_normal_(layer.weight)
            layer.weight.data *= 0.1
            layer.bias.data.zero_()
        self.Generators = nn.ModuleList([nn.Sequential(*stages)])
        self._native = _holder.Holder(n_rdb=n_residual_dense_blocks)

    @property
    def deferred_weight_grads(self) -> bool:
        """Opt-in: accumulate this module's weight gradients over ALL its calls of a backward pass in one packed buffer and write .grad once when
        the pass ends (a BiFPN calls the interpolator 28 times per step).  Not for torch DistributedDataParallel (its per-parameter hooks do not
        fire); see `_holder.Holder`."""
        return self._native.deferred

    @deferred_weight_grads.setter
    def deferred_weight_grads(self, on: bool) -> None:
        self._native.deferred = bool(on)

    def _params(self) -> List[nn.Parameter]:
        g = self.Generators[0]
        ps = [g[0][0].weight, g[0][0].bias]
        for rdb in g[1].RDBs:
            ps += rdb.weights()
        for i in (2, 3, 4):
            ps += [g[i][0].weight, g[i][0].bias]
        return ps

    def forward(self, features: torch.Tensor, out_hw: Optional[Tuple[int, int]] = None) -> torch.Tensor:
        """features [N,256,H,W] -> [N,256,2H,2W] (or its top-left out_hw crop: the _reshape_stage1 of the trainers)."""
        prec = native.PRECISIONS[self.precision or native.default_precision()]
        if self._use_graphs(features):
            return self._native.graphs.run(self._native, prec, self._params(), features, out_hw)
        return AFInterpolatorFn.apply(features, None, None, None, self._native, prec, out_hw, 1.0, torch.is_grad_enabled(), *self._params())

    def _use_graphs(self, x: torch.Tensor) -> bool:
        """Forward-only calls (inference: no autograd) replay CUDA graphs: eagerly issued they are host-bound (functional.InferenceGraphs).
        AFIGAN_INFER_GRAPH=0 or a stream capture already in progress keeps the eager path."""
        import os
        return (not torch.is_grad_enabled() and x.is_cuda and os.environ.get("AFIGAN_INFER_GRAPH", "1") != "0"
                and not torch.cuda.is_current_stream_capturing())

    def fuse(self, top_feature: torch.Tensor, cur_feature: torch.Tensor, weight: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Forward-only BiFPN fusion site in ONE library call (reference bifpn_sr.py:535-548): weight[0] * cur + weight[1] * self(top)."""
        from ...functional import afi_bifpn_fuse
        prec = native.PRECISIONS[self.precision or native.default_precision()]
        with torch.no_grad():
            if self._use_graphs(top_feature):
                oh, ow = cur_feature.shape[2:]
                return self._native.graphs.run(self._native, prec, self._params(), top_feature, (oh, ow), fuse=(native.boundary(cur_feature), weight))
            return afi_bifpn_fuse(top_feature, cur_feature, weight, self._native, prec, self._params())

    def merge(self, prev_features: torch.Tensor, bottom_up: torch.Tensor, lateral_weight: torch.Tensor,
              lateral_bias: Optional[torch.Tensor] = None, fuse_type: str = "sum") -> torch.Tensor:
        """The top-down merge of the AFI necks in ONE library call (reference fpn_sr.py:151-157, pafpn_sr.py:175-181):
        lateral_conv1x1(bottom_up) + self(prev_features), divided by 2 for fuse_type "avg".  The interpolated map is cropped to
        the lateral's size when 2H x 2W overshoots it (odd pyramid sizes)."""
        prec = native.PRECISIONS[self.precision or native.default_precision()]
        oh, ow = bottom_up.shape[2:]
        w2 = lateral_weight.reshape(lateral_weight.shape[0], lateral_weight.shape[1])
        if self._use_graphs(prev_features) and lateral_weight.is_contiguous() and lateral_weight.dtype == torch.float32:
            return self._native.graphs.run(self._native, prec, self._params(), prev_features, (oh, ow),
                                           lat=(native.boundary(bottom_up), w2, lateral_bias, 0.5 if fuse_type == "avg" else 1.0))
        return AFInterpolatorFn.apply(prev_features, bottom_up, w2, lateral_bias, self._native, prec, (oh, ow),
                                      0.5 if fuse_type == "avg" else 1.0, torch.is_grad_enabled(), *self._params())
