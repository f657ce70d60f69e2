"""PAFPN neck with the AF interpolator: drop-in for reference afigan/modeling/backbone/pafpn_sr.py (PAFPN_AFIGAN, :20-200).

Top-down merge = the library's fused Generator.merge (pafpn_sr.py:172-181); the PANet bottom-up augmentation's stride-2 3x3 convs
(:186-193) and the 3x3 output convs run on the library's tcgen05 GEMM engine when they are bare convs (FPN.NORM == ""); ReLU and adds are
elementwise torch ops."""
from __future__ import annotations

import torch.nn.functional as F

from ..._compat import BACKBONE_REGISTRY, ShapeSpec
from ...functional import conv3x3s2_autograd
from ... import native
from .fpn_sr import _AFINeck, _build, _resnest_builder, _resnet_builder, output_conv3x3

__all__ = ["build_resnet_pafpn_sr_backbone", "build_resnest_pafpn_sr_backbone", "PAFPN_AFIGAN"]


class PAFPN_AFIGAN(_AFINeck):
    """PANet neck: AFI top-down path, then a bottom-up augmentation of stride-2 3x3 convs (reference PAFPN_AFIGAN, pafpn_sr.py:20-200)."""

    output_prefix = "pafpn_output"

    def _extra_level_modules(self, stage, first, out_channels):
        if not first:      # every level but the finest gets a down-sampling conv (pafpn_sr.py:103-117)
            self.add_module(f"pafpn_downsample{stage}", self._conv(out_channels, out_channels, 3, stride=2))

    @property
    def downsample_convs(self):
        return [m for n, m in self.named_children() if n.startswith("pafpn_downsample")]

    def _down(self, conv, x):
        """Stride-2 down-sampling conv: native when bare (no norm), 64-aligned channels and a tensor-core operand mode."""
        prec = self.srf_module.precision or native.default_precision()
        bare = (getattr(conv, "norm", None) is None and getattr(conv, "activation", None) is None and conv.stride == (2, 2)
                and conv.kernel_size == (3, 3) and conv.padding == (1, 1) and conv.in_channels % 64 == 0 and conv.out_channels % 64 == 0)
        if bare and x.is_cuda and prec in ("bf16", "split") and min(x.shape[2:]) >= 2:
            return conv3x3s2_autograd(x, conv.weight, conv.bias, prec)
        return conv(x)

    def forward(self, x):
        bottom_up_features = self.bottom_up(x)
        merged = self._top_down(bottom_up_features)                    # finest first
        pa_prev = merged[0]
        results = [output_conv3x3(self._outputs_bottom_up[0], pa_prev, self.srf_module.precision)]
        for inter, down, out_conv in zip(merged[1:], self.downsample_convs, self._outputs_bottom_up[1:]):   # :186-193
            pa_prev = inter + F.relu_(self._down(down, pa_prev))
            if self._fuse_type == "avg":
                pa_prev = pa_prev / 2
            results.append(output_conv3x3(out_conv, pa_prev, self.srf_module.precision))
        return self._finish(bottom_up_features, results)


@BACKBONE_REGISTRY.register()
def build_resnet_pafpn_sr_backbone(cfg, input_shape: ShapeSpec):
    """pafpn_sr.py:237-258."""
    return _build(cfg, input_shape, _resnet_builder, PAFPN_AFIGAN)


@BACKBONE_REGISTRY.register()
def build_resnest_pafpn_sr_backbone(cfg, input_shape: ShapeSpec):
    """pafpn_sr.py:260-281."""
    return _build(cfg, input_shape, _resnest_builder, PAFPN_AFIGAN)
