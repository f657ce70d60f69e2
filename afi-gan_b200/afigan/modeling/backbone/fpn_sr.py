"""FPN neck with the AF interpolator in the top-down path: drop-in for reference afigan/modeling/backbone/fpn_sr.py.

Same constructor signature, parameter names (`fpn_lateral{2..5}`, `fpn_output{2..5}`, `srf_module`) and output contract as the
reference `FPN_AFIGAN` (fpn_sr.py:18-166).  The hot part -- `prev = lateral_conv(C_l) + srf_module(prev) [/2]` (fpn_sr.py:151-157) --
is ONE library call (Generator.merge: interpolator trunk, lateral 1x1 conv, bilinear skip, add and scale).  The 3x3 output convs,
the top lateral and the top block are plain torch ops ("next").
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

from ..._compat import BACKBONE_REGISTRY, Backbone, Conv2d, ShapeSpec, c2_xavier_fill, get_norm
from ..feat_interpol import generator_rdb as G_rdb

__all__ = ["build_resnet_fpn_sr_backbone", "build_resnest_fpn_sr_backbone", "FPN_AFIGAN", "LastLevelMaxPool"]


def _afi_freeze(cfg) -> bool:
    return bool(cfg is not None and getattr(getattr(cfg, "MODEL", None), "AFI_FREEZE", False))


def _assert_strides_are_log2_contiguous(strides):
    for i, stride in enumerate(strides[1:], 1):
        assert stride == 2 * strides[i - 1], f"Strides {stride} {strides[i - 1]} are not log2 contiguous"


def topdown_merge(srf_module, prev_features, features, lateral_conv, fuse_type):
    """One top-down step.  Bare 1x1 lateral (FPN.NORM == ""): fully fused in the library; with a norm in the lateral the conv+norm
    stay in torch and only the interpolator runs in the library (SURVEY.md §7 hard part 7)."""
    if getattr(lateral_conv, "norm", None) is None and getattr(lateral_conv, "activation", None) is None:
        return srf_module.merge(prev_features, features, lateral_conv.weight, lateral_conv.bias, fuse_type)
    lateral = lateral_conv(features)
    out = lateral + srf_module(prev_features, out_hw=tuple(lateral.shape[2:]))
    return out / 2 if fuse_type == "avg" else out


class FPN_AFIGAN(Backbone):
    def __init__(self, bottom_up, in_features, out_channels, norm="", top_block=None, fuse_type="sum", cfg=None):
        super().__init__()
        self.cfg = cfg
        input_shapes = bottom_up.output_shape()
        in_strides = [input_shapes[f].stride for f in in_features]
        in_channels = [input_shapes[f].channels for f in in_features]
        _assert_strides_are_log2_contiguous(in_strides)
        if out_channels != 256:
            raise ValueError("the AF interpolator kernels are specialised for 256-channel pyramids (MODEL.FPN.OUT_CHANNELS = 256)")

        self.srf_module = G_rdb.Generator(n_residual_dense_blocks=3)        # fpn_sr.py:65
        if _afi_freeze(cfg):                                                # fpn_sr.py:67-69
            for p in self.srf_module.parameters():
                p.requires_grad = False

        lateral_convs, output_convs = [], []
        use_bias = norm == ""
        stage = 0
        for idx, ch in enumerate(in_channels):
            lateral_conv = Conv2d(ch, out_channels, kernel_size=1, bias=use_bias, norm=get_norm(norm, out_channels))
            output_conv = Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=use_bias,
                                 norm=get_norm(norm, out_channels))
            c2_xavier_fill(lateral_conv)
            c2_xavier_fill(output_conv)
            stage = int(math.log2(in_strides[idx]))
            self.add_module(f"fpn_lateral{stage}", lateral_conv)
            self.add_module(f"fpn_output{stage}", output_conv)
            lateral_convs.append(lateral_conv)
            output_convs.append(output_conv)
        self.lateral_convs = lateral_convs[::-1]     # top-down order
        self.output_convs = output_convs[::-1]
        self.top_block = top_block
        self.in_features = in_features
        self.bottom_up = bottom_up
        self._out_feature_strides = {f"p{int(math.log2(s))}": s for s in in_strides}
        if self.top_block is not None:
            for s in range(stage, stage + self.top_block.num_levels):
                self._out_feature_strides[f"p{s + 1}"] = 2 ** (s + 1)
        self._out_features = list(self._out_feature_strides.keys())
        self._out_feature_channels = {k: out_channels for k in self._out_features}
        self._size_divisibility = in_strides[-1]
        assert fuse_type in {"avg", "sum"}
        self._fuse_type = fuse_type

    @property
    def size_divisibility(self):
        return self._size_divisibility

    def forward(self, x):
        bottom_up_features = self.bottom_up(x)
        feats = [bottom_up_features[f] for f in self.in_features[::-1]]
        results = []
        prev_features = self.lateral_convs[0](feats[0])
        results.append(self.output_convs[0](prev_features))
        for features, lateral_conv, output_conv in zip(feats[1:], self.lateral_convs[1:], self.output_convs[1:]):
            prev_features = topdown_merge(self.srf_module, prev_features, features, lateral_conv, self._fuse_type)
            results.insert(0, output_conv(prev_features))
        if self.top_block is not None:
            top_in = bottom_up_features.get(self.top_block.in_feature, None)
            if top_in is None:
                top_in = results[self._out_features.index(self.top_block.in_feature)]
            results.extend(self.top_block(top_in))
        assert len(self._out_features) == len(results)
        return dict(zip(self._out_features, results))

    def output_shape(self):
        return {name: ShapeSpec(channels=self._out_feature_channels[name], stride=self._out_feature_strides[name])
                for name in self._out_features}


class LastLevelMaxPool(nn.Module):
    """P6 from P5 by a stride-2 subsample (fpn_sr.py:187-199)."""

    def __init__(self):
        super().__init__()
        self.num_levels = 1
        self.in_feature = "p5"

    def forward(self, x):
        return [F.max_pool2d(x, kernel_size=1, stride=2, padding=0)]


def _build(cfg, input_shape, bottom_up_builder, neck_cls):
    bottom_up = bottom_up_builder(cfg, input_shape)
    return neck_cls(bottom_up=bottom_up, in_features=cfg.MODEL.FPN.IN_FEATURES, out_channels=cfg.MODEL.FPN.OUT_CHANNELS,
                    norm=cfg.MODEL.FPN.NORM, top_block=LastLevelMaxPool(), fuse_type=cfg.MODEL.FPN.FUSE_TYPE, cfg=cfg)


def _resnet_builder(cfg, input_shape):
    try:
        from detectron2.modeling.backbone.resnet import build_resnet_backbone
    except Exception as e:  # noqa: BLE001
        raise ImportError("build_resnet_fpn_sr_backbone needs detectron2's ResNet bottom-up backbone (out of the hot path's scope); "
                          "construct FPN_AFIGAN(bottom_up=...) directly with your own Backbone") from e
    return build_resnet_backbone(cfg, input_shape)


def _resnest_builder(cfg, input_shape):
    try:
        from afigan.modeling.backbone.resnest import build_resnest_backbone   # the reference's own ResNeSt (not re-implemented here)
    except Exception as e:  # noqa: BLE001
        raise ImportError("build_resnest_fpn_sr_backbone needs the reference's ResNeSt bottom-up backbone (out of scope: SURVEY.md §2 row 10)") from e
    return build_resnest_backbone(cfg, input_shape)


@BACKBONE_REGISTRY.register()
def build_resnet_fpn_sr_backbone(cfg, input_shape: ShapeSpec):
    """fpn_sr.py:201-222."""
    return _build(cfg, input_shape, _resnet_builder, FPN_AFIGAN)


@BACKBONE_REGISTRY.register()
def build_resnest_fpn_sr_backbone(cfg, input_shape: ShapeSpec):
    """fpn_sr.py:224-245."""
    return _build(cfg, input_shape, _resnest_builder, FPN_AFIGAN)
