"""PAFPN neck with the AF interpolator: drop-in for reference afigan/modeling/backbone/pafpn_sr.py (PAFPN_AFIGAN, :20-200).

Top-down merge = the library's fused Generator.merge (pafpn_sr.py:172-181); the PANet bottom-up augmentation (stride-2 3x3 convs,
ReLU, adds: :186-193) and the output convs are plain torch ops (out of the hot path)."""
from __future__ import annotations

import math

import torch.nn.functional as F

from ..._compat import BACKBONE_REGISTRY, Backbone, Conv2d, ShapeSpec, c2_xavier_fill, get_norm
from ..feat_interpol import generator_rdb as G_rdb
from .fpn_sr import LastLevelMaxPool, _afi_freeze, _assert_strides_are_log2_contiguous, _build, _resnest_builder, _resnet_builder, topdown_merge

__all__ = ["build_resnet_pafpn_sr_backbone", "build_resnest_pafpn_sr_backbone", "PAFPN_AFIGAN"]


class PAFPN_AFIGAN(Backbone):
    def __init__(self, bottom_up, in_features, out_channels, norm="", top_block=None, fuse_type="sum", cfg=None):
        super().__init__()
        self.cfg = cfg
        input_shapes = bottom_up.output_shape()
        in_strides = [input_shapes[f].stride for f in in_features]
        in_channels = [input_shapes[f].channels for f in in_features]
        _assert_strides_are_log2_contiguous(in_strides)
        if out_channels != 256:
            raise ValueError("the AF interpolator kernels are specialised for 256-channel pyramids")
        self.srf_module = G_rdb.Generator(n_residual_dense_blocks=3)        # pafpn_sr.py:67
        if _afi_freeze(cfg):
            for p in self.srf_module.parameters():
                p.requires_grad = False
        lateral_convs, output_convs, downsample_convs = [], [], []
        use_bias = norm == ""
        stage = 0
        for idx, ch in enumerate(in_channels):
            lateral_conv = Conv2d(ch, out_channels, kernel_size=1, bias=use_bias, norm=get_norm(norm, out_channels))
            output_conv = Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=use_bias, norm=get_norm(norm, out_channels))
            c2_xavier_fill(lateral_conv)
            c2_xavier_fill(output_conv)
            stage = int(math.log2(in_strides[idx]))
            self.add_module(f"fpn_lateral{stage}", lateral_conv)
            self.add_module(f"pafpn_output{stage}", output_conv)
            lateral_convs.append(lateral_conv)
            output_convs.append(output_conv)
            if idx > 0:
                down = Conv2d(out_channels, out_channels, kernel_size=3, stride=2, padding=1, bias=use_bias, norm=get_norm(norm, out_channels))
                c2_xavier_fill(down)
                self.add_module(f"pafpn_downsample{stage}", down)
                downsample_convs.append(down)
        self.lateral_convs = lateral_convs[::-1]
        self.output_convs = output_convs
        self.downsample_convs = downsample_convs
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
        prev_features = self.lateral_convs[0](feats[0])
        topdown = [prev_features]
        for features, lateral_conv in zip(feats[1:], self.lateral_convs[1:]):              # top-down pathway (:172-181)
            prev_features = topdown_merge(self.srf_module, prev_features, features, lateral_conv, self._fuse_type)
            topdown.insert(0, prev_features)
        pa_prev = topdown.pop(0)
        results = [self.output_convs[0](pa_prev)]
        for inter, down, out_conv in zip(topdown, self.downsample_convs, self.output_convs[1:]):   # bottom-up augmentation (:186-193)
            pa_prev = inter + F.relu_(down(pa_prev))
            if self._fuse_type == "avg":
                pa_prev = pa_prev / 2
            results.append(out_conv(pa_prev))
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


@BACKBONE_REGISTRY.register()
def build_resnet_pafpn_sr_backbone(cfg, input_shape: ShapeSpec):
    """pafpn_sr.py:237-258."""
    return _build(cfg, input_shape, _resnet_builder, PAFPN_AFIGAN)


@BACKBONE_REGISTRY.register()
def build_resnest_pafpn_sr_backbone(cfg, input_shape: ShapeSpec):
    """pafpn_sr.py:260-281."""
    return _build(cfg, input_shape, _resnest_builder, PAFPN_AFIGAN)
