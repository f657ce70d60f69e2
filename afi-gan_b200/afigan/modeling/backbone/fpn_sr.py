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
from ...functional import conv1x1_autograd, conv3x3_autograd
from ..feat_interpol import generator_rdb as G_rdb

__all__ = ["build_resnet_fpn_sr_backbone", "build_resnest_fpn_sr_backbone", "FPN_AFIGAN", "LastLevelMaxPool"]


def _afi_freeze(cfg) -> bool:
    return bool(cfg is not None and getattr(getattr(cfg, "MODEL", None), "AFI_FREEZE", False))


def _assert_strides_are_log2_contiguous(strides):
    for i, stride in enumerate(strides[1:], 1):
        assert stride == 2 * strides[i - 1], f"Strides {stride} {strides[i - 1]} are not log2 contiguous"


def output_conv3x3(conv, x, precision=None):
    """The necks' 3x3 output convs (fpn_sr.py:144-158): through the library's implicit-GEMM engine when the conv is bare (NORM == "",
    stride 1, 256 channels a multiple of 32), plain torch otherwise."""
    bare = (getattr(conv, "norm", None) is None and getattr(conv, "activation", None) is None and conv.stride == (1, 1)
            and conv.kernel_size == (3, 3) and conv.padding == (1, 1) and conv.in_channels % 32 == 0 and conv.out_channels % 32 == 0)
    if bare and x.is_cuda:
        return conv3x3_autograd(x, conv.weight, conv.bias, precision)
    return conv(x)


def lateral_conv1x1(conv, x, precision=None):
    """A neck's 1x1 lateral conv on its own -- the top-level lateral (fpn_sr.py:144-145), or any lateral that carries a norm (SyncBN
    configs): the 1x1 conv runs on the library's GEMM engine, norm / activation (if any) stay torch modules."""
    ok = (x.is_cuda and conv.kernel_size == (1, 1) and conv.stride == (1, 1) and conv.groups == 1 and conv.in_channels % 32 == 0
          and conv.out_channels % 32 == 0)
    if not ok:
        return conv(x)
    y = conv1x1_autograd(x, conv.weight, conv.bias, precision)
    if getattr(conv, "norm", None) is not None:
        y = conv.norm(y)
    if getattr(conv, "activation", None) is not None:
        y = conv.activation(y)
    return y


def topdown_merge(srf_module, prev_features, features, lateral_conv, fuse_type):
    """One top-down step.  Bare 1x1 lateral (FPN.NORM == ""): fully fused in the library; with a norm in the lateral the conv+norm
    stay in torch and only the interpolator runs in the library (SURVEY.md §7 hard part 7)."""
    if getattr(lateral_conv, "norm", None) is None and getattr(lateral_conv, "activation", None) is None:
        return srf_module.merge(prev_features, features, lateral_conv.weight, lateral_conv.bias, fuse_type)
    lateral = lateral_conv1x1(lateral_conv, features, srf_module.precision)
    out = lateral + srf_module(prev_features, out_hw=tuple(lateral.shape[2:]))
    return out / 2 if fuse_type == "avg" else out


class _AFINeck(Backbone):
    """Shared construction of the AFI necks: the interpolator (`srf_module`), one 1x1 lateral + one 3x3 output conv per input level,
    registered under the reference's attribute names so state dicts interchange."""

    output_prefix = "fpn_output"

    def __init__(self, bottom_up, in_features, out_channels, norm="", top_block=None, fuse_type="sum", cfg=None):
        super().__init__()
        if out_channels != 256:
            raise ValueError("the AF interpolator kernels are specialised for 256-channel pyramids (MODEL.FPN.OUT_CHANNELS = 256)")
        assert fuse_type in {"avg", "sum"}
        self.cfg = cfg
        shapes = bottom_up.output_shape()
        strides = [shapes[f].stride for f in in_features]
        _assert_strides_are_log2_contiguous(strides)
        self.srf_module = G_rdb.Generator(n_residual_dense_blocks=3)        # fpn_sr.py:65 / pafpn_sr.py:67
        if _afi_freeze(cfg):                                                # fpn_sr.py:67-69
            for p in self.srf_module.parameters():
                p.requires_grad = False
        self._norm, self._use_bias = norm, norm == ""
        laterals, outputs = [], []
        for f, stride in zip(in_features, strides):
            stage = int(math.log2(stride))
            lateral = self._conv(shapes[f].channels, out_channels, 1)
            output = self._conv(out_channels, out_channels, 3)
            self.add_module(f"fpn_lateral{stage}", lateral)
            self.add_module(f"{self.output_prefix}{stage}", output)
            laterals.append(lateral)
            outputs.append(output)
            self._extra_level_modules(stage, first=not laterals[:-1], out_channels=out_channels)
        self._laterals_bottom_up, self._outputs_bottom_up = laterals, outputs
        self.top_block, self.in_features, self.bottom_up = top_block, in_features, bottom_up
        self._out_feature_strides = {f"p{int(math.log2(s))}": s for s in strides}
        last = int(math.log2(strides[-1]))
        if top_block is not None:
            for s_ in range(last, last + top_block.num_levels):
                self._out_feature_strides[f"p{s_ + 1}"] = 2 ** (s_ + 1)
        self._out_features = list(self._out_feature_strides)
        self._out_feature_channels = {k: out_channels for k in self._out_features}
        self._size_divisibility = strides[-1]
        self._fuse_type = fuse_type

    def _conv(self, cin, cout, k, stride=1):
        conv = Conv2d(cin, cout, kernel_size=k, stride=stride, padding=k // 2, bias=self._use_bias, norm=get_norm(self._norm, cout))
        c2_xavier_fill(conv)
        return conv

    def _extra_level_modules(self, stage, first, out_channels):
        pass

    @property
    def size_divisibility(self):
        return self._size_divisibility

    def output_shape(self):
        return {name: ShapeSpec(channels=self._out_feature_channels[name], stride=self._out_feature_strides[name])
                for name in self._out_features}

    def _top_down(self, bottom_up_features):
        """Merged maps, finest first: prev = lateral(C_l) + srf_module(prev) [/2]   (fpn_sr.py:147-157, pafpn_sr.py:172-181)."""
        feats = [bottom_up_features[f] for f in self.in_features[::-1]]
        laterals = self._laterals_bottom_up[::-1]
        prev = lateral_conv1x1(laterals[0], feats[0], self.srf_module.precision)
        merged = [prev]
        for features, lateral_conv in zip(feats[1:], laterals[1:]):
            prev = topdown_merge(self.srf_module, prev, features, lateral_conv, self._fuse_type)
            merged.insert(0, prev)
        return merged

    def _finish(self, bottom_up_features, results):
        if self.top_block is not None:
            top_in = bottom_up_features.get(self.top_block.in_feature, None)
            if top_in is None:
                top_in = results[self._out_features.index(self.top_block.in_feature)]
            results.extend(self.top_block(top_in))
        assert len(self._out_features) == len(results)
        return dict(zip(self._out_features, results))


class FPN_AFIGAN(_AFINeck):
    """FPN whose top-down path up-samples with the AF interpolator (reference FPN_AFIGAN, fpn_sr.py:18-166)."""

    @property
    def lateral_convs(self):      # top-down order, like the reference attribute
        return self._laterals_bottom_up[::-1]

    @property
    def output_convs(self):
        return self._outputs_bottom_up[::-1]

    def forward(self, x):
        bottom_up_features = self.bottom_up(x)
        merged = self._top_down(bottom_up_features)
        results = [output_conv3x3(conv, m, self.srf_module.precision) for conv, m in zip(self._outputs_bottom_up, merged)]
        return self._finish(bottom_up_features, results)


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
