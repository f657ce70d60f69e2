"""BiFPN neck with the AF interpolator in every top-down fusion site: drop-in for reference afigan/modeling/backbone/bifpn_sr.py.

`BiFPN_AFIGAN` keeps the reference's constructor signature, attribute / parameter names (`before_bifpn.*`, `srf_module.*`, the seven
hand-unrolled layers `BiFPNLayer_{0..6}_conv{6,5,4,3}_up`, `..._conv{4,5,6,7}_down`, `..._p{6,5,4,3}_w1`, `..._p{4,5,6,7}_w2`;
bifpn_sr.py:283-517), creation order and arithmetic, including its quirks (App. D-9): `fpn_repeat` is ignored (always 7 layers), the
attention weights are used RAW (no ReLU, no normalisation; `_weight_act` is dead code), and the bottom-up path of EVERY layer takes its
skip inputs from the ORIGINAL laterals (`lateral_features`, bifpn_sr.py:596-598, 621, ...), only layer 0 using the extra p4 / p5 skip convs.

What runs where: the 28 fusion sites `w0 * cur + w1 * AFI(top)` go through the library (`bifpn_feature_fusion`: one call per site without
autograd, the interpolator's autograd Function with it); every 1x1 conv (input laterals, skips, the pointwise half of the 56 separable
convs) runs on the tcgen05 GEMM engine (`bifpn_layers.Conv2d`).  At inference (no autograd, eval-mode norms) the rest of a layer is native
too: `conv(swish(x))` is ONE library call per separable conv (swish + depthwise 3x3 + layout conversion in one HBM-bound pass, pointwise conv
with the BatchNorm folded into its weights on the tensor cores) and the bottom-up fusion sites (weighted sum + zero-padded max-pool) are one
elementwise pass each.  With autograd the depthwise 3x3, BatchNorm, swish and max-pool are torch ops.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from ..._compat import BACKBONE_REGISTRY, Backbone, ShapeSpec, c2_xavier_fill, get_norm
from ..bifpn_layers import Conv2d, MaxPool2d, MemoryEfficientSwish, SeparableConv2d
from ..feat_interpol import generator_rdb as G_rdb

__all__ = ["build_swint_bifpn_sr_backbone", "BiFPN_AFIGAN", "BeforeBiFPNLayer", "LastLevelP6P7", "ResampleFeature", "bifpn_feature_fusion"]


def bifpn_feature_fusion(srf_module, cur_feature: torch.Tensor, top_feature: torch.Tensor,
                         weight: Optional[torch.Tensor] = None, swish: bool = False) -> torch.Tensor:
    """`BiFPN_AFIGAN._feature_funsion` (bifpn_sr.py:542-548).  weight is the 2-element `..._w1` parameter (or None when attention is
    disabled): w[0]*cur + w[1]*AFI(top); swish=True also applies the x * sigmoid(x) the neck puts behind every fusion (bifpn_sr.py:591-594).
    Without autograd (inference, BASELINE config C5: 28 of these per image) the whole site is one library call; with autograd the
    interpolator is the library's autograd Function and the fusion (+ swish) ONE elementwise pass each way (`functional.FuseActFn`)."""
    needs_grad = torch.is_grad_enabled() and (cur_feature.requires_grad or top_feature.requires_grad or (weight is not None and weight.requires_grad)
                                              or any(p.requires_grad for p in srf_module.parameters()))
    if not needs_grad and top_feature.is_cuda and hasattr(srf_module, "fuse"):
        out = srf_module.fuse(top_feature, cur_feature, weight)
        return out * torch.sigmoid(out) if swish else out
    up = srf_module(top_feature, out_hw=tuple(cur_feature.shape[2:]))
    if up.is_cuda:
        from ...functional import bifpn_fuse_act
        return bifpn_fuse_act(cur_feature, up, weight, swish)
    out = cur_feature + up if weight is None else cur_feature * weight[0] + up * weight[1]
    return out * torch.sigmoid(out) if swish else out


class ResampleFeature(nn.Module):
    """bifpn_sr.py:740-750."""

    def __init__(self, in_channels, out_channels, kernel_size, norm):
        super().__init__()
        self.conv = Conv2d(in_channels, out_channels, kernel_size=1, stride=1, padding_mode="static_same")
        self.norm = get_norm(norm, out_channels) if norm != "" else (lambda x: x)
        self.resample = MaxPool2d(kernel_size=3, stride=2, padding_mode="static_same")
        c2_xavier_fill(self.conv)

    def forward(self, x):
        return self.resample(self.norm(self.conv(x)))


class LastLevelP6P7(nn.Module):
    """bifpn_sr.py:768-783: p6 = max-pool(norm(conv1x1(c5))), p7 = max-pool(p6)."""

    def __init__(self, in_channels, out_channels, norm=""):
        super().__init__()
        self.num_levels = 2
        self.p6 = ResampleFeature(in_channels, out_channels, 1, norm=norm)
        self.p7 = MaxPool2d(kernel_size=3, stride=2, padding_mode="static_same")

    def forward(self, p5):
        p6 = self.p6(p5)
        return [p6, self.p7(p6)]


class BeforeBiFPNLayer(nn.Module):
    """bifpn_sr.py:157-201: 1x1 conv + BatchNorm laterals for c3..c5, the top block for p6 / p7, and the p4 / p5 backbone skip convs."""

    def __init__(self, out_channels, in_channels=None, epsilon=1e-4, top_block=None):
        super().__init__()
        self.epsilon = epsilon
        mom, eps = 0.01, 1e-3

        def lat(cin):
            return nn.Sequential(Conv2d(cin, out_channels, 1, stride=1, padding_mode="static_same"), nn.BatchNorm2d(out_channels, momentum=mom, eps=eps))

        self.lateral3, self.lateral4, self.lateral5 = lat(in_channels[0]), lat(in_channels[1]), lat(in_channels[2])
        self.top_block = top_block
        self.p4_skip, self.p5_skip = lat(in_channels[1]), lat(in_channels[2])

    def forward(self, inputs):
        c3, c4, c5 = inputs
        c4_skip, c5_skip = self.p4_skip(c4), self.p5_skip(c5)
        c6, c7 = self.top_block(c5)
        return (self.lateral3(c3), self.lateral4(c4), self.lateral5(c5), c6, c7), (c4_skip, c5_skip)


class BiFPN_AFIGAN(Backbone):
    N_LAYERS = 7                                    # hand-unrolled in the reference; `fpn_repeat` is ignored (bifpn_sr.py:235)

    def __init__(self, bottom_up, in_features, out_channels, fpn_repeat, norm="SyncBN", top_block=None, fuse_type="sum", cfg=None):
        super().__init__()
        in_strides = [bottom_up._out_feature_strides[f] for f in in_features]
        in_channels = [bottom_up._out_feature_channels[f] for f in in_features]
        self.in_features, self.bottom_up, self.cfg = in_features, bottom_up, cfg
        self._out_feature_strides = {f"p{int(math.log2(s))}": s for s in in_strides}
        last_stage = int(math.log2(in_strides[-1]))
        for s in range(last_stage, last_stage + top_block.num_levels):
            in_strides.append(2 ** (s + 1))
            in_channels.append(out_channels)
            self._out_feature_strides[f"p{s + 1}"] = 2 ** (s + 1)
        for i, stride in enumerate(in_strides[1:], 1):
            assert stride == 2 * in_strides[i - 1], f"Strides {stride} {in_strides[i - 1]} are not log2 contiguous"
        self.before_bifpn = BeforeBiFPNLayer(out_channels, in_channels, top_block=top_block)
        srf_module = G_rdb.Generator(n_residual_dense_blocks=3)
        if cfg is not None and getattr(getattr(cfg, "MODEL", None), "AFI_FREEZE", False):
            for p in srf_module.parameters():
                p.requires_grad = False
        self.srf_module = srf_module
        self._downsample = MaxPool2d(3, 2)
        self._swish = MemoryEfficientSwish()
        self.attention = True
        self.epsilon = 1e-4
        mom, eps = 0.01, 1e-3
        for l in range(self.N_LAYERS):
            for name in ("conv6_up", "conv5_up", "conv4_up", "conv3_up", "conv4_down", "conv5_down", "conv6_down", "conv7_down"):
                setattr(self, f"BiFPNLayer_{l}_{name}", SeparableConv2d(out_channels, out_channels, 3, padding_mode="static_same", norm=norm,
                                                                         momentum=mom, eps=eps))
            for i in (6, 5, 4, 3):
                setattr(self, f"BiFPNLayer_{l}_p{i}_w1", nn.Parameter(torch.ones(2, dtype=torch.float32)))
            for i, n in ((4, 3), (5, 3), (6, 3), (7, 2)):
                setattr(self, f"BiFPNLayer_{l}_p{i}_w2", nn.Parameter(torch.ones(n, dtype=torch.float32)))
        self.bifpn = nn.Sequential()
        self._out_features = list(self._out_feature_strides.keys())
        self._out_feature_channels = {k: out_channels for k in self._out_features}
        self._size_divisibility = self._out_feature_strides[self._out_features[-1]]
        assert fuse_type in {"avg", "sum"}
        self._fuse_type = fuse_type

    @property
    def size_divisibility(self):
        return self._size_divisibility

    # ---- fusion sites (bifpn_sr.py:535-564)
    def _feature_funsion(self, layer_idx, cur_feature, top_feature, indice=-1, swish=False):
        w = getattr(self, f"BiFPNLayer_{layer_idx}_p{indice}_w1") if (self.attention and indice > 0) else None
        return bifpn_feature_fusion(self.srf_module, cur_feature, top_feature, w, swish)

    def _feature_funsion2(self, layer_idx, skip_feature, cur_feature, bottom_feature, indice=-1):
        if not torch.is_grad_enabled() and cur_feature.is_cuda and bottom_feature.size(2) >= 2 and bottom_feature.size(3) >= 2:
            # inference: weighted sum + zero-padded 3x3 / stride-2 max-pool of the level below in ONE library pass
            from ...functional import bifpn_fuse_down
            w = getattr(self, f"BiFPNLayer_{layer_idx}_p{indice}_w2") if (self.attention and indice > 0) else None
            if isinstance(skip_feature, torch.Tensor):
                return bifpn_fuse_down(skip_feature, cur_feature, bottom_feature, w)
            return bifpn_fuse_down(cur_feature, None, bottom_feature, w)
        down = self._downsample(bottom_feature)
        inputs = [skip_feature, cur_feature, down] if isinstance(skip_feature, torch.Tensor) else [cur_feature, down]
        if self.attention and indice > 0:
            w = getattr(self, f"BiFPNLayer_{layer_idx}_p{indice}_w2")
            assert len(inputs) == len(w)
            return sum(x_ * w_ for x_, w_ in zip(inputs, w))
        return sum(inputs)

    def _layer(self, l, laterals, down_skips):
        conv = lambda name: getattr(self, f"BiFPNLayer_{l}_{name}")          # noqa: E731
        p3_in, p4_in, p5_in, p6_in, p7_in = laterals
        # conv(swish(fused)).  Inference: the swish is handed to the separable conv (fused into its depthwise pass on the native path).
        # With autograd: fusion + swish are ONE elementwise pass each way (functional.FuseActFn) and the conv gets the activated map.
        tr = torch.is_grad_enabled()

        def td(name, cur, top, idx):
            return conv(name)(self._feature_funsion(l, cur, top, idx, swish=tr), pre_swish=not tr)

        p6_up = td("conv6_up", p6_in, p7_in, 6)
        p5_up = td("conv5_up", p5_in, p6_up, 5)
        p4_up = td("conv4_up", p4_in, p5_up, 4)
        p3_up = td("conv3_up", p3_in, p4_up, 3)
        s4, s5, s6, s7 = down_skips
        p4_out = conv("conv4_down")(self._feature_funsion2(l, s4, p4_up, p3_up, 4), pre_swish=True)
        p5_out = conv("conv5_down")(self._feature_funsion2(l, s5, p5_up, p4_out, 5), pre_swish=True)
        p6_out = conv("conv6_down")(self._feature_funsion2(l, s6, p6_up, p5_out, 6), pre_swish=True)
        p7_out = conv("conv7_down")(self._feature_funsion2(l, None, s7, p6_out, 7), pre_swish=True)
        return p3_up, p4_out, p5_out, p6_out, p7_out

    def forward(self, x):
        """bifpn_sr.py:566-731."""
        bottom_up_features = self.bottom_up(x)
        features = [bottom_up_features[f] for f in self.in_features]
        lateral_features, skip_features = self.before_bifpn(features)
        _, l4, l5, l6, l7 = lateral_features
        feats = lateral_features
        for l in range(self.N_LAYERS):
            # the bottom-up path of every layer reads the ORIGINAL laterals (p6 / p7 always; p4 / p5 through the extra skip convs in layer 0)
            down_skips = (skip_features[0], skip_features[1], l6, l7) if l == 0 else (l4, l5, l6, l7)
            feats = self._layer(l, feats, down_skips)
        assert len(self._out_features) == len(feats)
        return dict(zip(self._out_features, feats))


@BACKBONE_REGISTRY.register()
def build_swint_bifpn_sr_backbone(cfg, input_shape: ShapeSpec):
    """bifpn_sr.py:791-816.  The bottom-up network is whatever `build_swint_backbone` is registered in BACKBONE_REGISTRY (the reference's Swin
    Transformer needs timm and is out of this repository's scope); any Backbone with three contiguous-stride outputs works."""
    try:
        build_bottom_up = BACKBONE_REGISTRY.get("build_swint_backbone")
    except KeyError as e:
        raise ImportError("build_swint_bifpn_sr_backbone: no `build_swint_backbone` in BACKBONE_REGISTRY (the reference's Swin bottom-up, "
                          "afigan/modeling/backbone/swin_transformer.py:641, is not re-implemented here); register one, or construct "
                          "BiFPN_AFIGAN(bottom_up=...) directly") from e
    bottom_up = build_bottom_up(cfg, input_shape)
    in_features = cfg.MODEL.BIFPN.IN_FEATURES
    out_channels = cfg.MODEL.BIFPN.OUT_CHANNELS
    in_channels_p6p7 = bottom_up.output_shape()[in_features[-1]].channels
    return BiFPN_AFIGAN(bottom_up=bottom_up, in_features=in_features, out_channels=out_channels, fpn_repeat=cfg.MODEL.BIFPN.FPN_REPEAT,
                        norm=cfg.MODEL.BIFPN.NORM, top_block=LastLevelP6P7(in_channels_p6p7, out_channels, cfg.MODEL.BIFPN.NORM),
                        fuse_type=cfg.MODEL.BIFPN.FUSE_TYPE, cfg=cfg)
