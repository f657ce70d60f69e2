"""Bottom-up ResNet + plain FPN for the frozen GUIDE feature extractor (producer of the hot path's inputs).

The reference builds its guide backbone with detectron2's `build_resnet_fpn_backbone` (cfg.MODEL.GUIDE_BACKBONE.NAME, reference
afigan/config/defaults.py:16-22, afigan/modeling/meta_arch/rcnn_only.py:46-60) [upstream].  detectron2 is not in this image, so this file
provides the same builder NAME with the same module / parameter names (`bottom_up.stem.conv1`, `bottom_up.res{2..5}.{i}.conv{1,2,3}` +
`.shortcut`, each with `.norm.{weight,bias,running_mean,running_var}`; `fpn_lateral{2..5}`, `fpn_output{2..5}`), so a detectron2 model-zoo
R-50-FPN checkpoint maps onto it key for key.  It is registered only when detectron2 is absent; everything here is plain library convs
(cuDNN): the guide model runs under no_grad, outside the hot path, and feeds it.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from ..._compat import BACKBONE_REGISTRY, HAVE_DETECTRON2, Backbone, Conv2d, ShapeSpec, c2_xavier_fill


class FrozenBatchNorm2d(nn.Module):
    """detectron2.layers.FrozenBatchNorm2d [upstream]: y = x * weight * rsqrt(running_var + eps) + (bias - running_mean * scale); buffers only."""

    def __init__(self, num_features, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.register_buffer("weight", torch.ones(num_features))
        self.register_buffer("bias", torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features) - eps)

    def forward(self, x):
        scale = self.weight * (self.running_var + self.eps).rsqrt()
        bias = self.bias - self.running_mean * scale
        return x * scale.to(x.dtype).view(1, -1, 1, 1) + bias.to(x.dtype).view(1, -1, 1, 1)


def _conv(cin, cout, k, stride=1, pad=0):
    c = Conv2d(cin, cout, k, stride=stride, padding=pad, bias=False, norm=FrozenBatchNorm2d(cout))
    nn.init.kaiming_normal_(c.weight, mode="fan_out", nonlinearity="relu")      # c2_msra_fill
    return c


class BottleneckBlock(nn.Module):
    def __init__(self, cin, cout, mid, stride, stride_in_1x1=True):
        super().__init__()
        self.shortcut = _conv(cin, cout, 1, stride) if cin != cout else None
        s1, s3 = (stride, 1) if stride_in_1x1 else (1, stride)
        self.conv1 = _conv(cin, mid, 1, s1)
        self.conv2 = _conv(mid, mid, 3, s3, 1)
        self.conv3 = _conv(mid, cout, 1)

    def forward(self, x):
        out = F.relu_(self.conv1(x))
        out = F.relu_(self.conv2(out))
        out = self.conv3(out)
        return F.relu_(out + (self.shortcut(x) if self.shortcut is not None else x))


class BasicStem(nn.Module):
    def __init__(self, cin=3, cout=64):
        super().__init__()
        self.conv1 = _conv(cin, cout, 7, 2, 3)

    def forward(self, x):
        return F.max_pool2d(F.relu_(self.conv1(x)), kernel_size=3, stride=2, padding=1)


class ResNet(Backbone):
    """detectron2.modeling.backbone.ResNet [upstream] for depths 50 / 101 / 152, out_features res2..res5."""

    def __init__(self, depth=50, in_channels=3, stride_in_1x1=True, freeze_at=2):
        super().__init__()
        blocks = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}[depth]
        self.stem = BasicStem(in_channels, 64)
        cin, mid, cout = 64, 64, 256
        self._out_features, self._out_feature_channels, self._out_feature_strides = [], {}, {}
        for i, n in enumerate(blocks):
            stage = nn.Sequential(*[BottleneckBlock(cin if j == 0 else cout, cout, mid, 2 if (j == 0 and i > 0) else 1, stride_in_1x1)
                                    for j in range(n)])
            name = f"res{i + 2}"
            setattr(self, name, stage)
            self._out_features.append(name)
            self._out_feature_channels[name], self._out_feature_strides[name] = cout, 4 * 2 ** i
            cin, mid, cout = cout, mid * 2, cout * 2
        for i, m in enumerate([self.stem] + [getattr(self, n) for n in self._out_features]):     # FREEZE_AT: 1 = stem, 2 = + res2, ...
            if i < freeze_at:
                for p in m.parameters():
                    p.requires_grad = False

    def forward(self, x):
        out = {}
        x = self.stem(x)
        for name in self._out_features:
            x = getattr(self, name)(x)
            out[name] = x
        return out


class FPN(Backbone):
    """detectron2.modeling.backbone.FPN [upstream]: 1x1 laterals, nearest-2x top-down sum, 3x3 output convs, LastLevelMaxPool -> p2..p6."""

    def __init__(self, bottom_up, in_features, out_channels=256, fuse_type="sum"):
        super().__init__()
        self.bottom_up, self.in_features, self._fuse_type = bottom_up, list(in_features), fuse_type
        shapes = bottom_up.output_shape()
        strides = [shapes[f].stride for f in in_features]
        self._stages = []
        for f in in_features:
            stage = int(torch.tensor(float(shapes[f].stride)).log2())
            lat = Conv2d(shapes[f].channels, out_channels, 1)
            out = Conv2d(out_channels, out_channels, 3, padding=1)
            c2_xavier_fill(lat); c2_xavier_fill(out)
            self.add_module(f"fpn_lateral{stage}", lat)
            self.add_module(f"fpn_output{stage}", out)
            self._stages.append(stage)
        self._out_features = [f"p{s}" for s in self._stages] + [f"p{self._stages[-1] + 1}"]
        self._out_feature_strides = {f"p{s}": st for s, st in zip(self._stages, strides)}
        self._out_feature_strides[self._out_features[-1]] = strides[-1] * 2
        self._out_feature_channels = {k: out_channels for k in self._out_features}
        self._size_divisibility = strides[-1]

    @property
    def size_divisibility(self):
        return self._size_divisibility

    def forward(self, x):
        feats = self.bottom_up(x)
        res = {}
        prev = None
        for f, s in zip(self.in_features[::-1], self._stages[::-1]):
            lat = getattr(self, f"fpn_lateral{s}")(feats[f])
            if prev is not None:
                prev = lat + F.interpolate(prev, scale_factor=2.0, mode="nearest")
                if self._fuse_type == "avg":
                    prev = prev / 2
            else:
                prev = lat
            res[f"p{s}"] = getattr(self, f"fpn_output{s}")(prev)
        top = res[f"p{self._stages[-1]}"]
        res[self._out_features[-1]] = F.max_pool2d(top, kernel_size=1, stride=2, padding=0)       # LastLevelMaxPool
        return {k: res[k] for k in self._out_features}


def build_resnet_fpn_backbone(cfg, input_shape: ShapeSpec):
    """Stand-in for detectron2.modeling.backbone.fpn.build_resnet_fpn_backbone [upstream] (the default cfg.MODEL.GUIDE_BACKBONE.NAME)."""
    r = getattr(cfg.MODEL, "RESNETS", None)
    depth = int(getattr(r, "DEPTH", 50)) if r is not None else 50
    s1x1 = bool(getattr(r, "STRIDE_IN_1X1", True)) if r is not None else True
    freeze_at = int(getattr(getattr(cfg.MODEL, "GUIDE_BACKBONE", None), "FREEZE_AT", 2))
    bottom_up = ResNet(depth, input_shape.channels or 3, s1x1, freeze_at)
    in_features = list(getattr(cfg.MODEL.FPN, "IN_FEATURES", None) or ["res2", "res3", "res4", "res5"])
    return FPN(bottom_up, in_features, int(getattr(cfg.MODEL.FPN, "OUT_CHANNELS", 256)), getattr(cfg.MODEL.FPN, "FUSE_TYPE", "sum"))


if not HAVE_DETECTRON2:      # with detectron2 installed its own builder of this name is already in the registry
    BACKBONE_REGISTRY.register(build_resnet_fpn_backbone)
