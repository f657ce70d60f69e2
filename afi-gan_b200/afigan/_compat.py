"""detectron2 / fvcore interop.  When detectron2 is importable the drop-in modules subclass ITS Backbone and register into ITS
registries (so stage*_train.py / run_net.py find them); otherwise small local stand-ins keep the package importable and testable."""
from __future__ import annotations

from collections import namedtuple

from torch import nn

try:  # pragma: no cover - detectron2 is not installed in the build image
    from detectron2.layers import ShapeSpec
    from detectron2.modeling.backbone import Backbone
    from detectron2.modeling.backbone.build import BACKBONE_REGISTRY
    HAVE_DETECTRON2 = True
except Exception:  # noqa: BLE001
    HAVE_DETECTRON2 = False

    class ShapeSpec(namedtuple("_ShapeSpec", ["channels", "height", "width", "stride"])):
        def __new__(cls, *, channels=None, height=None, width=None, stride=None):
            return super().__new__(cls, channels, height, width, stride)

    class Backbone(nn.Module):
        """Minimal stand-in for detectron2.modeling.Backbone: forward(x) -> dict[str, Tensor], output_shape(), size_divisibility."""

        @property
        def size_divisibility(self):
            return 0

        def output_shape(self):
            return {name: ShapeSpec(channels=self._out_feature_channels[name], stride=self._out_feature_strides[name])
                    for name in self._out_features}

    class _Registry:
        def __init__(self, name):
            self._name, self._obj_map = name, {}

        def register(self, obj=None):
            if obj is None:
                return lambda fn: self.register(fn)
            self._obj_map[obj.__name__] = obj
            return obj

        def get(self, name):
            if name not in self._obj_map:
                raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
            return self._obj_map[name]

        def __contains__(self, name):
            return name in self._obj_map

    BACKBONE_REGISTRY = _Registry("BACKBONE")


def c2_xavier_fill(module: nn.Module) -> None:
    """fvcore.nn.weight_init.c2_xavier_fill [upstream]: kaiming_uniform_(a=1), bias 0."""
    nn.init.kaiming_uniform_(module.weight, a=1)
    if module.bias is not None:
        nn.init.constant_(module.bias, 0)


def get_norm(norm, out_channels):
    if norm is None or (isinstance(norm, str) and len(norm) == 0):
        return None
    if isinstance(norm, str):
        norm = {"BN": nn.BatchNorm2d, "SyncBN": nn.SyncBatchNorm, "GN": lambda c: nn.GroupNorm(32, c)}[norm]
    return norm(out_channels)


class Conv2d(nn.Conv2d):
    """detectron2.layers.Conv2d semantics: conv -> norm -> activation."""

    def __init__(self, *args, **kwargs):
        norm = kwargs.pop("norm", None)
        activation = kwargs.pop("activation", None)
        super().__init__(*args, **kwargs)
        self.norm = norm
        self.activation = activation

    def forward(self, x):
        x = super().forward(x)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x
