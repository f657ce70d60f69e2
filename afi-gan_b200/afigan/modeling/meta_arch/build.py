"""GUIDE_ARCH_REGISTRY / build_guide_model (reference afigan/modeling/meta_arch/build.py:5-20)."""
from __future__ import annotations

try:  # pragma: no cover
    from detectron2.utils.registry import Registry
    GUIDE_ARCH_REGISTRY = Registry("GUIDE_ARCH")
except Exception:  # noqa: BLE001
    from ..._compat import _Registry
    GUIDE_ARCH_REGISTRY = _Registry("GUIDE_ARCH")

META_ARCH_NAMES = ("GeneralizedRCNN_AFExtractor",)   # registered by the reference in detectron2's META_ARCH_REGISTRY (rcnn_extractor.py:21)


def build_guide_model(cfg):
    """Build the frozen guide feature extractor named by cfg.MODEL.GUIDE_ARCHITECTURE (build.py:14-20)."""
    return GUIDE_ARCH_REGISTRY.get(cfg.MODEL.GUIDE_ARCHITECTURE)(cfg)
