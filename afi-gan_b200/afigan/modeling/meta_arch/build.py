"""GUIDE_ARCH_REGISTRY / build_guide_model (reference afigan/modeling/meta_arch/build.py:5-20)."""
from __future__ import annotations

try:  # pragma: no cover
    from detectron2.utils.registry import Registry
    GUIDE_ARCH_REGISTRY = Registry("GUIDE_ARCH")
except Exception:  # noqa: BLE001
    from ..._compat import _Registry
    GUIDE_ARCH_REGISTRY = _Registry("GUIDE_ARCH")

META_ARCH_NAMES = ("GeneralizedRCNN_AFExtractor",)   # registered by the reference in detectron2's META_ARCH_REGISTRY (rcnn_extractor.py:21)


def _needs_reference(name: str, where: str):
    def factory(cfg):
        raise ImportError(f"{name} ({where}) feeds the AFI-GAN hot path with guide / detector features and is not re-implemented here "
                          "(SURVEY.md §2 row 9: out of scope); use the reference's implementation with detectron2 installed -- the interpolator, "
                          "discriminator, necks and loss blocks it calls are the drop-in modules of this package")
    factory.__name__ = name
    return factory


GUIDE_ARCH_REGISTRY.register(_needs_reference("RCNN_FPN_only", "afigan/modeling/meta_arch/rcnn_only.py:17"))


def build_guide_model(cfg):
    """Build the frozen guide feature extractor named by cfg.MODEL.GUIDE_ARCHITECTURE (build.py:14-20)."""
    return GUIDE_ARCH_REGISTRY.get(cfg.MODEL.GUIDE_ARCHITECTURE)(cfg)
