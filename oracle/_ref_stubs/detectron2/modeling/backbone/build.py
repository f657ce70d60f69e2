"""detectron2.modeling.backbone.build stand-in (test infrastructure)."""
from ...utils.registry import Registry

BACKBONE_REGISTRY = Registry("BACKBONE")
