"""Registry surface of the reference's meta-architectures (afigan/modeling/meta_arch/).  The guide model `RCNN_FPN_only` (rcnn_only.py:17) is
implemented here (it produces the hot path's inputs); the stage-2 detector `GeneralizedRCNN_AFExtractor`
(rcnn_extractor.py:21) is detectron2's GeneralizedRCNN with an AFI neck and stays the reference's."""
from .build import GUIDE_ARCH_REGISTRY, META_ARCH_NAMES, build_guide_model  # noqa: F401
from .rcnn_only import RCNN_FPN_only, pad_to_divisible  # noqa: F401
