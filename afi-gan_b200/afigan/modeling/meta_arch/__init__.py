"""Registry surface of the reference's meta-architectures (afigan/modeling/meta_arch/): the guide model `RCNN_FPN_only` (rcnn_only.py:17) and
the stage-2 detector `GeneralizedRCNN_AFExtractor` (rcnn_extractor.py:21) PRODUCE the hot path's inputs and are out of this repository's scope
(SURVEY.md §2 row 9).  The names stay importable and registered so that configs resolve; building them needs detectron2 and the
reference's own implementation."""
from .build import GUIDE_ARCH_REGISTRY, META_ARCH_NAMES, build_guide_model  # noqa: F401
