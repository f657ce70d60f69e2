from .feat_interpol import Generator, Discriminator  # noqa: F401
from .backbone import (BiFPN_AFIGAN, FPN_AFIGAN, PAFPN_AFIGAN, bifpn_feature_fusion, build_resnet_fpn_sr_backbone,  # noqa: F401
                       build_resnest_fpn_sr_backbone, build_resnet_pafpn_sr_backbone, build_resnest_pafpn_sr_backbone)
from .meta_arch import GUIDE_ARCH_REGISTRY, build_guide_model  # noqa: F401
