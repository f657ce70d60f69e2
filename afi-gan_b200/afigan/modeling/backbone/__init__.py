from .fpn_sr import FPN_AFIGAN, LastLevelMaxPool, build_resnet_fpn_sr_backbone, build_resnest_fpn_sr_backbone  # noqa: F401
from .pafpn_sr import PAFPN_AFIGAN, build_resnet_pafpn_sr_backbone, build_resnest_pafpn_sr_backbone  # noqa: F401
from .bifpn_sr import (BeforeBiFPNLayer, BiFPN_AFIGAN, LastLevelP6P7, ResampleFeature, bifpn_feature_fusion,  # noqa: F401
                       build_swint_bifpn_sr_backbone)
from .guide_resnet import FPN, ResNet, build_resnet_fpn_backbone  # noqa: F401
