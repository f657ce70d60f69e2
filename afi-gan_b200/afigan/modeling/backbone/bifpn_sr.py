"""The AF-interpolator merge site of the BiFPN neck (reference afigan/modeling/backbone/bifpn_sr.py:535-548).

`BiFPN_AFIGAN._feature_funsion(layer, cur, top, indice)` up-samples `top` with the shared `srf_module` and fuses it with `cur` using the
RAW (un-normalised, ReLU-less) attention weights `BiFPNLayer_{l}_p{i}_w1` (`_attention`, :535-537; `_weight_act` is dead code, App. D-9),
or a plain sum when attention is off.  The surrounding depthwise-separable convs / BN / swish / max-pool of the seven hand-unrolled layers
(:583-729) are out of the hot path and stay whatever the caller uses."""
from __future__ import annotations

from typing import Optional

import torch

from ..._compat import BACKBONE_REGISTRY, ShapeSpec


def bifpn_feature_fusion(srf_module, cur_feature: torch.Tensor, top_feature: torch.Tensor,
                         weight: Optional[torch.Tensor] = None) -> torch.Tensor:
    """weight is the 2-element `..._w1` parameter (or None when attention is disabled): w[0]*cur + w[1]*AFI(top).
    Without autograd (inference, BASELINE config C5: 28 of these per image) the whole site is one library call; with autograd the
    interpolator is the library's autograd Function and the two-term fusion stays in torch."""
    needs_grad = torch.is_grad_enabled() and (cur_feature.requires_grad or top_feature.requires_grad or (weight is not None and weight.requires_grad)
                                              or any(p.requires_grad for p in srf_module.parameters()))
    if not needs_grad and top_feature.is_cuda and hasattr(srf_module, "fuse"):
        return srf_module.fuse(top_feature, cur_feature, weight)
    up = srf_module(top_feature, out_hw=tuple(cur_feature.shape[2:]))
    if weight is None:
        return cur_feature + up
    return cur_feature * weight[0] + up * weight[1]


@BACKBONE_REGISTRY.register()
def build_swint_bifpn_sr_backbone(cfg, input_shape: ShapeSpec):
    """bifpn_sr.py:791-816.  The Swin bottom-up and the seven hand-unrolled BiFPN layers are not re-implemented;
    only their interpolator fusion site is native (`bifpn_feature_fusion`)."""
    raise ImportError("build_swint_bifpn_sr_backbone needs the reference's Swin backbone and BiFPN layer stack (out of the hot path's scope); "
                      "patch the reference BiFPN_AFIGAN._feature_funsion to call afigan.modeling.backbone.bifpn_sr.bifpn_feature_fusion")
