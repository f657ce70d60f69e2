"""The AF-interpolator merge site of the BiFPN neck (reference afigan/modeling/backbone/bifpn_sr.py:535-548).

`BiFPN_AFIGAN._feature_funsion(layer, cur, top, indice)` up-samples `top` with the shared `srf_module` and fuses it with `cur` using the
RAW (un-normalised, ReLU-less) attention weights `BiFPNLayer_{l}_p{i}_w1` (`_attention`, :535-537; `_weight_act` is dead code, App. D-9),
or a plain sum when attention is off.  The surrounding depthwise-separable convs / BN / swish / max-pool of the seven hand-unrolled layers
(:583-729) are out of the hot path and stay whatever the caller uses."""
from __future__ import annotations

from typing import Optional

import torch


def bifpn_feature_fusion(srf_module, cur_feature: torch.Tensor, top_feature: torch.Tensor,
                         weight: Optional[torch.Tensor] = None) -> torch.Tensor:
    """weight is the 2-element `..._w1` parameter (or None when attention is disabled): w[0]*cur + w[1]*AFI(top)."""
    up = srf_module(top_feature, out_hw=tuple(cur_feature.shape[2:]))
    if weight is None:
        return cur_feature + up
    return cur_feature * weight[0] + up * weight[1]
