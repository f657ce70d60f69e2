"""State-dict interop with the reference's checkpoints (reference afigan/engine/checkpoint.py:64-147).

The wire format is detectron2's `torch.save({"model": state_dict, ...})`; what is specific to AFI-GAN is the KEY handling when a trained
interpolator moves between stages:
  * stage 1 -> stage 2 (`load_AFIGEN_weight`, stage2_trainer.py:156-160): `Generators.*` is renamed to `backbone.srf_module.Generators.*`
    (checkpoint.py:94) and matched against the model by longest-suffix (checkpoint.py:136-147);
  * stage 2 -> stage 3 (`load_AFExtractor_weight`, stage3_trainer.py:103-107): only keys containing `srf_module` are kept (checkpoint.py:120).
Because this package keeps the reference's parameter names and shapes, these functions operate on plain dicts and the result can be fed to
`module.load_state_dict(..., strict=False)` of either implementation (weights trained here load in the reference and vice versa)."""
from __future__ import annotations

from typing import Dict, Mapping, Tuple

import torch


def strip_module_prefix(state: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """DDP-wrapped models are checkpointed with a `module.` prefix that fvcore strips on load [upstream]."""
    if state and all(k.startswith("module.") for k in state):
        return {k[len("module."):]: v for k, v in state.items()}
    return dict(state)


def convert_afi_names(weights: Mapping[str, torch.Tensor]) -> Tuple[Dict[str, torch.Tensor], Dict[str, str]]:
    """`Generators...` -> `backbone.srf_module.Generators...` (checkpoint.py:79-108). Returns (renamed weights, new -> original key)."""
    new, back = {}, {}
    for k in sorted(weights):
        r = k.replace("Generators", "backbone.srf_module.Generators")
        assert r not in new, f"key collision on {r}"
        new[r], back[r] = weights[k], k
    return new, back


def remain_only_afi_names(weights: Mapping[str, torch.Tensor]) -> Tuple[Dict[str, torch.Tensor], Dict[str, str]]:
    """Keep only the interpolator's keys (those containing `srf_module`; checkpoint.py:110-125)."""
    new = {k: v for k, v in weights.items() if "srf_module" in k}
    return new, {k: k for k in new}


def align_by_suffix(model_state: Mapping[str, torch.Tensor], ckpt_state: Mapping[str, torch.Tensor]) -> Dict[str, str]:
    """model key -> checkpoint key, where the checkpoint key equals the model key or is a complete ('.'-delimited) suffix of it; the
    longest match wins (checkpoint.py:136-147, detectron2's align_and_update_state_dicts heuristic)."""
    out = {}
    ckpt_keys = sorted(ckpt_state)
    for mk in sorted(model_state):
        best = ""
        for ck in ckpt_keys:
            if (mk == ck or mk.endswith("." + ck)) and len(ck) > len(best):
                best = ck
        if best and tuple(model_state[mk].shape) == tuple(ckpt_state[best].shape):
            out[mk] = best
    return out


def load_generator_into_extractor(model: torch.nn.Module, generator_ckpt: Mapping[str, torch.Tensor]) -> int:
    """Stage 1 -> stage 2: load a stage-1 `Generator` checkpoint ('model' dict) into a detector whose backbone owns `srf_module`."""
    ckpt, _ = convert_afi_names(strip_module_prefix(generator_ckpt))
    state = model.state_dict()
    mapping = align_by_suffix(state, ckpt)
    with torch.no_grad():
        for mk, ck in mapping.items():
            state[mk].copy_(ckpt[ck])
    return len(mapping)


def load_extractor_into_detector(model: torch.nn.Module, extractor_ckpt: Mapping[str, torch.Tensor]) -> int:
    """Stage 2 -> stage 3: copy only the `srf_module` weights of an AF-extractor checkpoint into the target detector."""
    ckpt, _ = remain_only_afi_names(strip_module_prefix(extractor_ckpt))
    state = model.state_dict()
    mapping = align_by_suffix(state, ckpt)
    with torch.no_grad():
        for mk, ck in mapping.items():
            state[mk].copy_(ckpt[ck])
    return len(mapping)
