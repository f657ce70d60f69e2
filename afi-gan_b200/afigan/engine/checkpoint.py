"""State-dict interop with the reference's checkpoints (reference afigan/engine/checkpoint.py:64-147).

The wire format is detectron2's `torch.save({"model": state_dict, ...})`; what is specific to AFI-GAN is the KEY handling when a trained
interpolator moves between stages:
  * stage 1 -> stage 2 (`load_AFIGEN_weight`, stage2_trainer.py:156-160): `Generators.*` is renamed to `backbone.srf_module.Generators.*`
    (checkpoint.py:94) and matched against the model by longest-suffix (checkpoint.py:136-147);
  * stage 2 -> stage 3 (`load_AFExtractor_weight`, stage3_trainer.py:103-107): only keys containing `srf_module` are kept (checkpoint.py:120).
Because this package keeps the reference's parameter names and shapes, these functions operate on plain dicts and the result can be fed to
`module.load_state_dict(..., strict=False)` of either implementation (weights trained here load in the reference and vice versa)."""
from __future__ import annotations

from typing import Dict, Mapping, Optional, Sequence, Tuple

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


class AFCheckpointer:
    """File-level counterpart of the reference's `AF_DetectionCheckpointer` (reference afigan/engine/checkpoint.py:14-76, on top of fvcore's
    `Checkpointer` [upstream]) without the fvcore / detectron2 dependency.  Wire format = `torch.save({"model": state_dict, **checkpointables})`
    with a `last_checkpoint` text file next to it; `module.` prefixes of DDP-wrapped models are stripped on save and tolerated on load; `.pkl`
    files in detectron2's model-zoo format (`{"model": {name: ndarray}, "__author__": ...}`) load too.  Files written here load with the
    reference's checkpointers into the reference's modules and vice versa (tests/test_checkpoint_interop.py)."""

    def __init__(self, model: torch.nn.Module, save_dir: str = "", save_to_disk: bool = True, **checkpointables):
        self.model = model.module if hasattr(model, "module") and isinstance(model.module, torch.nn.Module) else model
        self.save_dir, self.save_to_disk, self.checkpointables = save_dir, save_to_disk, dict(checkpointables)

    # ---- save / bookkeeping (fvcore Checkpointer.save / get_checkpoint_file / tag_last_checkpoint [upstream])
    def save(self, name: str, **extra) -> str:
        import os
        if not self.save_dir or not self.save_to_disk:
            return ""
        data = {"model": self.model.state_dict()}
        for k, obj in self.checkpointables.items():
            data[k] = obj.state_dict() if hasattr(obj, "state_dict") else obj
        data.update(extra)
        os.makedirs(self.save_dir, exist_ok=True)
        path = os.path.join(self.save_dir, f"{name}.pth")
        torch.save(data, path)
        with open(os.path.join(self.save_dir, "last_checkpoint"), "w") as f:
            f.write(os.path.basename(path))
        return path

    def has_checkpoint(self) -> bool:
        import os
        return bool(self.save_dir) and os.path.exists(os.path.join(self.save_dir, "last_checkpoint"))

    def get_checkpoint_file(self) -> str:
        import os
        try:
            with open(os.path.join(self.save_dir, "last_checkpoint")) as f:
                return os.path.join(self.save_dir, f.read().strip())
        except OSError:
            return ""

    # ---- load
    def _load_file(self, filename: str) -> Dict:
        """checkpoint.py:29-47."""
        if filename.endswith(".pkl"):
            import pickle
            with open(filename, "rb") as f:
                data = pickle.load(f, encoding="latin1")
            if "model" in data and "__author__" in data:
                data = dict(data)
                data["model"] = {k: torch.as_tensor(v) for k, v in data["model"].items()}
                return data
            raise NotImplementedError("Caffe2 / Detectron1 model-zoo pickles need detectron2's name-matching heuristics (c2_model_loading) [upstream]")
        loaded = torch.load(filename, map_location="cpu", weights_only=False)
        if "model" not in loaded:
            loaded = {"model": loaded}
        return loaded

    def _load_model(self, checkpoint: Dict) -> Dict[str, list]:
        """fvcore Checkpointer._load_model [upstream]: strip `module.`, skip shape mismatches, load non-strictly, report what did not match."""
        state = strip_module_prefix(checkpoint["model"])
        own = self.model.state_dict()
        incorrect = [(k, tuple(v.shape), tuple(own[k].shape)) for k, v in state.items() if k in own and tuple(v.shape) != tuple(own[k].shape)]
        for k, _, _ in incorrect:
            state.pop(k)
        res = self.model.load_state_dict(state, strict=False)
        return {"missing_keys": list(res.missing_keys), "unexpected_keys": list(res.unexpected_keys), "incorrect_shapes": incorrect}

    def load(self, path: str, checkpointables: Optional[Sequence[str]] = None) -> Dict:
        """Checkpointer.load [upstream]: the model plus the named checkpointables (optimizer, scheduler, ...); returns what is left (`iteration`)."""
        if not path:
            return {}
        checkpoint = self._load_file(path)
        self.last_incompatible = self._load_model(checkpoint)
        for k in (self.checkpointables if checkpointables is None else checkpointables):
            if k in checkpoint and hasattr(self.checkpointables.get(k), "load_state_dict"):
                self.checkpointables[k].load_state_dict(checkpoint.pop(k))
        checkpoint.pop("model", None)
        return checkpoint

    def resume_or_load(self, path: str, resume: bool = True) -> Dict:
        """Checkpointer.resume_or_load [upstream] as stage1_trainer.py:157-174 uses it."""
        if resume and self.has_checkpoint():
            return self.load(self.get_checkpoint_file())
        return self.load(path, checkpointables=[])

    def _load_AFExtractor_weights_file(self, filename: str) -> int:
        """checkpoint.py:64-69: a stage-1 generator checkpoint into a detector whose backbone owns `srf_module` (stage 1 -> stage 2)."""
        return load_generator_into_extractor(self.model, self._load_file(filename)["model"])

    def _load_TargetDetector_weights_file(self, filename: str) -> int:
        """checkpoint.py:71-76: only the `srf_module` weights of an AF-extractor checkpoint into the target detector (stage 2 -> stage 3)."""
        return load_extractor_into_detector(self.model, self._load_file(filename)["model"])
