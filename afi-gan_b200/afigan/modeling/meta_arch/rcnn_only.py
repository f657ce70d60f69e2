"""The frozen guide feature extractor of stage 1 / stage 2: drop-in for reference afigan/modeling/meta_arch/rcnn_only.py:17-60.

`RCNN_FPN_only(cfg).forward(batched_inputs, img_dict_name)` normalises the named image of every dataset dict, pads the batch to the
backbone's size divisibility (detectron2 ImageList.from_tensors semantics [upstream]: top-left aligned, zero padded) and returns
`[{"features": {"p2": ..., ..., "p6": ...}}]` -- the producer of the hot path's inputs (stage1_trainer.py:320-327).
`extract_pair` is the fast path a stage-1 step needs: both scales under no_grad, bf16 channels_last convs, results detached fp32 lists.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
from torch import nn

from ..._compat import BACKBONE_REGISTRY, Backbone, ShapeSpec
from .build import GUIDE_ARCH_REGISTRY

__all__ = ["RCNN_FPN_only", "pad_to_divisible"]


def pad_to_divisible(images: Sequence[torch.Tensor], size_divisibility: int, pad_value: float = 0.0) -> torch.Tensor:
    """detectron2.structures.ImageList.from_tensors [upstream]: stack [C,H,W] images into [N,C,Hmax,Wmax] rounded up to the divisibility."""
    h = max(int(t.shape[-2]) for t in images)
    w = max(int(t.shape[-1]) for t in images)
    if size_divisibility > 1:
        h = (h + size_divisibility - 1) // size_divisibility * size_divisibility
        w = (w + size_divisibility - 1) // size_divisibility * size_divisibility
    out = images[0].new_full((len(images), images[0].shape[0], h, w), pad_value)
    for o, t in zip(out, images):
        o[:, : t.shape[-2], : t.shape[-1]].copy_(t)
    return out


@GUIDE_ARCH_REGISTRY.register()
class RCNN_FPN_only(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.device = torch.device(cfg.MODEL.DEVICE)
        self.backbone = self.build_backbone(cfg)
        self.input_format = getattr(getattr(cfg, "INPUT", None), "FORMAT", "BGR")
        assert len(cfg.MODEL.PIXEL_MEAN) == len(cfg.MODEL.PIXEL_STD)
        c = len(cfg.MODEL.PIXEL_MEAN)
        self.register_buffer("pixel_mean", torch.tensor(cfg.MODEL.PIXEL_MEAN, dtype=torch.float32).view(c, 1, 1), persistent=False)
        self.register_buffer("pixel_std", torch.tensor(cfg.MODEL.PIXEL_STD, dtype=torch.float32).view(c, 1, 1), persistent=False)
        self.to(self.device)

    def normalizer(self, x):
        return (x - self.pixel_mean) / self.pixel_std

    def _batch(self, batched_inputs, img_dict_name):
        images = [self.normalizer(x[img_dict_name].to(self.device, non_blocking=True).float()) for x in batched_inputs]
        return pad_to_divisible(images, self.backbone.size_divisibility)

    def forward(self, batched_inputs: List[Dict], img_dict_name: str = "image"):
        """rcnn_only.py:34-44."""
        features = self.backbone(self._batch(batched_inputs, img_dict_name))
        return [{"features": features}]

    @torch.no_grad()
    def extract_pair(self, batched_inputs: List[Dict], hr_name: str = "image", lr_name: str = "image_x0.5",
                     levels: Sequence[str] = ("p2", "p3", "p4", "p5", "p6"), bf16: bool = True) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
        """(lr_features, hr_features) of stage1_trainer.py:320-327 as detached fp32 lists p2..p6: two guide forwards under no_grad (the
        reference runs them with autograd on and detaches, App. D-6), channels_last and bf16 autocast on CUDA."""
        out = []
        for name in (lr_name, hr_name):
            x = self._batch(batched_inputs, name)
            if x.is_cuda:
                x = x.contiguous(memory_format=torch.channels_last)
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16):
                    f = self.backbone(x)
            else:
                f = self.backbone(x)
            out.append([f[k].float().contiguous() for k in levels])
        return out[0], out[1]

    def build_backbone(self, cfg, input_shape=None):
        """rcnn_only.py:47-60: the backbone named by cfg.MODEL.GUIDE_BACKBONE.NAME."""
        if input_shape is None:
            input_shape = ShapeSpec(channels=len(cfg.MODEL.PIXEL_MEAN))
        backbone = BACKBONE_REGISTRY.get(cfg.MODEL.GUIDE_BACKBONE.NAME)(cfg, input_shape)
        assert isinstance(backbone, Backbone)
        return backbone
