"""Paired-scale dataset mapper: the data side of stage 1 / stage 2 (reference afigan/engine/dataset_mapper.py:70-193 and the transform
plumbing of afigan/engine/transform_gen.py:456-559).

For every image the reference emits the usual detectron2 fields plus a HALF-scale copy that shares the random flip:
  image        = flip?(ResizeShortestEdge(min_size, max_size)(img))                           [upstream detectron2 transforms]
  image_x0.5   = flip?(resize(img, (int(0.5 * new_h), int(0.5 * new_w))))                     transform_gen.py:540-552
  width_x0.5 / heigth_x0.5 (sic, dataset_mapper.py:121), instances / instances_x0.5 (boxes transformed with each scale's transforms).
Both resizes start from the ORIGINAL image (the LR image is not a down-sampling of the HR one) and use PIL bilinear interpolation like
detectron2's ResizeTransform.  This is CPU-side loader code (numpy / PIL), deliberately free of detectron2 imports."""
from __future__ import annotations

import copy
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


def shortest_edge_size(h: int, w: int, size: int, max_size: int) -> Tuple[int, int]:
    """detectron2 ResizeShortestEdge.get_transform [upstream]: scale so the short side is `size`, cap the long side at max_size, round half up."""
    scale = size * 1.0 / min(h, w)
    newh, neww = (size, scale * w) if h < w else (scale * h, size)
    if max(newh, neww) > max_size:
        s = max_size * 1.0 / max(newh, neww)
        newh, neww = newh * s, neww * s
    return int(newh + 0.5), int(neww + 0.5)


def resize_image(img: np.ndarray, new_h: int, new_w: int) -> np.ndarray:
    """detectron2 ResizeTransform.apply_image [upstream] for uint8 HWC images: PIL bilinear."""
    from PIL import Image
    assert img.dtype == np.uint8 and img.ndim == 3
    pil = Image.fromarray(img if img.shape[2] != 1 else img[:, :, 0])
    out = np.asarray(pil.resize((new_w, new_h), Image.BILINEAR))
    return out if out.ndim == 3 else out[:, :, None]


class PairedScaleMapper:
    def __init__(self, min_size: Sequence[int] = (800,), max_size: int = 1333, scale_ratio: Sequence[float] = (0.5,), flip_prob: float = 0.5,
                 is_train: bool = True, sample_style: str = "choice", rng: Optional[np.random.Generator] = None):
        assert list(scale_ratio) == [0.5], "the reference hard-codes the 0.5x copy (dataset_mapper.py:104, transform_gen.py:540-552)"
        self.min_size, self.max_size, self.scale_ratio = tuple(min_size), max_size, tuple(scale_ratio)
        self.flip_prob, self.is_train, self.sample_style = flip_prob if is_train else 0.0, is_train, sample_style
        self.rng = rng or np.random.default_rng()

    @classmethod
    def from_config(cls, cfg, is_train=True):
        """DatasetMapper(cfg, [0.5], is_train) as stage1_trainer.py:602 builds it."""
        inp = cfg.INPUT
        return cls(inp.MIN_SIZE_TRAIN if is_train else (inp.MIN_SIZE_TEST,), inp.MAX_SIZE_TRAIN if is_train else inp.MAX_SIZE_TEST, (0.5,),
                   0.5 if is_train else 0.0, is_train, getattr(inp, "MIN_SIZE_TRAIN_SAMPLING", "choice"))

    def _pick_size(self) -> int:
        if self.sample_style == "range":
            return int(self.rng.integers(self.min_size[0], self.min_size[-1] + 1))
        return int(self.min_size[int(self.rng.integers(0, len(self.min_size)))])

    @staticmethod
    def _boxes(annos: List[dict], sx: float, sy: float, flip: bool, width: int) -> torch.Tensor:
        """XYXY_ABS boxes through resize (+ horizontal flip) -- transform_instance_annotations [upstream] restricted to boxes."""
        out = []
        for a in annos:
            if a.get("iscrowd", 0):
                continue
            x0, y0, x1, y1 = a["bbox"]
            if a.get("bbox_mode", "XYXY_ABS") in ("XYWH_ABS", 1):
                x1, y1 = x0 + x1, y0 + y1
            x0, x1, y0, y1 = x0 * sx, x1 * sx, y0 * sy, y1 * sy
            if flip:
                x0, x1 = width - x1, width - x0
            out.append([x0, y0, x1, y1])
        return torch.tensor(out, dtype=torch.float32).reshape(-1, 4)

    def __call__(self, dataset_dict: Dict) -> Dict:
        d = copy.deepcopy(dataset_dict)
        img = d.pop("image_array") if "image_array" in d else None
        if img is None:
            from PIL import Image
            img = np.asarray(Image.open(d["file_name"]).convert("RGB"))[:, :, ::-1]       # BGR, cfg.INPUT.FORMAT default
        img = np.ascontiguousarray(img)
        h0, w0 = img.shape[:2]
        new_h, new_w = shortest_edge_size(h0, w0, self._pick_size(), self.max_size)
        lr_h, lr_w = int(new_h * 0.5), int(new_w * 0.5)                                    # transform_gen.py:540-543
        flip = bool(self.rng.random() < self.flip_prob)
        hr, lr = resize_image(img, new_h, new_w), resize_image(img, lr_h, lr_w)            # both from the ORIGINAL image (dataset_mapper.py:79, 92)
        if flip:
            hr, lr = hr[:, ::-1], lr[:, ::-1]                                              # the LR copy re-uses the HR flip (transform_gen.py:545-552)
        d["image"] = torch.as_tensor(np.ascontiguousarray(hr.transpose(2, 0, 1)))
        if self.is_train:
            d["image_x0.5"] = torch.as_tensor(np.ascontiguousarray(lr.transpose(2, 0, 1)))
        # (the reference stores int(C*0.5) / int(H*0.5) of the CHW tensor under these names, dataset_mapper.py:120-121 -- kept verbatim)
        d["width_x0.5"], d["heigth_x0.5"] = int(d["image"].shape[1] * 0.5), int(d["image"].shape[2] * 0.5)
        if not self.is_train:
            d.pop("annotations", None)
            return d
        if "annotations" in d:
            annos = d.pop("annotations")
            d["instances"] = {"gt_boxes": self._boxes(annos, new_w / w0, new_h / h0, flip, new_w), "image_size": (new_h, new_w)}
            d["instances_x0.5"] = {"gt_boxes": self._boxes(annos, lr_w / w0, lr_h / h0, flip, lr_w), "image_size": (lr_h, lr_w)}
        return d
