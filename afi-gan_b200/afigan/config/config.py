"""get_cfg(): reference afigan/config/config.py:3-12.  With detectron2 installed this is detectron2's default CfgNode plus the AFI-GAN keys;
without it a small attribute-dict CfgNode carries the same key names so the drop-in modules and YAML key checks keep working."""
from __future__ import annotations

from .defaults import CfgNode, add_afigan_config


def get_cfg():
    try:  # pragma: no cover
        from detectron2.config import get_cfg as d2_get_cfg
        cfg = d2_get_cfg()
    except Exception:  # noqa: BLE001
        cfg = CfgNode()
        cfg.MODEL = CfgNode()
        cfg.MODEL.DEVICE = "cuda"
        cfg.MODEL.PIXEL_MEAN = [103.530, 116.280, 123.675]      # detectron2 defaults [upstream]
        cfg.MODEL.PIXEL_STD = [1.0, 1.0, 1.0]
        cfg.INPUT = CfgNode()
        cfg.INPUT.FORMAT = "BGR"
        cfg.INPUT.MIN_SIZE_TRAIN = (800,)
        cfg.INPUT.MAX_SIZE_TRAIN = 1333
        cfg.INPUT.MIN_SIZE_TRAIN_SAMPLING = "choice"
        cfg.MODEL.RESNETS = CfgNode()
        cfg.MODEL.FPN = CfgNode()
        cfg.MODEL.RESNETS.DEPTH = 50
        cfg.MODEL.RESNETS.STRIDE_IN_1X1 = True
        cfg.MODEL.FPN.IN_FEATURES = []
        cfg.MODEL.FPN.OUT_CHANNELS = 256
        cfg.MODEL.FPN.NORM = ""
        cfg.MODEL.FPN.FUSE_TYPE = "sum"
        cfg.SOLVER = CfgNode()
        cfg.SOLVER.BASE_LR = 0.001
        cfg.SOLVER.MOMENTUM = 0.9
        cfg.SOLVER.WEIGHT_DECAY = 0.0001
        cfg.SOLVER.WEIGHT_DECAY_NORM = 0.0
        cfg.SOLVER.IMS_PER_BATCH = 16
    return add_afigan_config(cfg)
