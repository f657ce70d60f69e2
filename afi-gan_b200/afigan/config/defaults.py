"""The AFI-GAN configuration keys (names and defaults of reference afigan/config/defaults.py:5-94; the key NAMES are part of the drop-in
surface, the mechanism is yacs / detectron2's)."""
from __future__ import annotations


class CfgNode(dict):
    """Tiny attribute-dict used only when detectron2 (and its yacs CfgNode) is not installed."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def merge_from_list(self, kv):
        assert len(kv) % 2 == 0
        for k, v in zip(kv[0::2], kv[1::2]):
            node = self
            *path, leaf = k.split(".")
            for p in path:
                node = node[p]
            if leaf not in node:
                raise KeyError(f"Non-existent config key: {k}")
            node[leaf] = v
        return self


def _node(cfg):
    return type(cfg)() if type(cfg).__name__ == "CfgNode" else CfgNode()


def add_afigan_config(cfg):
    m = cfg.MODEL
    m.GUIDE_ARCHITECTURE = ""            # defaults.py:5
    m.GUIDE_WEIGHTS = ""                 # :7
    m.AFI_GEN_WEIGHTS = ""               # :8
    m.AFI_DIS_WEIGHTS = ""               # :9
    m.AF_EXTRACTOR_WEIGHTS = ""          # :10
    m.AFI_FREEZE = False                 # :11  read by the necks at construction (fpn_sr.py:67)
    m.GUIDE_BACKBONE = _node(cfg)        # :16-22
    m.GUIDE_BACKBONE.NAME = "build_resnet_fpn_backbone"
    m.GUIDE_BACKBONE.FREEZE_AT = 2
    if "RESNETS" not in m:
        m.RESNETS = _node(cfg)
    m.RESNETS.RADIX = 1                  # :32-41 (ResNeSt)
    m.RESNETS.BOTTLENECK_WIDTH = 64
    m.RESNETS.DEEP_STEM = False
    m.RESNETS.AVD = False
    m.RESNETS.AVG_DOWN = False
    m.BIFPN = _node(cfg)                 # :47-59
    m.BIFPN.IN_FEATURES = []
    m.BIFPN.OUT_CHANNELS = 256
    m.BIFPN.FPN_REPEAT = 3
    m.BIFPN.NORM = "SyncBN"
    m.BIFPN.FUSE_TYPE = "sum"
    m.SWINT = _node(cfg)                 # :65-73
    m.SWINT.EMBED_DIM = 96
    m.SWINT.OUT_FEATURES = ["stage2", "stage3", "stage4", "stage5"]
    m.SWINT.DEPTHS = [2, 2, 6, 2]
    m.SWINT.NUM_HEADS = [3, 6, 12, 24]
    m.SWINT.WINDOW_SIZE = 7
    m.SWINT.MLP_RATIO = 4
    m.SWINT.DROP_PATH_RATE = 0.2
    m.SWINT.APE = False
    s = cfg.SOLVER
    s.OPTIMIZER = "SGD"                  # :81-94 (never read by the reference's own code)
    if "AMP" not in s:
        s.AMP = _node(cfg)
        s.AMP.ENABLED = False
    if "CLIP_GRADIENTS" not in s:
        s.CLIP_GRADIENTS = _node(cfg)
        s.CLIP_GRADIENTS.ENABLED = False
        s.CLIP_GRADIENTS.CLIP_TYPE = "value"
        s.CLIP_GRADIENTS.CLIP_VALUE = 1.0
        s.CLIP_GRADIENTS.NORM_TYPE = 2.0
    return cfg
