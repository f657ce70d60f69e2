"""Runs the reference's stage-1 step on the CPU with the reference's OWN module files (oracle/_ref).  TEST INFRASTRUCTURE:
only bench.py's cpu_baseline / --impl reference legs and tests/ import this.

The trainer (afigan/engine/stage1_trainer.py) cannot be imported -- it pulls in detectron2's engine / data / solver -- so its
run_step body is restated here line by line on top of the unmodified Generator / Discriminator classes:
  build_model          stage1_trainer.py:495-514   Generator(n_residual_dense_blocks=3), Discriminator()
  optimisers           stage1_trainer.py:109-114   SGD over Generators[0] / Discriminators[0]; detectron2 build_optimizer defaults
                                                    [upstream]: lr BASE_LR, momentum 0.9, weight decay 1e-4, norm parameters 0
  D phase              stage1_trainer.py:334-381
  G phase              stage1_trainer.py:384-433
  crop                 stage1_trainer.py:437-443
"""
import importlib.util
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in ("generator_rdb.py", "feature_patch_discriminator.py"))


def load_modules():
    """(generator_rdb module, feature_patch_discriminator module) loaded from oracle/_ref through the detectron2 / fvcore stand-ins."""
    stubs = os.path.join(HERE, "_ref_stubs")
    if stubs not in sys.path:
        sys.path.insert(0, stubs)
    mods = []
    for name in ("generator_rdb", "feature_patch_discriminator"):
        spec = importlib.util.spec_from_file_location("afigan_ref_" + name, os.path.join(REF_DIR, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods.append(m)
    return tuple(mods)


def _reshape(t, size):      # stage1_trainer.py:437-443
    if size[2] != t.size(2) or size[3] != t.size(3):
        return t[:, :, 0:min(size[2], t.size(2)), 0:min(size[3], t.size(3))]
    return t


class ReferenceStage1:
    """The reference modules + optimisers of one trainer process (seed 0: G first, then D, stage1_trainer.py:505-506)."""

    def __init__(self, lr=1e-3, momentum=0.9, weight_decay=1e-4, weight_decay_norm=0.0, seed=0):
        gen_mod, dis_mod = load_modules()
        torch.manual_seed(seed)
        self.G = gen_mod.Generator(n_residual_dense_blocks=3)
        self.D = dis_mod.Discriminator()
        self.G.Generators[0].train()
        self.D.Discriminators[0].train()

        def groups(module):
            norm, rest = [], []
            for m in module.modules():
                for p in m.parameters(recurse=False):
                    (norm if isinstance(m, nn.BatchNorm2d) else rest).append(p)
            g = [{"params": rest, "weight_decay": weight_decay}]
            if norm:
                g.append({"params": norm, "weight_decay": weight_decay_norm})
            return g

        self.opt_g = torch.optim.SGD(groups(self.G.Generators[0]), lr=lr, momentum=momentum)
        self.opt_d = torch.optim.SGD(groups(self.D.Discriminators[0]), lr=lr, momentum=momentum)
        self.crit = nn.BCEWithLogitsLoss()

    def run_step(self, lr_feats, hr_feats):
        G, D0, crit = self.G, self.D.Discriminators[0], self.crit
        d_loss = {}
        for lv, (lo, hi) in enumerate(zip(lr_feats, hr_feats), 2):
            tr = G(lo).detach()
            tr = _reshape(tr, hi.size())
            hi = _reshape(hi, tr.size())
            real, fake = D0(hi), D0(tr)
            d_loss[f"d_loss_p{lv}"] = crit(real, torch.ones_like(real)) + crit(fake, torch.zeros_like(fake))
        self.opt_d.zero_grad()
        sum(d_loss.values()).backward()
        self.opt_d.step()
        g_loss = {}
        for lv, (lo, hi) in enumerate(zip(lr_feats, hr_feats), 2):
            tr = G(lo)
            tr = _reshape(tr, hi.size())
            hi = _reshape(hi, tr.size())
            fake = D0(tr).detach()
            _ = D0(hi)
            g_loss[f"g_loss_p{lv}"] = crit(fake, torch.ones_like(fake)) * 1e-3 + F.l1_loss(tr, hi)
        self.opt_g.zero_grad()
        sum(g_loss.values()).backward()
        self.opt_g.step()
        return {k: float(v.detach()) for k, v in {**d_loss, **g_loss}.items()}
