"""CPU oracle for the AFI-GAN hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module, and only as the checker or the timed CPU baseline.  The product path
(afi-gan_b200/afigan + libafigan_b200.so) never imports it and has no CPU fallback.

What it is: a functional restatement (plain torch CPU ops on explicit state dicts) of

  * Generator.__init__/forward            /root/reference/afigan/modeling/feat_interpol/generator_rdb.py:75-130
  * ResidualInResidual / ResidualDenseBlock                                  generator_rdb.py:15-71
  * Discriminator.__init__ / Discriminators[0]   .../feat_interpol/feature_patch_discriminator.py:18-55
  * the stage-1 loss block                       /root/reference/afigan/engine/stage1_trainer.py:334-443
  * the FPN top-down merge                       /root/reference/afigan/modeling/backbone/fpn_sr.py:147-158
  * the stage-2 loss block's data flow           /root/reference/afigan/engine/stage2_trainer.py:298-384

Third-party arithmetic: conv2d / conv_transpose2d / batch_norm / leaky_relu / bilinear upsample /
BCE-with-logits / L1 all live in PyTorch, which the reference does not pin (requirements.txt pins
only timm and dataclasses; README recommends a detectron2 v0.1.1-era stack).  The oracle therefore is
torch (2.11.0 in this image) CPU fp32 -- or fp64 when dtype=torch.float64 is passed -- executing the
math of App. A of SURVEY.md.

Pinning: the reference ships NO tests, golden vectors or fixtures for this path, so
parity is pinned the second way the task allows: tests/golden/make_golden.py imports the UNMODIFIED
reference modules (through oracle/_ref_stubs) in the dev container, checks that this restatement
reproduces their state dicts, outputs, losses and gradients, and commits small fixtures
(tests/golden/*.npz) that tests/test_oracle.py re-checks everywhere (incl. the GPU box, where
/root/reference does not exist).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

LRELU = 0.2
BN_EPS = 1e-5
BN_MOMENTUM = 0.1
StateDict = Dict[str, torch.Tensor]


# ------------------------------------------------------------------------------------------------
# Initialisation: consumes the global torch RNG in exactly the reference's order
# ------------------------------------------------------------------------------------------------
def _default_conv_init(weight: torch.Tensor, bias: Optional[torch.Tensor]) -> None:
    """torch.nn.modules.conv._ConvNd.reset_parameters (what nn.Conv2d/ConvTranspose2d do on construction)."""
    torch.nn.init.kaiming_uniform_(weight, a=math.sqrt(5))
    if bias is not None:
        fan_in = weight.size(1) * weight[0][0].numel()
        if fan_in != 0:
            bound = 1 / math.sqrt(fan_in)
            torch.nn.init.uniform_(bias, -bound, bound)


def init_generator_state(in_channels: int = 256, n_rdb: int = 3, growth: int = 32) -> "OrderedDict[str, torch.Tensor]":
    """State dict of Generator(in_channels, n_rdb, growth) as the reference would build it under the
    current global RNG state.  generator_rdb.py:34-62 (RDB init), :75-121 (Generator init)."""
    C, g = in_channels, growth
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def new_conv(cout, cin, k, bias):
        w = torch.empty(cout, cin, k, k)
        b = torch.empty(cout) if bias else None
        _default_conv_init(w, b)
        return w, b

    # [0] head conv (d2 Conv2d, bias=True)                                     generator_rdb.py:91-93
    w00, b00 = new_conv(C, C, 3, True)
    # [1] RiR: per RDB 5 default-initialised convs, then kaiming_normal_*0.1 in modules() order  :34-62
    rdb = []
    for r in range(n_rdb):
        ws = [new_conv(g, C + i * g, 3, False)[0] for i in range(4)]
        ws.append(new_conv(C, C + 4 * g, 3, False)[0])
        for w in ws:
            torch.nn.init.kaiming_normal_(w)
            w.mul_(0.1)
        rdb.append(ws)
    # [2] post conv, [3] transposed conv (weight [C_in, C_out, 6, 6]), [4] output conv          :97-108
    w20, b20 = new_conv(C, C, 3, True)
    w30 = torch.empty(C, C, 6, 6)
    b30 = torch.empty(C)
    _default_conv_init(w30, b30)
    w40, b40 = new_conv(C, C, 3, True)
    # re-init loop over the non-RiR stages in construction order                                :110-118
    for w, b in ((w00, b00), (w20, b20), (w30, b30), (w40, b40)):
        torch.nn.init.kaiming_normal_(w)
        w.mul_(0.1)
        b.zero_()

    sd["Generators.0.0.0.weight"], sd["Generators.0.0.0.bias"] = w00, b00
    for r in range(n_rdb):
        for i in range(4):
            sd[f"Generators.0.1.RDBs.{r}.conv{i + 1}.0.weight"] = rdb[r][i]
        sd[f"Generators.0.1.RDBs.{r}.conv5.weight"] = rdb[r][4]
    sd["Generators.0.2.0.weight"], sd["Generators.0.2.0.bias"] = w20, b20
    sd["Generators.0.3.0.weight"], sd["Generators.0.3.0.bias"] = w30, b30
    sd["Generators.0.4.0.weight"], sd["Generators.0.4.0.bias"] = w40, b40
    return sd


D_CHANNELS = (256, 512, 1024, 1024)


def init_discriminator_state() -> "OrderedDict[str, torch.Tensor]":
    """State dict of Discriminator() under the current RNG state.  feature_patch_discriminator.py:18-49:
    three d2 Conv2d(+bias, norm=BN) then Conv2d 1024->1, all default-initialised at construction, then
    c2_msra_fill (kaiming_normal_ fan_out/relu, bias 0) on each in order."""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    convs = []
    for n in range(3):
        cin, cout = D_CHANNELS[n], D_CHANNELS[n + 1]
        w, b = torch.empty(cout, cin, 3, 3), torch.empty(cout)
        _default_conv_init(w, b)
        convs.append((w, b))
    w, b = torch.empty(1, D_CHANNELS[3], 3, 3), torch.empty(1)
    _default_conv_init(w, b)
    convs.append((w, b))
    for w, b in convs:
        torch.nn.init.kaiming_normal_(w, mode="fan_out", nonlinearity="relu")
        b.zero_()
    for n in range(3):
        cout = D_CHANNELS[n + 1]
        p = f"Discriminators.0.{n}.0."
        sd[p + "weight"], sd[p + "bias"] = convs[n]
        sd[p + "norm.weight"] = torch.ones(cout)
        sd[p + "norm.bias"] = torch.zeros(cout)
        sd[p + "norm.running_mean"] = torch.zeros(cout)
        sd[p + "norm.running_var"] = torch.ones(cout)
        sd[p + "norm.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    sd["Discriminators.0.3.0.weight"], sd["Discriminators.0.3.0.bias"] = convs[3]
    return sd


def init_states(seed: int = 0) -> Tuple[StateDict, StateDict]:
    """G then D under torch.manual_seed(seed): the construction order of stage1_trainer.py:505-506."""
    torch.manual_seed(seed)
    g = init_generator_state()
    d = init_discriminator_state()
    return g, d


def generator_param_keys(n_rdb: int = 3) -> List[str]:
    keys = ["Generators.0.0.0.weight", "Generators.0.0.0.bias"]
    for r in range(n_rdb):
        keys += [f"Generators.0.1.RDBs.{r}.conv{i}.0.weight" for i in range(1, 5)]
        keys.append(f"Generators.0.1.RDBs.{r}.conv5.weight")
    for s in (2, 3, 4):
        keys += [f"Generators.0.{s}.0.weight", f"Generators.0.{s}.0.bias"]
    return keys


def discriminator_param_keys() -> List[str]:
    keys = []
    for n in range(3):
        p = f"Discriminators.0.{n}.0."
        keys += [p + "weight", p + "bias", p + "norm.weight", p + "norm.bias"]
    keys += ["Discriminators.0.3.0.weight", "Discriminators.0.3.0.bias"]
    return keys


# ------------------------------------------------------------------------------------------------
# Forward math
# ------------------------------------------------------------------------------------------------
def _n_rdb(sd: StateDict) -> int:
    n = 0
    while f"Generators.0.1.RDBs.{n}.conv5.weight" in sd:
        n += 1
    return n


def generator_branch(sd: StateDict, f: torch.Tensor) -> torch.Tensor:
    """Generators[0](f): the learned branch without the bilinear skip.  generator_rdb.py:91-108, :27-30, :64-71."""
    L = lambda t: F.leaky_relu(t, LRELU)
    h0 = L(F.conv2d(f, sd["Generators.0.0.0.weight"], sd["Generators.0.0.0.bias"], padding=1))
    x = h0
    for r in range(_n_rdb(sd)):
        p = f"Generators.0.1.RDBs.{r}."
        feats = [x]
        for i in range(1, 5):
            feats.append(L(F.conv2d(torch.cat(feats, 1), sd[p + f"conv{i}.0.weight"], None, padding=1)))
        c5 = F.conv2d(torch.cat(feats, 1), sd[p + "conv5.weight"], None, padding=1)
        x = x + c5 * LRELU  # residual_scale == 0.2 (same constant as the slope, generator_rdb.py:75)
    h1 = x * 0.2 + h0
    h2 = L(F.conv2d(h1, sd["Generators.0.2.0.weight"], sd["Generators.0.2.0.bias"], padding=1))
    h3 = L(F.conv_transpose2d(h2, sd["Generators.0.3.0.weight"], sd["Generators.0.3.0.bias"], stride=2, padding=2))
    return F.conv2d(h3, sd["Generators.0.4.0.weight"], sd["Generators.0.4.0.bias"], padding=1)


def bilinear2x(f: torch.Tensor) -> torch.Tensor:
    """F.interpolate(f, scale_factor=2, mode='bilinear') restated by its closed form (align_corners=False):
    out[2i] = 1/4 x[i-1] + 3/4 x[i], out[2i+1] = 3/4 x[i] + 1/4 x[i+1], indices edge-clamped, separable.
    generator_rdb.py:125; identity verified in SURVEY.md App. G."""
    def up(t: torch.Tensor, dim: int) -> torch.Tensor:
        n = t.size(dim)
        idx = torch.arange(n)
        prev = t.index_select(dim, (idx - 1).clamp(min=0))
        nxt = t.index_select(dim, (idx + 1).clamp(max=n - 1))
        even = 0.25 * prev + 0.75 * t
        odd = 0.75 * t + 0.25 * nxt
        out = torch.stack((even, odd), dim=dim + 1)
        shape = list(t.shape)
        shape[dim] = 2 * n
        return out.reshape(shape)

    return up(up(f, 2), 3)


def generator_forward(sd: StateDict, f: torch.Tensor) -> torch.Tensor:
    """Generator.forward: Generators[0](f) + bilinear 2x skip.  generator_rdb.py:123-130."""
    return generator_branch(sd, f) + F.interpolate(f, scale_factor=2, mode="bilinear")


def discriminator_forward(sd: StateDict, x: torch.Tensor, training: bool = True,
                          update_running: bool = True) -> torch.Tensor:
    """Discriminators[0](x).  feature_patch_discriminator.py:32-41.  In training mode BatchNorm uses the
    statistics of THIS call and (update_running) updates sd's running buffers in place like nn.BatchNorm2d."""
    a = x
    for n in range(3):
        p = f"Discriminators.0.{n}.0."
        z = F.conv2d(a, sd[p + "weight"], sd[p + "bias"], padding=1)
        rm, rv = sd[p + "norm.running_mean"], sd[p + "norm.running_var"]
        if training and not update_running:
            rm, rv = rm.clone(), rv.clone()
        z = F.batch_norm(z, rm, rv, sd[p + "norm.weight"], sd[p + "norm.bias"], training, BN_MOMENTUM, BN_EPS)
        if training and update_running:
            sd[p + "norm.num_batches_tracked"] += 1
        a = F.leaky_relu(z, LRELU)
    return F.conv2d(a, sd["Discriminators.0.3.0.weight"], sd["Discriminators.0.3.0.bias"], padding=1)


def crop_to_min(a: torch.Tensor, b: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """_reshape_stage1 applied both ways (stage1_trainer.py:346-347, 437-443): top-left crop of both to the
    element-wise min H, W."""
    h = min(a.size(2), b.size(2))
    w = min(a.size(3), b.size(3))
    return a[:, :, :h, :w], b[:, :, :h, :w]


def bce_logits_mean(x: torch.Tensor, target: float) -> torch.Tensor:
    """nn.BCEWithLogitsLoss() with a constant target map: mean(max(x,0) - x t + log1p(exp(-|x|)))."""
    return (x.clamp(min=0) - x * target + torch.log1p(torch.exp(-x.abs()))).mean()


# ------------------------------------------------------------------------------------------------
# Stage-1 step (stage1_trainer.py:334-433)
# ------------------------------------------------------------------------------------------------
def _as_params(sd: StateDict, keys: Sequence[str], dtype) -> StateDict:
    out = dict(sd)
    for k in keys:
        out[k] = sd[k].detach().to(dtype).clone().requires_grad_(True)
    return out


def sgd_update(sd: StateDict, grads: StateDict, mom: StateDict, lr: float, momentum: float = 0.9,
               weight_decay: float = 1e-4, weight_decay_norm: float = 0.0) -> None:
    """torch.optim.SGD as detectron2's build_optimizer configures it [upstream]: d_p = g + wd*p;
    buf = momentum*buf + d_p (buf = d_p on first use); p -= lr*buf.  Norm (BN) parameters use
    WEIGHT_DECAY_NORM = 0.0; biases use WEIGHT_DECAY (BIAS_LR_FACTOR 1, WEIGHT_DECAY_BIAS = WEIGHT_DECAY)."""
    for k, g in grads.items():
        wd = weight_decay_norm if ".norm." in k else weight_decay
        d_p = g + wd * sd[k].detach()
        if k not in mom:
            mom[k] = d_p.clone()
        else:
            mom[k].mul_(momentum).add_(d_p)
        sd[k] = (sd[k].detach() - lr * mom[k]).to(sd[k].dtype)


def stage1_step(g_sd: StateDict, d_sd: StateDict, lr_feats: Sequence[torch.Tensor],
                hr_feats: Sequence[torch.Tensor], lr: Optional[float] = None,
                g_mom: Optional[StateDict] = None, d_mom: Optional[StateDict] = None,
                dtype=torch.float32, want_outputs: bool = False) -> Dict[str, object]:
    """One AFIGAN_Trainer.run_step on pre-extracted features (the guide model is out of scope).

    D phase (stage1_trainer.py:334-381): per level tr = G(lr).detach(); crop; D0(hr) BEFORE D0(tr);
      d_l = BCE(real,1)+BCE(fake,0); sum; backward into D; SGD step on D (if lr is given).
    G phase (:384-433): per level tr = G(lr); crop; D0(tr).detach() BEFORE D0(hr) (BN side effects only);
      g_l = 1e-3*BCE(fake,1) [no grad] + L1(tr,hr); sum; backward into G; SGD step on G.
    BN running buffers of d_sd are updated in place in the reference's call order.  With lr=None the
    optimiser steps are skipped (the D the G phase sees is then the un-stepped one).
    Returns losses (python floats), gradients (name -> tensor) and optionally tr/logit tensors."""
    gk, dk = generator_param_keys(_n_rdb(g_sd)), discriminator_param_keys()
    levels = range(2, 2 + len(lr_feats))
    out: Dict[str, object] = {}
    lr_feats = [t.to(dtype) for t in lr_feats]
    hr_feats = [t.to(dtype) for t in hr_feats]

    # ---- D phase
    gp = {k: v.to(dtype) if v.is_floating_point() else v for k, v in g_sd.items()}
    dp = _as_params({k: (v.to(dtype) if v.is_floating_point() else v) for k, v in d_sd.items()}, dk, dtype)
    d_losses, saved = [], {}
    for lv, lo, hi in zip(levels, lr_feats, hr_feats):
        with torch.no_grad():
            tr = generator_forward(gp, lo)
        tr, hi_c = crop_to_min(tr, hi)
        logit_real = discriminator_forward(dp, hi_c, True)
        logit_fake = discriminator_forward(dp, tr, True)
        d_l = bce_logits_mean(logit_real, 1.0) + bce_logits_mean(logit_fake, 0.0)
        d_losses.append(d_l)
        if want_outputs:
            saved[f"tr_p{lv}"] = tr.detach().clone()
            saved[f"logit_real_p{lv}"] = logit_real.detach().clone()
            saved[f"logit_fake_p{lv}"] = logit_fake.detach().clone()
    d_total = sum(d_losses)
    d_grads_t = torch.autograd.grad(d_total, [dp[k] for k in dk])
    d_grads = {k: g.detach() for k, g in zip(dk, d_grads_t)}
    for k, v in dp.items():  # carry BN buffers back
        if "running_" in k or "num_batches" in k:
            d_sd[k] = v.detach().to(d_sd[k].dtype) if v.is_floating_point() else v
    if lr is not None:
        sgd_update(d_sd, {k: d_grads[k].to(d_sd[k].dtype) for k in dk}, d_mom if d_mom is not None else {}, lr)

    # ---- G phase
    gp = _as_params(gp, gk, dtype)
    dp = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in d_sd.items()}
    g_losses, adv_l, l1_l = [], [], []
    for lv, lo, hi in zip(levels, lr_feats, hr_feats):
        tr = generator_forward(gp, lo)
        tr, hi_c = crop_to_min(tr, hi)
        with torch.no_grad():
            logit_fake = discriminator_forward(dp, tr.detach(), True)
            _ = discriminator_forward(dp, hi_c, True)
        adv = bce_logits_mean(logit_fake, 1.0)
        l1 = (tr - hi_c).abs().mean()
        g_losses.append(adv * 1e-3 + l1)
        adv_l.append(float(adv.detach()))
        l1_l.append(float(l1.detach()))
        if want_outputs:
            saved[f"g_logit_fake_p{lv}"] = logit_fake.detach().clone()
    g_total = sum(g_losses)
    g_grads_t = torch.autograd.grad(g_total, [gp[k] for k in gk])
    g_grads = {k: g.detach() for k, g in zip(gk, g_grads_t)}
    for k, v in dp.items():
        if "running_" in k or "num_batches" in k:
            d_sd[k] = v.detach().to(d_sd[k].dtype) if v.is_floating_point() else v
    if lr is not None:
        sgd_update(g_sd, {k: g_grads[k].to(g_sd[k].dtype) for k in gk}, g_mom if g_mom is not None else {}, lr)

    out["d_loss"] = {f"d_loss_p{lv}": float(v.detach()) for lv, v in zip(levels, d_losses)}
    out["g_loss"] = {f"g_loss_p{lv}": float(v.detach()) for lv, v in zip(levels, g_losses)}
    out["adv_loss"] = {f"adv_loss_p{lv}": v for lv, v in zip(levels, adv_l)}
    out["content_loss"] = {f"content_loss_p{lv}": v for lv, v in zip(levels, l1_l)}
    out["d_total"], out["g_total"] = float(d_total.detach()), float(g_total.detach())
    out["d_grads"], out["g_grads"] = d_grads, g_grads
    out["saved"] = saved
    return out


# ------------------------------------------------------------------------------------------------
# Merge sites (fpn_sr.py:147-158, pafpn_sr.py:172-181, bifpn_sr.py:542-548) and stage-2 'real' features
# ------------------------------------------------------------------------------------------------
def fpn_topdown_merge(g_sd: StateDict, prev: torch.Tensor, bottom_up: torch.Tensor, lateral_w: torch.Tensor,
                      lateral_b: Optional[torch.Tensor], fuse_type: str = "sum") -> torch.Tensor:
    """prev_features = lateral_conv(features) + srf_module(prev_features) [/2 if avg].  fpn_sr.py:151-157.
    (NORM == "" case: the lateral conv is a bare 1x1 conv with bias.)"""
    top_down = generator_forward(g_sd, prev)
    lateral = F.conv2d(bottom_up, lateral_w, lateral_b)
    out = lateral + top_down
    if fuse_type == "avg":
        out = out / 2
    return out


def bifpn_fusion(g_sd: StateDict, cur: torch.Tensor, top: torch.Tensor, w0: torch.Tensor, w1: torch.Tensor) -> torch.Tensor:
    """_feature_funsion for two inputs: swish(w0*cur + w1*AFI(top)) with the raw weights.  bifpn_sr.py:535-548."""
    s = w0 * cur + w1 * generator_forward(g_sd, top)
    return s * torch.sigmoid(s)


def nearest_half(x: torch.Tensor) -> torch.Tensor:
    """F.interpolate(x, scale_factor=0.5) (nearest): out[i,j] = x[2i,2j], size floor(H/2).  stage2_trainer.py:302."""
    h, w = x.size(2) // 2, x.size(3) // 2
    return x[:, :, : 2 * h : 2, : 2 * w : 2]


# ------------------------------------------------------------------------------------------------
# Shapes and synthetic data of the BASELINE configs
# ------------------------------------------------------------------------------------------------
C1_LR_SHAPES = ((104, 168), (52, 84), (26, 42), (13, 21), (7, 11))
C1_HR_SHAPES = ((200, 336), (100, 168), (50, 84), (25, 42), (13, 21))


def synthetic_features(batch: int, rank: int = 0, lr_shapes=C1_LR_SHAPES, hr_shapes=C1_HR_SHAPES,
                       channels: int = 256, seed: int = 1234) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
    """N(0,1) fp32 features, generator seeded with seed + rank; all HR levels drawn first, then all LR."""
    gen = torch.Generator().manual_seed(seed + rank)
    hr = [torch.randn(batch, channels, h, w, generator=gen) for h, w in hr_shapes]
    lr = [torch.randn(batch, channels, h, w, generator=gen) for h, w in lr_shapes]
    return lr, hr


G_FWD_FLOP_PER_INPUT_PX = 19_206_144
D_FWD_FLOP_PER_PX = 30_689_280


def stage1_step_flops(lr_px: int, hr_px: int) -> float:
    """Algorithmic FLOPs of one stage-1 step for lr_px generator-input pixels and hr_px discriminator
    pixels: 2 G fwd + 1 G bwd (no input grad) + 4 D fwd + 2 D bwd (no input grad)."""
    g, d = G_FWD_FLOP_PER_INPUT_PX, D_FWD_FLOP_PER_PX
    return 2 * g * lr_px + (2 * g - 1_179_648) * lr_px + 4 * d * hr_px + 2 * (2 * d - 2_359_296) * hr_px
