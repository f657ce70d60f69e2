"""Merge sites and autograd surface: input gradient of the interpolator, the fused FPN / PAFPN top-down merge (fpn_sr.py:147-158,
pafpn_sr.py:172-181), the BiFPN fusion (bifpn_sr.py:535-548) and the stage-2 loss block (stage2_trainer.py:298-384) against the oracle."""
import pytest
import torch
import torch.nn.functional as F

from gpu_util import TOL, cosine, rel
from oracle import afigan_oracle as O

pytestmark = pytest.mark.gpu


def _gen(precision):
    from afigan.modeling import Generator
    torch.manual_seed(0)
    G = Generator(n_residual_dense_blocks=3, precision=precision).cuda()
    g_sd, _ = O.init_states(0)
    return G, g_sd


@pytest.mark.parametrize("precision", ["fp32", "split", "bf16"])
def test_input_gradient(precision):
    G, g_sd = _gen(precision)
    gen = torch.Generator().manual_seed(21)
    x = torch.randn(2, 256, 6, 9, generator=gen)
    dy = torch.randn(2, 256, 11, 17, generator=gen)          # loss only sees a top-left crop
    xr = x.clone().requires_grad_(True)
    (O.generator_forward(g_sd, xr)[:, :, :11, :17] * dy).sum().backward()
    xg = x.cuda().requires_grad_(True)
    (G(xg, out_hw=(11, 17)) * dy.cuda()).sum().backward()
    r = rel(xg.grad, xr.grad)
    # the bilinear-skip adjoint dominates dx at init, so even bf16 operands agree closely
    assert r < (1e-5 if precision in ("fp32", "split") else 2e-3), r


@pytest.mark.parametrize("precision", ["fp32", "split", "bf16"])
@pytest.mark.parametrize("fuse_type,lat_c", [("sum", 256), ("avg", 512)])
def test_fpn_topdown_merge_forward_backward(precision, fuse_type, lat_c):
    G, g_sd = _gen(precision)
    tol = TOL[precision]
    gen = torch.Generator().manual_seed(22)
    prev = torch.randn(2, 256, 7, 11, generator=gen)
    feat = torch.randn(2, lat_c, 13, 21, generator=gen)       # odd lateral size: the 14x22 interpolated map is cropped to it
    lw = (torch.randn(256, lat_c, 1, 1, generator=gen) * 0.05)
    lb = torch.randn(256, generator=gen) * 0.1
    keys = O.generator_param_keys()
    params = dict(g_sd)
    for k in keys:
        params[k] = g_sd[k].clone().requires_grad_(True)
    pr, fr, lwr, lbr = (t.clone().requires_grad_(True) for t in (prev, feat, lw, lb))
    td = O.generator_forward(params, pr)[:, :, :13, :21]
    ref = F.conv2d(fr, lwr, lbr) + td
    if fuse_type == "avg":
        ref = ref / 2
    pc, fc, lwc, lbc = (t.cuda().requires_grad_(True) for t in (prev, feat, lw, lb))
    out = G.merge(pc, fc, lwc, lbc, fuse_type)
    assert out.shape == ref.shape
    assert rel(out, ref) < (1e-5 if precision in ("fp32", "split") else 8e-3), rel(out, ref)
    dy = torch.randn(ref.shape, generator=gen)
    (ref * dy).sum().backward()
    (out * dy.cuda()).sum().backward()
    gt = 1e-4 if precision in ("fp32", "split") else 1.5e-2
    assert rel(pc.grad, pr.grad) < gt, rel(pc.grad, pr.grad)
    assert rel(fc.grad, fr.grad) < gt, rel(fc.grad, fr.grad)
    assert rel(lwc.grad, lwr.grad) < gt, rel(lwc.grad, lwr.grad)
    assert rel(lbc.grad, lbr.grad) < gt
    for k, p in zip(keys, G._params()):
        r, c = rel(p.grad, params[k].grad), cosine(p.grad, params[k].grad)
        assert r < tol["grad"] and c > tol["cos"], f"{k}: {r:.3e}"


class _ToyBottomUp(torch.nn.Module):
    """Stand-in for the ResNet bottom-up (out of scope): strided 1x1 convs producing res2..res5 with R-50 channel counts."""

    def __init__(self):
        super().__init__()
        from afigan._compat import ShapeSpec
        self.convs = torch.nn.ModuleList([torch.nn.Conv2d(3, c, 1) for c in (256, 512, 1024, 2048)])
        self._shapes = {f"res{i + 2}": ShapeSpec(channels=c, stride=2 ** (i + 2)) for i, c in enumerate((256, 512, 1024, 2048))}

    def output_shape(self):
        return self._shapes

    def forward(self, x):
        return {f"res{i + 2}": conv(F.avg_pool2d(x, 2 ** (i + 2))) for i, conv in enumerate(self.convs)}


@pytest.mark.parametrize("neck", ["fpn", "pafpn"])
def test_neck_matches_torch_composition(neck):
    from afigan.config import get_cfg
    from afigan.modeling import FPN_AFIGAN, PAFPN_AFIGAN
    from afigan.modeling.backbone import LastLevelMaxPool
    torch.manual_seed(1)
    torch.backends.cudnn.allow_tf32 = False        # the neck's torch-side convs (out of the hot path) would otherwise run in TF32
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = get_cfg()
    cls = FPN_AFIGAN if neck == "fpn" else PAFPN_AFIGAN
    m = cls(_ToyBottomUp(), ["res2", "res3", "res4", "res5"], 256, norm="", top_block=LastLevelMaxPool(), fuse_type="sum", cfg=cfg).cuda()
    m.srf_module.precision = "fp32"
    names = [n for n, _ in m.named_parameters()]
    assert "fpn_lateral2.weight" in names and "srf_module.Generators.0.0.0.weight" in names
    assert (f"{'fpn' if neck == 'fpn' else 'pafpn'}_output5.bias") in names
    x = torch.randn(1, 3, 96, 128).cuda()
    out = m(x)
    assert list(out) == ["p2", "p3", "p4", "p5", "p6"] and m.size_divisibility == 32
    assert out["p2"].shape == (1, 256, 24, 32) and out["p6"].shape == (1, 256, 2, 2)
    # reference composition with the oracle interpolator
    g_sd = {k: v.detach().cpu() for k, v in m.srf_module.state_dict().items()}
    bu = {k: v.detach().cpu() for k, v in m.bottom_up(x).items()}
    cpu = lambda mod: (mod.weight.detach().cpu(), mod.bias.detach().cpu())
    feats = [bu[f] for f in ["res5", "res4", "res3", "res2"]]
    prev = F.conv2d(feats[0], *cpu(m.fpn_lateral5))
    tds = [prev]
    for f, st in zip(feats[1:], (4, 3, 2)):
        prev = F.conv2d(f, *cpu(getattr(m, f"fpn_lateral{st}"))) + O.generator_forward(g_sd, prev)
        tds.insert(0, prev)
    if neck == "fpn":
        ref_p2 = F.conv2d(tds[0], *cpu(m.fpn_output2), padding=1)
    else:
        ref_p2 = F.conv2d(tds[0], *cpu(m.pafpn_output2), padding=1)
    assert rel(out["p2"], ref_p2) < 1e-4, rel(out["p2"], ref_p2)
    out["p2"].square().mean().backward()
    assert m.fpn_lateral5.weight.grad is not None and m.srf_module.Generators[0][0][0].weight.grad is not None
    assert m.bottom_up.convs[3].weight.grad is not None        # gradient flows through the interpolator's input


def test_afi_freeze_flag():
    from afigan.config import get_cfg
    from afigan.modeling import FPN_AFIGAN
    cfg = get_cfg()
    cfg.MODEL.AFI_FREEZE = True
    m = FPN_AFIGAN(_ToyBottomUp(), ["res2", "res3", "res4", "res5"], 256, cfg=cfg)
    assert all(not p.requires_grad for p in m.srf_module.parameters())       # fpn_sr.py:67-69
    assert m.fpn_lateral3.weight.requires_grad


def test_bifpn_fusion():
    from afigan.modeling import bifpn_feature_fusion
    G, g_sd = _gen("fp32")
    gen = torch.Generator().manual_seed(23)
    cur, top = torch.randn(1, 256, 8, 12, generator=gen), torch.randn(1, 256, 4, 6, generator=gen)
    w = torch.tensor([0.7, 1.3])
    ref = w[0] * cur + w[1] * O.generator_forward(g_sd, top)
    out = bifpn_feature_fusion(G, cur.cuda(), top.cuda(), w.cuda())
    assert rel(out, ref) < 1e-5
    assert rel(bifpn_feature_fusion(G, cur.cuda(), top.cuda()), cur + O.generator_forward(g_sd, top)) < 1e-5
    # the forward-only site (no autograd: one library call, fusion folded into the output pass) against the same oracle,
    # incl. a cropped `cur` (odd pyramid sizes) and a strided (channels_last) one
    with torch.no_grad():
        fused = bifpn_feature_fusion(G, cur.cuda(), top.cuda(), w.cuda())
        assert rel(fused, ref) < 1e-5
        assert rel(bifpn_feature_fusion(G, cur.cuda(), top.cuda()), cur + O.generator_forward(g_sd, top)) < 1e-5
        cur_odd = torch.randn(1, 256, 7, 11, generator=gen)
        ref_odd = w[0] * cur_odd + w[1] * O.generator_forward(g_sd, top)[:, :, :7, :11]
        assert rel(bifpn_feature_fusion(G, cur_odd.cuda().to(memory_format=torch.channels_last), top.cuda(), w.cuda()), ref_odd) < 1e-5


@pytest.mark.parametrize("precision", ["fp32", "split", "bf16"])
def test_stage2_loss_block(precision):
    from afigan.engine import stage2_discriminator_losses, stage2_generator_losses
    from afigan.modeling import Discriminator
    tol = TOL[precision]
    torch.manual_seed(0)
    from afigan.modeling import Generator
    Generator(n_residual_dense_blocks=3)                      # consume the RNG like the reference does before building D
    D = Discriminator(precision=precision).cuda()
    D.Discriminators[0].train()
    _, d_sd = O.init_states(0)
    gen = torch.Generator().manual_seed(24)
    guide = [torch.randn(2, 256, 26, 42, generator=gen), torch.randn(2, 256, 13, 21, generator=gen)]      # HR-image pyramid
    model = [torch.randn(2, 256, 13, 21, generator=gen), torch.randn(2, 256, 7, 11, generator=gen)]       # AFI-FPN outputs on the 0.5x image
    # ---- oracle
    keys = O.discriminator_param_keys()
    dp = dict(d_sd)
    for k in keys:
        dp[k] = d_sd[k].clone().requires_grad_(True)
    d_ref = []
    for hr, up in zip(guide, model):
        real, fake = O.crop_to_min(O.nearest_half(hr), up)
        d_ref.append(O.bce_logits_mean(O.discriminator_forward(dp, real, True), 1.0) + O.bce_logits_mean(O.discriminator_forward(dp, fake, True), 0.0))
    sum(d_ref).backward()
    ups = [u.clone().requires_grad_(True) for u in model]
    g_ref = []
    for hr, up in zip(guide, ups):
        real, fake = O.crop_to_min(O.nearest_half(hr), up)
        with torch.no_grad():
            lf = O.discriminator_forward(dp, fake, True)
            O.discriminator_forward(dp, real, True)
        g_ref.append(O.bce_logits_mean(lf, 1.0) * 1e-3 + (fake - real).abs().mean())
    sum(g_ref).backward()
    # ---- library
    d_loss = stage2_discriminator_losses(D, [g.cuda() for g in guide], [m.cuda() for m in model])
    assert list(d_loss) == ["d_loss_p2", "d_loss_p3"]
    for a, b in zip(d_loss.values(), d_ref):
        assert abs(float(a.detach()) - float(b.detach())) < tol["loss"] * abs(float(b.detach())) + 1e-6
    sum(d_loss.values()).backward()
    for k, p in zip(keys, D.Discriminators[0]._params()):
        if k.endswith("0.bias") and ".3." not in k:
            continue
        assert rel(p.grad, dp[k].grad) < tol["dgrad"], k
    ups_c = [u.cuda().requires_grad_(True) for u in model]
    g_loss = stage2_generator_losses(D, [g.cuda() for g in guide], ups_c)
    for a, b in zip(g_loss.values(), g_ref):
        assert abs(float(a.detach()) - float(b.detach())) < 1e-4 + tol["loss"] * 1e-2
    sum(g_loss.values()).backward()
    for a, b in zip(ups_c, ups):
        assert rel(a.grad, b.grad) < 1e-5          # sign(fake-real)/numel inside the crop, zero outside
    assert int(D.state_dict()["Discriminators.0.0.0.norm.num_batches_tracked"]) == 8   # (2 + 2) D calls x 2 levels


def test_inference_graphs_match_the_eager_path():
    """Forward-only calls replay CUDA graphs with static buffers (functional.InferenceGraphs): same numbers as the eager autograd path, across
    repeated calls with fresh inputs on the same shape, for the plain forward, the fused FPN merge and the BiFPN fusion site, and after an
    in-place weight update (the packed weights are refreshed in place, the graph keeps pointing at them)."""
    G, _ = _gen("fp32")
    gen = torch.Generator().manual_seed(31)
    lat_w = (torch.randn(256, 512, 1, 1, generator=gen) * 0.05).cuda()
    lat_b = torch.randn(256, generator=gen).cuda()
    w = torch.tensor([0.6, 1.4]).cuda()
    for rep in range(3):
        x = torch.randn(1, 256, 7, 11, generator=gen).cuda()
        c = torch.randn(1, 512, 13, 21, generator=gen).cuda()
        cur = torch.randn(1, 256, 13, 21, generator=gen).cuda()
        eager = (G(x), G(x, out_hw=(13, 21)), G.merge(x, c, lat_w, lat_b, "avg"))           # autograd enabled: eager launches
        with torch.no_grad():
            graphed = (G(x), G(x, out_hw=(13, 21)), G.merge(x, c, lat_w, lat_b, "avg"))
            fused = G.fuse(x, cur, w)
        for a, b in zip(graphed, eager):
            assert rel(a, b.detach()) < 1e-6
        assert rel(fused, (w[0] * cur + w[1] * eager[1]).detach()) < 1e-5
        if rep == 1:                                # in-place update of a weight: later calls must see it
            with torch.no_grad():
                G.Generators[0][4][0].weight.mul_(1.5)
    assert len(G._native.graphs.entries) == 4       # one graph per (shape, operands) -- not one per call
