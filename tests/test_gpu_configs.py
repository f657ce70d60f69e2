"""Parity / property cases on the shapes of the other BASELINE.json configs:
  C3  stage-2 BiFPN shapes: LR pyramid 64x96 .. 4x6 (AFI on p7..p4), D on 56x88 .. 3x5
  C4  stage-3 PAFPN top-down merge, batch 16 per GPU, AFI on 25x42, 50x84, 100x168 with lateral adds, fwd + full bwd
  C5  inference sweep: BiFPN AFI calls at short side 400 (512x768 padded) and 1200 (1280x2048 padded)
Small cases are checked against the CPU oracle; at sizes the oracle cannot finish in seconds the fp32 (parity) mode of the library --
itself pinned to the oracle on the small cases -- is the yardstick for the bf16 (tensor-core) mode."""
import pytest
import torch
import torch.nn.functional as F

from gpu_util import cosine, rel
from oracle import afigan_oracle as O

pytestmark = pytest.mark.gpu


def _gen(precision):
    from afigan.modeling import Generator
    torch.manual_seed(0)
    return Generator(n_residual_dense_blocks=3, precision=precision).cuda()


@pytest.mark.parametrize("precision", ["fp32", "split", "bf16"])
def test_c3_bifpn_pyramid_inference_vs_oracle(precision):
    """C3/C5 (short side 400): one BiFPN top-down sweep p7 -> p3 with the shared interpolator, eval mode, vs the oracle."""
    from afigan.modeling import bifpn_feature_fusion
    G = _gen(precision).eval()
    g_sd, _ = O.init_states(0)
    gen = torch.Generator().manual_seed(31)
    shapes = [(4, 6), (8, 12), (16, 24), (32, 48), (64, 96)]             # p7 .. p3 of a 512x768 input
    feats = [torch.randn(1, 256, h, w, generator=gen) for h, w in shapes]
    w1 = torch.tensor([1.1, 0.9])
    with torch.no_grad():
        top_ref, top = feats[0], feats[0].cuda()
        for f in feats[1:]:
            s = w1[0] * f + w1[1] * O.generator_forward(g_sd, top_ref)
            top_ref = s * torch.sigmoid(s)                                 # swish after the fusion (bifpn_sr.py:591-594)
            s2 = bifpn_feature_fusion(G, f.cuda(), top, w1.cuda())
            top = s2 * torch.sigmoid(s2)
    assert top.shape == (1, 256, 64, 96)
    assert rel(top, top_ref) < (1e-5 if precision in ("fp32", "split") else 4e-3), rel(top, top_ref)


def test_c5_largest_inference_shape_bf16_vs_fp32_mode():
    """C5 (short side 1200 -> 1280x2048 padded): AFI on the p4 map 80x128 -> 160x256, batch 1; bf16 mode against the library's fp32 mode,
    plus the exact-2x and zero-weight identities at this size."""
    gen = torch.Generator().manual_seed(32)
    x = torch.randn(1, 256, 80, 128, generator=gen).cuda()
    with torch.no_grad():
        y32 = _gen("fp32").eval()(x)
        y16 = _gen("bf16").eval()(x)
    assert y32.shape == (1, 256, 160, 256)
    assert rel(y16, y32) < 1e-3
    skip = F.interpolate(x, scale_factor=2, mode="bilinear")
    assert rel(y16 - skip, y32 - skip) < 3e-2                              # the learned branch alone (bf16 operand rounding)


def test_c4_pafpn_topdown_batch16_bf16_vs_fp32_mode():
    """C4: three AFI calls of the PAFPN top-down path (25x42 -> 50x84 -> 100x168 -> 200x336 is out of the neck; here the three merges up to
    100x168) at batch 16 with 1024/512/256-channel laterals, forward + full backward (input, lateral and interpolator gradients)."""
    N = 16
    gen = torch.Generator().manual_seed(33)
    c5 = torch.randn(N, 256, 13, 21, generator=gen).cuda()                 # top lateral output (prev_features)
    lat_in = [torch.randn(N, c, h, w, generator=gen).cuda() for c, h, w in ((1024, 25, 42), (512, 50, 84), (256, 100, 168))]
    lat_w = [(torch.randn(256, c, 1, 1, generator=gen) * (1.0 / c) ** 0.5).cuda() for c in (1024, 512, 256)]
    lat_b = [torch.zeros(256).cuda() for _ in range(3)]
    outs, grads = {}, {}
    for precision in ("fp32", "bf16"):
        G = _gen(precision)
        prev = c5.clone().requires_grad_(True)
        ws = [w.clone().requires_grad_(True) for w in lat_w]
        xs = [t.clone().requires_grad_(True) for t in lat_in]
        cur = prev
        for x_, w_, b_ in zip(xs, ws, lat_b):
            cur = G.merge(cur, x_, w_, b_, "sum")
        assert cur.shape == (N, 256, 100, 168)
        cur.square().mean().backward()
        outs[precision] = cur.detach()
        grads[precision] = [prev.grad, ws[0].grad, xs[2].grad, G.Generators[0][4][0].weight.grad, G.Generators[0][0][0].weight.grad]
        del G
        torch.cuda.empty_cache()
    assert rel(outs["bf16"], outs["fp32"]) < 8e-3
    for a, b in zip(grads["bf16"], grads["fp32"]):
        assert cosine(a, b) > 0.97 and rel(a, b) < 0.2, (rel(a, b), cosine(a, b))


@pytest.mark.parametrize("precision", ["fp32", "split", "bf16"])
def test_c3_stage2_discriminator_shapes(precision):
    """C3: the stage-2 D-phase on the cropped BiFPN-pyramid sizes 56x88 .. 3x5 (ten grouped calls, per-call BatchNorm statistics)."""
    from afigan.engine import stage2_discriminator_losses
    from afigan.modeling import Discriminator, Generator
    torch.manual_seed(0)
    Generator(n_residual_dense_blocks=3)
    D = Discriminator(precision=precision).cuda()
    D.Discriminators[0].train()
    _, d_sd = O.init_states(0)
    gen = torch.Generator().manual_seed(34)
    sizes = [(56, 88), (28, 44), (14, 22), (7, 11), (3, 5)]
    guide = [torch.randn(1, 256, 2 * h + 1, 2 * w, generator=gen) for h, w in sizes]       # nearest-half then crop to (h, w)
    model = [torch.randn(1, 256, h, w, generator=gen) for h, w in sizes]
    ref = []
    for hr, up in zip(guide, model):
        real, fake = O.crop_to_min(O.nearest_half(hr), up)
        ref.append(float(O.bce_logits_mean(O.discriminator_forward(d_sd, real, True), 1.0) + O.bce_logits_mean(O.discriminator_forward(d_sd, fake, True), 0.0)))
    out = stage2_discriminator_losses(D, [g.cuda() for g in guide], [m.cuda() for m in model])
    got = [float(v.detach()) for v in out.values()]
    tol = 1e-5 if precision in ("fp32", "split") else 6e-3
    for a, b in zip(got, ref):
        assert abs(a - b) <= tol * abs(b) + 1e-5, (got, ref)


@pytest.mark.parametrize("precision", ["split", "bf16"])
def test_stage2_grouped_step_matches_the_per_level_autograd_path(precision):
    """`Stage2Step` (grouped launches over the five levels, D gradients in one flat buffer, SGD in the library) against the per-level
    autograd composition of the same loss block (`stage2_discriminator_losses` / `stage2_generator_losses`, themselves checked against
    the oracle above): D losses, D gradients, the parameters after one SGD step, the G-side losses and their gradient w.r.t. the model's
    features, and the BatchNorm buffers (20 calls) -- reference stage2_trainer.py:298-384."""
    from afigan.engine import Stage2Step, stage2_discriminator_losses, stage2_generator_losses
    from afigan.modeling import Discriminator, Generator
    gen = torch.Generator().manual_seed(41)
    sizes = [(28, 44), (14, 22), (7, 11), (3, 5)]
    guide = [torch.randn(2, 256, 2 * h + 1, 2 * w, generator=gen).cuda() for h, w in sizes]
    model = [torch.randn(2, 256, h + (l == 0), w, generator=gen).cuda() for l, (h, w) in enumerate(sizes)]     # level 0 over-sized: cropped
    outs = {}
    for kind in ("grouped", "autograd"):
        torch.manual_seed(0)
        Generator(n_residual_dense_blocks=3)
        D = Discriminator(precision=precision).cuda()
        D.Discriminators[0].train()
        feats = [m.clone().requires_grad_(True) for m in model]
        if kind == "grouped":
            step = Stage2Step(D, lr=1e-2, precision=precision)
            d_row = step.d_phase(guide, feats)
            d_loss = [float(v) for v in d_row[:4].cpu()]
            d_grads = [p.grad.clone() for p in step.d_params]
            g = step.g_losses(guide, feats)
        else:
            params = D.Discriminators[0]._params()
            opt = torch.optim.SGD([{"params": [p for i, p in enumerate(params) if not (i % 4 >= 2 and i < 12)], "weight_decay": 1e-4},
                                   {"params": [p for i, p in enumerate(params) if i % 4 >= 2 and i < 12], "weight_decay": 0.0}], lr=1e-2, momentum=0.9)
            dl = stage2_discriminator_losses(D, guide, feats)
            d_loss = [float(v.detach()) for v in dl.values()]
            opt.zero_grad()
            sum(dl.values()).backward()
            d_grads = [p.grad.clone() for p in params]
            opt.step()
            g = stage2_generator_losses(D, guide, feats)
        sum(g.values()).backward()
        outs[kind] = dict(d_loss=d_loss, d_grads=d_grads, g=[float(v.detach()) for v in g.values()], fgrad=[f.grad.clone() for f in feats],
                          params=[p.detach().clone() for p in D.Discriminators[0]._params()], sd={k: v.clone() for k, v in D.state_dict().items() if "running" in k or "num_batches" in k})
    a, b = outs["grouped"], outs["autograd"]
    tol = 1e-5 if precision == "split" else 2e-2
    assert list(outs["grouped"]["g"]) and len(a["g"]) == 4
    for x, y in zip(a["d_loss"], b["d_loss"]):
        assert abs(x - y) <= tol * abs(y)
    for x, y in zip(a["g"], b["g"]):
        assert abs(x - y) <= tol * abs(y) + 1e-6
    for i, (x, y) in enumerate(zip(a["d_grads"], b["d_grads"])):
        if float(y.norm()) < 1e-6:
            continue
        assert rel(x, y) < (2e-3 if precision == "split" else 0.1), i
    for x, y in zip(a["params"], b["params"]):
        assert rel(x, y) < (1e-5 if precision == "split" else 2e-3)
    for x, y in zip(a["fgrad"], b["fgrad"]):
        assert rel(x, y) < 1e-6
    for k in a["sd"]:
        assert rel(a["sd"][k].float(), b["sd"][k].float()) < (1e-5 if precision == "split" else 2e-2), k
    assert int(a["sd"]["Discriminators.0.0.0.norm.num_batches_tracked"]) == 16
