"""BiFPN neck (reference afigan/modeling/backbone/bifpn_sr.py:203-731) against the golden fixture made by running the
UNMODIFIED reference class on the CPU (tests/golden/make_golden_bifpn.py): same state-dict keys / shapes, and -- with weights derived from
the keys on both sides -- the same five output maps, with the 28 interpolator fusion sites and every 1x1 conv running in the library."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_bifpn import keyed_state, sample, tiny_bottom_up  # noqa: E402

FX = np.load(os.path.join(HERE, "golden", "bifpn_small.npz"))


def _build(precision=None):
    from afigan._compat import Backbone
    from afigan.modeling import BiFPN_AFIGAN
    from afigan.modeling.backbone import LastLevelP6P7
    torch.manual_seed(0)
    net = BiFPN_AFIGAN(tiny_bottom_up(Backbone), ["s2", "s3", "s4"], 256, 3, norm="BN", top_block=LastLevelP6P7(128, 256, "BN"))
    if precision is not None:
        net.srf_module.precision = precision
        for m in net.modules():
            if hasattr(m, "precision") and m is not net.srf_module:
                m.precision = precision
    return net


def test_bifpn_state_dict_matches_the_reference_key_for_key():
    net = _build()
    sd = net.state_dict()
    assert list(sd.keys()) == list(FX["keys"])
    assert [str(tuple(v.shape)) for v in sd.values()] == list(FX["shapes"])
    assert net.size_divisibility == 128 and net._out_features == ["p3", "p4", "p5", "p6", "p7"]
    from afigan._compat import BACKBONE_REGISTRY
    assert callable(BACKBONE_REGISTRY.get("build_swint_bifpn_sr_backbone"))


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["split", "fp32", "bf16"])
def test_bifpn_forward_vs_reference_golden(precision):
    net = _build(precision)
    net.load_state_dict(keyed_state(net.state_dict()), strict=True)
    net = net.cuda().eval()
    x = torch.randn(1, 3, 128, 128, generator=torch.Generator().manual_seed(77)).cuda()
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False          # the depthwise convs are cuDNN's: keep them fp32 like the CPU reference (TF32 costs 1.2e-4 here)
    try:
        with torch.no_grad():
            out = net(x)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    tol = 2e-5 if precision in ("split", "fp32") else 2e-2
    worst = 0.0
    for k, v in out.items():
        v = v.float().cpu()
        assert tuple(v.shape) == tuple(FX[f"shape/{k}"])
        if f"full/{k}" in FX:
            ref = torch.from_numpy(FX[f"full/{k}"])
            e = float((v - ref).norm() / ref.norm())
        else:
            ref = FX[f"sample/{k}"]
            e = float(np.linalg.norm(sample(v) - ref) / np.linalg.norm(ref))
            assert abs(float(v.norm()) - float(FX[f"norm/{k}"])) <= tol * float(FX[f"norm/{k}"])
        worst = max(worst, e)
        assert e <= tol, (k, e)
    print(f"[{precision}] BiFPN_AFIGAN forward vs the reference golden: worst rel err {worst:.2e}")


@pytest.mark.gpu
def test_bifpn_training_step_has_gradients_everywhere():
    """Autograd through the neck (stage 3 trains a detector on top of it): every parameter that takes part gets a finite gradient, and the
    interpolator's input gradient path (grad w.r.t. the bottom-up features) is live."""
    net = _build("split")
    net.load_state_dict(keyed_state(net.state_dict()), strict=True)
    net = net.cuda().train()
    x = torch.randn(2, 3, 128, 128, generator=torch.Generator().manual_seed(78)).cuda()
    out = net(x)
    sum(v.square().mean() for v in out.values()).backward()
    missing = [k for k, p in net.named_parameters() if p.grad is None or not torch.isfinite(p.grad).all()]
    assert not missing, missing[:8]
    assert float(net.bottom_up.c[0].weight.grad.abs().sum()) > 0 and float(net.srf_module.Generators[0][0][0].weight.grad.abs().sum()) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["split", "bf16"])
def test_conv1x1_autograd_vs_torch(precision):
    from afigan.functional import conv1x1_autograd
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 512, 13, 21, generator=g).cuda().requires_grad_(True)
    w = (torch.randn(256, 512, 1, 1, generator=g) * 0.05).cuda().requires_grad_(True)
    b = torch.randn(256, generator=g).cuda().requires_grad_(True)
    dy = torch.randn(2, 256, 13, 21, generator=g).cuda()
    conv1x1_autograd(x, w, b, precision).backward(dy)
    got = [conv1x1_autograd(x, w, b, precision).detach(), x.grad.clone(), w.grad.clone(), b.grad.clone()]
    x2, w2, b2 = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    y2 = F.conv2d(x2, w2, b2)
    y2.backward(dy.double())
    ref = [y2.detach(), x2.grad, w2.grad, b2.grad]
    tol = 3e-5 if precision == "split" else 1e-2
    for a, r in zip(got, ref):
        assert float((a.double() - r).norm() / r.norm()) < tol


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["split", "bf16"])
@pytest.mark.parametrize("shape", [(2, 256, 256, 25, 42), (1, 256, 256, 26, 43), (1, 64, 128, 13, 21), (2, 256, 256, 100, 168), (1, 256, 256, 2, 2)],
                         ids=lambda s: "x".join(map(str, s)))
def test_conv3x3_stride2_vs_torch(shape, precision):
    """The PANet neck's down-sampling conv as a 9-tap implicit GEMM over the four sub-pixel phase views (odd sizes: the phases differ by a
    row / column and the padding comes from each view's own TMA bounds): forward, input gradient, weight and bias gradients vs torch fp64."""
    from afigan.functional import conv3x3s2_autograd
    import torch.nn.functional as F
    n, cin, cout, h, w = shape
    g = torch.Generator().manual_seed(6 + h)
    x = torch.randn(n, cin, h, w, generator=g).cuda().requires_grad_(True)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) * 0.05).cuda().requires_grad_(True)
    b = torch.randn(cout, generator=g).cuda().requires_grad_(True)
    y = conv3x3s2_autograd(x, wt, b, precision)
    dy = torch.randn(y.shape, generator=g).cuda()
    y.backward(dy)
    x2, w2, b2 = (t.detach().double().requires_grad_(True) for t in (x, wt, b))
    y2 = F.conv2d(x2, w2, b2, stride=2, padding=1)
    assert y.shape == y2.shape
    y2.backward(dy.double())
    tol = 3e-5 if precision == "split" else 1e-2
    for name, a, r in (("y", y.detach(), y2.detach()), ("dx", x.grad, x2.grad), ("dw", wt.grad, w2.grad), ("db", b.grad, b2.grad)):
        e = float((a.double() - r).norm() / r.norm())
        assert e < tol, (name, e)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["split", "bf16"])
@pytest.mark.parametrize("shape", [(1, 256, 16, 16), (2, 256, 7, 11), (1, 256, 1, 1), (1, 64, 5, 40)], ids=lambda s: "x".join(map(str, s)))
def test_native_sepconv_and_fuse_down_vs_torch(shape, precision):
    """The inference-only building blocks of the BiFPN neck against their torch composition (the reference's arithmetic): swish -> depthwise
    3x3 ('static_same') -> pointwise 1x1 -> eval BatchNorm as ONE library call, and the bottom-up fusion site with the zero-padded max-pool."""
    import torch.nn.functional as F
    from afigan.functional import bifpn_fuse_down
    from afigan.modeling.bifpn_layers import MaxPool2d, SeparableConv2d
    n, c, h, w = shape
    torch.manual_seed(11)
    m = SeparableConv2d(c, c, 3, padding_mode="static_same", norm="BN", momentum=0.01, eps=1e-3, precision=precision).cuda().eval()
    with torch.no_grad():
        m.norm.running_mean.normal_(0, 0.2); m.norm.running_var.uniform_(0.5, 1.5); m.norm.weight.normal_(1, 0.2); m.norm.bias.normal_(0, 0.1)
    x = torch.randn(n, c, h, w, device="cuda")
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            got = m(x, pre_swish=True)                                   # native path (no autograd, eval norm)
            xs = (x * torch.sigmoid(x)).double()
            ref = F.conv2d(F.pad(xs, (1, 1, 1, 1)), m.depthwise.weight.double(), None, 1, 0, 1, c)
            ref = F.conv2d(ref, m.pointwise.weight.double(), m.pointwise.bias.double())
            ref = F.batch_norm(ref, m.norm.running_mean.double(), m.norm.running_var.double(), m.norm.weight.double(), m.norm.bias.double(), False, 0.0, m.norm.eps)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    e = float((got.double() - ref).norm() / ref.norm())
    assert e < (2e-5 if precision == "split" else 1e-2), e
    if h >= 1 and w >= 1:
        a, b = torch.randn(n, c, h, w, device="cuda"), torch.randn(n, c, h, w, device="cuda")
        for dh, dw in ((2 * h, 2 * w), (2 * h + 1, 2 * w + 1)):        # even and odd source sizes give the same h x w output
            dn = torch.randn(n, c, dh, dw, device="cuda") - 1.5          # mostly negative: the ZERO padding wins the max at the border
            pool = MaxPool2d(3, 2)(dn)
            assert pool.shape[2:] == (h, w)
            w3, w2 = torch.tensor([0.4, 0.7, 0.9], device="cuda"), torch.tensor([0.6, 0.8], device="cuda")
            assert torch.allclose(bifpn_fuse_down(a, b, dn, w3), w3[0] * a + w3[1] * b + w3[2] * pool, atol=1e-6)
            assert torch.allclose(bifpn_fuse_down(a, None, dn, w2), w2[0] * a + w2[1] * pool, atol=1e-6)
            assert torch.allclose(bifpn_fuse_down(a, b, dn, None), a + b + pool, atol=1e-6)


@pytest.mark.gpu
def test_deferred_weight_grads_equal_per_call_accumulation():
    """`Generator.deferred_weight_grads`: 8 interpolator calls in one graph (two BiFPN-like top-down sweeps), two backward passes in a row
    (so that the second one ADDS into existing .grad): same parameter gradients and same input gradients as autograd's per-call accumulation."""
    from afigan.modeling import Generator, bifpn_feature_fusion
    shapes = [(4, 6), (8, 12), (16, 24), (32, 48), (64, 96)]
    gen = torch.Generator().manual_seed(9)
    base = [torch.randn(1, 256, h, w, generator=gen).cuda() for h, w in shapes]
    res = {}
    for deferred in (False, True):
        torch.manual_seed(0)
        G = Generator(n_residual_dense_blocks=3, precision="split").cuda()
        G.deferred_weight_grads = deferred
        feats = [t.clone().requires_grad_(True) for t in base]
        wt = torch.tensor([0.7, 1.3], device="cuda", requires_grad=True)
        for rep in range(2):
            top = feats[0]
            for sweep in range(2):
                for f in feats[1:]:
                    s = bifpn_feature_fusion(G, f, top, wt)
                    top = s * torch.sigmoid(s)
                top = top[:, :, ::16, ::16]                      # back to the 4 x 6 level for the second sweep
            top.square().mean().backward()
        res[deferred] = ([p.grad.clone() for p in G.parameters()], [f.grad.clone() for f in feats], wt.grad.clone())
    for a, b in zip(res[True][0], res[False][0]):
        assert float((a - b).norm()) <= 1e-5 * float(b.norm()) + 1e-12
    for a, b in zip(res[True][1], res[False][1]):
        assert float((a - b).norm()) <= 1e-6 * float(b.norm())
    assert torch.allclose(res[True][2], res[False][2], rtol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("swish", [False, True])
@pytest.mark.parametrize("weighted", [False, True])
def test_fused_topdown_site_matches_the_torch_composition(swish, weighted):
    """`functional.FuseActFn` (w0 * cur + w1 * up, optional swish; reference bifpn_sr.py:542-548, 591-594) against the torch ops it replaces:
    output, both input gradients and the gradient of the 2-element fusion weight; `cur` as a channels_last crop, `up` contiguous."""
    from afigan.functional import bifpn_fuse_act
    gen = torch.Generator().manual_seed(21)
    cur0 = torch.randn(2, 256, 15, 23, generator=gen).cuda().contiguous(memory_format=torch.channels_last)[:, :, :13, :21]
    up0 = torch.randn(2, 256, 13, 21, generator=gen).cuda()
    w0 = torch.tensor([0.7, 1.3], device="cuda")
    dy = torch.randn(2, 256, 13, 21, generator=gen).cuda()
    res = []
    for fused in (True, False):
        cur, up = cur0.clone().requires_grad_(True), up0.clone().requires_grad_(True)
        w = w0.clone().requires_grad_(True) if weighted else None
        if fused:
            out = bifpn_fuse_act(cur, up, w, swish)
        else:
            s = (cur.double() * w[0].double() + up.double() * w[1].double()) if weighted else cur.double() + up.double()
            out = (s * torch.sigmoid(s) if swish else s)
        out.backward(dy.to(out.dtype))
        res.append((out.detach().double(), cur.grad.double(), up.grad.double(), None if w is None else w.grad.double()))
    for a, b in zip(res[0][:3], res[1][:3]):
        assert float((a - b).norm()) <= 2e-6 * float(b.norm())
    if weighted:
        assert float((res[0][3] - res[1][3]).norm()) <= 1e-4 * float(res[1][3].norm())


@pytest.mark.gpu
def test_bifpn_class_training_path_uses_one_pass_per_fusion_site():
    """BiFPN_AFIGAN with autograd: the top-down sites run through FuseActFn (fusion + swish fused), the result equals the composition
    `conv(swish(w0 * cur + w1 * AFI(top)))` evaluated with torch ops around the same interpolator call."""
    from afigan.modeling import Generator, bifpn_feature_fusion
    torch.manual_seed(0)
    G = Generator(n_residual_dense_blocks=3, precision="split").cuda()
    gen = torch.Generator().manual_seed(22)
    cur = torch.randn(1, 256, 16, 24, generator=gen).cuda().requires_grad_(True)
    top = torch.randn(1, 256, 8, 12, generator=gen).cuda().requires_grad_(True)
    w = torch.tensor([0.9, 1.1], device="cuda", requires_grad=True)
    a = bifpn_feature_fusion(G, cur, top, w, swish=True)
    a.square().mean().backward()
    ga = [cur.grad.clone(), top.grad.clone(), w.grad.clone(), G.Generators[0][4][0].weight.grad.clone()]
    for t in (cur, top, w, *G.parameters()):
        t.grad = None
    up = G(top, out_hw=(16, 24))
    s = cur * w[0] + up * w[1]
    b = s * torch.sigmoid(s)
    b.square().mean().backward()
    gb = [cur.grad, top.grad, w.grad, G.Generators[0][4][0].weight.grad]
    assert float((a - b).detach().norm()) <= 2e-6 * float(b.detach().norm())
    for x, y in zip(ga, gb):
        assert float((x - y).norm()) <= 2e-5 * float(y.norm())


@pytest.mark.gpu
def test_deferred_weight_grads_survive_a_backward_pass_that_raised():
    """The packed accumulator of `deferred_weight_grads` is keyed by the autograd graph task: a backward pass that raises after some
    interpolator calls (its end-of-pass callback never runs) must not leak into the next pass."""
    from afigan.modeling import Generator
    gen = torch.Generator().manual_seed(23)
    x = torch.randn(1, 256, 8, 12, generator=gen).cuda()

    def run(G, poison):
        xin = x.clone().requires_grad_(True)
        y = G(G(xin)[:, :, ::2, ::2])                      # two interpolator calls in one graph
        if poison:
            xin.register_hook(lambda g: (_ for _ in ()).throw(RuntimeError("boom")))
        (y.square().mean()).backward()

    torch.manual_seed(0)
    G = Generator(n_residual_dense_blocks=3, precision="split").cuda()
    G.deferred_weight_grads = True
    with pytest.raises(RuntimeError, match="boom"):
        run(G, True)
    for p in G.parameters():
        p.grad = None
    run(G, False)
    got = [p.grad.clone() for p in G.parameters()]
    torch.manual_seed(0)
    R = Generator(n_residual_dense_blocks=3, precision="split").cuda()
    run(R, False)
    for a, b in zip(got, [p.grad for p in R.parameters()]):
        assert float((a - b).norm()) <= 1e-5 * float(b.norm()) + 1e-12
