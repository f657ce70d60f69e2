"""The boundary accepts fp32 AND bf16 input views with any strides (ABI 4, afi_view4.dtype): a bf16 tensor (autocast activations, config 4 of
BASELINE.json) enters the kernels without an up-cast copy, a channels_last tensor without a transpose.  bf16 -> fp32 is exact, so every result
must be BIT-identical to the one obtained from the fp32 copy of the same values."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _gen(precision):
    from afigan.modeling import Generator
    torch.manual_seed(0)
    return Generator(n_residual_dense_blocks=3, precision=precision).cuda()


@pytest.mark.parametrize("precision", ["split", "bf16"])
@pytest.mark.parametrize("layout", ["contiguous", "channels_last"])
def test_generator_bf16_and_channels_last_inputs_are_bit_identical(precision, layout):
    gen = torch.Generator().manual_seed(71)
    x16 = torch.randn(2, 256, 13, 21, generator=gen).cuda().bfloat16()
    if layout == "channels_last":
        x16 = x16.contiguous(memory_format=torch.channels_last)
    res = []
    for x in (x16.float().contiguous(), x16, x16.float()):          # fp32 contiguous (yardstick), bf16 in `layout`, fp32 in `layout`
        G = _gen(precision)
        xin = x.clone(memory_format=torch.preserve_format).requires_grad_(True)
        y = G(xin)
        y.square().mean().backward()
        res.append((y.detach(), xin.grad, [p.grad.clone() for p in G.parameters()]))
    y0, dx0, g0 = res[0]
    for (y, dx, g), name in zip(res[1:], ("bf16", "fp32-layout")):
        assert y.dtype == torch.float32 and torch.equal(y, y0), name
        assert dx.dtype == (torch.bfloat16 if name == "bf16" else torch.float32)
        assert torch.equal(dx.float(), dx0.to(dx.dtype).float()), name       # autograd casts the fp32 input gradient to the input's dtype
        # weight gradients: identical operands, but the split-K reductions use fp32 atomics (run-to-run rounding noise)
        for a, b in zip(g, g0):
            assert float((a - b).norm()) <= 1e-5 * float(b.norm()) + 1e-12, name


@pytest.mark.parametrize("precision", ["split", "bf16"])
def test_merge_and_fuse_with_bf16_operands(precision):
    """PAFPN / FPN merge with a bf16 bottom-up map (1x1 lateral input) and the BiFPN fusion site with a bf16 current-level map, training
    (autograd) and inference (CUDA-graph) paths."""
    from afigan.modeling import bifpn_feature_fusion
    gen = torch.Generator().manual_seed(72)
    prev = torch.randn(2, 256, 13, 21, generator=gen).cuda().bfloat16()
    bottom = torch.randn(2, 512, 25, 42, generator=gen).cuda().bfloat16()      # odd size: the interpolated map is cropped 26 -> 25
    cur = torch.randn(2, 256, 26, 42, generator=gen).cuda().bfloat16()
    lat_w = (torch.randn(256, 512, 1, 1, generator=gen) * 0.04).cuda()
    lat_b = torch.randn(256, generator=gen).cuda() * 0.1
    wt = torch.tensor([0.7, 1.3], device="cuda")
    out = {}
    for kind in ("fp32", "bf16"):
        cast = (lambda t: t.float()) if kind == "fp32" else (lambda t: t.clone())
        G = _gen(precision)
        p, b = cast(prev).requires_grad_(True), cast(bottom).requires_grad_(True)
        w_ = lat_w.clone().requires_grad_(True)
        m = G.merge(p, b, w_, lat_b, "avg")
        m.square().mean().backward()
        c = cast(cur).requires_grad_(True)
        f = bifpn_feature_fusion(G, c, cast(prev), wt)
        f.square().mean().backward()
        G.eval()
        with torch.no_grad():
            m_eval = G.merge(cast(prev), cast(bottom), lat_w, lat_b, "avg")
            f_eval = bifpn_feature_fusion(G, cast(cur), cast(prev), wt)
        out[kind] = dict(m=m.detach(), f=f.detach(), m_eval=m_eval, f_eval=f_eval, dp=p.grad.float(), db=b.grad.float(), dc=c.grad.float(), dw=w_.grad)
    a, b = out["bf16"], out["fp32"]
    for k in ("m", "m_eval", "f_eval"):
        assert a[k].dtype == torch.float32 and torch.equal(a[k], b[k]), k
    # (with autograd the two-term fusion w0 * cur + w1 * up is torch code: a bf16 `cur` is scaled in bf16 there)
    assert float((a["f"].float() - b["f"]).norm()) <= 4e-3 * float(b["f"].norm())
    assert float((a["dc"] - b["dc"]).norm()) <= 8e-3 * float(b["dc"].norm())
    for k in ("dp", "db"):
        assert torch.equal(a[k], b[k].bfloat16().float()), k
    assert float((a["dw"] - b["dw"]).norm()) <= 1e-5 * float(b["dw"].norm())


@pytest.mark.parametrize("precision", ["split", "bf16"])
def test_discriminator_and_single_convs_with_bf16_inputs(precision):
    from afigan.functional import conv1x1_autograd, conv3x3s2_autograd
    from afigan.modeling import Discriminator, Generator
    gen = torch.Generator().manual_seed(73)
    x16 = torch.randn(2, 256, 14, 22, generator=gen).cuda().bfloat16().contiguous(memory_format=torch.channels_last)
    w1 = (torch.randn(256, 256, 1, 1, generator=gen) * 0.06).cuda()
    w3 = (torch.randn(256, 256, 3, 3, generator=gen) * 0.02).cuda()
    res = {}
    for kind in ("fp32", "bf16"):
        x = x16.float().contiguous() if kind == "fp32" else x16
        torch.manual_seed(0)
        Generator(n_residual_dense_blocks=3)
        D = Discriminator(precision=precision).cuda()
        D.Discriminators[0].train()
        xin = x.clone(memory_format=torch.preserve_format).requires_grad_(True)
        lg = D.Discriminators[0](xin)
        lg.square().mean().backward()
        with torch.no_grad():
            y1 = conv1x1_autograd(x, w1, None, precision)
            y2 = conv3x3s2_autograd(x, w3, None, precision)
        res[kind] = (lg.detach(), xin.grad.float(), y1, y2, {k: v.clone() for k, v in D.state_dict().items() if "running" in k})
    a, b = res["bf16"], res["fp32"]
    assert torch.equal(a[0], b[0])
    assert torch.equal(a[1], b[1].bfloat16().float())
    assert torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])
    for k in a[4]:
        assert torch.equal(a[4][k], b[4][k]), k


def test_l1_loss_on_bf16_and_strided_views():
    import ctypes as C
    from afigan import native as N
    gen = torch.Generator().manual_seed(74)
    a16 = torch.randn(2, 256, 9, 15, generator=gen).cuda().bfloat16()
    b32 = torch.randn(2, 256, 10, 16, generator=gen).cuda()[:, :, :9, :15]          # a top-left crop: strided view
    out = torch.zeros(2, device="cuda")
    N.check(N.lib().afi_l1_loss(N.view4(a16), N.view4(b32), 2, 256, 9, 15, out[0:].data_ptr(), None, C.c_float(1.0), None, C.c_float(0.0), N.stream_ptr()))
    ref = (a16.float() - b32).abs().mean()
    assert abs(float(out[0]) - float(ref)) <= 1e-6 * float(ref)


def test_unknown_view_dtype_is_rejected():
    from afigan import native as N
    x = torch.zeros(1, 256, 4, 6, device="cuda")
    v = N.view4(x)
    v.dtype = 7
    y = torch.zeros(2, device="cuda")
    import ctypes as C
    rc = N.lib().afi_l1_loss(v, N.view4(x), 1, 256, 4, 6, y.data_ptr(), None, C.c_float(1.0), None, C.c_float(0.0), N.stream_ptr())
    assert rc != 0 and b"dtype" in N.lib().afi_last_error()
    with pytest.raises(TypeError):
        N.view4(x.half())
