"""Per-kernel parity of the implicit-GEMM engines against torch CPU fp32 conv on the SAME bf16-representable inputs:
every product is then exact in fp32, only the accumulation order differs (SURVEY.md App. F guidance (1)), so all three
operand modes must agree with the oracle op to ~1e-5 -- a tight, well-posed gate for the tcgen05 descriptors."""
import pytest
import torch
import torch.nn.functional as F

from gpu_util import bf16_round, rel

pytestmark = pytest.mark.gpu

SHAPES = [
    # n, cin, cout, h, w
    (1, 256, 256, 8, 16),      # exactly one 128-pixel tile
    (2, 256, 256, 13, 21),     # ragged (p6-sized level)
    (1, 256, 32, 7, 11),       # dense-block growth conv (N = 32)
    (2, 288, 32, 9, 10),       # K not a multiple of 64
    (1, 384, 256, 12, 20),     # dense-block fusion conv
    (1, 32, 352, 6, 9),        # dense-block dgrad shape (K = 32, N split 2 x 176)
    (1, 256, 512, 25, 42),     # discriminator layer 1
    (1, 512, 1024, 10, 12),    # discriminator layer 2 (4 N tiles)
]


@pytest.mark.parametrize("precision", ["fp32", "bf16_simt", "bf16", "split"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv3x3_forward(shape, precision):
    from afigan.functional import conv3x3
    n, cin, cout, h, w = shape
    g = torch.Generator().manual_seed(hash(shape) % 1000)
    x = bf16_round(torch.randn(n, cin, h, w, generator=g))
    wt = bf16_round(torch.randn(cout, cin, 3, 3, generator=g) * 0.05)
    b = torch.randn(cout, generator=g)
    ref = F.leaky_relu(F.conv2d(x, wt, b, padding=1), 0.2)
    y = conv3x3(x.cuda(), wt.cuda(), b.cuda(), True, precision).cpu()
    tol = 2e-5 if precision in ("fp32", "split") else 6e-3   # bf16 modes round the OUTPUT to bf16 (2^-9 relative)
    assert rel(y, ref) < tol, f"{precision} {shape}: rel err {rel(y, ref):.3e}"
    if precision not in ("fp32", "split"):
        assert rel(y, bf16_round(ref)) < 2e-3


@pytest.mark.parametrize("precision", ["fp32", "bf16_simt", "bf16", "split"])
@pytest.mark.parametrize("shape", SHAPES[:6], ids=lambda s: "x".join(map(str, s)))
def test_conv3x3_backward(shape, precision):
    from afigan.functional import conv3x3_backward
    n, cin, cout, h, w = shape
    g = torch.Generator().manual_seed(1 + hash(shape) % 1000)
    x = bf16_round(torch.randn(n, cin, h, w, generator=g)).requires_grad_(True)
    wt = bf16_round(torch.randn(cout, cin, 3, 3, generator=g) * 0.05).requires_grad_(True)
    dy = bf16_round(torch.randn(n, cout, h, w, generator=g))
    F.conv2d(x, wt, None, padding=1).backward(dy)
    dw, dx = conv3x3_backward(x.detach().cuda(), dy.cuda(), wt.detach().cuda(), precision)
    assert rel(dw, wt.grad) < 3e-5, f"wgrad {precision} {shape}: {rel(dw, wt.grad):.3e}"
    assert rel(dx, x.grad) < 3e-5, f"dgrad {precision} {shape}: {rel(dx, x.grad):.3e}"


SPLIT_SHAPES = SHAPES + [(1, 1024, 1024, 25, 42), (3, 256, 256, 17, 9), (2, 64, 64, 5, 40), (1, 352, 128, 12, 20)]


@pytest.mark.parametrize("pairs", ["default", "six"])
@pytest.mark.parametrize("shape", SPLIT_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_split_precision_matches_fp64_on_full_fp32_operands(shape, pairs, monkeypatch):
    """The split mode multiplies ARBITRARY fp32 operands (not bf16-representable ones) as bf16 plane-pair products with fp32 TMEM
    accumulation.  Six pairs (forward convs of a training step; everything with AFIGAN_SPLIT_PAIRS=6) must match an fp64 evaluation as
    closely as fp32 arithmetic does: the CUDA-core fp32 mode is evaluated next to it.  The tensor cores add into the TMEM accumulator
    with truncation, which biases a chain of n MMAs by ~2e-8 n (measured 1e-5 at K = 9216 with one accumulator); the kernels alternate
    between two accumulators, so the split mode may be up to 6x worse than FFMA fp32 and must stay below 2e-6 x sqrt(K / 256).
    Three pairs (hi hi + hi mid + mid hi: the default for dgrad / wgrad, on which results depend linearly) are accurate to ~2^-17."""
    from afigan.functional import conv3x3, conv3x3_backward
    if pairs == "six":
        monkeypatch.setenv("AFIGAN_SPLIT_PAIRS", "6")
    n, cin, cout, h, w = shape
    g = torch.Generator().manual_seed(11 + hash(shape) % 1000)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) * 0.05
    b = torch.randn(cout, generator=g)
    dy = torch.randn(n, cout, h, w, generator=g)
    xd, wd = x.double().requires_grad_(True), wt.double().requires_grad_(True)
    ref = F.conv2d(xd, wd, b.double(), padding=1)
    ref.backward(dy.double())
    bound = 2e-6 * max(1.0, (9 * max(cin, cout) / 256) ** 0.5)
    errs = {}
    for precision in ("fp32", "split"):
        y = conv3x3(x.cuda(), wt.cuda(), b.cuda(), False, precision).cpu()
        dw, dx = conv3x3_backward(x.cuda(), dy.cuda(), wt.cuda(), precision)
        errs[precision] = (rel(y, ref), rel(dx, xd.grad), rel(dw, wd.grad))
    print(f"{shape} [{pairs}]: fp32 {errs['fp32']}, split {errs['split']}")
    for i, (e32, esp) in enumerate(zip(errs["fp32"], errs["split"])):
        if i == 0 or pairs == "six":
            assert esp < bound and esp < 6 * e32 + 2e-7, (shape, errs)
        else:
            assert esp < 3e-5, (shape, errs)


def test_strided_and_channels_last_inputs():
    from afigan.functional import conv3x3
    g = torch.Generator().manual_seed(3)
    big = bf16_round(torch.randn(2, 256, 14, 22, generator=g))
    wt = bf16_round(torch.randn(256, 256, 3, 3, generator=g) * 0.05)
    crop = big[:, :, :13, :21]                     # the _reshape_stage1 top-left crop: a strided view, no copy
    ref = F.conv2d(crop, wt, None, padding=1)
    y = conv3x3(big.cuda()[:, :, :13, :21], wt.cuda(), None, False, "fp32").cpu()
    assert rel(y, ref) < 2e-5
    y2 = conv3x3(crop.contiguous().cuda().to(memory_format=torch.channels_last), wt.cuda(), None, False, "fp32").cpu()
    assert rel(y2, ref) < 2e-5


# The three tcgen05 convolution kernels must agree with the oracle op on every shape, whichever the library would pick by default:
# per-tap tiles (AFIGAN_CONV_HALO=0), halo tiles on single CTAs (1), halo tiles on CTA pairs for every K (2 + AFIGAN_PAIR_ALL).
VARIANTS = {
    "per_tap": {"AFIGAN_CONV_HALO": "0"},
    "halo": {"AFIGAN_CONV_HALO": "1"},
    "pair": {"AFIGAN_CONV_HALO": "2", "AFIGAN_PAIR_ALL": "1"},
}
VARIANT_SHAPES = SHAPES + [
    (1, 512, 512, 16, 8),       # exactly one 16 x 8 patch: a pair with a masked duplicate tile
    (1, 512, 256, 8, 16),       # one 8 x 16 patch (the other orientation)
    (3, 256, 256, 17, 9),       # odd tile counts, ragged in both directions
    (1, 1024, 1024, 25, 42),    # long K, four N tiles, 8 x 16 orientation
    (2, 64, 64, 5, 40),         # wide and flat
]


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("shape", VARIANT_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv_kernel_variants(shape, variant, monkeypatch):
    from afigan.functional import conv3x3, conv3x3_backward
    for k, v in VARIANTS[variant].items():
        monkeypatch.setenv(k, v)
    n, cin, cout, h, w = shape
    g = torch.Generator().manual_seed(7 + hash(shape) % 1000)
    x = bf16_round(torch.randn(n, cin, h, w, generator=g)).requires_grad_(True)
    wt = bf16_round(torch.randn(cout, cin, 3, 3, generator=g) * 0.05).requires_grad_(True)
    b = torch.randn(cout, generator=g)
    dy = bf16_round(torch.randn(n, cout, h, w, generator=g))
    ref = F.leaky_relu(F.conv2d(x, wt, b, padding=1), 0.2)
    y = conv3x3(x.detach().cuda(), wt.detach().cuda(), b.cuda(), True, "bf16").cpu()
    assert rel(y, bf16_round(ref)) < 2e-3, f"{variant} {shape}: forward rel err {rel(y, bf16_round(ref)):.3e}"
    F.conv2d(x, wt, None, padding=1).backward(dy)
    dw, dx = conv3x3_backward(x.detach().cuda(), dy.cuda(), wt.detach().cuda(), "bf16")
    assert rel(dx, x.grad) < 3e-5, f"{variant} {shape}: dgrad rel err {rel(dx, x.grad):.3e}"
    assert rel(dw, wt.grad) < 3e-5
