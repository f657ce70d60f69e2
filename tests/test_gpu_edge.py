"""Edge cases and size-independent properties at BASELINE.json's full sizes (config 1: batch 2, 800x1333 R-50-FPN shapes)."""
import pytest
import torch
import torch.nn.functional as F

from gpu_util import rel
from oracle import afigan_oracle as O

pytestmark = pytest.mark.gpu


def _gd(precision, n_rdb=3):
    from afigan.modeling import Discriminator, Generator
    torch.manual_seed(0)
    G = Generator(n_residual_dense_blocks=n_rdb, precision=precision).cuda()
    D = Discriminator(precision=precision).cuda()
    return G, D


@pytest.mark.parametrize("shape", [(1, 1, 1), (1, 2, 3), (3, 5, 4), (1, 17, 33)])
@pytest.mark.parametrize("precision", ["fp32", "split", "bf16"])
def test_generator_tiny_and_odd_shapes(shape, precision):
    n, h, w = shape
    G, _ = _gd(precision)
    g_sd, _ = O.init_states(0)
    x = torch.randn(n, 256, h, w, generator=torch.Generator().manual_seed(h * 100 + w))
    y = G(x.cuda())
    assert y.shape == (n, 256, 2 * h, 2 * w)                      # exactly 2x for any size (k6 s2 p2), SURVEY §8c (i)
    assert rel(y, O.generator_forward(g_sd, x)) < (1e-5 if precision in ("fp32", "split") else 1e-3)


def test_generator_two_dense_blocks_default_ctor():
    """Generator() defaults to n_residual_dense_blocks=2 (generator_rdb.py:75); every caller passes 3, both must work."""
    from afigan.modeling import Generator
    torch.manual_seed(3)
    G = Generator(precision="fp32").cuda()
    assert G.n_residual_dense_blocks == 2 and len(G.Generators[0][1].RDBs) == 2
    sd = {k: v.detach().cpu() for k, v in G.state_dict().items()}
    x = torch.randn(2, 256, 6, 5)
    assert rel(G(x.cuda()), O.generator_forward(sd, x)) < 1e-5


@pytest.mark.parametrize("precision", ["fp32", "split", "bf16"])
def test_discriminator_batch_one_and_eval(precision):
    _, D = _gd(precision)
    _, d_sd = O.init_states(0)
    stack = D.Discriminators[0]
    x = torch.randn(1, 256, 9, 7, generator=torch.Generator().manual_seed(5))
    tol = 2e-4 if precision in ("fp32", "split") else 3e-2
    stack.train()
    with torch.no_grad():
        assert rel(stack(x.cuda()), O.discriminator_forward(d_sd, x, True)) < tol
        stack.eval()
        assert rel(stack(x.cuda()), O.discriminator_forward(d_sd, x, False)) < tol
    assert D(x.cuda()).shape == (1, 1, 9, 7)                      # Discriminator.forward dispatches on current_step == 0


def test_linearity_of_the_conv_engine_full_size():
    """Size-independent property at the full p2 size: conv(a x + b z) == a conv(x) + b conv(z) for the bias-free, activation-free layer
    (exact in fp32 up to rounding; checks tiling/halo handling over all 525 x 2 tiles without needing a CPU oracle at this size)."""
    from afigan.functional import conv3x3
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 256, 200, 336, generator=g).cuda()
    z = torch.randn(2, 256, 200, 336, generator=g).cuda()
    w = (torch.randn(256, 256, 3, 3, generator=g) * 0.02).cuda()
    for precision, tol in (("fp32", 2e-6), ("bf16", 8e-3)):
        lhs = conv3x3(0.5 * x + 2.0 * z, w, None, False, precision)
        rhs = 0.5 * conv3x3(x, w, None, False, precision) + 2.0 * conv3x3(z, w, None, False, precision)
        assert rel(lhs, rhs) < tol, (precision, rel(lhs, rhs))
    # and the two engines agree with each other and with cuDNN (library cross-check, fp32, TF32 off)
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d(x, w, None, padding=1)
    assert rel(conv3x3(x, w, None, False, "fp32"), ref) < 1e-5
    assert rel(conv3x3(x, w, None, False, "bf16"), ref) < 6e-3


def test_full_size_step_properties():
    """Config-1 full-size stage-1 step: (a) bf16 and fp32 operand modes agree on every loss, (b) zero generator weights give exactly the
    bilinear skip, (c) after one step each D BatchNorm has seen 4 calls x 5 levels = 20 batches."""
    from afigan.engine import Stage1Step
    lr_f, hr_f = O.synthetic_features(2, 0)
    lr_c, hr_c = [t.cuda() for t in lr_f], [t.cuda() for t in hr_f]
    losses = {}
    for precision in ("fp32", "bf16"):
        G, D = _gd(precision)
        step = Stage1Step(G, D, precision=precision)
        step.run_step(lr_c, hr_c, apply_updates=False)
        losses[precision] = step.metrics()
        assert int(D.state_dict()["Discriminators.0.2.0.norm.num_batches_tracked"]) == 20
        assert all(torch.isfinite(p.grad).all() for p in step.g_params + step.d_params)
        del step, G, D
        torch.cuda.empty_cache()
    for k, v in losses["fp32"].items():
        assert abs(losses["bf16"][k] - v) <= 2e-3 * abs(v) + 1e-4, (k, v, losses["bf16"][k])
    # survey probe at random init: d_loss ~ 25 per level, L1 ~ 0.945+ per level
    assert 20 < losses["fp32"]["d_loss_p2"] < 30 and 0.9 < losses["fp32"]["content_loss_p2"] < 1.0
    G, _ = _gd("bf16")
    with torch.no_grad():
        for p in G.parameters():
            p.zero_()
        y = G(lr_c[0])
    assert torch.allclose(y, F.interpolate(lr_c[0], scale_factor=2, mode="bilinear"), atol=1e-5)
