"""The drop-in nn.Modules (Generator / Discriminator, autograd path) against the oracle on the same seeded inputs."""
import pytest
import torch

from gpu_util import TOL, cosine, rel
from oracle import afigan_oracle as O

pytestmark = pytest.mark.gpu


def _modules(precision):
    from afigan.modeling import Discriminator, Generator
    torch.manual_seed(0)
    G = Generator(n_residual_dense_blocks=3, precision=precision)
    D = Discriminator(precision=precision)
    g_sd, d_sd = O.init_states(0)
    assert all(torch.equal(G.state_dict()[k], v) for k, v in g_sd.items())       # same seed => same weights as the reference
    assert all(torch.equal(D.state_dict()[k], v) for k, v in d_sd.items())
    return G.cuda(), D.cuda(), g_sd, d_sd


@pytest.mark.parametrize("precision", ["fp32", "split", "bf16_simt", "bf16"])
def test_generator_forward_backward(precision):
    G, _, g_sd, _ = _modules(precision)
    tol = TOL[precision]
    gen = torch.Generator().manual_seed(11)
    x = torch.randn(2, 256, 7, 11, generator=gen)
    hr = torch.randn(2, 256, 13, 21, generator=gen)
    keys = O.generator_param_keys()
    params = dict(g_sd)
    for k in keys:
        params[k] = g_sd[k].clone().requires_grad_(True)
    y_ref = O.generator_forward(params, x)
    assert y_ref.shape == (2, 256, 14, 22)
    y = G(x.cuda())
    assert y.shape == (2, 256, 14, 22)                                            # exactly 2x for odd sizes too
    assert rel(y, y_ref) < tol["feat"], rel(y, y_ref)
    # the learned branch alone (the skip dominates the output at init: SURVEY.md §7 hard part 1)
    br, br_ref = y.cpu() - O.bilinear2x(x), y_ref.detach() - O.bilinear2x(x)
    assert rel(br, br_ref) < (1e-3 if precision in ("fp32", "split") else 3e-2), rel(br, br_ref)
    # stage-1 style loss on the top-left 13x21 crop
    yc = G(x.cuda(), out_hw=(13, 21))
    assert rel(yc, y_ref[:, :, :13, :21]) < tol["feat"]
    loss = (yc - hr.cuda()).abs().mean()
    loss_ref = (y_ref[:, :, :13, :21] - hr).abs().mean()
    assert abs(float(loss.detach()) - float(loss_ref.detach())) < 1e-4
    loss.backward()
    loss_ref.backward()
    worst = 0.0
    for k, p in zip(keys, G._params()):
        r, c = rel(p.grad, params[k].grad), cosine(p.grad, params[k].grad)
        worst = max(worst, r)
        assert r < tol["grad"] and c > tol["cos"], f"{precision} {k}: rel {r:.3e} cos {c:.6f}"
    print(f"[{precision}] G grads worst rel err {worst:.3e}")


def test_generator_zero_weights_is_bilinear():
    G, _, _, _ = _modules("fp32")
    with torch.no_grad():
        for p in G.parameters():
            p.zero_()
    x = torch.randn(1, 256, 5, 6)
    y = G(x.cuda()).cpu()
    assert torch.allclose(y, torch.nn.functional.interpolate(x, scale_factor=2, mode="bilinear"), atol=1e-6)


@pytest.mark.parametrize("precision", ["fp32", "split", "bf16_simt", "bf16"])
def test_discriminator_forward_backward(precision):
    _, D, _, d_sd = _modules(precision)
    tol = TOL[precision]
    gen = torch.Generator().manual_seed(12)
    x = torch.randn(2, 256, 13, 21, generator=gen)
    keys = O.discriminator_param_keys()
    params = dict(d_sd)
    for k in keys:
        params[k] = d_sd[k].clone().requires_grad_(True)
    stack = D.Discriminators[0]
    stack.train()
    logit_ref = O.discriminator_forward(params, x, True)
    logit = stack(x.cuda())
    assert logit.shape == (2, 1, 13, 21)
    assert rel(logit, logit_ref) < tol["logit"], rel(logit, logit_ref)
    loss_ref = O.bce_logits_mean(logit_ref, 0.0)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logit, torch.zeros_like(logit))
    assert abs(float(loss.detach()) - float(loss_ref.detach())) < tol["loss"] * abs(float(loss_ref.detach())) + 1e-6
    loss.backward()
    loss_ref.backward()
    for k, p in zip(keys, stack._params()):
        if k.endswith("0.bias") and ".3." not in k:
            assert float(p.grad.abs().max()) < 2e-3       # bias feeds train-mode BN: true gradient is 0 (App. D-4); fp32 sum noise only
            continue
        r, c = rel(p.grad, params[k].grad), cosine(p.grad, params[k].grad)
        assert r < tol["dgrad"] and c > tol["cos"], f"{precision} {k}: rel {r:.3e} cos {c:.6f}"
    sd = D.state_dict()
    for n in range(3):
        p = f"Discriminators.0.{n}.0.norm."
        assert rel(sd[p + "running_mean"], params[p + "running_mean"]) < (1e-5 if precision in ("fp32", "split") else 2e-2)
        assert rel(sd[p + "running_var"], params[p + "running_var"]) < (1e-5 if precision in ("fp32", "split") else 2e-2)
        assert int(sd[p + "num_batches_tracked"]) == 1
    # eval mode uses the running statistics
    stack.eval()
    with torch.no_grad():
        ev = stack(x.cuda())
    ev_ref = O.discriminator_forward(params, x, False)
    assert rel(ev, ev_ref.detach()) < tol["logit"]


def test_cpu_tensor_is_a_hard_error():
    from afigan.modeling import Generator
    G = Generator(n_residual_dense_blocks=3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        G(torch.zeros(1, 256, 4, 4))


@pytest.mark.parametrize("precision", ["fp32", "split", "bf16"])
def test_discriminator_input_gradient(precision):
    """Gradient w.r.t. the discriminator's input (the trainers detach it, but the module must be a well-behaved autograd citizen)."""
    _, D, _, d_sd = _modules(precision)
    stack = D.Discriminators[0]
    stack.train()
    x = torch.randn(2, 256, 9, 12, generator=torch.Generator().manual_seed(41))
    xr = x.clone().requires_grad_(True)
    O.bce_logits_mean(O.discriminator_forward({k: v.clone() for k, v in d_sd.items()}, xr, True), 1.0).backward()
    xc = x.cuda().requires_grad_(True)
    logit = stack(xc)
    torch.nn.functional.binary_cross_entropy_with_logits(logit, torch.ones_like(logit)).backward()
    r = rel(xc.grad, xr.grad)
    assert r < (5e-3 if precision in ("fp32", "split") else 0.15) and cosine(xc.grad, xr.grad) > (0.9999 if precision in ("fp32", "split") else 0.985), r


@pytest.mark.parametrize("precision", ["fp32", "split", "bf16"])
def test_discriminator_eval_mode_backward(precision):
    """Backward through an eval-mode discriminator (running statistics are constants: no mean terms, conv biases DO get a gradient)."""
    _, D, _, d_sd = _modules(precision)
    tol = TOL[precision]
    stack = D.Discriminators[0]
    gen = torch.Generator().manual_seed(42)
    sd = {k: v.clone() for k, v in d_sd.items()}
    for n in range(3):      # non-trivial running statistics
        sd[f"Discriminators.0.{n}.0.norm.running_mean"] = 0.1 * torch.randn(sd[f"Discriminators.0.{n}.0.norm.running_mean"].shape, generator=gen)
        sd[f"Discriminators.0.{n}.0.norm.running_var"] = 1 + 50 * torch.rand(sd[f"Discriminators.0.{n}.0.norm.running_var"].shape, generator=gen)
    D.load_state_dict(sd)
    stack.eval()
    x = torch.randn(2, 256, 8, 10, generator=gen)
    keys = O.discriminator_param_keys()
    params = dict(sd)
    for k in keys:
        params[k] = sd[k].clone().requires_grad_(True)
    O.bce_logits_mean(O.discriminator_forward(params, x, False), 0.0).backward()
    logit = stack(x.cuda())
    torch.nn.functional.binary_cross_entropy_with_logits(logit, torch.zeros_like(logit)).backward()
    for k, p in zip(keys, stack._params()):
        r = rel(p.grad, params[k].grad)
        assert r < tol["dgrad"], f"{precision} {k}: {r:.3e}"
    assert int(D.state_dict()["Discriminators.0.0.0.norm.num_batches_tracked"]) == 0          # eval mode leaves the buffers alone


@pytest.mark.parametrize("precision", ["fp32", "split", "bf16"])
def test_conv3x3_autograd_function(precision):
    from afigan.functional import conv3x3_autograd
    gen = torch.Generator().manual_seed(43)
    x = torch.randn(2, 256, 12, 17, generator=gen)
    w = torch.randn(256, 256, 3, 3, generator=gen) * 0.03
    b = torch.randn(256, generator=gen)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    ref = torch.nn.functional.conv2d(xr, wr, br, padding=1)
    dy = torch.randn(ref.shape, generator=gen)
    (ref * dy).sum().backward()
    xc, wc, bc = (t.cuda().requires_grad_(True) for t in (x, w, b))
    out = conv3x3_autograd(xc, wc, bc, precision)
    (out * dy.cuda()).sum().backward()
    t = 1e-5 if precision in ("fp32", "split") else 1e-2
    assert rel(out, ref) < t and rel(xc.grad, xr.grad) < t and rel(wc.grad, wr.grad) < t and rel(bc.grad, br.grad) < t
