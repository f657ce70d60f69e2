"""The kernel-sequence model (tests/kernel_model.py, which csrc/api.cu transliterates) against the oracle."""
import torch

import kernel_model as KM
from oracle import afigan_oracle as O


def _rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def _to64(sd):
    return {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}


def test_generator_sequence_matches_oracle_fp64():
    g_sd, _ = O.init_states(0)
    g_sd = _to64(g_sd)
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(2, 256, 5, 7, generator=gen, dtype=torch.float64)
    keys = O.generator_param_keys()
    params = {k: g_sd[k].clone().requires_grad_(True) for k in keys}
    xg = x.clone().requires_grad_(True)
    y_ref = O.generator_forward(params, xg)
    y, saved = KM.g_forward(g_sd, x)
    assert _rel(y, y_ref.detach()) < 1e-12
    # ragged crop like stage 1: gradient only inside the top-left 9x13 window
    dy = torch.randn(2, 256, 9, 13, generator=gen, dtype=torch.float64)
    (y_ref[:, :, :9, :13] * dy).sum().backward()
    grads, dx = KM.g_backward(g_sd, saved, dy, need_dx=True)
    for k in keys:
        assert _rel(grads[k], params[k].grad) < 1e-10, k
    assert _rel(dx, xg.grad) < 1e-10


def test_discriminator_sequence_matches_oracle_fp64():
    _, d_sd = O.init_states(0)
    d_sd = _to64(d_sd)
    gen = torch.Generator().manual_seed(6)
    for k in d_sd:   # non-trivial BN affine so gamma/beta paths are exercised
        if k.endswith("norm.weight"):
            d_sd[k] = 1 + 0.1 * torch.randn(d_sd[k].shape, generator=gen, dtype=torch.float64)
        if k.endswith("norm.bias") or (k.endswith("0.bias")):
            d_sd[k] = 0.1 * torch.randn(d_sd[k].shape, generator=gen, dtype=torch.float64)
    x = torch.randn(2, 256, 6, 5, generator=gen, dtype=torch.float64)
    keys = O.discriminator_param_keys()
    params = dict(d_sd)
    for k in keys:
        params[k] = d_sd[k].clone().requires_grad_(True)
    logit_ref = O.discriminator_forward(params, x, True, update_running=False)
    logit, saved = KM.d_forward(d_sd, x)
    assert _rel(logit, logit_ref.detach()) < 1e-11
    loss = O.bce_logits_mean(logit_ref, 1.0)
    loss.backward()
    dlogit = (torch.sigmoid(logit) - 1.0) / logit.numel()
    grads = KM.d_backward(d_sd, saved, dlogit)
    for k in keys:
        ref = params[k].grad
        if k.endswith("0.bias") and ".3." not in k:
            assert grads[k].abs().max() < 1e-12 and ref.abs().max() < 1e-12  # feeds train-mode BN: zero
        else:
            assert _rel(grads[k], ref) < 1e-9, k
    # running-stat update rule (momentum 0.1, unbiased variance)
    d2 = {k: v.clone() for k, v in d_sd.items()}
    O.discriminator_forward(d2, x, True)
    mean, _, var_unb = saved["stats"][0]
    assert _rel(0.1 * mean, d2["Discriminators.0.0.0.norm.running_mean"]) < 1e-10
    assert _rel(0.9 + 0.1 * var_unb, d2["Discriminators.0.0.0.norm.running_var"]) < 1e-10
