"""The oracle against the committed fixtures (made by tests/golden/make_golden.py from the unmodified reference)."""
import os

import numpy as np
import torch

from oracle import afigan_oracle as O

FX = np.load(os.path.join(os.path.dirname(__file__), "golden", "stage1_small.npz"))


def _sample(t, n=257):
    f = t.detach().reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, min(n, f.numel())).long()
    return f[idx].numpy()


def test_init_matches_reference_bit_exact():
    g_sd, d_sd = O.init_states(0)
    assert sum(v.numel() for v in g_sd.values()) == 7_834_624
    assert sum(d_sd[k].numel() for k in O.discriminator_param_keys()) == 15_352_321
    assert np.array_equal(np.concatenate([_sample(v, 17) for v in g_sd.values()]), FX["init_g_sample"])
    assert np.array_equal(np.concatenate([_sample(v.float(), 17) for v in d_sd.values()]), FX["init_d_sample"])
    assert list(g_sd)[:2] == ["Generators.0.0.0.weight", "Generators.0.0.0.bias"]
    assert g_sd["Generators.0.3.0.weight"].shape == (256, 256, 6, 6)
    assert d_sd["Discriminators.0.2.0.norm.running_var"].shape == (1024,)


def test_forward_matches_reference():
    g_sd, d_sd = O.init_states(0)
    gen = torch.Generator().manual_seed(99)
    f = torch.randn(2, 256, 7, 11, generator=gen)
    with torch.no_grad():
        y = O.generator_forward(g_sd, f)
        assert y.shape == (2, 256, 14, 22)
        np.testing.assert_allclose(y.numpy(), FX["g_fwd_out"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(O.generator_branch(g_sd, f).numpy(), FX["g_fwd_branch"], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(O.bilinear2x(f).numpy(), (y - O.generator_branch(g_sd, f)).numpy(), atol=1e-5)
        x = torch.randn(2, 256, 13, 21, generator=gen)
        lg = O.discriminator_forward(d_sd, x, True)
        np.testing.assert_allclose(lg.numpy(), FX["d_fwd_logits"], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(d_sd["Discriminators.0.0.0.norm.running_mean"].numpy(), FX["d_fwd_running_mean0"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(d_sd["Discriminators.0.2.0.norm.running_var"].numpy(), FX["d_fwd_running_var2"], rtol=1e-4)
        ev = O.discriminator_forward(d_sd, x, False)
        np.testing.assert_allclose(ev.numpy(), FX["d_eval_logits"], rtol=1e-4, atol=1e-4)


def test_known_answers():
    g_sd, _ = O.init_states(0)
    zero = {k: torch.zeros_like(v) for k, v in g_sd.items()}
    f = torch.randn(1, 256, 3, 5)
    assert torch.allclose(O.generator_forward(zero, f), O.bilinear2x(f), atol=1e-6)
    row = torch.arange(6.0).reshape(1, 1, 1, 6)
    assert torch.allclose(O.bilinear2x(row)[0, 0, 0], torch.tensor([0, .25, .75, 1.25, 1.75, 2.25, 2.75, 3.25, 3.75, 4.25, 4.75, 5.0]))
    assert torch.equal(O.nearest_half(torch.arange(7.0).reshape(1, 1, 1, 7).expand(1, 1, 2, 7))[0, 0, 0], torch.tensor([0., 2., 4.]))


def test_stage1_step_matches_reference():
    g_sd, d_sd = O.init_states(0)
    lr_shapes = tuple(map(tuple, FX["s1_lr_shapes"]))
    hr_shapes = tuple(map(tuple, FX["s1_hr_shapes"]))
    lr_f, hr_f = O.synthetic_features(2, 0, lr_shapes, hr_shapes, seed=4321)
    res = O.stage1_step(g_sd, d_sd, lr_f, hr_f, lr=None, want_outputs=True)
    np.testing.assert_allclose(list(res["d_loss"].values()), FX["s1_d_loss"], rtol=2e-6)
    np.testing.assert_allclose(list(res["g_loss"].values()), FX["s1_g_loss"], rtol=2e-6)
    np.testing.assert_allclose(res["saved"]["tr_p4"].numpy(), FX["s1_tr_p4"], rtol=1e-5, atol=1e-6)
    for lv in (2, 3, 4):
        np.testing.assert_allclose(res["saved"][f"logit_fake_p{lv}"].numpy(), FX[f"s1_logit_fake_p{lv}"], rtol=1e-4, atol=1e-4)
    for k, g in res["d_grads"].items():
        short = k[len("Discriminators.0."):]
        ref_norm = float(FX["s1_dgrad_norm/" + short])
        if short.endswith("0.bias") and not short.startswith("3."):
            assert float(g.norm()) < 1e-5
            continue
        assert abs(float(g.norm()) - ref_norm) <= 1e-4 * ref_norm, k
        np.testing.assert_allclose(_sample(g), FX["s1_dgrad_sample/" + short], rtol=1e-3, atol=1e-5 * ref_norm)
    for k, g in res["g_grads"].items():
        short = k[len("Generators.0."):]
        ref_norm = float(FX["s1_ggrad_norm/" + short])
        assert abs(float(g.norm()) - ref_norm) <= 1e-4 * ref_norm, k
        np.testing.assert_allclose(_sample(g), FX["s1_ggrad_sample/" + short], rtol=1e-3, atol=1e-5 * ref_norm)
    assert int(d_sd["Discriminators.0.1.0.norm.num_batches_tracked"]) == 12   # 4 D calls x 3 levels
    for n in range(3):
        for b in ("running_mean", "running_var"):
            k = f"Discriminators.0.{n}.0.norm.{b}"
            np.testing.assert_allclose(d_sd[k].numpy(), FX["s1_bn/" + k], rtol=1e-4, atol=1e-6)
