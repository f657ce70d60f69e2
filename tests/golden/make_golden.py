"""Pins oracle/afigan_oracle.py against the UNMODIFIED reference and writes the committed fixtures.

Run in the dev container only (needs /root/reference):   python tests/golden/make_golden.py
It (1) imports generator_rdb.py / feature_patch_discriminator.py from /root/reference through
oracle/_ref_stubs, (2) asserts that the oracle reproduces the reference's init (bit-exact state dicts
under the same seed), forward outputs, stage-1 losses and every parameter gradient, and (3) stores
small fixtures in tests/golden/*.npz: inputs are re-derivable from seeds, so only expected outputs are
kept (full small tensors, and for the 23 M parameter gradients a norm + a strided sample each).
"""
import importlib.util
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref_stubs"))
sys.path.insert(0, ROOT)
from oracle import afigan_oracle as O  # noqa: E402

REF = "/root/reference/afigan/modeling/feat_interpol"


def load_ref(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def sample(t, n=257):
    f = t.detach().reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, min(n, f.numel())).long()
    return f[idx].numpy().copy()


def ref_stage1(G, D, lr_feats, hr_feats):
    """stage1_trainer.py:334-433 executed with the reference modules (optimiser steps omitted)."""
    crit = nn.BCEWithLogitsLoss()

    def reshape(t, size):
        if size[2] != t.size(2) or size[3] != t.size(3):
            return t[:, :, 0:min(size[2], t.size(2)), 0:min(size[3], t.size(3))]
        return t

    d_loss = {}
    for lv, (lo, hi) in enumerate(zip(lr_feats, hr_feats), 2):
        tr = G(lo).detach()
        tr = reshape(tr, hi.size())
        hi = reshape(hi, tr.size())
        real = D.Discriminators[0](hi)
        fake = D.Discriminators[0](tr)
        d_loss[f"d_loss_p{lv}"] = crit(real, torch.ones_like(real)) + crit(fake, torch.zeros_like(fake))
    D.zero_grad()
    sum(d_loss.values()).backward()
    g_loss = {}
    for lv, (lo, hi) in enumerate(zip(lr_feats, hr_feats), 2):
        tr = G(lo)
        tr = reshape(tr, hi.size())
        hi = reshape(hi, tr.size())
        fake = D.Discriminators[0](tr).detach()
        _ = D.Discriminators[0](hi)
        adv = crit(fake, torch.ones_like(fake))
        g_loss[f"g_loss_p{lv}"] = adv * 1e-3 + F.l1_loss(tr, hi)
    G.zero_grad()
    sum(g_loss.values()).backward()
    return d_loss, g_loss


def main():
    gen_mod, dis_mod = load_ref("generator_rdb"), load_ref("feature_patch_discriminator")
    torch.set_num_threads(os.cpu_count())

    # ---- 1. init parity (bit exact)
    torch.manual_seed(0)
    G = gen_mod.Generator(n_residual_dense_blocks=3)
    D = dis_mod.Discriminator()
    g_sd, d_sd = O.init_states(0)
    rg, rd = G.state_dict(), D.state_dict()
    assert list(rg.keys()) == list(g_sd.keys()), "G key order"
    assert list(rd.keys()) == list(d_sd.keys()), "D key order"
    for k in rg:
        assert torch.equal(rg[k], g_sd[k]), k
    for k in rd:
        assert torch.equal(rd[k], d_sd[k]), k
    assert sum(v.numel() for v in rg.values()) == 7_834_624
    assert sum(p.numel() for p in D.parameters()) == 15_352_321
    print("init parity: bit-exact;  G params 7834624, D params 15352321")

    fx = {}
    fx["init_g_sample"] = np.concatenate([sample(v, 17) for v in g_sd.values()])
    fx["init_d_sample"] = np.concatenate([sample(v.float(), 17) for v in d_sd.values()])

    # ---- 2. forward parity on a small ragged case
    gen = torch.Generator().manual_seed(99)
    f = torch.randn(2, 256, 7, 11, generator=gen)
    with torch.no_grad():
        y_ref = G(f)
        y_or = O.generator_forward(g_sd, f)
        assert torch.equal(y_ref, y_or), float((y_ref - y_or).abs().max())
        assert torch.allclose(O.bilinear2x(f), F.interpolate(f, scale_factor=2, mode="bilinear"), atol=1e-6)
        br = O.generator_branch(g_sd, f)
    fx["g_fwd_out"] = y_ref.numpy()
    fx["g_fwd_branch"] = br.numpy()
    x = torch.randn(2, 256, 13, 21, generator=gen)
    D.train()
    d_tmp = {k: v.clone() for k, v in d_sd.items()}
    with torch.no_grad():
        l_ref = D.Discriminators[0](x)
        l_or = O.discriminator_forward(d_tmp, x, True)
    assert torch.equal(l_ref, l_or)
    for k, v in D.state_dict().items():
        assert torch.equal(v, d_tmp[k]), k
    fx["d_fwd_logits"] = l_ref.numpy()
    fx["d_fwd_running_mean0"] = d_tmp["Discriminators.0.0.0.norm.running_mean"].numpy()
    fx["d_fwd_running_var2"] = d_tmp["Discriminators.0.2.0.norm.running_var"].numpy()
    D.eval()
    with torch.no_grad():
        e_ref = D.Discriminators[0](x)
        e_or = O.discriminator_forward(d_tmp, x, False)
    assert torch.equal(e_ref, e_or)
    fx["d_eval_logits"] = e_ref.numpy()
    D.train()
    print("forward parity: bit-exact (G 7x11 -> 14x22, D 13x21 train+eval)")

    # ---- 3. stage-1 step parity (losses + all grads + BN buffers), small 3-level ragged pyramid
    torch.manual_seed(0)
    G = gen_mod.Generator(n_residual_dense_blocks=3)
    D = dis_mod.Discriminator()
    g_sd, d_sd = O.init_states(0)
    lr_shapes, hr_shapes = ((13, 21), (7, 11), (4, 6)), ((25, 42), (13, 21), (7, 11))
    lr_f, hr_f = O.synthetic_features(2, 0, lr_shapes, hr_shapes, seed=4321)
    d_loss, g_loss = ref_stage1(G, D, lr_f, hr_f)
    res = O.stage1_step(g_sd, d_sd, lr_f, hr_f, lr=None, want_outputs=True)
    for k, v in d_loss.items():
        assert abs(float(v) - res["d_loss"][k]) <= 1e-6 * abs(float(v)), (k, float(v), res["d_loss"][k])
    for k, v in g_loss.items():
        assert abs(float(v) - res["g_loss"][k]) <= 1e-6 * abs(float(v)), (k, float(v), res["g_loss"][k])
    worst = 0.0
    for k, p in D.Discriminators[0].named_parameters():
        a, b = p.grad, res["d_grads"]["Discriminators.0." + k]
        if k.endswith("0.bias") and not k.startswith("3."):
            assert a.abs().max() < 1e-5 and b.abs().max() < 1e-5  # true gradient is 0 (bias feeds train-mode BN)
            continue
        worst = max(worst, float((a - b).norm() / a.norm()))
    for k, p in G.Generators[0].named_parameters():
        a, b = p.grad, res["g_grads"]["Generators.0." + k]
        worst = max(worst, float((a - b).norm() / a.norm()))
    assert worst < 1e-5, worst
    for k, v in D.state_dict().items():
        if "running" in k or "num_batches" in k:
            assert torch.allclose(v.float(), d_sd[k].float(), rtol=1e-6, atol=1e-7), k
    assert int(d_sd["Discriminators.0.0.0.norm.num_batches_tracked"]) == 4 * 3
    print(f"stage-1 parity: losses <=1e-6 rel, worst grad rel err {worst:.2e}, BN buffers equal")

    fx["s1_lr_shapes"], fx["s1_hr_shapes"] = np.array(lr_shapes), np.array(hr_shapes)
    fx["s1_d_loss"] = np.array([float(v) for v in d_loss.values()], dtype=np.float64)
    fx["s1_g_loss"] = np.array([float(v) for v in g_loss.values()], dtype=np.float64)
    for lv in (2, 3, 4):
        fx[f"s1_logit_real_p{lv}"] = res["saved"][f"logit_real_p{lv}"].numpy()
        fx[f"s1_logit_fake_p{lv}"] = res["saved"][f"logit_fake_p{lv}"].numpy()
    fx["s1_tr_p4"] = res["saved"]["tr_p4"].numpy()
    for k, p in D.Discriminators[0].named_parameters():
        fx["s1_dgrad_norm/" + k] = np.array(float(p.grad.norm()))
        fx["s1_dgrad_sample/" + k] = sample(p.grad)
    for k, p in G.Generators[0].named_parameters():
        fx["s1_ggrad_norm/" + k] = np.array(float(p.grad.norm()))
        fx["s1_ggrad_sample/" + k] = sample(p.grad)
    for n in range(3):
        for b in ("running_mean", "running_var"):
            k = f"Discriminators.0.{n}.0.norm.{b}"
            fx["s1_bn/" + k] = D.state_dict()[k].numpy()

    # ---- 4. fp64 run of the same step: the oracle's own rounding floor (reported in DESIGN.md)
    g64, d64 = O.init_states(0)
    r64 = O.stage1_step(g64, d64, lr_f, hr_f, lr=None, dtype=torch.float64)
    floor = max(float((res["d_grads"][k].double() - r64["d_grads"][k]).norm() / r64["d_grads"][k].norm())
                for k in O.discriminator_param_keys() if not (k.endswith("0.bias") and ".3." not in k))
    floor_g = max(float((res["g_grads"][k].double() - r64["g_grads"][k]).norm() / r64["g_grads"][k].norm())
                  for k in O.generator_param_keys())
    print(f"fp32-vs-fp64 oracle floor: D grads worst {floor:.2e}, G grads worst {floor_g:.2e}, "
          f"d_loss diff {abs(res['d_total'] - r64['d_total']):.2e}")
    fx["s1_fp64_d_total"], fx["s1_fp64_g_total"] = np.array(r64["d_total"]), np.array(r64["g_total"])
    fx["s1_fp32_vs_fp64_dgrad_floor"], fx["s1_fp32_vs_fp64_ggrad_floor"] = np.array(floor), np.array(floor_g)

    out = os.path.join(HERE, "stage1_small.npz")
    np.savez_compressed(out, **fx)
    print("wrote", out, os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()
