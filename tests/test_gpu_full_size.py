"""One stage-1 step at BASELINE.json's FULL config-1 size (batch 2, five levels, 200x336 ... 13x21, crops 208 -> 200 and
14x22 -> 13x21) against the golden fixture tests/golden/stage1_full.npz, which tests/golden/make_golden_full.py made by running the
UNMODIFIED reference modules (generator_rdb.py, feature_patch_discriminator.py) through the loss block of stage1_trainer.py:334-433
on torch CPU fp32.  Inputs are the tensors bench.py feeds rank 0 (seed 1234)."""
import os

import numpy as np
import pytest
import torch

from gpu_util import TOL
from oracle import afigan_oracle as O

pytestmark = pytest.mark.gpu
FX = np.load(os.path.join(os.path.dirname(__file__), "golden", "stage1_full.npz"))


def _sample(t, n=257):
    f = t.detach().reshape(-1).cpu()
    idx = torch.linspace(0, f.numel() - 1, min(n, f.numel())).long()
    return f[idx].double().numpy()


@pytest.mark.parametrize("precision", ["split", "fp32", "bf16"])
def test_stage1_full_size_vs_golden(precision):
    from afigan.engine import Stage1Step
    from afigan.modeling import Discriminator, Generator
    tol = TOL[precision]
    torch.manual_seed(0)
    G = Generator(n_residual_dense_blocks=3, precision=precision).cuda()
    D = Discriminator(precision=precision).cuda()
    D.Discriminators[0].train()
    step = Stage1Step(G, D, lr=1e-3, precision=precision)
    lr_shapes, hr_shapes = tuple(map(tuple, FX["lr_shapes"])), tuple(map(tuple, FX["hr_shapes"]))
    lr_f, hr_f = O.synthetic_features(int(FX["batch"]), 0, lr_shapes, hr_shapes, seed=int(FX["seed"]))
    step.run_step([t.cuda() for t in lr_f], [t.cuda() for t in hr_f], apply_updates=False)
    m = step.metrics(5)
    d_loss = np.array([m[f"d_loss_p{l}"] for l in range(2, 7)])
    g_loss = np.array([m[f"g_loss_p{l}"] for l in range(2, 7)])
    worst = {"d_loss": float(np.max(np.abs(d_loss - FX["d_loss"]) / FX["d_loss"])), "g_loss": float(np.max(np.abs(g_loss - FX["g_loss"])))}
    np.testing.assert_allclose(d_loss, FX["d_loss"], rtol=tol["loss"])
    np.testing.assert_allclose(g_loss, FX["g_loss"], rtol=tol["loss"], atol=1e-4 if precision == "bf16" else 1e-5)
    # gradients: norm of every parameter tensor and a 257-point strided sample of it
    worst["d_norm"] = worst["g_norm"] = worst["d_sample"] = worst["g_sample"] = 0.0
    for name, p in D.Discriminators[0].named_parameters():
        if name.endswith("0.bias") and not name.startswith("3."):
            assert float(p.grad.abs().max()) < 2e-3        # true gradient 0: the bias feeds a batch-statistics BatchNorm
            continue
        ref_norm = float(FX["dgrad_norm/" + name])
        e = abs(float(p.grad.norm()) - ref_norm) / ref_norm
        smp, ref = _sample(p.grad), FX["dgrad_sample/" + name].astype(np.float64)
        es = float(np.linalg.norm(smp - ref) / np.linalg.norm(ref))
        worst["d_norm"], worst["d_sample"] = max(worst["d_norm"], e), max(worst["d_sample"], es)
        assert e <= tol["grad"], (name, e)
        assert es <= (tol["dgrad"] if precision != "bf16" else 0.25), (name, es)
    for name, p in G.Generators[0].named_parameters():
        ref_norm = float(FX["ggrad_norm/" + name])
        e = abs(float(p.grad.norm()) - ref_norm) / ref_norm
        smp, ref = _sample(p.grad), FX["ggrad_sample/" + name].astype(np.float64)
        es = float(np.linalg.norm(smp - ref) / np.linalg.norm(ref))
        worst["g_norm"], worst["g_sample"] = max(worst["g_norm"], e), max(worst["g_sample"], es)
        assert e <= tol["grad"], (name, e)
        assert es <= (tol["grad"] if precision != "bf16" else 0.25), (name, es)
    sd = D.state_dict()
    worst["bn"] = 0.0
    for n in range(3):
        for b in ("running_mean", "running_var"):
            k = f"Discriminators.0.{n}.0.norm.{b}"
            ref = torch.from_numpy(FX["bn/" + k]).double()
            e = float((sd[k].double().cpu() - ref).norm() / ref.norm())
            worst["bn"] = max(worst["bn"], e)
            assert e < (1e-4 if precision != "bf16" else 3e-2), (k, e)
        assert int(sd[f"Discriminators.0.{n}.0.norm.num_batches_tracked"]) == int(FX[f"bn/nbt{n}"]) == 20
    print(f"[{precision}] full-size config-1 step vs the reference golden: " + ", ".join(f"{k} {v:.2e}" for k, v in worst.items()))
