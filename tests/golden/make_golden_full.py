"""Full-size config-1 golden: ONE stage-1 step of the UNMODIFIED reference modules at BASELINE.json's shapes.

Run in the dev container only (needs /root/reference; ~2-3 min and ~20 GB of host memory):
    python tests/golden/make_golden_full.py
Batch 2, five pyramid levels (LR 104x168 ... 7x11 -> HR 200x336 ... 13x21, crops 208 -> 200 and 14x22 -> 13x21 as
stage1_trainer.py:437-443 does them), weights under torch.manual_seed(0) (G first, then D: stage1_trainer.py:505-506),
features from torch.Generator().manual_seed(1234) -- exactly the tensors bench.py feeds rank 0.  The reference step
(stage1_trainer.py:334-433, optimiser updates omitted) runs through oracle/_ref_stubs on torch CPU fp32; what is kept
(tests/golden/stage1_full.npz, < 200 KB): the ten losses, for each of the 33 parameter tensors the gradient norm and a
257-point strided sample, and the BatchNorm running buffers.  tests/test_gpu_full_size.py compares the CUDA path with it.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref_stubs"))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_golden import load_ref, ref_stage1, sample  # noqa: E402
from oracle import afigan_oracle as O  # noqa: E402

LR_SHAPES = ((104, 168), (52, 84), (26, 42), (13, 21), (7, 11))
HR_SHAPES = ((200, 336), (100, 168), (50, 84), (25, 42), (13, 21))


def main():
    gen_mod, dis_mod = load_ref("generator_rdb"), load_ref("feature_patch_discriminator")
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    G = gen_mod.Generator(n_residual_dense_blocks=3)
    D = dis_mod.Discriminator()
    D.train()
    lr_f, hr_f = O.synthetic_features(2, 0, LR_SHAPES, HR_SHAPES, seed=1234)
    t0 = time.time()
    d_loss, g_loss = ref_stage1(G, D, lr_f, hr_f)
    print(f"reference step at config-1 size: {time.time() - t0:.1f} s on {os.cpu_count()} threads")
    fx = {"lr_shapes": np.array(LR_SHAPES), "hr_shapes": np.array(HR_SHAPES), "seed": np.array(1234), "batch": np.array(2)}
    fx["d_loss"] = np.array([float(v) for v in d_loss.values()], dtype=np.float64)
    fx["g_loss"] = np.array([float(v) for v in g_loss.values()], dtype=np.float64)
    for k, p in D.Discriminators[0].named_parameters():
        fx["dgrad_norm/" + k] = np.array(float(p.grad.norm()))
        fx["dgrad_sample/" + k] = sample(p.grad)
    for k, p in G.Generators[0].named_parameters():
        fx["ggrad_norm/" + k] = np.array(float(p.grad.norm()))
        fx["ggrad_sample/" + k] = sample(p.grad)
    sd = D.state_dict()
    for n in range(3):
        for b in ("running_mean", "running_var"):
            k = f"Discriminators.0.{n}.0.norm.{b}"
            fx["bn/" + k] = sd[k].numpy()
        fx[f"bn/nbt{n}"] = np.array(int(sd[f"Discriminators.0.{n}.0.norm.num_batches_tracked"]))
    print("d_loss", fx["d_loss"], "g_loss", fx["g_loss"])
    out = os.path.join(HERE, "stage1_full.npz")
    np.savez_compressed(out, **fx)
    print("wrote", out, os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()
