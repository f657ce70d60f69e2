"""Inference-side cost of the AF interpolator (BASELINE config C5 shapes): the 28 fusion sites of a 7-layer BiFPN top-down path on one image,
wall clock vs device time (how much of it is host overhead), plus the three merge sites of an FPN."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "afi-gan_b200"))
import torch  # noqa: E402

from afigan.modeling import Generator, bifpn_feature_fusion  # noqa: E402

short = int(sys.argv[1]) if len(sys.argv) > 1 else 800
long_ = (short * 1333 // 800 + 127) // 128 * 128
short_p = (short + 127) // 128 * 128
torch.manual_seed(0)
G = Generator(n_residual_dense_blocks=3, precision="bf16").cuda().eval()
levels = [(short_p // s, long_ // s) for s in (8, 16, 32, 64, 128)]          # p3 .. p7
feats = [torch.randn(1, 256, h, w, device="cuda") for h, w in levels]
wts = torch.tensor([0.7, 1.3], device="cuda")


def one_image():
    with torch.no_grad():
        for _ in range(7):                       # seven BiFPN layers, four top-down fusion sites each
            top = feats[4]
            for l in (3, 2, 1, 0):
                top = bifpn_feature_fusion(G, feats[l], top, wts)
    return top


for _ in range(3):
    one_image()
torch.cuda.synchronize()
n = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(n):
    one_image()
e1.record()
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t_wall = time.perf_counter() - t0
print(f"short side {short}: levels {levels}; per image: host issue {1e3 * t_issue / n:.2f} ms, wall {1e3 * t_wall / n:.2f} ms, "
      f"device span {e0.elapsed_time(e1) / n:.2f} ms  (28 interpolator calls)")
