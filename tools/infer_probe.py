"""One BiFPN top-down sweep (4 AF-interpolator fusion calls, eval mode, batch 1) for an ncu launch list or event timing.
usage: infer_probe.py short_side [precision]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "afi-gan_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from afigan.modeling import bifpn_feature_fusion  # noqa: E402

short = int(sys.argv[1])
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
dev = torch.device("cuda")
G, _ = bench._models(prec, dev)
G.eval()
long_ = (short * 1333 // 800 + 127) // 128 * 128
short_p = (short + 127) // 128 * 128
levels = [(short_p // s, long_ // s) for s in (8, 16, 32, 64, 128)]
gen = torch.Generator().manual_seed(36)
feats = [torch.randn(1, 256, h, w, generator=gen).to(dev) for h, w in levels]
wts = torch.tensor([0.7, 1.3], device=dev)


def sweep():
    with torch.no_grad():
        top = feats[4]
        for l in (3, 2, 1, 0):
            top = bifpn_feature_fusion(G, feats[l], top, wts)
    return top


for _ in range(3):
    sweep()
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
e[0].record()
for _ in range(20):
    sweep()
e[1].record()
torch.cuda.synchronize()
print(f"[{prec}] short {short}: {e[0].elapsed_time(e[1]) / 20 * 1e3:.1f} us per 4-call sweep, levels {levels}")
