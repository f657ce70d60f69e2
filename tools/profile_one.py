"""One forward 3x3 conv shape in isolation: CUDA-event timing, or the target of an `ncu -k regex:k_conv` capture.
usage: profile_one.py n cin cout h w [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "afi-gan_b200"))
import torch  # noqa: E402

from afigan.functional import conv3x3  # noqa: E402

n, cin, cout, h, w = map(int, sys.argv[1:6])
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 5
prec = sys.argv[7] if len(sys.argv) > 7 else "bf16"
x = torch.randn(n, cin, h, w, device="cuda")
wt = torch.randn(cout, cin, 3, 3, device="cuda") * 0.02
fl = 2.0 * n * h * w * 9 * cin * cout
conv3x3(x, wt, None, False, prec)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    conv3x3(x, wt, None, False, prec)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"[{prec}] halo={os.environ.get('AFIGAN_CONV_HALO', 'default')} n{n} {cin}->{cout} {h}x{w}: {ms:8.3f} ms incl. layout passes ({fl / ms / 1e9:7.1f} TFLOP/s incl.)")
