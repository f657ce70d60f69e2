"""A/B helper: CUDA-event time of the weight-gradient kernel alone for one 3x3 conv shape.  usage: ab_wgrad.py n cin cout h w [iters] [precision]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "afi-gan_b200"))
import torch  # noqa: E402

from afigan import native  # noqa: E402
from afigan.functional import conv3x3_backward  # noqa: E402

n, cin, cout, h, w = map(int, sys.argv[1:6])
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 20
prec = sys.argv[7] if len(sys.argv) > 7 else "bf16"
x = torch.randn(n, cin, h, w, device="cuda")
dy = torch.randn(n, cout, h, w, device="cuda")
wt = torch.randn(cout, cin, 3, 3, device="cuda") * 0.02
for _ in range(3):
    conv3x3_backward(x, dy, wt, prec, need_dx=False)
torch.cuda.synchronize()
lib = native.lib()
native.check(lib.afi_profile_begin(256))
for _ in range(iters):
    conv3x3_backward(x, dy, wt, prec, need_dx=False)
cnt = C.c_int()
native.check(lib.afi_profile_end(C.byref(cnt)))
kind, fl, ms = C.c_int(), C.c_double(), C.c_float()
ts = []
for i in range(cnt.value):
    lib.afi_profile_get(i, C.byref(kind), C.byref(fl), C.byref(ms), None, None, None)
    if kind.value == 1:
        ts.append((ms.value, fl.value))
ts.sort()
med = ts[len(ts) // 2]
print(f"[{prec}] wgrad n{n} {cin}->{cout} {h}x{w}: median {med[0]:.4f} ms = {med[1] / med[0] / 1e9:7.1f} TFLOP/s")
