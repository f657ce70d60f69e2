"""A/B helper: CUDA-event time of the implicit-GEMM kernel alone (afi_profile_*) for one 3x3 conv shape.
usage: ab_conv.py n cin cout h w [iters] [precision]   (AFIGAN_LIB_PATH selects a build variant)"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "afi-gan_b200"))
import torch  # noqa: E402

from afigan import native  # noqa: E402
from afigan.functional import conv3x3  # noqa: E402

n, cin, cout, h, w = map(int, sys.argv[1:6])
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 20
prec = sys.argv[7] if len(sys.argv) > 7 else "bf16"
x = torch.randn(n, cin, h, w, device="cuda")
wt = torch.randn(cout, cin, 3, 3, device="cuda") * 0.02
for _ in range(3):
    conv3x3(x, wt, None, False, prec)
torch.cuda.synchronize()
lib = native.lib()
native.check(lib.afi_profile_begin(256))
for _ in range(iters):
    conv3x3(x, wt, None, False, prec)
cnt = C.c_int()
native.check(lib.afi_profile_end(C.byref(cnt)))
kind, fl, ms = C.c_int(), C.c_double(), C.c_float()
ts = []
for i in range(cnt.value):
    lib.afi_profile_get(i, C.byref(kind), C.byref(fl), C.byref(ms), None, None, None)
    ts.append((ms.value, fl.value))
ts.sort()
med = ts[len(ts) // 2]
print(f"[{prec}] {os.path.basename(os.environ.get('AFIGAN_LIB_PATH', 'default'))} n{n} {cin}->{cout} {h}x{w}: median {med[0]:.4f} ms = {med[1] / med[0] / 1e9:7.1f} TFLOP/s, "
      f"best {ts[0][0]:.4f} ms = {ts[0][1] / ts[0][0] / 1e9:7.1f}")
