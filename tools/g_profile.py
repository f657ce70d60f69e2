"""AF interpolator alone (bench.py --workload g_only): per-shape table of its implicit-GEMM launches (CUDA events around every launch)
next to the wall time of the whole forward + backward, i.e. what the layout / elementwise passes and launch gaps add.
usage: g_profile.py [precision]"""
import collections
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "afi-gan_b200"))
import torch  # noqa: E402

import bench as B  # noqa: E402
from afigan import native  # noqa: E402
from afigan.engine import Stage1Step  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
dev = torch.device("cuda")
G, D = B._models(precision, dev)
step = Stage1Step(G, D, precision=precision, distributed=False)
lr_h, hr_h = B.synthetic_features(B.PER_GPU_BATCH, 0)
lr_d, hr_d = [t.to(dev) for t in lr_h], [t.to(dev) for t in hr_h]
trs = step._g_forward(lr_d, hr_d, True, "g")
dys = [torch.randn_like(t) / t.numel() for t in trs]
lib = native.lib()


def one():
    native.check(lib.afi_zero(step.g_acc.data_ptr(), step.g_acc.numel(), native.stream_ptr()))
    step._g_forward(lr_d, hr_d, True, "g")
    step._g_backward(lr_d, hr_d, dys, "g")


for _ in range(3):
    one()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    one()
e1.record()
torch.cuda.synchronize()
wall = e0.elapsed_time(e1) / 10
native.check(lib.afi_profile_begin(4096))
one()
n = C.c_int()
native.check(lib.afi_profile_end(C.byref(n)))
agg = collections.OrderedDict()
kind, fl, ms, cin, cout, px = C.c_int(), C.c_double(), C.c_float(), C.c_int(), C.c_int(), C.c_longlong()
for i in range(n.value):
    lib.afi_profile_get(i, C.byref(kind), C.byref(fl), C.byref(ms), C.byref(cin), C.byref(cout), C.byref(px))
    key = (kind.value, cin.value, cout.value, px.value, round(fl.value / (2.0 * px.value * cin.value * cout.value)))
    a = agg.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1; a[1] += fl.value; a[2] += ms.value
names = {0: "conv_tc", 1: "wgrad_tc", 2: "conv_simt", 3: "wgrad_simt", 4: "conv_pair", 5: "conv_halo"}
tot = sum(v[2] for v in agg.values())
totf = sum(v[1] for v in agg.values())
print(f"[{precision}] wall {wall:.3f} ms per fwd+bwd; GEMM launches {n.value}, event-bracketed total {tot:.3f} ms = {totf / tot / 1e9:.1f} TFLOP/s executed")
print(f"{'kernel':10s} {'cin':>5s} {'cout':>5s} {'pixels':>8s} {'taps':>4s} {'n':>4s} {'ms':>8s} {'%':>6s} {'TFLOP/s':>8s}")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    print(f"{names[key[0]]:10s} {key[1]:5d} {key[2]:5d} {key[3]:8d} {key[4]:4d} {v[0]:4d} {v[2]:8.3f} {100 * v[2] / tot:6.1f} {v[1] / v[2] / 1e9:8.1f}")
