"""Per-shape table of the implicit-GEMM launches of one stage-1 step (CUDA events around every launch)."""
import collections
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "afi-gan_b200"))
import torch  # noqa: E402

from afigan import native  # noqa: E402
from afigan.engine import Stage1Step  # noqa: E402
from afigan.modeling import Discriminator, Generator  # noqa: E402
import bench as O  # noqa: E402  (workload definition: shapes, synthetic features, FLOP count)

precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
torch.manual_seed(0)
G = Generator(n_residual_dense_blocks=3, precision=precision).cuda()
D = Discriminator(precision=precision).cuda()
step = Stage1Step(G, D, precision=precision, overlap=False)   # single stream: per-launch event times do not overlap
lr_f, hr_f = O.synthetic_features(2, 0)
lr_f, hr_f = [t.cuda() for t in lr_f], [t.cuda() for t in hr_f]
for _ in range(2):
    step.run_step(lr_f, hr_f)
torch.cuda.synchronize()
lib = native.lib()
native.check(lib.afi_profile_begin(4096))
step.run_step(lr_f, hr_f)
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
print(f"GEMM launches {n.value}, total {tot:.2f} ms")
print(f"{'kernel':10s} {'cin':>5s} {'cout':>5s} {'pixels':>8s} {'taps':>4s} {'n':>4s} {'ms':>8s} {'%':>6s} {'TFLOP/s':>8s}")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    print(f"{names[key[0]]:10s} {key[1]:5d} {key[2]:5d} {key[3]:8d} {key[4]:4d} {v[0]:4d} {v[2]:8.3f} {100 * v[2] / tot:6.1f} {v[1] / v[2] / 1e9:8.1f}")
