"""Quick device timing of the full config-1 stage-1 step (not the bench: see bench.py)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "afi-gan_b200"))
import torch  # noqa: E402

from afigan import native  # noqa: E402
from afigan.engine import Stage1Step  # noqa: E402
from afigan.modeling import Discriminator, Generator  # noqa: E402
import bench as O  # noqa: E402  (workload definition: shapes, synthetic features, FLOP count)

precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
G = Generator(n_residual_dense_blocks=3, precision=precision).cuda()
D = Discriminator(precision=precision).cuda()
step = Stage1Step(G, D, precision=precision, overlap=os.environ.get('AFIGAN_OVERLAP', '1') == '1')
lr_f, hr_f = O.synthetic_features(2, 0)
lr_f, hr_f = [t.cuda() for t in lr_f], [t.cuda() for t in hr_f]
step.run_step(lr_f, hr_f)
torch.cuda.synchronize()
print("warm-up losses", {k: round(v, 5) for k, v in step.metrics().items() if "d_loss" in k or "g_loss" in k})
native.lib().afi_launch_count(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    step.run_step(lr_f, hr_f)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
fl = O.stage1_step_flops(2 * 23282, 2 * 89523, 1 if step.reuse_g_forward else 2)
print(f"[{precision}] {ms:.2f} ms/step  {2000.0 / ms:.2f} img/s  {fl / ms / 1e9:.1f} TFLOP/s  launches/step {native.lib().afi_launch_count(0) / steps:.0f}"
      f"  mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
