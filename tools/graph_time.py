"""Stage-1 step under CUDA-graph replay vs eager launches (device time per step)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "afi-gan_b200"))
import torch  # noqa: E402

from afigan.engine import Stage1Step  # noqa: E402
from afigan.modeling import Discriminator, Generator  # noqa: E402
import bench as O  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
torch.manual_seed(0)
G = Generator(n_residual_dense_blocks=3, precision="bf16").cuda()
D = Discriminator(precision="bf16").cuda()
step = Stage1Step(G, D, precision="bf16")
lr_f, hr_f = O.synthetic_features(2, 0)
lr_f, hr_f = [t.cuda() for t in lr_f], [t.cuda() for t in hr_f]
for _ in range(3):
    step.run_step(lr_f, hr_f)
torch.cuda.synchronize()


def timeit(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


print(f"eager  {timeit(lambda: step.run_step(lr_f, hr_f)):.2f} ms/step")
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    step.run_step(lr_f, hr_f)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        step.run_step(lr_f, hr_f)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
print(f"graph  {timeit(g.replay):.2f} ms/step")
print("losses", [round(float(v), 5) for v in step.losses[0, :5].cpu()])
