"""Multi-GPU check of the data-parallel stage-1 step (run under torchrun, one rank per GPU):
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 tools/ddp_check.py
(1) replicas built from DIFFERENT seeds are identical after Stage1Step's constructor (rank 0's parameters are broadcast, DDP semantics);
(2) after two steps on different per-rank batches every rank holds the same parameters;
(3) the overlapped issue order (all-reduces hidden behind independent compute) gives the same parameters as the literal order."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "afi-gan_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench as B  # noqa: E402
from afigan.engine import Stage1Step  # noqa: E402
from afigan.modeling import Discriminator, Generator  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
shapes_lr, shapes_hr = ((26, 42), (13, 21), (7, 11)), ((50, 84), (25, 42), (13, 21))
lr_f, hr_f = B.synthetic_features(2, rank, shapes_lr, shapes_hr)
lr_d, hr_d = [t.to(dev) for t in lr_f], [t.to(dev) for t in hr_f]


def flat(step):
    return torch.cat([p.detach().reshape(-1) for p in step.g_params + step.d_params])


results = {}
for overlap_comm in (True, False):
    torch.manual_seed(1000 + rank)                      # every rank its own initial weights, like detectron2
    G = Generator(n_residual_dense_blocks=3, precision=precision).to(dev)
    D = Discriminator(precision=precision).to(dev)
    step = Stage1Step(G, D, lr=1e-2, precision=precision, overlap_comm=overlap_comm)
    p0 = flat(step)
    ref = p0.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(p0, ref), "parameters differ across ranks after construction"
    for _ in range(2):
        step.run_step(lr_d, hr_d)
    torch.cuda.synchronize()
    p = flat(step)
    ref = p.clone()
    dist.broadcast(ref, 0)
    err = float((p - ref).norm() / ref.norm())
    assert err < 1e-6, f"replicas diverged: {err}"
    assert float((p - p0).norm()) > 0
    results[overlap_comm] = p
d = float((results[True] - results[False]).norm() / results[False].norm())
assert d < (5e-3 if precision == "bf16" else 1e-4), f"overlapped vs literal issue order: {d}"
if rank == 0:
    print(f"ddp_check[{precision}] world {world}: broadcast OK, replicas identical after 2 steps, overlapped vs literal order rel diff {d:.2e}")
dist.destroy_process_group()
