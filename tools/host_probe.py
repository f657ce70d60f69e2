import sys, time, os
sys.path.insert(0, "."); sys.path.insert(0, "afi-gan_b200")
import torch, bench
from afigan.engine import Stage2Step
from afigan.modeling import bifpn_feature_fusion
import argparse
dev = torch.device("cuda")
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
G, D = bench._models(prec, dev)
G.deferred_weight_grads = os.environ.get('DEFER', '1') == '1'
N = 2
gen = torch.Generator().manual_seed(35)
feats = [torch.randn(N, 256, h, w, generator=gen).to(dev).requires_grad_(True) for h, w in bench.C3_LEVELS]
wts = [torch.tensor([0.7, 1.3], device=dev, requires_grad=True) for _ in range(28)]
guide = [torch.randn(N, 256, 2 * h + 1, 2 * w, generator=gen).to(dev) for h, w in bench.C3_D_SIZES]
s2 = Stage2Step(D, precision=prec, distributed=False)
def one():
    outs = []; k = 0
    for layer in range(7):
        top = feats[4]
        for l in (3, 2, 1, 0):
            top = bifpn_feature_fusion(G, feats[l], top, wts[k], swish=True); k += 1
            if layer == 6: outs.append(top)
    model = [o[:, :, :h, :w] for o, (h, w) in zip(outs[::-1], bench.C3_D_SIZES[:4])] + [feats[4][:, :, :3, :5]]
    s2.d_phase(guide, model)
    g = s2.g_losses(guide, model)
    sum(g.values()).backward()
    for t in feats + wts + list(G.parameters()): t.grad = None
WARM = int(os.environ.get('WARM', '3')); ITERS = int(os.environ.get('ITERS', '10'))
for _ in range(WARM): one()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(ITERS): one()
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"[{prec}] stage2_c3: host issue {t_issue*1e3/ITERS:.2f} ms/step, wall {t_all*1e3/ITERS:.2f} ms/step")
