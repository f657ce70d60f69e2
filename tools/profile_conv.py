"""Runs the dominant GEMM shapes in isolation (for `ncu --set full -k regex:k_conv_tc|k_wgrad_tc`) and prints CUDA-event timings.
Shape = discriminator conv3 (1024 -> 1024, 3x3) on the p2 level of config 1: [2, 1024, 200, 336] -> M = 134400, N = 1024, K = 9216."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "afi-gan_b200"))
import torch  # noqa: E402

from afigan.functional import conv3x3, conv3x3_backward  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
shapes = [(2, 1024, 1024, 200, 336), (2, 512, 1024, 200, 336), (2, 256, 512, 200, 336), (2, 256, 256, 104, 168), (2, 256, 256, 208, 336)]
if len(sys.argv) > 3:
    shapes = shapes[:int(sys.argv[3])]
for (n, cin, cout, h, w) in shapes:
    x = torch.randn(n, cin, h, w, device="cuda")
    wt = torch.randn(cout, cin, 3, 3, device="cuda") * 0.02
    dy = torch.randn(n, cout, h, w, device="cuda")
    fl = 2.0 * n * h * w * 9 * cin * cout
    for name, fn in (("conv fwd", lambda: conv3x3(x, wt, None, False, prec)), ("wgrad+dgrad", lambda: conv3x3_backward(x, dy, wt, prec, True))):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        mult = 1 if name == "conv fwd" else 2
        print(f"[{prec}] {name:12s} n{n} {cin}->{cout} {h}x{w}: {ms:8.3f} ms incl. layout passes  ({mult * fl / ms / 1e9:7.1f} TFLOP/s)")
