import torch


def rel(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cosine(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-300))


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


# tolerances per operand mode for END-TO-END quantities: (features, logits, loss_rel, grads)
TOL = {
    # dgrad (fp32): two fp32 implementations of a conv differ by ~3e-6 in a pre-activation; that flips the LeakyReLU slope
    # of ~1e-6 of the elements and each flip moves the norm-wise D gradient error by ~sqrt(1/#elements): 1e-3 is AT the
    # fp32-vs-fp32 noise floor for the discriminator, so the gate is 5e-3 + a cosine bound.
    "fp32": dict(feat=1e-4, logit=2e-4, loss=1e-5, grad=1e-3, dgrad=5e-3, cos=0.99998),
    # split: fp32 storage, bf16x3 six-term products on the tensor cores -- gated at the fp32 row
    "split": dict(feat=1e-4, logit=2e-4, loss=1e-5, grad=1e-3, dgrad=5e-3, cos=0.99998),
    "bf16": dict(feat=1e-3, logit=3e-2, loss=5e-3, grad=0.15, dgrad=0.15, cos=0.985),
    "bf16_simt": dict(feat=1e-3, logit=3e-2, loss=5e-3, grad=0.15, dgrad=0.15, cos=0.985),
}
