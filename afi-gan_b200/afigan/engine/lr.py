"""detectron2's WarmupMultiStepLR [upstream] as a pure function, for Stage1Step (the stage-1 recipe: BASE_LR 0.001, STEPS (270000,),
GAMMA 0.1, linear warm-up over WARMUP_ITERS 1000 from WARMUP_FACTOR 0.001; configs/step1_afigan_training/*.yaml:16-20 + d2 defaults)."""
from __future__ import annotations

from bisect import bisect_right
from typing import Sequence


def warmup_multistep_lr(it: int, base_lr: float = 1e-3, steps: Sequence[int] = (270000,), gamma: float = 0.1, warmup_factor: float = 1e-3,
                        warmup_iters: int = 1000, warmup_method: str = "linear") -> float:
    if it >= warmup_iters:
        wf = 1.0
    elif warmup_method == "constant":
        wf = warmup_factor
    elif warmup_method == "linear":
        alpha = it / warmup_iters
        wf = warmup_factor * (1 - alpha) + alpha
    else:
        raise ValueError(f"Unknown warmup method: {warmup_method}")
    return base_lr * wf * gamma ** bisect_right(list(steps), it)
