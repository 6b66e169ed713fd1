"""PSNR exactly as the reference measures it (metrics/metrics.py:23-34, metrics/base.py `_normalize`): inputs in
`input_range` are mapped to [0,1], PSNR = 10 log10(1 / MSE).  Used as the fp16/bf16 acceptance gate (>= 35 dB)."""
from __future__ import annotations

import torch


def psnr(pred: torch.Tensor, target: torch.Tensor, input_range=(-1.0, 1.0)) -> float:
    lo, hi = input_range
    p = (pred.float() - lo) / (hi - lo)
    t = (target.float() - lo) / (hi - lo)
    mse = torch.mean((p - t) ** 2)
    return float(10 * torch.log10(1 / mse))
