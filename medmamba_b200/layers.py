"""Small host-side layers the reference takes from timm (MedMamba.py:11), which is not a dependency here."""
from __future__ import annotations

import torch
import torch.nn as nn


class DropPath(nn.Module):
    """Per-sample stochastic depth on the residual branch (identity in eval / at rate 0)."""

    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = float(drop_prob)
        self.scale_by_keep = scale_by_keep

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self.training or self.drop_prob == 0.0:
            return x
        keep = 1.0 - self.drop_prob
        mask = torch.empty((x.shape[0],) + (1,) * (x.dim() - 1), dtype=x.dtype, device=x.device).bernoulli_(keep)
        if self.scale_by_keep and keep > 0.0:
            mask = mask / keep
        return x * mask

    def extra_repr(self) -> str:
        return f"drop_prob={self.drop_prob:.3f}"


def trunc_normal_(tensor: torch.Tensor, mean: float = 0.0, std: float = 1.0, a: float = -2.0, b: float = 2.0):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)
