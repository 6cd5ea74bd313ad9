"""Counterpart of vq_ae/layers/misc.py: ``SELayer`` (squeeze-excite) with the reference's constructor
signature and parameter names.  Inside an MBConv its arithmetic is part of the block's kernels
(csrc/mbconv.cu)."""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn


def make_divisible(value: float, divisor: int, divide: bool = True, min_value: Optional[int] = None):
    """utils/train_helpers.py:11-24."""
    floor = divisor if min_value is None else min_value
    return max(floor, int(value + divisor / 2)) // (divisor if divide else 1)


class SELayer(nn.Module):
    """x * sigmoid(FC(SiLU(FC(mean_hw(x)))))  (layers/misc.py:7-30)."""

    def __init__(self, in_channels: int, out_channels: int, bottleneck_divisor: int):
        super().__init__()
        hidden = make_divisible(in_channels, bottleneck_divisor, divide=True)
        self.fc = nn.Sequential(nn.Linear(in_channels, hidden), nn.SiLU(),
                                nn.Linear(hidden, out_channels), nn.Sigmoid())

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        # the reference builds SELayers only inside MBConv branches (conv_block.py:296-300), where the
        # squeeze is fused into the depthwise kernel and the gate into the projection GEMM
        raise NotImplementedError("SELayer: built as part of MBConv (csrc/mbconv.cu); a standalone "
                                  "squeeze-excite layer has no B200 kernel")
