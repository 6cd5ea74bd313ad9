"""Counterpart of vq_ae/layers/conv.py."""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from .. import _lib as L
from .. import engine as E


class ResizeConv2D(nn.Conv2d):
    """bicubic x2 upsample (align_corners=False) followed by the conv (vq_ae/layers/conv.py:4-11).

    Inside ``PreActFixupResBlock`` (mode 'up') the pair is executed fused by the block kernel
    sequence; this standalone ``forward`` exists for API parity and supports the shipped
    configuration (1x1 kernel, stride 1, groups 1, no bias -- up2dresize.yaml).
    """

    def __init__(self, *conv_args, **conv_kwargs):
        super().__init__(*conv_args, **conv_kwargs)
        # kept for attribute parity; has no parameters, so state_dict keys are unchanged
        self.upsample = nn.Upsample(mode="bicubic", scale_factor=2, align_corners=False)

    def forward(self, data: torch.Tensor) -> torch.Tensor:
        E.require_cuda(data, "ResizeConv2D.forward")
        if (self.kernel_size != (1, 1) or self.stride != (1, 1) or self.groups != 1
                or self.bias is not None):
            raise NotImplementedError("ResizeConv2D: only the shipped 1x1 configuration is built")
        lib = L.load()
        x, cl = E.to_nhwc(data)
        b, h, w, ci = x.shape
        co = self.out_channels
        wp = E.pack_conv_weight(self.weight)
        lo = torch.empty(b, h, w, co, dtype=torch.float32, device=x.device)
        out = torch.empty(b, 2 * h, 2 * w, co, dtype=torch.float32, device=x.device)
        st = E._stream(x.device)
        L.check(lib.vqae_conv_f32(L.CONV_1x1, E._ptr(x), E._ptr(wp), E._ptr(lo), None, b, h, w,
                                  ci, co, 0.0, 0, 0.0, 1.0, 0.0, st), "vqae_conv_f32")
        L.check(lib.vqae_bicubic_up2_f32(E._ptr(lo), E._ptr(out), b, h, w, co, 0.0, st),
                "vqae_bicubic_up2_f32")
        return E.from_nhwc(out, cl)
