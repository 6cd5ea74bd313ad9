"""Counterpart of vq_ae/layers/conv_block.py: DownBlock, UpBlock, EnvelopBlock and
PreActFixupResBlock with the reference's constructor signatures, parameter names and
initialisation; eval-mode ``forward`` of the residual block runs the fused kernels.  ``MBConv`` (the
alternative EfficientNetV2 block of conf/model/encoder/efficientnetv2.yaml, scope row f-4) runs the fp32
kernels of csrc/mbconv.cu.
"""
from __future__ import annotations

from collections.abc import Iterable, Sequence
from functools import partial
from math import isclose
from typing import Optional

import numpy as np
import torch
from torch import nn

from .. import engine as E
from .. import mbconv as M
from .. import plan as P
from .._instantiate import instantiate


class DownBlock(nn.Module):
    """n_down x [pre 'same' blocks, 'down' block, post 'same' blocks] (conv_block.py:18-52)."""
    out_channels: int

    def __init__(self, in_channels: int, n_down: int, conv_conf, n_pre_layers: Optional[int],
                 n_post_layers: Optional[int]):
        super().__init__()
        pre_layers, post_layers = [
            [{**conv_conf, **{'mode': 'same'}}] * n_layers
            for n_layers in (n_pre_layers, n_post_layers)
        ]
        self.layers = nn.Sequential(*(
            EnvelopBlock(envelop_conf={**conv_conf, **{'mode': 'down'}}, in_channels=in_c,
                         out_channels=out_c, pre_layers=pre_layers, post_layers=post_layers)
            for in_c, out_c in ((in_channels * (2 ** j), in_channels * (2 ** (j + 1)))
                                for j in range(n_down))
        ))
        self.out_channels = in_channels * 2 ** n_down

    def forward(self, x):
        return self.layers(x)


class UpBlock(nn.Module):
    """DownBlock in reverse (conv_block.py:55-91)."""
    in_channels: int

    def __init__(self, out_channels: int, n_up: int, conv_conf, n_pre_layers: Optional[int],
                 n_post_layers: Optional[int]):
        super().__init__()
        pre_layers, post_layers = [
            [{**conv_conf, **{'mode': 'same'}}] * n_layers
            for n_layers in (n_pre_layers, n_post_layers)
        ]
        self.layers = nn.Sequential(*(
            EnvelopBlock(envelop_conf={**conv_conf, **{'mode': 'up'}}, in_channels=in_c,
                         out_channels=out_c, pre_layers=pre_layers, post_layers=post_layers)
            for in_c, out_c in ((out_channels * (2 ** (j + 1)), out_channels * (2 ** j))
                                for j in range(n_up - 1, -1, -1))
        ))
        self.in_channels = out_channels * 2 ** n_up

    def forward(self, x):
        return self.layers(x)


class EnvelopBlock(nn.Module):
    """pre layers -> envelop (down/up) layer -> post layers (conv_block.py:94-129)."""

    def __init__(self, envelop_conf, in_channels: int, out_channels: int, pre_layers=None,
                 post_layers=None):
        super().__init__()

        def instantiate_layers(layers, in_channels: int, out_channels: int) -> Iterable:
            if layers is None:
                return ()
            if isinstance(layers, Sequence) and not isinstance(layers, (str, dict)):
                if len(layers) == 2 and isinstance(layers[1], int):
                    layers = [layers[0]] * layers[1]
            else:
                layers = (layers,)
            return map(partial(instantiate, in_channels=in_channels, out_channels=out_channels),
                       filter(None, layers))

        self.layers = nn.Sequential(
            *instantiate_layers(pre_layers, in_channels, in_channels),
            instantiate(envelop_conf, in_channels=in_channels, out_channels=out_channels),
            *instantiate_layers(post_layers, out_channels, out_channels),
        )

    def forward(self, x):
        return self.layers(x)


class PreActFixupResBlock(nn.Module):
    """Pre-activation Fixup residual block (conv_block.py:132-237).

    forward (conv_block.py:196-216):
        out = conv1(act(x + bias1a) + bias1b)
        out = conv2(act(out + bias2a) + bias2b)
        out = conv3(act(out + bias3a) + bias3b)
        out = out * scale + bias4
        out = out + (skip_conv(x + bias1c) + bias1d  if skip_conv else  x)
    """

    def __init__(self, in_channels: int, out_channels: int, mode: str, bottleneck_divisor: float,
                 activation, conv_conf, n_layers: Optional[int] = None):
        super().__init__()
        assert mode in ("down", "same", "up", "out")
        self.mode = mode
        conv_conf = conv_conf[mode]

        max_channels = max(in_channels, out_channels)
        assert isclose(max_channels % bottleneck_divisor, 0), (
            f"residual channels: {max_channels} not divisible by bottleneck divisor: "
            f"{bottleneck_divisor}!")
        branch_channels = max(round(max_channels / bottleneck_divisor), 1)

        self.activation = instantiate(activation)
        (self.bias1a, self.bias1b, self.bias2a, self.bias2b, self.bias3a, self.bias3b,
         self.bias4) = (nn.Parameter(torch.zeros(1)) for _ in range(7))
        self.scale = nn.Parameter(torch.ones(1))

        self.branch_conv1 = instantiate(conv_conf['branch_conv1'], in_channels=in_channels,
                                        out_channels=branch_channels)
        self.branch_conv2 = instantiate(conv_conf['branch_conv2'], in_channels=branch_channels,
                                        out_channels=branch_channels)
        self.branch_conv3 = instantiate(conv_conf['branch_conv3'], in_channels=branch_channels,
                                        out_channels=out_channels)
        if not (mode in ("same", "out") and in_channels == out_channels):
            self.bias1c, self.bias1d = (nn.Parameter(torch.zeros(1)) for _ in range(2))
            self.skip_conv = instantiate(conv_conf['skip_conv'], in_channels=in_channels,
                                         out_channels=out_channels)
        else:
            self.skip_conv = None

        if n_layers is not None:
            self.initialize_weights(n_layers)

    # -- B200 path -------------------------------------------------------------------------
    def check_supported(self) -> None:
        E.check_block_supported(self)

    def packed(self) -> E.PackedFixup:
        st = P.state(self)
        key = E.block_version(self)
        if st.packed is None or key != st.packed_key:
            self.check_supported()
            st.packed = E.pack_blocks([self])[0]
            st.packed_key = key
        return st.packed

    def forward(self, inp: torch.Tensor) -> torch.Tensor:
        return P.block_forward(self, inp)

    @torch.no_grad()
    def initialize_weights(self, num_layers):
        """Fixup initialisation (conv_block.py:218-237)."""
        weight = self.branch_conv1.weight
        nn.init.normal_(weight, mean=0,
                        std=np.sqrt(2 / (weight.shape[0] * np.prod(weight.shape[2:])))
                        * num_layers ** (-0.5))
        nn.init.kaiming_normal_(self.branch_conv2.weight)
        nn.init.constant_(self.branch_conv3.weight, val=0)
        if self.skip_conv is not None:
            nn.init.xavier_normal_(self.skip_conv.weight)


class MBConv(nn.Module):
    """expand 1x1 -> depthwise -> squeeze-excite -> project 1x1, BatchNorm after every conv, plus a skip
    path (conv_block.py:240-321).  ``branch`` holds the same modules at the same indices as the
    reference's (absent BatchNorm / SELayer configs leave no gap), which is the ``state_dict`` contract."""

    def __init__(self, in_channels: int, out_channels: int, mode: str, expand_ratio: float,
                 activation_conf, conv_conf, batchnorm_conf, se_conf):
        super().__init__()
        assert mode in ("down", "same", "up", "out")
        conv_conf = conv_conf[mode]
        widest = max(in_channels, out_channels)
        assert isclose(widest * expand_ratio % 1, 0), (
            f"max_channels: {widest} x expand_ratio: {expand_ratio} % 1 !\u2248 0!")
        mid = round(widest * expand_ratio)
        stages = (
            (conv_conf['branch_conv1'], dict(in_channels=in_channels, out_channels=mid)),
            (batchnorm_conf, dict(num_features=mid)),
            (activation_conf, {}),
            (conv_conf['branch_conv2'], dict(in_channels=mid, out_channels=mid, groups=mid)),
            (batchnorm_conf, dict(num_features=mid)),
            (activation_conf, {}),
            (se_conf, dict(in_channels=mid, out_channels=mid)),
            (conv_conf['branch_conv3'], dict(in_channels=mid, out_channels=out_channels)),
            (batchnorm_conf, dict(num_features=out_channels)),
        )
        built = (instantiate(conf, **kw) for conf, kw in stages)
        self.branch = nn.Sequential(*(m for m in built if m is not None))
        needs_skip_conv = not (mode in ("same", "out") and in_channels == out_channels)
        self.skip_conv = (instantiate(conv_conf['skip_conv'], in_channels=in_channels,
                                      out_channels=out_channels) if needs_skip_conv else None)
        with torch.no_grad():
            self.branch[-1].weight *= 0            # last BatchNorm's gamma starts at zero (:312-313)

    def forward(self, x):
        return M.block_forward(self, x)
