"""MBConv / SELayer in eval mode over the C-ABI kernels of csrc/mbconv.cu (scope row f-4).

Duck-typed like plan.py: a block is read through what the reference's class defines -- ``branch`` (an
nn.Sequential of conv / BatchNorm2d / activation / SELayer modules, layers/conv_block.py:264-309) and
``skip_conv`` (:311-318) -- so the same code serves this package's mirrored classes and instances of the
reference's own.  BatchNorm in eval mode is a per-channel affine map; it is folded into (scale, shift)
once per parameter version.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch
from torch import Tensor, nn

from . import _lib as L
from . import engine as E

PW_DIRECT, PW_S2D, PW_CONVT = 0, 1, 2
DW_SAME, DW_DOWN, DW_UP = 0, 1, 2


def is_mbconv(m) -> bool:
    return isinstance(getattr(m, "branch", None), nn.Sequential) and hasattr(m, "skip_conv") \
        and not hasattr(m, "bias1a")


def _is_se(m) -> bool:
    return isinstance(getattr(m, "fc", None), nn.Sequential) and len(m.fc) == 4


def _fold_bn(bn: Optional[nn.BatchNorm2d], channels: int, dev) -> Tuple[Optional[Tensor], Optional[Tensor]]:
    if bn is None:
        return None, None
    if bn.running_var is None:
        raise NotImplementedError("MBConv: BatchNorm2d without running statistics has no eval-mode fold")
    inv = torch.rsqrt(bn.running_var.detach().float() + bn.eps)
    gamma = bn.weight.detach().float() if bn.weight is not None else torch.ones(channels, device=dev)
    beta = bn.bias.detach().float() if bn.bias is not None else torch.zeros(channels, device=dev)
    scale = gamma * inv
    return scale.contiguous(), (beta - bn.running_mean.detach().float() * scale).contiguous()


def version_key(block) -> Tuple:
    return tuple((t.data_ptr(), t._version) for t in list(block.parameters()) + list(block.buffers()))


class PackedMBConv:
    """Weights of one MBConv in the layouts the kernels read."""

    def __init__(self, blk):
        mods = list(blk.branch)
        convs = [m for m in mods if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d))]
        if len(convs) != 3:
            raise NotImplementedError(f"MBConv: expected 3 convs in branch, found {len(convs)}")
        for m in mods:
            if not isinstance(m, (nn.Conv2d, nn.ConvTranspose2d, nn.BatchNorm2d, nn.SiLU)) and not _is_se(m):
                raise NotImplementedError(
                    f"MBConv: {type(m).__name__} in branch has no B200 kernel (the reference's mbconv.yaml "
                    "uses SiLU, BatchNorm2d and SELayer)")
        c1, c2, c3 = convs
        dev = c1.weight.device
        E.require_cuda(c1.weight, "MBConv")

        def bn_after(conv):
            i = mods.index(conv)
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            return nxt if isinstance(nxt, nn.BatchNorm2d) else None

        def act_follows(conv):
            i = mods.index(conv) + (2 if bn_after(conv) is not None else 1)
            return i < len(mods) and isinstance(mods[i], nn.SiLU)

        if not (act_follows(c1) and act_follows(c2)) or act_follows(c3):
            raise NotImplementedError("MBConv: activation placement differs from conv_block.py:264-309")
        for conv in (c1, c3):
            if not isinstance(conv, nn.Conv2d) or conv.kernel_size != (1, 1) or conv.groups != 1 \
                    or conv.stride != (1, 1):
                raise NotImplementedError("MBConv: branch_conv1 / branch_conv3 must be 1x1 convs")
        self.c_in, self.c_mid, self.c_out = c1.in_channels, c1.out_channels, c3.out_channels
        if (self.c_in | self.c_mid | self.c_out) & 3:
            raise NotImplementedError("MBConv: channel counts must be multiples of 4")

        def with_bias(conv, scale, shift, n):
            """A conv bias in front of the affine map: scale * (y + b) + shift."""
            if conv.bias is None:
                return scale, shift
            b = conv.bias.detach().float()
            if scale is None:
                return None, b.contiguous()
            return scale, (shift + scale * b).contiguous()

        self.w1 = c1.weight.detach().float().reshape(self.c_mid, self.c_in).contiguous()
        self.s1, self.h1 = with_bias(c1, *_fold_bn(bn_after(c1), self.c_mid, dev), self.c_mid)
        # depthwise stage
        k, s = c2.kernel_size, c2.stride
        if c2.groups != self.c_mid or c2.in_channels != self.c_mid or c2.out_channels != self.c_mid:
            raise NotImplementedError("MBConv: branch_conv2 must be depthwise (groups = channels)")
        if isinstance(c2, nn.ConvTranspose2d):
            if k != (2, 2) or s != (2, 2) or c2.padding != (0, 0) or c2.output_padding != (0, 0):
                raise NotImplementedError("MBConv 'up': only the 2x2 stride-2 transposed conv of up2d.yaml")
            self.dw_mode = DW_UP
        elif k == (3, 3) and s == (1, 1):
            if c2.padding_mode != "circular" or c2.padding != (1, 1):
                raise NotImplementedError("MBConv 'same': only the circular 3x3 of mbconv.yaml")
            self.dw_mode = DW_SAME
        elif k == (2, 2) and s == (2, 2) and c2.padding == (0, 0):
            self.dw_mode = DW_DOWN
        else:
            raise NotImplementedError(f"MBConv: depthwise conv {k} stride {s} is not built")
        taps = k[0] * k[1]
        self.w2 = c2.weight.detach().float().reshape(self.c_mid, taps).t().contiguous()     # [taps][C]
        self.s2, self.h2 = with_bias(c2, *_fold_bn(bn_after(c2), self.c_mid, dev), self.c_mid)
        # squeeze-excite
        se = next((m for m in mods if _is_se(m)), None)
        self.se = None
        if se is not None:
            f1, a1, f2, a2 = se.fc
            if not (isinstance(f1, nn.Linear) and isinstance(a1, nn.SiLU) and isinstance(f2, nn.Linear)
                    and isinstance(a2, nn.Sigmoid)) or f2.out_features != self.c_mid:
                raise NotImplementedError("SELayer: expected Linear-SiLU-Linear-Sigmoid (misc.py:16-21)")
            self.se = tuple(t.detach().float().contiguous() for t in (f1.weight, f1.bias, f2.weight, f2.bias))
            self.c_hidden = f1.out_features
        self.w3 = c3.weight.detach().float().reshape(self.c_out, self.c_mid).contiguous()
        self.s3, self.h3 = with_bias(c3, *_fold_bn(bn_after(c3), self.c_out, dev), self.c_out)
        # skip path
        sk = blk.skip_conv
        self.skip_mode = None
        if sk is not None:
            if sk.groups != 1:
                raise NotImplementedError("MBConv: grouped skip conv")
            if isinstance(sk, nn.ConvTranspose2d) and sk.kernel_size == (2, 2) and sk.stride == (2, 2):
                self.skip_mode = PW_CONVT                       # weight [C_in, C_out, 2, 2]
                self.w_skip = sk.weight.detach().float().permute(2, 3, 1, 0).reshape(
                    4, self.c_out, self.c_in).contiguous()
            elif isinstance(sk, nn.Conv2d) and sk.kernel_size == (2, 2) and sk.stride == (2, 2):
                self.skip_mode = PW_S2D                         # k = (dy * 2 + dx) * C_in + ci
                self.w_skip = sk.weight.detach().float().permute(0, 2, 3, 1).reshape(
                    self.c_out, 4 * self.c_in).contiguous()
            elif isinstance(sk, nn.Conv2d) and sk.kernel_size == (1, 1) and sk.stride == (1, 1):
                self.skip_mode = PW_DIRECT
                self.w_skip = sk.weight.detach().float().reshape(self.c_out, self.c_in).contiguous()
            else:
                raise NotImplementedError(f"MBConv: skip conv {type(sk).__name__} {sk.kernel_size}")
            self.b_skip = sk.bias.detach().float().contiguous() if sk.bias is not None else None
        elif self.c_in != self.c_out or self.dw_mode != DW_SAME:
            raise NotImplementedError("MBConv: identity skip needs equal shapes")


def pointwise_conv(a: Tensor, w: Tensor, n_out: int, mode: int = PW_DIRECT, scale=None, shift=None,
                   gate=None, res=None, act: bool = False) -> Tensor:
    b, h, wd, c = a.shape
    ho, wo = (h // 2, wd // 2) if mode == PW_S2D else ((2 * h, 2 * wd) if mode == PW_CONVT else (h, wd))
    out = torch.empty(b, ho, wo, n_out, dtype=torch.float32, device=a.device)
    L.check(L.load().vqae_pointwise_conv_f32(
        E._ptr(a), E._ptr(w), E._ptr(scale), E._ptr(shift), E._ptr(gate), E._ptr(res), E._ptr(out), b, h,
        wd, c, n_out, mode, int(act), E._stream(a.device)), "vqae_pointwise_conv_f32")
    return out


def depthwise_conv(a: Tensor, w_taps: Tensor, mode: int, scale=None, shift=None, want_row_sums: bool = True
                   ) -> Tuple[Tensor, Optional[Tensor]]:
    b, h, wd, c = a.shape
    ho, wo = (h // 2, wd // 2) if mode == DW_DOWN else ((2 * h, 2 * wd) if mode == DW_UP else (h, wd))
    out = torch.empty(b, ho, wo, c, dtype=torch.float32, device=a.device)
    rows = torch.empty(b, ho, c, dtype=torch.float32, device=a.device) if want_row_sums else None
    L.check(L.load().vqae_depthwise_conv_f32(
        E._ptr(a), E._ptr(w_taps), E._ptr(scale), E._ptr(shift), E._ptr(out), E._ptr(rows), b, h, wd, c,
        mode, E._stream(a.device)), "vqae_depthwise_conv_f32")
    return out, rows


def se_gate(row_sums: Tensor, pixels_per_image: int, se, c_hidden: int) -> Tensor:
    b, rows, c = row_sums.shape
    gate = torch.empty(b, c, dtype=torch.float32, device=row_sums.device)
    w1, b1, w2, b2 = se
    L.check(L.load().vqae_se_gate_f32(E._ptr(row_sums), b, rows, pixels_per_image, c, E._ptr(w1), E._ptr(b1),
                                      E._ptr(w2), E._ptr(b2), c_hidden, E._ptr(gate),
                                      E._stream(row_sums.device)), "vqae_se_gate_f32")
    return gate


def forward_nhwc(pk: PackedMBConv, x: Tensor) -> Tensor:
    """MBConv.forward (conv_block.py:315-321) on a contiguous NHWC fp32 tensor."""
    if x.shape[-1] != pk.c_in:
        raise ValueError(f"MBConv: input has {x.shape[-1]} channels, block expects {pk.c_in}")
    x = x.float().contiguous()
    t1 = pointwise_conv(x, pk.w1, pk.c_mid, PW_DIRECT, pk.s1, pk.h1, act=True)
    t2, rows = depthwise_conv(t1, pk.w2, pk.dw_mode, pk.s2, pk.h2, want_row_sums=pk.se is not None)
    gate = se_gate(rows, t2.shape[1] * t2.shape[2], pk.se, pk.c_hidden) if pk.se is not None else None
    skip = x if pk.skip_mode is None else pointwise_conv(x, pk.w_skip, pk.c_out, pk.skip_mode,
                                                         shift=pk.b_skip)
    return pointwise_conv(t2, pk.w3, pk.c_out, PW_DIRECT, pk.s3, pk.h3, gate=gate, res=skip)


def packed(block) -> PackedMBConv:
    st = block.__dict__.setdefault("_b200_mb", {})
    key = version_key(block)
    if st.get("key") != key:
        st["packed"], st["key"] = PackedMBConv(block), key
    return st["packed"]


def block_forward(block, inp: Tensor) -> Tensor:
    """One MBConv on an NCHW / channels_last tensor."""
    if block.training:
        raise RuntimeError("MBConv: training-mode forward (batch statistics) is outside the B200 "
                           "inference path; call .eval()")
    E.require_cuda(inp, "MBConv.forward")
    x, cl = E.to_nhwc(inp)
    return E.from_nhwc(forward_nhwc(packed(block), x), cl)
