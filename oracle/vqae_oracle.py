"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's inference hot path.

This is the *oracle* ("port" kind): a module-free, functional restatement, on torch CPU
tensors, of what the reference's nn.Modules compute on the encode -> quantise -> decode
path.  It takes a reference-layout ``state_dict`` and walks it explicitly.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import it -- and only as the checker / the timed CPU baseline,
never on the product path.

Parity pinning: the reference ships no tests or golden vectors, so
this file is pinned against the reference *itself*: ``tests/test_oracle_vs_reference.py``
(container only, through ``oracle/ref_shim.py``) runs the unmodified reference modules
and this restatement on the same weights and inputs, and ``oracle/make_golden.py`` writes
the reference's outputs to ``tests/golden/`` so the pin travels to the GPU box.

All ``file:line`` citations are relative to the reference checkout.  Third-party
arithmetic: every FLOP of the reference runs inside PyTorch (pinned torch==1.11.0+cu115,
pyproject.toml:9); here it runs on torch 2.11 CPU kernels.  Input normalisation is
albumentations 1.1.0 ``Normalize`` (pdm.lock:35-36), restated from its published formula.
"""
from __future__ import annotations

import re
from typing import Optional, Dict, List, Mapping, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Mapping[str, Tensor]

# conf/transforms/camelyon16_transforms.yaml:15-23
CAMELYON16_MEAN = (0.7279, 0.5955, 0.7762)
CAMELYON16_STD = (0.2419, 0.3083, 0.1741)


# ------------------------------------------------------------------------------------
# a-N  input normalisation (albumentations.Normalize + ToTensorV2)
# ------------------------------------------------------------------------------------
def normalize_u8(img_hwc_u8: np.ndarray, mean=CAMELYON16_MEAN, std=CAMELYON16_STD,
                 max_pixel_value: float = 255.0) -> np.ndarray:
    """``(u8 - 255*mean) * (1/(255*std))`` in fp32, then HWC -> CHW.

    Follows albumentations 1.1.0 ``augmentations/functional.py::normalize`` (numpy branch)
    and ``pytorch/transforms.py::ToTensorV2`` as configured by
    conf/transforms/camelyon16_transforms.yaml:1-23 and transforms/normalize.yaml:1-11.
    Accepts [H,W,3] or [B,H,W,3]; returns [3,H,W] / [B,3,H,W] float32.
    """
    mean_a = np.array(mean, dtype=np.float32) * np.float32(max_pixel_value)
    std_a = np.array(std, dtype=np.float32) * np.float32(max_pixel_value)
    denom = np.reciprocal(std_a, dtype=np.float32)
    img = img_hwc_u8.astype(np.float32)
    img -= mean_a
    img *= denom
    return np.moveaxis(img, -1, -3).copy()


# ------------------------------------------------------------------------------------
# building blocks
# ------------------------------------------------------------------------------------
def elu(x: Tensor) -> Tensor:
    """nn.ELU(alpha=1) (conf/model/layers/activation/elu.yaml)."""
    return torch.where(x > 0, x, torch.expm1(x))


def conv2d_circular(x: Tensor, w: Tensor, stride: int = 1, padding: int = 0) -> Tensor:
    """nn.Conv2d(padding_mode='circular') = circular F.pad then a valid conv
    (pre_activation_fixup.yaml:40,60)."""
    if padding:
        x = F.pad(x, (padding,) * 4, mode="circular")
    return F.conv2d(x, w, None, stride=stride)


_A = -0.75  # torch's cubic convolution coefficient


def _cubic_weights(t: float) -> Tuple[float, float, float, float]:
    def c1(x):  # |x| <= 1
        return ((_A + 2.0) * x - (_A + 3.0)) * x * x + 1.0

    def c2(x):  # 1 < |x| < 2
        return ((_A * x - 5.0 * _A) * x + 8.0 * _A) * x - 4.0 * _A

    return c2(t + 1.0), c1(t), c1(1.0 - t), c2(2.0 - t)


def bicubic_up2(x: Tensor) -> Tensor:
    """nn.Upsample(mode='bicubic', scale_factor=2, align_corners=False) (layers/conv.py:8).

    Explicit 4x4-tap restatement: output o reads source coordinate (o+0.5)/2-0.5, taps
    floor-1..floor+2 with *clamped* indices, Keys weights A=-0.75.  For scale 2 the phase
    is 0.75 (even o) or 0.25 (odd o): weights (-9,67,225,-27)/256 and (-27,225,67,-9)/256.
    """
    B, C, H, W = x.shape

    def taps(n: int):
        o = torch.arange(2 * n, dtype=torch.float64)
        src = (o + 0.5) / 2.0 - 0.5
        fl = torch.floor(src)
        t = (src - fl)
        idx = torch.stack([(fl.long() + k).clamp(0, n - 1) for k in (-1, 0, 1, 2)], 0)  # [4, 2n]
        w = torch.empty(4, 2 * n, dtype=torch.float64)
        for j in range(2 * n):
            w[:, j] = torch.tensor(_cubic_weights(float(t[j])), dtype=torch.float64)
        return idx, w.to(x.dtype)

    iy, wy = taps(H)
    ix, wx = taps(W)
    # interpolate along x for every source row, then along y (torch's loop order)
    rows = sum(x[:, :, :, ix[k]] * wx[k] for k in range(4))          # [B,C,H,2W]
    out = sum(rows[:, :, iy[k], :] * wy[k][:, None] for k in range(4))  # [B,C,2H,2W]
    return out


def _blk(sd: StateDict, prefix: str) -> Dict[str, Tensor]:
    plen = len(prefix)
    return {k[plen:]: v for k, v in sd.items() if k.startswith(prefix)}


def fixup_block(x: Tensor, p: Mapping[str, Tensor]) -> Tensor:
    """PreActFixupResBlock.forward (layers/conv_block.py:196-216); mode is inferred from
    the weight shapes: branch_conv2 3x3 -> 'same' (circular pad 1), 2x2 -> 'down'
    (stride 2), 1x1 -> 'up' (bicubic x2 then 1x1, layers/conv.py:10-11)."""
    w1, w2, w3 = p["branch_conv1.weight"], p["branch_conv2.weight"], p["branch_conv3.weight"]
    k2 = w2.shape[-1]
    out = elu(x + p["bias1a"])
    out = F.conv2d(out + p["bias1b"], w1)
    out = elu(out + p["bias2a"])
    out = out + p["bias2b"]
    if k2 == 3:
        out = conv2d_circular(out, w2, stride=1, padding=1)
    elif k2 == 2:
        out = F.conv2d(out, w2, stride=2)
    else:
        out = F.conv2d(bicubic_up2(out), w2)
    out = elu(out + p["bias3a"])
    out = F.conv2d(out + p["bias3b"], w3)
    out = out * p["scale"] + p["bias4"]
    if "skip_conv.weight" in p:
        ws = p["skip_conv.weight"]
        s = x + p["bias1c"]
        if ws.shape[-1] == 2:
            s = F.conv2d(s, ws, stride=2)
        elif k2 == 1:
            s = F.conv2d(bicubic_up2(s), ws)
        else:
            s = F.conv2d(s, ws)
        return out + (s + p["bias1d"])
    return out + x


def mbconv_block(x: Tensor, p: Mapping[str, Tensor], mode: Optional[str] = None) -> Tensor:
    """MBConv.forward in eval mode (layers/conv_block.py:240-321) with SELayer (layers/misc.py:7-30):
    ``branch`` = conv1x1 [BN] SiLU depthwise [BN] SiLU [SE] conv1x1 [BN] at consecutive indices, absent
    stages leaving no gap.  A state_dict cannot tell a depthwise 2x2 conv from a depthwise transposed 2x2
    conv (same shapes), so the mode follows the channel change like the reference's pyramids do
    (DownBlock doubles, UpBlock halves) unless given."""
    stages = []
    for i in range(9):                                                   # at most nine stages (:264-309)
        if f"branch.{i}.running_mean" in p:
            stages.append(("bn", i))
        elif f"branch.{i}.fc.0.weight" in p:
            stages.append(("se", i))
        elif f"branch.{i}.weight" in p:
            stages.append(("conv", i))                                   # activations hold no tensors
    convs = [j for kind, j in stages if kind == "conv"]
    assert len(convs) == 3
    c_in, c_out = p[f"branch.{convs[0]}.weight"].shape[1], p[f"branch.{convs[2]}.weight"].shape[0]
    k2 = p[f"branch.{convs[1]}.weight"].shape[-1]
    if mode is None:
        mode = "same" if k2 == 3 else ("down" if c_out > c_in else "up")
        assert k2 == 3 or c_in != c_out, "2x2 depthwise with equal widths: pass mode="

    def bn(h, j):
        return F.batch_norm(h, p[f"branch.{j}.running_mean"], p[f"branch.{j}.running_var"],
                            p.get(f"branch.{j}.weight"), p.get(f"branch.{j}.bias"), False, 0.0, 1e-5)

    h, n_conv = x, 0
    for pos, (kind, j) in enumerate(stages):
        if kind == "conv":
            w = p[f"branch.{j}.weight"]
            b = p.get(f"branch.{j}.bias")
            if n_conv == 1:                                              # depthwise stage
                c = w.shape[0]
                if mode == "same":
                    h = F.conv2d(F.pad(h, (1, 1, 1, 1), mode="circular"), w, b, groups=c)
                elif mode == "down":
                    h = F.conv2d(h, w, b, stride=2, groups=c)
                else:
                    h = F.conv_transpose2d(h, w, b, stride=2, groups=c)
            else:
                h = F.conv2d(h, w, b)
            n_conv += 1
            nxt = stages[pos + 1][0] if pos + 1 < len(stages) else None
            if nxt != "bn" and n_conv < 3:
                h = F.silu(h)
        elif kind == "bn":
            h = bn(h, j)
            if n_conv < 3:
                h = F.silu(h)
        else:                                                            # squeeze-excite
            y = h.mean(dim=(2, 3))
            y = F.silu(F.linear(y, p[f"branch.{j}.fc.0.weight"], p[f"branch.{j}.fc.0.bias"]))
            y = torch.sigmoid(F.linear(y, p[f"branch.{j}.fc.2.weight"], p[f"branch.{j}.fc.2.bias"]))
            h = h * y[:, :, None, None]
    if "skip_conv.weight" in p:
        ws, bs = p["skip_conv.weight"], p.get("skip_conv.bias")
        if ws.shape[-1] == 1:
            s = F.conv2d(x, ws, bs)
        elif mode == "down":
            s = F.conv2d(x, ws, bs, stride=2)
        else:
            s = F.conv_transpose2d(x, ws, bs, stride=2)
        return h + s
    return h + x


def any_block(x: Tensor, p: Mapping[str, Tensor]) -> Tensor:
    return mbconv_block(x, p) if "branch.0.weight" in p else fixup_block(x, p)


def _indexed_children(sd: StateDict, prefix: str) -> List[str]:
    pat = re.compile(re.escape(prefix) + r"(\d+)\.")
    idx = sorted({int(m.group(1)) for k in sd for m in [pat.match(k)] if m})
    return [f"{prefix}{i}." for i in idx]


def block_sequence(x: Tensor, sd: StateDict, prefix: str) -> Tensor:
    """nn.Sequential of PreActFixupResBlocks stored under ``prefix{i}.``."""
    for child in _indexed_children(sd, prefix):
        x = any_block(x, _blk(sd, child))
    return x


def envelop_pyramid(x: Tensor, sd: StateDict, prefix: str) -> Tensor:
    """DownBlock / UpBlock: ``prefix{j}.layers.{i}.`` (layers/conv_block.py:18-91,94-129)."""
    for level in _indexed_children(sd, prefix):
        x = block_sequence(x, sd, level + "layers.")
    return x


# ------------------------------------------------------------------------------------
# a-Q / a-P  quantiser
# ------------------------------------------------------------------------------------
NEAR_TIE_REL_GAP = 16.0 * 2.0 ** -23  # ~1.9e-6, SURVEY.md section 7 hard part 2


def l4_distances_unrooted(flat: Tensor, embed: Tensor) -> Tensor:
    """sum_d |z_d - e_kd|^4 accumulated in d order in fp32 (the un-rooted cdist(p=4))."""
    agg = torch.zeros(flat.shape[0], embed.shape[0], dtype=flat.dtype)
    for d in range(flat.shape[1]):
        diff = flat[:, d:d + 1] - embed[None, :, d]
        sq = diff * diff
        agg = agg + sq * sq
    return agg


def quantize_flat(flat: Tensor, embed: Tensor, chunk: int = 32768
                  ) -> Tuple[Tensor, Tensor, Tensor]:
    """argmin_k cdist_{p=4}(flat, embed) (layers/vq.py:121-129; p = inputs.dim() = 4,
    vq.py:97), first index wins ties (torch.argmin).  Returns (idx int64 [N],
    rel_gap fp32 [N] = (d2-d1)/d2 of the un-rooted sums, d1 [N])."""
    idxs, gaps, d1s = [], [], []
    for s in range(0, flat.shape[0], chunk):
        agg = l4_distances_unrooted(flat[s:s + chunk].float(), embed.float())
        # the reference compares rooted values; x -> x^(1/4) is monotone, so the argmin is
        # the same except where rooting merges values closer than ~4 ulp (reported as
        # near ties through rel_gap).
        top2 = torch.topk(agg, 2, dim=1, largest=False)
        idx = torch.argmin(torch.pow(agg, 0.25), dim=1)
        d1 = agg.gather(1, idx[:, None])[:, 0]
        other = torch.where(top2.indices[:, 0] == idx, top2.values[:, 1], top2.values[:, 0])
        gaps.append((other - d1).abs() / other.clamp_min(1e-30))
        idxs.append(idx)
        d1s.append(d1)
    return torch.cat(idxs), torch.cat(gaps), torch.cat(d1s)


def ema_quantizer_forward(inputs: Tensor, embed: Tensor, commitment_cost: float = 1.0
                          ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """EMAVectorQuantizer.forward in eval mode (layers/vq.py:96-154).
    Returns (quantized [B,D,*sp], indices int64 [B,*sp], loss 0-dim, rel_gap [B,*sp])."""
    ndim = inputs.dim()
    assert ndim >= 3                                                   # vq.py:98
    if inputs.shape[1] != embed.shape[1]:                              # vq.py:100-104
        raise NotImplementedError("VQ dim != channel dim not supported")
    channel_last = inputs.permute(0, *range(2, ndim), 1)               # vq.py:107-113
    shape = channel_last.shape
    flat = channel_last.reshape(-1, embed.shape[1])                    # vq.py:116
    idx, gap, _ = quantize_flat(flat, embed)
    quantized = F.embedding(idx, embed).reshape(shape)                 # vq.py:130, 44-45
    quantized = quantized.permute(0, -1, *range(1, ndim - 1))          # vq.py:139
    loss = F.mse_loss(inputs, quantized) * commitment_cost             # vq.py:143
    quantized = inputs + (quantized - inputs)                          # vq.py:146 (STE)
    return quantized, idx.reshape(shape[:-1]), loss, gap.reshape(shape[:-1])


def projected_quantizer_forward(inputs: Tensor, sd: StateDict, prefix: str,
                                commitment_cost: float = 1.0):
    """ProjectedEMAVectorQuantizer2d.forward (layers/vq.py:190-192)."""
    z = F.conv2d(inputs, sd[prefix + "proj_in.weight"], sd[prefix + "proj_in.bias"])
    q, idx, loss, gap = ema_quantizer_forward(z, sd[prefix + "embed"], commitment_cost)
    out = F.conv2d(q, sd[prefix + "proj_out.weight"], sd[prefix + "proj_out.bias"])
    return out, idx, loss, gap, z


def embed_code(idx: Tensor, embed: Tensor) -> Tensor:
    """EMAVectorQuantizer.embed_code (layers/vq.py:44-45)."""
    return F.embedding(idx, embed)


# ------------------------------------------------------------------------------------
# a-E / a-De / a-V  encoder, decoder, full model (single VQ level, as shipped)
# ------------------------------------------------------------------------------------
def encoder_forward(x: Tensor, sd: StateDict, prefix: str = "encoder.", with_aux: bool = False):
    """Encoder.forward (model.py:189-217) for one VQ level: in_stem (3x3, zero pad, bias)
    -> DownBlock -> 50-block trunk -> projected quantiser.
    Returns ((enc,), (idx,), (loss,)) like the reference; with_aux adds (rel_gap, z, pre_vq)."""
    h = F.conv2d(x, sd[prefix + "in_stem.weight"], sd[prefix + "in_stem.bias"], padding=1)
    h = envelop_pyramid(h, sd, prefix + "down_layers.0.layers.")
    h = block_sequence(h + 0, sd, prefix + "pre_enc_layers.0.")         # model.py:208 (+ 0)
    enc, idx, loss, gap, z = projected_quantizer_forward(h, sd, prefix + "vq_layers.0.")
    out = ((enc,), (idx,), (loss,))
    return out + ((gap, z, h),) if with_aux else out


def decoder_forward(encs: Sequence[Tensor], sd: StateDict, prefix: str = "decoder.") -> Tensor:
    """Decoder.forward (model.py:274-291) for one level: trunk -> UpBlock -> out_stem."""
    h = block_sequence(0 + encs[0], sd, prefix + "post_enc_layers.0.")
    h = envelop_pyramid(0 + h, sd, prefix + "up_layers.0.layers.")
    return F.conv2d(h, sd[prefix + "out_stem.weight"], sd[prefix + "out_stem.bias"], padding=1)


def vqae_forward(x: Tensor, sd: StateDict):
    """VQAE.forward (model.py:41-48): returns (out, (loss,))."""
    encs, _idx, loss = encoder_forward(x, sd)
    return decoder_forward(encs, sd), loss


def decode_from_codes(idx: Tensor, sd: StateDict) -> Tensor:
    """Decompress stored codes: embed_code -> proj_out -> Decoder."""
    p = "encoder.vq_layers.0."
    q = embed_code(idx, sd[p + "embed"]).permute(0, 3, 1, 2)
    enc = F.conv2d(q, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
    return decoder_forward((enc,), sd)


# ------------------------------------------------------------------------------------
# f-4  multi-level hierarchy (model.py:144-148, 189-217, 274-291) and training-mode EMA update
# ------------------------------------------------------------------------------------
def module_chain(x: Tensor, sd: StateDict, prefix: str) -> Tensor:
    """Whatever chain of PreActFixupResBlocks lives under ``prefix``: a bare block, an nn.Sequential of
    blocks (``prefix{i}.``) or a DownBlock / UpBlock (``prefix`` + ``layers.{j}.layers.{i}.``)."""
    if not any(k.startswith(prefix) for k in sd):
        return x                                   # n_down = 0 / n_up = 0: an empty nn.Sequential
    if prefix + "bias1a" in sd or prefix + "branch.0.weight" in sd:
        return any_block(x, _blk(sd, prefix))
    if any(k.startswith(prefix + "layers.") for k in sd):
        return envelop_pyramid(x, sd, prefix + "layers.")
    return block_sequence(x, sd, prefix)


def _quantizer_any(h: Tensor, sd: StateDict, prefix: str):
    if prefix + "proj_in.weight" in sd:
        enc, idx, loss, _gap, _z = projected_quantizer_forward(h, sd, prefix)
        return enc, idx, loss
    enc, idx, loss, _gap = ema_quantizer_forward(h, sd[prefix + "embed"])
    return enc, idx, loss


def encoder_forward_levels(x: Tensor, sd: StateDict, prefix: str = "encoder."):
    """Encoder.forward (model.py:189-217) for any number of levels.  Module lists are stored low-res
    first (model.py:182-187) except ``down_layers`` (execution order); ``shortcut_layers.{i}`` is absent
    from the state_dict where the reference holds ``None``."""
    h = F.conv2d(x, sd[prefix + "in_stem.weight"], sd[prefix + "in_stem.bias"], padding=1)
    n = len(_indexed_children(sd, prefix + "vq_layers."))
    downs = []
    for i in range(n):
        h = module_chain(h, sd, prefix + f"down_layers.{i}.")
        downs.append(h)
    outs, aux = [], None
    for i, down in enumerate(reversed(downs)):                          # model.py:203-215
        sc = prefix + f"shortcut_layers.{i}."
        has_sc = any(k.startswith(sc) for k in sd)
        h = down + (module_chain(aux, sd, sc) if has_sc else 0)
        h = block_sequence(h, sd, prefix + f"pre_enc_layers.{i}.")
        enc, idx, loss = _quantizer_any(h, sd, prefix + f"vq_layers.{i}.")
        aux = enc
        outs.append((enc, idx, loss))
    return tuple(zip(*outs))


def decoder_forward_levels(encs: Sequence[Tensor], sd: StateDict, prefix: str = "decoder.") -> Tensor:
    """Decoder.forward (model.py:274-291) for any number of levels, encodings low-res first."""
    prev_up, aux = 0, None
    for i, enc in enumerate(encs):
        sc = prefix + f"shortcut_layers.{i}."
        has_sc = any(k.startswith(sc) for k in sd)
        h = (module_chain(aux, sd, sc) if has_sc else 0) + enc          # model.py:285-286
        aux = enc
        h = block_sequence(h, sd, prefix + f"post_enc_layers.{i}.")
        prev_up = module_chain(prev_up + h, sd, prefix + f"up_layers.{i}.")
    return F.conv2d(prev_up, sd[prefix + "out_stem.weight"], sd[prefix + "out_stem.bias"], padding=1)


def ema_init(flat: Tensor, embed: Tensor, cluster_size: Tensor, world: int = 1):
    """EMAVectorQuantizer._init_ema (vq.py:76-94) on one rank's rows; returns the new
    (embed, embed_avg, cluster_size)."""
    mean, std = flat.mean(dim=0), flat.std(dim=0)                       # vq.py:77-78
    e = embed * std + mean                                              # vq.py:90-91
    return e, e.clone(), cluster_size + flat.shape[0] * world / embed.shape[0]


def ema_update(flat: Tensor, idx: Tensor, embed_avg: Tensor, cluster_size: Tensor, decay: float,
               laplace_alpha: float):
    """EMAVectorQuantizer._update_ema (vq.py:47-74) without the one-hot matrix; returns the new
    (embed, embed_avg, cluster_size)."""
    k, d = embed_avg.shape
    counts = torch.bincount(idx.reshape(-1), minlength=k).to(flat.dtype)            # one_hot.sum(0)
    dw = torch.zeros(k, d, dtype=torch.float64).index_add_(0, idx.reshape(-1), flat.double())
    cs = cluster_size * decay + counts * (1 - decay)                                # vq.py:60-62
    avg = embed_avg * decay + dw.to(flat.dtype) * (1 - decay)                       # vq.py:64
    n = cs.sum()
    smoothed = n * ((cs + laplace_alpha) / (n + k * laplace_alpha))                 # vq.py:67-71
    return avg / smoothed.unsqueeze(-1), avg, cs


# ------------------------------------------------------------------------------------
# a-X  slide tiling / code-map stitching (host-side geometry)
# ------------------------------------------------------------------------------------
def slide_grid(level_shape: Tuple[int, int], patch: int) -> Tuple[int, int]:
    """rows, cols = level_shape // patch_size, remainder dropped
    (datamodules/camelyon16.py:160-168)."""
    return level_shape[0] // patch, level_shape[1] // patch


def patch_rc(patch_index: int, cols: int) -> Tuple[int, int]:
    """row-major patch index -> (row, col) (datamodules/camelyon16.py:184-190)."""
    return patch_index // cols, patch_index % cols


def stitch_code_map(codes: np.ndarray, rows: int, cols: int) -> np.ndarray:
    """Place [P,h,w] code tiles at (r*h, c*w) of a [rows*h, cols*w] map and narrow to the
    smallest dtype (scripts/extract_embeddings/extract_embeddings.py:47-59,75-89)."""
    P, h, w = codes.shape
    out = np.empty((rows * h, cols * w), dtype=codes.dtype)
    for p in range(P):
        r, c = patch_rc(p, cols)
        out[r * h:(r + 1) * h, c * w:(c + 1) * w] = codes[p]
    return cast_to_lowest_dtype(out)


def cast_to_lowest_dtype(array: np.ndarray) -> np.ndarray:
    """extract_embeddings.py:54-59."""
    amin, amax = array.min(), array.max()
    if amin == 0 and amax == 1:
        return array.astype(bool)
    return array.astype(np.result_type(np.min_scalar_type(amin), np.min_scalar_type(amax)))
