"""Deterministic synthetic weights and inputs (there are no checkpoints or slides offline).

``make_state_dict`` fills a reference-layout ``state_dict`` key by key from generators seeded
with ``crc32(key) ^ seed``, so the values depend neither on module construction order nor on
which implementation (reference or this package) built the template.  Two regimes
(hard part 6):

  "fixup"      what the reference's own initialisation produces in distribution:
               branch_conv3 = 0, scalar biases 0, scale 1 -> every 'same' block is the identity.
  "perturbed"  non-degenerate: scalar biases ~ N(0, 0.05), scale ~ N(1, 0.1), branch_conv3
               ~ N(0, 2/fan_out / n_layers) -- every conv of every block is observable.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Mapping, Optional

import torch

Tensor = torch.Tensor


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _normal(shape, std: float, g: torch.Generator, mean: float = 0.0) -> Tensor:
    return torch.randn(shape, generator=g) * std + mean


def _uniform(shape, bound: float, g: torch.Generator) -> Tensor:
    return (torch.rand(shape, generator=g) * 2.0 - 1.0) * bound


def make_state_dict(template: Mapping[str, Tensor], seed: int = 0, regime: str = "perturbed",
                    n_layers: Optional[int] = None) -> Dict[str, Tensor]:
    """New CPU fp32 state_dict with the template's keys/shapes and deterministic values."""
    assert regime in ("fixup", "perturbed")
    if n_layers is None:
        n_layers = sum(1 for k in template if k.endswith("branch_conv3.weight"))
    out: Dict[str, Tensor] = {}
    embed_keys = []
    for key, ref in template.items():
        g = _gen(key, seed)
        shape = tuple(ref.shape)
        leaf = key.rsplit(".", 1)[-1]
        parent = key.rsplit(".", 2)[-2] if key.count(".") >= 1 else ""
        if parent in ("branch_conv1", "branch_conv2", "branch_conv3", "skip_conv"):
            o, i, kh, kw = shape
            if parent == "branch_conv1":
                v = _normal(shape, math.sqrt(2.0 / (o * kh * kw)) * n_layers ** -0.5, g)
            elif parent == "branch_conv2":
                v = _normal(shape, math.sqrt(2.0 / (i * kh * kw)), g)
            elif parent == "branch_conv3":
                v = (torch.zeros(shape) if regime == "fixup"
                     else _normal(shape, math.sqrt(2.0 / (o * kh * kw)) * n_layers ** -0.5, g))
            else:
                v = _normal(shape, math.sqrt(2.0 / ((i + o) * kh * kw)), g)
        elif leaf.startswith("bias") and shape == (1,):
            v = torch.zeros(1) if regime == "fixup" else _normal(shape, 0.05, g)
        elif leaf == "scale":
            v = torch.ones(1) if regime == "fixup" else _normal(shape, 0.1, g, mean=1.0)
        elif parent in ("in_stem", "out_stem", "proj_in", "proj_out"):
            if leaf == "weight":
                fan_in = shape[1] * shape[2] * shape[3]
                v = _uniform(shape, 1.0 / math.sqrt(fan_in), g)
            else:
                wshape = template[key[:-4] + "weight"].shape
                v = _uniform(shape, 1.0 / math.sqrt(wshape[1] * wshape[2] * wshape[3]), g)
        elif leaf == "embed":
            v = torch.randn(shape, generator=g)
            embed_keys.append(key)
        elif leaf == "embed_avg":
            v = None  # filled from embed below
        elif leaf == "cluster_size":
            v = torch.zeros(shape)
        elif leaf == "first_pass":
            v = torch.as_tensor(1)
        elif ".branch." in key or key.startswith("branch."):
            # MBConv / SELayer tensors (layers/conv_block.py:264-309, layers/misc.py:16-21): convs,
            # BatchNorm2d affine + running statistics, the two Linear layers of the squeeze-excite
            stem = key.rsplit(".", 1)[0]
            if len(shape) == 4:
                v = _normal(shape, math.sqrt(2.0 / (shape[1] * shape[2] * shape[3])), g)
            elif len(shape) == 2:
                v = _uniform(shape, 1.0 / math.sqrt(shape[1]), g)
            elif leaf == "bias" and ".fc." in key:
                v = _uniform(shape, 1.0 / math.sqrt(template[stem + ".weight"].shape[1]), g)
            elif leaf == "weight":
                # BatchNorm gamma; the block's LAST BatchNorm (the reference starts it at zero,
                # conv_block.py:312-313) stays small so that deep stacks of blocks keep O(1) activations
                head, _, idx = stem.rpartition(".")
                last = not any(k.startswith(head + ".") and k[len(head) + 1:].split(".")[0].isdigit()
                               and int(k[len(head) + 1:].split(".")[0]) > int(idx) for k in template)
                v = _normal(shape, 0.05, g, mean=0.25) if last else _normal(shape, 0.1, g, mean=1.0)
            elif leaf == "bias":
                v = _normal(shape, 0.05, g)                       # BatchNorm beta
            elif leaf == "running_mean":
                v = _normal(shape, 0.1, g)
            elif leaf == "running_var":
                v = torch.rand(shape, generator=g) + 0.5
            elif leaf == "num_batches_tracked":
                v = torch.zeros(shape, dtype=torch.int64)
            else:
                raise KeyError(f"make_state_dict: no rule for {key} {shape}")
        else:
            raise KeyError(f"make_state_dict: no rule for {key} {shape}")
        out[key] = v if v is None else v.to(ref.dtype)
    for embed_key in embed_keys:                      # one per VQ level
        out[embed_key[:-5] + "embed_avg"] = out[embed_key].clone()
    return out


def rescale_codebook(embed: Tensor, latents: Tensor) -> Tensor:
    """What ``_init_ema`` does on the first training batch (vq_ae/layers/vq.py:76-94):
    embed * std(latents) + mean(latents), per distance-space dimension.  latents: [N, D]."""
    return embed * latents.std(dim=0) + latents.mean(dim=0)


def synthetic_patches(batch: int, size: int, seed: int, device="cpu", dtype=torch.float32
                      ) -> Tensor:
    """Stand-in for normalised patches: N(0,1) [B,3,size,size] from a CPU generator."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randn(batch, 3, size, size, generator=g, dtype=torch.float32).to(
        device=device, dtype=dtype)


def synthetic_patches_u8(batch: int, size: int, seed: int, device="cpu") -> Tensor:
    """Stand-in for raw RGB tiles: uint8 [B,size,size,3]."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randint(0, 256, (batch, size, size, 3), generator=g, dtype=torch.uint8).to(device)
