"""The code-extraction caller on the B200 path (scope row f-1).

Counterpart of ``run_eval`` / ``get_encodings`` in
scripts/extract_embeddings/extract_embeddings.py:43-138: patches in, per-slide u8 code map
out.  Differences by design: tiles arrive as raw uint8 NHWC (normalisation is fused into the
stem kernel instead of running in DataLoader workers), the quantised tensor is never written
(only indices are used, extract_embeddings.py:125-137), and tiles are placed into the slide
map by a kernel instead of advanced indexing.
"""
from __future__ import annotations

from typing import Iterable, Optional, Tuple

import numpy as np
import torch

from . import engine as E
from .sharding import gather_code_tiles, shard_range, slide_grid


def cast_to_lowest_dtype(array: np.ndarray) -> np.ndarray:
    """extract_embeddings.py:54-59."""
    amin, amax = array.min(), array.max()
    if amin == 0 and amax == 1:
        return array.astype(bool)
    return array.astype(np.result_type(np.min_scalar_type(amin), np.min_scalar_type(amax)))


@torch.no_grad()
def encode_patches(encoder, patches: torch.Tensor, mean=None, std=None) -> torch.Tensor:
    """patches: uint8 [B,H,W,3] or float [B,3,H,W] on a CUDA device -> int64 codes [B,h,w]."""
    _, idx, _, _, _ = encoder.encode(patches, mean=mean, std=std, want_quantized=False)
    return idx


@torch.no_grad()
def compress_slide(encoder, batches: Iterable[Tuple[int, torch.Tensor]], grid: Tuple[int, int],
                   code_hw: Tuple[int, int], device: torch.device, mean=None, std=None
                   ) -> torch.Tensor:
    """Encode (first_patch_index, patch batch) pairs of one slide and place the code tiles into a
    u8 map [rows*h, cols*w] on ``device`` (extract_embeddings.py:47-52,75-89)."""
    rows, cols = grid
    th, tw = code_hw
    code_map = torch.zeros(rows * th, cols * tw, dtype=torch.uint8, device=device)
    for first_patch, patches in batches:
        idx = encode_patches(encoder, patches.to(device, non_blocking=True), mean, std)
        E.codemap_place(idx, first_patch, cols, code_map)
    return code_map


def tiles_to_map(tiles_u8: torch.Tensor, grid: Tuple[int, int]) -> torch.Tensor:
    """[P,th,tw] tiles in row-major patch order -> [rows*th, cols*tw] map (pure view ops)."""
    rows, cols = grid
    p, th, tw = tiles_u8.shape
    assert p == rows * cols
    return tiles_u8.view(rows, cols, th, tw).permute(0, 2, 1, 3).reshape(rows * th, cols * tw)
