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


class StreamingEncoder:
    """Double-buffered host -> device -> host code extraction.

    The copy of batch i+1 from pinned host memory runs on a side stream while batch i is encoded,
    and the code indices of batch i return to pinned host memory asynchronously -- the B200
    counterpart of the reference's DataLoader(pin_memory) + ``imgs.to(device, non_blocking=True)``
    loop (extract_embeddings.py:95-101,118-122).  ``encode_stream`` yields one pinned int64
    ``[B,h,w]`` tensor per input batch, in order; a yielded tensor is valid until two more
    batches have been yielded."""

    def __init__(self, encoder, device: torch.device, mean=None, std=None):
        self.encoder, self.device, self.mean, self.std = encoder, device, mean, std
        self.copy_stream = torch.cuda.Stream(device)
        self._dev_in = [None, None]
        self._host_out = [None, None]
        self._h2d = [torch.cuda.Event() for _ in range(2)]
        self._free = [torch.cuda.Event() for _ in range(2)]     # device input buffer consumed
        self._d2h = [torch.cuda.Event() for _ in range(2)]

    def _stage(self, slot: int, host_batch: torch.Tensor, first_use: bool) -> None:
        if self._dev_in[slot] is None or self._dev_in[slot].shape != host_batch.shape:
            self._dev_in[slot] = torch.empty(host_batch.shape, dtype=host_batch.dtype,
                                             device=self.device)
            first_use = True
        with torch.cuda.stream(self.copy_stream):
            if not first_use:
                self.copy_stream.wait_event(self._free[slot])
            self._dev_in[slot].copy_(host_batch, non_blocking=True)
            self._h2d[slot].record(self.copy_stream)

    @torch.no_grad()
    def encode_stream(self, host_batches):
        main = torch.cuda.current_stream(self.device)
        it = iter(host_batches)
        nxt = next(it, None)
        if nxt is None:
            return
        self._stage(0, nxt, True)
        i = 0
        pending = []
        while nxt is not None:
            slot = i & 1
            cur, nxt = nxt, next(it, None)
            if nxt is not None:
                self._stage(slot ^ 1, nxt, i == 0)
            main.wait_event(self._h2d[slot])
            idx = encode_patches(self.encoder, self._dev_in[slot], self.mean, self.std)
            self._free[slot].record(main)
            if self._host_out[slot] is None or self._host_out[slot].shape != idx.shape:
                self._host_out[slot] = torch.empty(idx.shape, dtype=idx.dtype).pin_memory()
            elif len(pending) == 2:
                pass
            self._host_out[slot].copy_(idx, non_blocking=True)
            self._d2h[slot].record(main)
            pending.append(slot)
            if len(pending) == 2:                  # results lag one batch behind the submission
                done = pending.pop(0)
                self._d2h[done].synchronize()
                yield self._host_out[done]
            i += 1
        for done in pending:
            self._d2h[done].synchronize()
            yield self._host_out[done]


def tiles_to_map(tiles_u8: torch.Tensor, grid: Tuple[int, int]) -> torch.Tensor:
    """[P,th,tw] tiles in row-major patch order -> [rows*th, cols*tw] map (pure view ops)."""
    rows, cols = grid
    p, th, tw = tiles_u8.shape
    assert p == rows * cols
    return tiles_u8.view(rows, cols, th, tw).permute(0, 2, 1, 3).reshape(rows * th, cols * tw)
