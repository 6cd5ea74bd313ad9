"""Patch sharding across ranks and the optional whole-slide code-map gather.

Patches are independent units (circular padding keeps every conv halo inside its patch,
pre_activation_fixup.yaml:40,60), so the data path has NO collective: each rank encodes a
contiguous block of the row-major patch list.  Only assembling one slide's code map on rank 0
uses a collective -- one ``all_gather`` of equal-padded u8 code tiles (NCCL on GPUs, gloo in
the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def slide_grid(level_shape: Tuple[int, int], patch_size: int) -> Tuple[int, int]:
    """(rows, cols) = level_shape // patch_size, remainder dropped
    (datamodules/camelyon16.py:160-168)."""
    return level_shape[0] // patch_size, level_shape[1] // patch_size


def patch_rc(patch_index: int, cols: int) -> Tuple[int, int]:
    """Row-major patch index -> (row, col) (datamodules/camelyon16.py:184-190)."""
    return patch_index // cols, patch_index % cols


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [start, stop) of rank: ceil(n/world) items each, last ranks may be short
    or empty."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    per = -(-n_items // world_size)
    start = min(rank * per, n_items)
    return start, min(start + per, n_items)


def gather_code_tiles(tiles: torch.Tensor, n_total: int, group: Optional[dist.ProcessGroup] = None
                      ) -> torch.Tensor:
    """All-gather every rank's u8 code tiles [P_r,th,tw] (rank r holds shard_range(n_total, r, W))
    and return the full [n_total,th,tw] tensor in patch order on every rank."""
    if not (dist.is_available() and dist.is_initialized()):
        assert tiles.shape[0] == n_total
        return tiles
    world = dist.get_world_size(group)
    per = -(-n_total // world)
    th, tw = tiles.shape[1:]
    padded = torch.zeros(per, th, tw, dtype=tiles.dtype, device=tiles.device)
    padded[: tiles.shape[0]] = tiles
    out = torch.empty(world * per, th, tw, dtype=tiles.dtype, device=tiles.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return out[:n_total]
