"""Patch sharding across ranks and the optional whole-slide code-map gather.

Patches are independent units (circular padding keeps every conv halo inside its patch,
pre_activation_fixup.yaml:40,60), so the data path has NO collective: each rank encodes a
contiguous block of the row-major patch list.  Only assembling one slide's code map on rank 0
uses a collective -- one ``all_gather`` of equal-padded u8 code tiles (NCCL on GPUs, gloo in
the CPU tests).
"""
from __future__ import annotations

import os
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


def bind_host_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off, so that pinned host
    buffers allocated afterwards (first touch) and the staging thread are node-local.  With one rank
    per GPU all ranks otherwise stage their uint8 batches through whichever node the scheduler put
    them on: at 8 ranks x 50 MB per 4 ms step the cross-socket hops cost ~3 % of the end-to-end rate.
    Best effort: returns the node, or None (and changes nothing) when sysfs does not say."""
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None
