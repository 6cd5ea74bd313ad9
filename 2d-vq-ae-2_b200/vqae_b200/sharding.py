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


def _parse_cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def _gpu_local_cpus(device_index: int):
    """(cpus, where) of the NUMA node the GPU hangs off: sysfs first, `nvidia-smi topo -m`'s CPU-affinity
    column second (containers often hide the sysfs node); (None, None) when neither says."""
    import re
    import subprocess
    try:
        pr = torch.cuda.get_device_properties(device_index)
        path = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node >= 0:
            return _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read()), f"node{node}"
    except (OSError, ValueError, AttributeError, RuntimeError):
        pass
    try:
        # CUDA_VISIBLE_DEVICES may renumber: match the row by PCI bus id through nvidia-smi's own query
        q = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id", "--format=csv,noheader"],
                           capture_output=True, text=True, timeout=20).stdout
        pr = torch.cuda.get_device_properties(device_index)
        want = f"{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0".lower()
        smi_index = None
        for line in q.splitlines():
            idx, _, bus = line.partition(",")
            if bus.strip().lower().endswith(want):
                smi_index = int(idx)
        if smi_index is None:
            return None, None
        topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        for line in topo.splitlines():
            cols = line.split()
            if cols and cols[0] == f"GPU{smi_index}":
                for tok in cols[1:]:
                    if re.fullmatch(r"\d+(-\d+)?(,\d+(-\d+)?)*", tok) and ("-" in tok or "," in tok):
                        return _parse_cpulist(tok), "nvidia-smi topo"
    except (OSError, ValueError, subprocess.SubprocessError):
        pass
    return None, None


def bind_host_to_gpu_numa_node(device_index: int) -> Optional[str]:
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off, so that pinned host
    buffers allocated afterwards (first touch) and the staging thread are node-local.  With one rank
    per GPU all ranks otherwise stage their uint8 batches through whichever node the scheduler put
    them on: at 8 ranks x 50 MB per 4 ms step the cross-socket hops cost several % of the end-to-end
    rate.  Best effort: returns where the CPU list came from, or None (and changes nothing)."""
    try:
        cpus, where = _gpu_local_cpus(device_index)
        if not cpus:
            return None
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return f"{where}: {len(cpus)} cpus"
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None
