"""BASELINE.json config 5: a synthetic whole slide (50 000^2 px -> 195 x 195 = 38 025 patches of 256^2)
compressed to a 6 240 x 6 240 u8 code map, patches sharded by contiguous blocks over the ranks, code
tiles collected with ONE NCCL all-gather (vqae_b200.sharding.gather_code_tiles) and placed by
vqae_codemap_place_u8.  Pixels are generated on the device per batch from (seed, first patch index), so
the map is a function of the patch index only: its checksum must not depend on the number of ranks.

    python profiles/slide_bench.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29512 profiles/slide_bench.py                  # 8 GPUs
"""
import hashlib
import json
import os
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "2d-vq-ae-2_b200")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import vqae_b200  # noqa: E402
from vqae_b200 import engine as E  # noqa: E402
from vqae_b200 import synthetic as S  # noqa: E402
from vqae_b200.extract import encode_patches  # noqa: E402
from vqae_b200.sharding import gather_code_tiles, shard_range, slide_grid  # noqa: E402

BATCH, PATCH, LEVEL = 256, 256, (50_000, 50_000)


def device_patches(first: int, count: int, dev) -> torch.Tensor:
    """uint8 [count,256,256,3] tiles that depend only on the patch index."""
    out = torch.empty(count, PATCH, PATCH, 3, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev)
    for i in range(count):                       # one generator state per patch: shard-independent
        g.manual_seed(1234 + first + i)
        out[i] = torch.randint(0, 256, (PATCH, PATCH, 3), dtype=torch.uint8, device=dev, generator=g)
    return out


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rows, cols = slide_grid(LEVEL, PATCH)
    n_total = rows * cols
    lo, hi = shard_range(n_total, rank, world)
    model = vqae_b200.build_vqae(n_down=3).eval()
    model.load_state_dict(S.make_state_dict(model.state_dict(), seed=1, regime="perturbed"))
    model = vqae_b200.set_precision(model.to(dev), "fp16")
    enc = model.encoder
    # inputs are generated before the timed region (H2D / generation reported separately)
    t0 = time.perf_counter()
    batches = [(s, device_patches(s, min(BATCH, hi - s), dev)) for s in range(lo, hi, BATCH)]
    torch.cuda.synchronize(dev)
    t_gen = time.perf_counter() - t0
    with torch.no_grad():
        encode_patches(enc, batches[0][1])        # warm-up (weight packing, module load)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    with torch.no_grad():
        tiles = torch.cat([encode_patches(enc, p).to(torch.uint8) for _, p in batches])
    e1.record()
    full = gather_code_tiles(tiles, n_total)      # [n_total,32,32] u8 on every rank
    code_map = torch.zeros(rows * 32, cols * 32, dtype=torch.uint8, device=dev)
    E.codemap_place(full.long(), 0, cols, code_map)
    e2.record()
    torch.cuda.synchronize(dev)
    t_enc, t_all = e0.elapsed_time(e1), e0.elapsed_time(e2)
    times = torch.tensor([t_enc, t_all], device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    if rank == 0:
        digest = hashlib.sha256(code_map.cpu().numpy().tobytes()).hexdigest()[:16]
        per8 = -(-n_total // 8)
        parts = [hashlib.sha256(full[k * per8:(k + 1) * per8].cpu().numpy().tobytes()).hexdigest()[:8]
                 for k in range(8)]
        print(json.dumps({
            "workload": f"synthetic slide {LEVEL[0]}x{LEVEL[1]} px -> {rows}x{cols} patches -> "
                        f"{code_map.shape[0]}x{code_map.shape[1]} u8 code map",
            "n_gpus": world, "patches": n_total, "encode_ms": float(times[0]), "total_ms": float(times[1]),
            "gather_and_place_ms": float(times[1] - times[0]),
            "patches_per_s": n_total / float(times[1]) * 1e3,
            "input_generation_s_rank0": t_gen, "code_map_sha256_16": digest,
            "tile_sha256_8_per_eighth": parts,
            "gathered_bytes_per_rank": int(tiles.numel())}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
