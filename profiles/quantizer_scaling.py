"""tcgen05 quantiser: time per call against N (fixed overhead vs per-tile cost), CUDA events."""
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "2d-vq-ae-2_b200")]
import torch  # noqa: E402

from vqae_b200 import engine as E  # noqa: E402
from vqae_b200.layers.vq import ProjectedEMAVectorQuantizer2d  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
pq = ProjectedEMAVectorQuantizer2d(256, C, 1.0, 0.99, 1e-5, 8).eval().to(dev)
packed = pq.packed()
print("N tiles_per_sm us_per_call GB/s")
for b in (37, 74, 148, 296, 512, 1024, 2048):
    n = b * 1024
    xs = [torch.randn(b, 1024, C, device=dev) for _ in range(3)]
    bufs = E.QuantizeBuffers(packed, n, torch.float32, dev)
    for i in range(5):
        E.quantize_into(packed, xs[i % 3], bufs, b, 1024)
    torch.cuda.synchronize()
    reps = 30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        E.quantize_into(packed, xs[i % 3], bufs, b, 1024)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"{n:8d} {n / 128 / 148:6.1f} {us:8.2f} {n * (2 * C * 4 + 8) / us / 1e3:8.1f}")
