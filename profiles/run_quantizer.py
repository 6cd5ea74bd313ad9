"""Time the projected quantiser at BASELINE config 2 ([512,64,32,32] fp32 NHWC, K=256, D=8):
tcgen05 filter kernel vs the exact CUDA-core kernel; a few launches each (also the ncu target)."""
import os
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "2d-vq-ae-2_b200")]
import torch  # noqa: E402

from vqae_b200 import engine as E  # noqa: E402
from vqae_b200.layers.vq import ProjectedEMAVectorQuantizer2d  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
torch.manual_seed(0)
pq = ProjectedEMAVectorQuantizer2d(256, 64, 1.0, 0.99, 1e-5, 8).eval().to(dev)
xs = [torch.randn(B, 1024, 64, device=dev) for _ in range(3)]   # 3 x 134 MB in, rotating (> L2)
packed = pq.packed()
n = B * 1024
bytes_alg = n * 64 * 4 * 2 + n * 8
for mode in ("1", "0"):
    os.environ["VQAE_QUANT_TC"] = mode
    for want_out in (True, False):
        for i in range(3):
            E.quantize(packed, xs[i % 3], True, True, B, 1024, want_out=want_out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(REPS):
            out, idx, loss, ties, _ = E.quantize(packed, xs[i % 3], True, True, B, 1024, want_out=want_out)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / REPS
        by = bytes_alg if want_out else n * 64 * 4 + n * 8
        print(f"tc={mode} want_out={want_out}: {us:8.1f} us/call  {by / us * 1e-3:7.1f} GB/s algorithmic "
              f"({by / us * 1e-3 / 6551.7 * 100:.1f}% of 6551.7)  ties={int(ties)} loss={float(loss):.6f} "
              f"idxsum={int(idx.sum())}")
