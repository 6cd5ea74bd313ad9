import os, sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "2d-vq-ae-2_b200")]
import torch
from vqae_b200 import engine as E
from vqae_b200.layers.vq import ProjectedEMAVectorQuantizer2d
dev = torch.device("cuda:0")
torch.manual_seed(0)
pq = ProjectedEMAVectorQuantizer2d(256, 64, 1.0, 0.99, 1e-5, 8).eval().to(dev)
B = 512
xs = [torch.randn(B, 1024, 64, device=dev) for _ in range(3)]
packed = pq.packed()
for want_out in (True, False, True):
    for i in range(3):
        E.quantize(packed, xs[i % 3], True, True, B, 1024, want_out=want_out)
    torch.cuda.synchronize()
    ts = []
    for i in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = E.quantize(packed, xs[i % 3], True, True, B, 1024, want_out=want_out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print("sync each, want_out", want_out, " ".join(f"{t:.0f}" for t in ts))
    ts = []
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(13)]
    evs[0].record()
    keep = []
    for i in range(12):
        keep.append(E.quantize(packed, xs[i % 3], True, True, B, 1024, want_out=want_out))
        evs[i + 1].record()
    torch.cuda.synchronize()
    print("async,     want_out", want_out, " ".join(f"{evs[i].elapsed_time(evs[i+1])*1e3:.0f}" for i in range(12)))
    del keep
