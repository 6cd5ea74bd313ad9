"""Time the image-resident trunk kernel alone (CUDA events).  usage: time_resident.py [n_blocks=54] [reps=5] [batch=256] [C=64]"""
import sys
import types
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "2d-vq-ae-2_b200")]
import torch  # noqa: E402
from vqae_b200 import _lib as L  # noqa: E402
from vqae_b200 import engine as E  # noqa: E402

nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 54
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
C = int(sys.argv[4]) if len(sys.argv) > 4 else 64
H = W = {64: 32, 128: 32, 32: 64}[C]
dev = torch.device("cuda:0")
lib = L.load()
st = E._stream(dev)
gen = torch.Generator().manual_seed(7)
packs = []
for i in range(nblk):
    ws = [(torch.randn(C, C, k, k, generator=gen) * 0.05).to(dev) for k in (1, 3, 1)]
    pk = torch.empty(11 * C * C, dtype=torch.bfloat16, device=dev)
    L.check(lib.vqae_pack_resident_block_f16(E._ptr(ws[0]), E._ptr(ws[1]), E._ptr(ws[2]), C, 0.2,
                                              E._ptr(pk), st), "pack")
    packs.append(pk)
chain = types.SimpleNamespace(weights=torch.cat(packs), n=nblk, scalars=torch.tensor(
    [[0.01, 0.02, -0.01, 0.03, 0.02, -0.02, 0.01, 0.2]] * nblk, dtype=torch.float32).to(dev))
xs = [torch.randn(B, H, W, C, device=dev) for _ in range(4)]
y = torch.empty(B, H, W, C, device=dev)
for i in range(3):
    E.trunk_resident(xs[i % 4], y, chain)
torch.cuda.synchronize()
ts = []
for i in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    E.trunk_resident(xs[i % 4], y, chain)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
print(f"{L.library_path().name}: resident {nblk} blocks C={C} batch {B}: median {ms:.4f} ms  min {min(ts):.4f}  "
      f"{2.0 * B * H * W * C * C * 11 * nblk / ms / 1e9:.1f} TFLOP/s  checksum {float(y.double().sum()):.6e}")
