"""Launch the fused tcgen05 'same' block kernel a few times at the bench shape (for ncu)."""
import ctypes
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "2d-vq-ae-2_b200")]
import torch  # noqa: E402

from vqae_b200 import _lib as L  # noqa: E402
from vqae_b200 import engine as E  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
HW = int(sys.argv[2]) if len(sys.argv) > 2 else 32
B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
dev = torch.device("cuda:0")
lib = L.load()
xs = [torch.randn(B, HW, HW, C, device=dev) for _ in range(2)]
ys = [torch.empty_like(xs[0]) for _ in range(2)]
ws = [torch.randn(C, C, k, k, device=dev) * 0.05 for k in (1, 3, 1)]
cp = max(C, 16)
packed = torch.empty(11 * cp * cp, dtype=torch.bfloat16, device=dev)
st = E._stream(dev)
L.check(lib.vqae_pack_same_block_f16(E._ptr(ws[0]), E._ptr(ws[1]), E._ptr(ws[2]), C, E._ptr(packed), st), "pack")
sc = (ctypes.c_float * 8)(0.01, 0.02, -0.01, 0.03, 0.02, -0.02, 0.01, 0.9)
for i in range(8):
    L.check(lib.vqae_same_block_f16(E._ptr(xs[i % 2]), E._ptr(ys[i % 2]), E._ptr(packed), sc, B, HW, HW, C, st), "run")
torch.cuda.synchronize()
print("ok")
