"""Phase-level clock64() profile of the fused tcgen05 'same' block kernel (first tile of each CTA).
usage: python profiles/phase_clocks.py [C] [HW] [batch]"""
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
x = torch.randn(B, HW, HW, C, device=dev)
y = torch.empty_like(x)
ws = [torch.randn(C, C, k, k, device=dev) * 0.05 for k in (1, 3, 1)]
cp = max(C, 16)
packed = torch.empty(11 * cp * cp, dtype=torch.bfloat16, device=dev)
st = E._stream(dev)
L.check(lib.vqae_pack_same_block_f16(E._ptr(ws[0]), E._ptr(ws[1]), E._ptr(ws[2]), C, E._ptr(packed), st), "pack")
sc = (ctypes.c_float * 8)(0.01, 0.02, -0.01, 0.03, 0.02, -0.02, 0.01, 0.9)
prof = torch.zeros(148 * 4, 8, dtype=torch.int64, device=dev)
for _ in range(3):
    L.check(L.load_testaids().vqae_same_block_f16_profile(E._ptr(x), E._ptr(y), E._ptr(packed), sc, B, HW, HW, C,
                                             E._ptr(prof), st), "profile")
torch.cuda.synchronize()
p = prof.cpu()
p = p[p[:, 7] > 0]
d = (p[:, 1:] - p[:, :-1]).double()
names = ["P (1st tile)", "G1 (+wait)", "E1 + sync", "P(next)||G2", "G2 remainder", "E2+sync+G3", "E3 + sync"]
print(f"# C={C} HW={HW} batch={B}: {p.shape[0]} CTAs; cycles per phase of each CTA's first tile")
tot = (p[:, 7] - p[:, 0]).double()
for i, n in enumerate(names):
    print(f"{n:14s} median {d[:, i].median():9.0f}  min {d[:, i].min():9.0f}  max {d[:, i].max():9.0f}  share {d[:, i].median() / tot.median():6.1%}")
print(f"{'tile total':12s} median {tot.median():9.0f}")
