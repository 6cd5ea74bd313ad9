import sys
sys.path[:0] = ["/root/repo/2d-vq-ae-2_b200"]
import torch
from vqae_b200 import _lib as L
from vqae_b200 import engine as E
dev = torch.device("cuda:0"); lib = L.load(); st = E._stream(dev)
C, HW, B, nblk = 32, 64, 256, 5
gen = torch.Generator().manual_seed(7)
packs = []
for i in range(nblk):
    ws = [(torch.randn(C, C, k, k, generator=gen) * 0.05).to(dev) for k in (1, 3, 1)]
    pk = torch.empty(11 * C * C, dtype=torch.bfloat16, device=dev)
    L.check(lib.vqae_pack_resident_block_f16(E._ptr(ws[0]), E._ptr(ws[1]), E._ptr(ws[2]), C, 0.2, E._ptr(pk), st), "pack")
    packs.append(pk)
w_all = torch.cat(packs)
scal = torch.tensor([[0.01, 0.02, -0.01, 0.03, 0.02, -0.02, 0.01, 0.2]] * nblk, dtype=torch.float32).to(dev)
x = torch.randn(B, HW, HW, C, device=dev)
outs = []
for i in range(401):
    y = torch.empty_like(x)
    E.trunk_resident(x, y, chain)
    outs.append(y)
torch.cuda.synchronize()
for i, o in enumerate(outs[1:], 1):
    if not torch.equal(o, outs[0]):
        d = torch.nonzero(o != outs[0])
        imgs = sorted(set(d[:, 0].tolist()))
        print("launch", i, "differing elements", d.shape[0], "images", imgs)
        for im in imgs:
            dd = d[d[:, 0] == im]
            px = sorted(set((int(r), int(c)) for r, c in dd[:, 1:3].tolist()))
            print("  image", im, "pixels (row, col):", px, "channels", sorted(set(dd[:, 3].tolist()))[:40])
            print("  max abs diff", float((o[im] - outs[0][im]).abs().max()))
