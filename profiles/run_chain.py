"""Launch the persistent 50-block trunk kernel (vqae_same_chain_f16) at the bench shape (for ncu)."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "2d-vq-ae-2_b200")]
import torch  # noqa: E402
from vqae_b200 import _lib as L  # noqa: E402
from vqae_b200 import engine as E  # noqa: E402

B, H, W, C = 256, 32, 32, 64
nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 50
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
lib = L.load()
st = E._stream(dev)
gen = torch.Generator().manual_seed(7)
packs = []
for i in range(nblk):
    ws = [(torch.randn(C, C, k, k, generator=gen) * 0.05).to(dev) for k in (1, 3, 1)]
    pk = torch.empty(11 * C * C, dtype=torch.bfloat16, device=dev)
    L.check(lib.vqae_pack_same_block_f16(E._ptr(ws[0]), E._ptr(ws[1]), E._ptr(ws[2]), C, E._ptr(pk), st), "pack")
    packs.append(pk)
w_all = torch.cat(packs)
scal = torch.tensor([[0.01, 0.02, -0.01, 0.03, 0.02, -0.02, 0.01, 0.2]] * nblk, dtype=torch.float32).to(dev)
fbytes = lib.vqae_same_chain_flag_bytes(nblk, B)
flags = torch.empty(fbytes, dtype=torch.uint8, device=dev)
xs = [torch.randn(B, H, W, C, device=dev) for _ in range(2)]
ys = [torch.empty(B, H, W, C, device=dev) for _ in range(2)]
for i in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(lib.vqae_same_chain_f16(E._ptr(xs[i % 2]), E._ptr(ys[0]), E._ptr(ys[1]), E._ptr(w_all), E._ptr(scal),
                                     E._ptr(flags), fbytes, nblk, B, H, W, C, st), "chain")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"chain {nblk} blocks: {ms:.3f} ms, {ms / nblk * 1e3:.1f} us/block, {2.0 * B * H * W * C * C * 11 * nblk / ms / 1e9:.1f} TFLOP/s")
