"""Launch the image-resident trunk kernel (vqae_trunk_resident_f16) at the bench shape (timing / ncu).
usage: python profiles/run_resident.py [n_blocks=54] [reps=3] [batch=256] [C=64]"""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "2d-vq-ae-2_b200")]
import torch  # noqa: E402
from vqae_b200 import _lib as L  # noqa: E402
from vqae_b200 import engine as E  # noqa: E402

nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 54
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
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
w_all = torch.cat(packs)
scal = torch.tensor([[0.01, 0.02, -0.01, 0.03, 0.02, -0.02, 0.01, 0.2]] * nblk, dtype=torch.float32).to(dev)
import types  # noqa: E402
chain = types.SimpleNamespace(weights=w_all, scalars=scal, n=nblk)
print("resident clusters per device:", lib.vqae_trunk_resident_max_clusters())
xs = [torch.randn(B, H, W, C, device=dev) for _ in range(2)]
y = torch.empty(B, H, W, C, device=dev)
for i in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    E.trunk_resident(xs[i % 2], y, chain)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"resident {nblk} blocks, batch {B}: {ms:.3f} ms, {ms / nblk * 1e3:.1f} us/block, "
          f"{2.0 * B * H * W * C * C * 11 * nblk / ms / 1e9:.1f} TFLOP/s")

# phase clocks of CTA 0, eight steady-state half-rounds
prof = torch.zeros(8 * 32, dtype=torch.int64, device=dev)
L.load_testaids().vqae_trunk_resident_set_profile(E._ptr(prof))
E.trunk_resident(xs[0], y, chain)
torch.cuda.synchronize()
L.load_testaids().vqae_trunk_resident_set_profile(None)
p = prof.cpu().view(8, 32)
t0 = int(p[0, 0])
print("issuer (slot 0, M-tile 0), even rows: step_start  A1wait_done  G1_issued  Uwait_done  taps_issued  Vwait_done  G3_issued")
print("workers, every row (half-round): start  E2m0_go  E2m1_go  E2_done  Pm0_go  Pm1_go  P_done  E1m0_go  E1m1_go  E1_done")
for r in range(8):
    if r % 2 == 0:
        print("step%2d issuer " % (r // 2), " ".join("%7d" % (int(v) - t0) for v in p[r, 0:7]),
              "| G2_complete %d, cumulative ring-wait cycles %d" % (int(p[r, 7]) - t0, int(p[r, 20])))
        print("        tap issue starts (after halo waits):", " ".join("%7d" % (int(v) - t0) for v in p[r, 23:32]))
        print("        pusher (slot 0): U_seen  free_waits_done  pushes_issued:",
              " ".join("%7d" % (int(v) - t0) for v in (p[r, 18], p[r, 19], p[r, 21])))
    print("   hr%2d workers" % r, " ".join("%7d" % (int(v) - t0) for v in p[r, 8:18]))

