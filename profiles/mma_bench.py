"""tcgen05.mma issue-rate microbenchmark: cycles per 128xNx16 bf16 MMA from shared memory."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "2d-vq-ae-2_b200")]
import torch  # noqa: E402
from vqae_b200 import _lib as L  # noqa: E402
from vqae_b200 import engine as E  # noqa: E402
lib = L.load()
dev = torch.device("cuda:0")
out = torch.zeros(2, dtype=torch.int64, device=dev)
print("N layout n_acc reps cycles/MMA ideal(128*N/256)")
for n in (64, 128, 256):
    for layout in (0, 2):
        for nacc in (1, 2, 4, 8):
            if nacc * n > 512:
                continue
            reps = 1000
            L.check(L.load_testaids().vqae_tc_mma_bench(n, layout | (nacc << 4), reps, 128, E._ptr(out), E._stream(dev)), "bench")
            torch.cuda.synchronize()
            c, r = out.tolist()
            print(f"{n:4d} {layout:3d} {nacc:5d} {reps:5d} {c / r:8.1f} {128 * n / 256:6.0f}")
