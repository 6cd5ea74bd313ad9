"""tcgen05.mma throughput per SM vs issuing warps per CTA and CTAs per SM (vqae_tc_mma_bench2)."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "2d-vq-ae-2_b200")]
import torch  # noqa: E402
from vqae_b200 import _lib as L  # noqa: E402
from vqae_b200 import engine as E  # noqa: E402
lib = L.load()
dev = torch.device("cuda:0")
out = torch.zeros(148 * 4, dtype=torch.int64, device=dev)
reps = 2000
print("M N issuers ctas/SM mode(1: own A/B per issuer, 2: 160 B row groups) : cycles per MMA per issuer | per SM")
for m, n in ((128, 64), (128, 16), (128, 128), (64, 64), (64, 128), (64, 256), (128, 256)):
    for iss, cps in ((1, 1), (2, 1), (4, 1), (1, 2), (2, 2), (1, 4)):
        for mode in (0, 1, 3, 7):
            rc = L.load_testaids().vqae_tc_mma_bench2(m, n, reps, iss, cps, mode, E._ptr(out), E._stream(dev))
            if rc != 0:
                continue
            torch.cuda.synchronize()
            c = int(out[: (148 if mode < 4 else 144) * cps].max())
            print(f"M={m} N={n} issuers={iss} ctas/SM={cps} mode={mode}: {c / reps:.1f} | "
                  f"{c / reps / (iss * cps):.1f}")
