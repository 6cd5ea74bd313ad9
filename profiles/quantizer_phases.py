"""Phase timeline (clock64) of the tcgen05 quantiser: CTA 0, tiles 10..13, one lane per role."""
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "2d-vq-ae-2_b200")]
import torch  # noqa: E402

from vqae_b200 import _lib as L  # noqa: E402
from vqae_b200 import engine as E  # noqa: E402
from vqae_b200.layers.vq import ProjectedEMAVectorQuantizer2d  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
DT = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}[sys.argv[2] if len(sys.argv) > 2 else "fp32"]
pq = ProjectedEMAVectorQuantizer2d(256, C, 1.0, 0.99, 1e-5, 8).eval().to(dev)
x = torch.randn(512, 1024, C, device=dev).to(DT)
packed = pq.packed()
prof = torch.zeros(64, dtype=torch.int64, device=dev)
lib = L.load()
for _ in range(3):
    E.quantize(packed, x, True, True, 512, 1024)
L.load_testaids().vqae_quantize_tc_set_profile(prof.data_ptr())
E.quantize(packed, x, True, True, 512, 1024)
torch.cuda.synchronize()
L.load_testaids().vqae_quantize_tc_set_profile(None)
p = prof.view(4, 16).cpu()
t0 = int(p[0][p[0] > 0].min())
names = ["mma_go", "proj_top", "proj_x_ready", "proj_z_done", "proj_bufs_free", "proj_done",
         "epi_top", "epi_acc_ready", "epi_pass1", "epi_pass2_rel", "epi_exact", "epi_out"]
for t in range(4):
    print(f"tile {10 + t}: " + "  ".join(f"{n}={int(p[t][i]) - t0 if p[t][i] else -1}" for i, n in enumerate(names)))
