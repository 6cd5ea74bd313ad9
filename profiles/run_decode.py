"""One bf16 decode of 64 code grids (256-model) -- target for ncu captures of the decoder kernels."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "2d-vq-ae-2_b200")]
import torch  # noqa: E402
import vqae_b200  # noqa: E402
from vqae_b200 import synthetic as S  # noqa: E402

dev = torch.device("cuda:0")
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
m = vqae_b200.build_vqae(n_down=3).eval()
m.load_state_dict(S.make_state_dict(m.state_dict(), seed=1, regime="perturbed"))
m = vqae_b200.set_precision(m.to(dev), "fp16")
enc = torch.randn(batch, 64, 32, 32, device=dev)
with torch.no_grad():
    for _ in range(3):
        out = m.decoder((enc,))
torch.cuda.synchronize()
print("decoded", tuple(out.shape))
