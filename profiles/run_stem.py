"""Launch the stem kernels at the bench shape (timing / ncu).  usage: python profiles/run_stem.py [in|out] [reps]"""
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "2d-vq-ae-2_b200")]
import torch  # noqa: E402

import vqae_b200  # noqa: E402
from vqae_b200 import engine as E  # noqa: E402
from vqae_b200 import synthetic as S  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "in"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda:0")
m = vqae_b200.build_vqae(n_down=3).eval()
m.load_state_dict(S.make_state_dict(m.state_dict(), seed=1, regime="perturbed"))
m = m.to(dev)
if which == "in":
    xs = [S.synthetic_patches_u8(256, 256, 40 + i).to(dev) for i in range(2)]
    run = lambda x: E.stem_in(x, m.encoder.in_stem.weight, m.encoder.in_stem.bias, precision="fp16")
    byts = 256 * 65536 * (3 + 32)
else:
    xs = [torch.randn(256, 256, 256, 8, device=dev) for _ in range(2)]
    run = lambda x: E.stem_out(x, m.decoder.out_stem.weight, m.decoder.out_stem.bias, False, "fp16")
    byts = 256 * 65536 * (32 + 12)
for i in range(3):
    run(xs[i % 2])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for i in range(reps):
    run(xs[i % 2])
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
print(f"stem_{which} (split-operand MMA kernel), batch 256 of 256^2: {us:.1f} us, {byts / us / 1e3:.0f} GB/s of algorithmic traffic")
