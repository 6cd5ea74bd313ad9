"""Launch the fused encoder front end (csrc/mma_front.cu) at the bench shape (timing / ncu).
usage: python profiles/run_front.py [reps=6] [fused=1]"""
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "2d-vq-ae-2_b200")]
import torch  # noqa: E402

import vqae_b200  # noqa: E402
from vqae_b200 import engine as E  # noqa: E402
from vqae_b200 import plan as P  # noqa: E402
from vqae_b200 import synthetic as S  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
E.FRONT_FUSED = (int(sys.argv[2]) if len(sys.argv) > 2 else 1) != 0
dev = torch.device("cuda:0")
m = vqae_b200.build_vqae(n_down=3).eval()
m.load_state_dict(S.make_state_dict(m.state_dict(), seed=1, regime="perturbed"))
enc = m.to(dev).encoder
xs = [S.synthetic_patches_u8(256, 256, 40 + i).to(dev) for i in range(2)]
packed = P.Plan().get((P.flat_blocks(enc.down_layers) + P.flat_blocks(enc.pre_enc_layers))[:2])


def run(x):
    h, used = E.encoder_front(x, enc.in_stem.weight, enc.in_stem.bias, None, None, packed, "fp16")
    return E.run_blocks_nhwc(packed[used:], h, "fp16")


with torch.no_grad():
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
print(f"front end (fused={E.FRONT_FUSED}), batch 256 of 256^2: {us:.1f} us; algorithmic bytes 50 MB in + 268 MB "
      f"out = {318.8e6 / us / 1e3:.0f} GB/s")
