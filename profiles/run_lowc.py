"""Launch one low-channel 'same' block (mma_same.cu or the tcgen05 tile kernel) at the bench shape.
usage: python profiles/run_lowc.py C HW [mma|tc|split] [reps]      (split = precision "fp32tc", tc_split.cu)"""
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "2d-vq-ae-2_b200"), str(REPO / "tests")]
import torch  # noqa: E402

from vqae_b200 import engine as E  # noqa: E402

C, HW = int(sys.argv[1]), int(sys.argv[2])
kind = sys.argv[3] if len(sys.argv) > 3 else "mma"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
B = 256
dev = torch.device("cuda:0")
from vqae_b200.config import pre_activation_fixup  # noqa: E402
from vqae_b200.layers.conv_block import PreActFixupResBlock  # noqa: E402
from vqae_b200 import synthetic as S  # noqa: E402
conf = pre_activation_fixup(n_layers=12)
for k in ("_target_", "_recursive_", "in_channels", "out_channels", "mode"):
    conf.pop(k)
blk = PreActFixupResBlock(in_channels=C, out_channels=C, mode="same", **conf).eval()
blk.load_state_dict(S.make_state_dict(blk.state_dict(), seed=3, regime="perturbed", n_layers=12))
pk = blk.to(dev).packed()
E.LOWC_MMA = {8, 16, 32} if kind == "mma" else set()
prec = "fp32tc" if kind == "split" else "fp16"
xs = [torch.randn(B, HW, HW, C, device=dev) for _ in range(2)]
ys = [torch.empty_like(xs[0]) for _ in range(2)]
for i in range(3):
    E.fixup_forward_nhwc(pk, xs[i % 2], out=ys[i % 2], precision=prec)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for i in range(reps):
    E.fixup_forward_nhwc(pk, xs[i % 2], out=ys[i % 2], precision=prec)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
byts = 2 * xs[0].numel() * 4
print(f"C={C} HW={HW} {kind}: {us:.1f} us per block, {byts / us / 1e3:.0f} GB/s of algorithmic traffic")
