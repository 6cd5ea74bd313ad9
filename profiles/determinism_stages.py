"""Repeat every stage of one bf16 encode on a fixed input and count repeats whose output differs
bitwise from the first one (a kernel with an internal race shows up here)."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "2d-vq-ae-2_b200")]
import torch  # noqa: E402
import vqae_b200  # noqa: E402
from vqae_b200 import engine as E  # noqa: E402
from vqae_b200 import synthetic as S  # noqa: E402
from vqae_b200.model import _flat_blocks  # noqa: E402

dev = torch.device("cuda:0")
precision = sys.argv[1] if len(sys.argv) > 1 else "fp16"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
m = vqae_b200.build_vqae(n_down=3).eval()
m.load_state_dict(S.make_state_dict(m.state_dict(), seed=1, regime="perturbed"))
m = m.to(dev)
enc = m.encoder
x = S.synthetic_patches_u8(256, 256, 42).to(dev)
packed = E.pack_blocks(_flat_blocks(enc.down_layers) + _flat_blocks(enc.pre_enc_layers))
pq = enc.vq_layers[0].packed()
names = {0: "same", 1: "down", 2: "up"}


def repeat(label, fn):
    first = fn()
    torch.cuda.synchronize()
    bad, worst = 0, 0
    outs = [fn() for _ in range(reps)]          # back to back
    torch.cuda.synchronize()
    for o in outs:
        d = int((o != first).sum())
        bad += d > 0
        worst = max(worst, d)
    print(f"{label:28s} differing repeats {bad}/{reps}, max differing elements {worst}")
    return first


with torch.no_grad():
    a = repeat("stem_in", lambda: E.stem_in(x, enc.in_stem.weight, enc.in_stem.bias))
    runs = dict(E._chain_runs(packed, a.shape[1], a.shape[2], a.shape[0])) if precision == "fp16" else {}
    i = 0
    while i < len(packed):
        pk = packed[i]
        if i in runs:
            j = runs[i]
            a = repeat(f"run {j - i}x same C{pk.c_in}", lambda: E.run_blocks_nhwc(packed[i:j], a, precision, {}))
            i = j
        else:
            a = repeat(f"{i} {names[pk.mode]} C{pk.c_in}->{pk.c_out} @{a.shape[1]}",
                       lambda: E.fixup_forward_nhwc(pk, a, precision=precision))
            i += 1
    b, hh, ww, c = a.shape
    repeat("quantize (codes)", lambda: E.quantize(pq, a, True, True, b, hh * ww, want_out=False)[1])
