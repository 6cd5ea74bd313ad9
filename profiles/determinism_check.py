"""Run-to-run reproducibility probe: hashes of every intermediate of one bf16 encode of 256 fixed
patches.  Run twice (two processes) and diff the output; any line that differs names the first
kernel whose result depends on something other than its inputs."""
import hashlib
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "2d-vq-ae-2_b200"), str(REPO / "profiles")]
import torch  # noqa: E402
import vqae_b200  # noqa: E402
from vqae_b200 import engine as E  # noqa: E402
from vqae_b200 import synthetic as S  # noqa: E402
from vqae_b200.model import _flat_blocks  # noqa: E402

dev = torch.device("cuda:0")
precision = sys.argv[1] if len(sys.argv) > 1 else "fp16"
# leave different garbage in freed device memory from run to run
junk = torch.empty(int(sys.argv[2]) if len(sys.argv) > 2 else 1, 1 << 20, device=dev).normal_()
del junk


def h(t):
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()[:12]


m = vqae_b200.build_vqae(n_down=3).eval()
m.load_state_dict(S.make_state_dict(m.state_dict(), seed=1, regime="perturbed"))
m = m.to(dev)
enc = m.encoder
x = S.synthetic_patches_u8(256, 256, 42).to(dev)
print("input", h(x))
packed = E.pack_blocks(_flat_blocks(enc.down_layers) + _flat_blocks(enc.pre_enc_layers))
pq = enc.vq_layers[0].packed()
names = {0: "same", 1: "down", 2: "up"}
with torch.no_grad():
    a = E.stem_in(x, enc.in_stem.weight, enc.in_stem.bias)
    print("stem_in", h(a))
    runs = dict(E._chain_runs(packed, a.shape[1], a.shape[2], a.shape[0])) if precision == "fp16" else {}
    i = 0
    while i < len(packed):
        pk = packed[i]
        if i in runs:
            j = runs[i]
            a = E.run_blocks_nhwc(packed[i:j], a, precision, {})
            print(f"run {j - i}x same C{pk.c_in}", h(a))
            i = j
        else:
            a = E.fixup_forward_nhwc(pk, a, precision=precision)
            print(f"{i} {names[pk.mode]} C{pk.c_in}->{pk.c_out} @{a.shape[1]}", h(a))
            i += 1
    b, hh, ww, c = a.shape
    out, idx, loss, ties, z = E.quantize(pq, a, True, True, b, hh * ww, want_out=False)
    print("codes", h(idx), "loss", float(loss))
