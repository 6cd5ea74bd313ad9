"""Repeat the tcgen05 quantiser on fixed latents and compare every repeat with the CUDA-core
quantiser (exact reference of the same arithmetic): counts wrong codes per repeat."""
import os
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "2d-vq-ae-2_b200")]
import torch  # noqa: E402
import vqae_b200  # noqa: E402
from vqae_b200 import _lib as L  # noqa: E402
from vqae_b200 import engine as E  # noqa: E402
from vqae_b200 import synthetic as S  # noqa: E402

dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
m = vqae_b200.build_vqae(n_down=3).eval()
m.load_state_dict(S.make_state_dict(m.state_dict(), seed=1, regime="perturbed"))
m = vqae_b200.set_precision(m.to(dev), "fp16")
enc = m.encoder
x = S.synthetic_patches_u8(256, 256, 42).to(dev)
with torch.no_grad():
    _, _, _, _, z = enc.encode(x, want_quantized=False, want_latents=True)
pq = enc.vq_layers[0].packed()
# latents = input of the quantiser: recompute them through the plan
from vqae_b200.model import _flat_blocks  # noqa: E402
with torch.no_grad():
    h = E.stem_in(x, enc.in_stem.weight, enc.in_stem.bias)
    h = E.run_blocks_nhwc(E.pack_blocks(_flat_blocks(enc.down_layers) + _flat_blocks(enc.pre_enc_layers)), h, "fp16")
b, hh, ww, c = h.shape
os.environ["VQAE_QUANT_TC"] = "0"
_r = E.quantize(pq, h, True, True, b, hh * ww, want_out=False, want_z=True)
ref, zref = _r[1].clone(), _r[4].clone()
del os.environ["VQAE_QUANT_TC"]
torch.cuda.synchronize()
for mode in ("back to back", "synchronised"):
    outs, zs = [], []
    for _ in range(reps):
        _o = E.quantize(pq, h, True, True, b, hh * ww, want_out=False, want_z=True)
        outs.append(_o[1])
        zs.append(_o[4])
        if mode == "synchronised":
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    bad = [(i, int((o != ref).sum())) for i, o in enumerate(outs) if not torch.equal(o, ref)]
    print(f"{mode}: {len(bad)}/{reps} repeats differ from the CUDA-core quantiser; (repeat, wrong codes):", bad[:12])
    if bad:
        o = outs[bad[0][0]]
        pos = torch.nonzero(o != ref).flatten()[:4].tolist()
        print("  first differing vectors:", pos, "tc:", o[pos].tolist(), "ref:", ref[pos].tolist())
        zt = zs[bad[0][0]]
        emb = pq.embed
        for q in pos:
            d_tc = float(((zref[q] - emb[o[q]]) ** 4).sum()); d_ref = float(((zref[q] - emb[ref[q]]) ** 4).sum())
            print(f"   vector {q} (tile row {q % 128}): z identical to reference: {bool(torch.equal(zt[q], zref[q]))}, "
                  f"L4 of tc code {d_tc:.6g} vs of reference code {d_ref:.6g}")
