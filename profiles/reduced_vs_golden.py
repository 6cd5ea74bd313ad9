"""Reduced-precision (tcgen05) path against the goldens written by the unmodified reference: the
numbers the bars in tests/test_gpu_golden_bf16.py are set from.
usage: python profiles/bf16_vs_golden.py [precision ...]"""
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "2d-vq-ae-2_b200"), str(REPO / "tests"), str(REPO / "oracle")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import helpers as H  # noqa: E402
import vqae_b200  # noqa: E402

DEV = "cuda:0"
precisions = sys.argv[1:] or ["fp16"]
for prec in precisions:
    for tag in sorted(H.MODEL_CASES):
        g = H.golden(tag)
        m, sd, x = H.model_and_state(tag)
        m = vqae_b200.set_precision(m.to(DEV), prec)
        with torch.no_grad():
            (enc,), (idx,), (loss,) = m.encoder(x.to(DEV))
            recon, _ = m(x.to(DEV))
            dec = m.decode_codes(torch.from_numpy(g["idx"].astype(np.int64)).to(DEV))
            _, _, _, ties, z = m.encoder.encode(x.to(DEV), want_latents=True)
        ref_idx = g["idx"].astype(np.int64).reshape(-1)
        idx_np = idx.cpu().numpy().reshape(-1)
        same = idx_np == ref_idx
        z_ref = torch.from_numpy(g["z"])
        z_err = float((z.cpu().reshape(-1, 8) - z_ref).abs().max() / z_ref.abs().max())
        enc_sub = enc.cpu()[:, ::8, ::4, ::4]
        rec_sub = recon.cpu()[:, :, ::8, ::8]
        # positions whose code matches: the float outputs must agree there
        msk = torch.from_numpy(same.reshape(g["idx"].shape))[:, ::4, ::4]
        e_ref = torch.from_numpy(g["enc_sub"])
        e_err_masked = float(((enc_sub - e_ref).abs() * msk[:, None]).max() / e_ref.abs().max())
        print(f"{prec:6s} {tag:26s} agree {same.mean():.4f} z_err {z_err:.2e} "
              f"loss {loss.item():.6f} ref {float(g['loss']):.6f} "
              f"enc_masked {e_err_masked:.2e} recon {H.rel_err(rec_sub, torch.from_numpy(g['recon_sub'])):.2e} "
              f"dec(ref codes) {H.rel_err(dec.cpu()[:, :, ::8, ::8], torch.from_numpy(g['decode_codes_sub'])):.2e} "
              f"gap-weighted: mismatches with gap>1e-2: {int((~same & (g['gap'] > 1e-2)).sum())}, >1e-1: {int((~same & (g['gap'] > 1e-1)).sum())}")
        m.cpu()
        vqae_b200.set_precision(m, "fp32")
