"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python oracle/make_golden.py

The reference has no tests or golden vectors of its own, so the pin is
the reference's own nn.Modules, imported through oracle/ref_shim.py, run on deterministic
synthetic weights (vqae_b200/synthetic.py: values depend only on state_dict key + seed) and
seeded inputs.  Every array written here is an output of reference code (torch 2.11 CPU
kernels; the reference pins torch 1.11) -- the oracle restatement and the CUDA path are both
tested against these files, which travel to the GPU box where /root/reference does not exist.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
REPO = HERE.parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(REPO / "2d-vq-ae-2_b200"))

import ref_shim  # noqa: E402
from vqae_b200 import synthetic as S  # noqa: E402
from vqae_b200.config import (compose_efficientnetv2_conf, compose_multilevel_conf,  # noqa: E402
                              compose_vqae_conf, mbconv, pre_activation_fixup)

GOLDEN = REPO / "tests" / "golden"


def _np(t):
    return t.detach().cpu().numpy().copy()   # copy: buffers are overwritten in place later


def top2_gap(z_flat: torch.Tensor, embed: torch.Tensor):
    """relative top-2 gap of the un-rooted L4 sums, from the reference's own cdist."""
    d = torch.cdist(z_flat, embed, 4, compute_mode='donot_use_mm_for_euclid_dist') ** 4
    t2 = torch.topk(d, 2, dim=1, largest=False).values
    return ((t2[:, 1] - t2[:, 0]) / t2[:, 1].clamp_min(1e-30)).float()


def stats(t: torch.Tensor) -> np.ndarray:
    t = t.double()
    return np.array([t.mean().item(), t.std().item(), t.abs().sum().item()])


@torch.no_grad()
def quantizer_cases(vq_mod):
    out = {}
    # (1) bare EMAVectorQuantizer on [4,8,16,16], constructor-default randn codebook
    torch.manual_seed(42)
    q = vq_mod.EMAVectorQuantizer(256, 8, 1.0, 0.99, 1e-5).eval()
    g = torch.Generator().manual_seed(100)
    q.embed.copy_(torch.randn(256, 8, generator=g))
    x = torch.randn(4, 8, 16, 16, generator=g)
    quant, idx, loss = q(x)
    out.update(bare_embed=_np(q.embed), bare_x=_np(x), bare_idx=_np(idx).astype(np.int16),
               bare_quant=_np(quant), bare_loss=_np(loss),
               bare_gap=_np(top2_gap(x.permute(0, 2, 3, 1).reshape(-1, 8), q.embed)))
    # (2) tie case: duplicated codebook rows -> lowest index wins (torch.argmin)
    emb2 = q.embed.clone()
    emb2[200] = emb2[3]
    emb2[77] = emb2[3]
    q.embed.copy_(emb2)
    quant2, idx2, loss2 = q(x)
    out.update(tie_embed=_np(emb2), tie_idx=_np(idx2).astype(np.int16), tie_loss=_np(loss2))
    # (3) channels_last input: output strides follow input
    xcl = x.contiguous(memory_format=torch.channels_last)
    qcl, icl, _ = q(xcl)
    out.update(tie_cl_is_channels_last=np.array(
        qcl.is_contiguous(memory_format=torch.channels_last) and not qcl.is_contiguous()),
        tie_cl_idx_equal=np.array(bool((icl == idx2).all())))
    # (4) projected quantiser, C = 64 and C = 128
    for c in (64, 128):
        pq = vq_mod.ProjectedEMAVectorQuantizer2d(256, c, 1.0, 0.99, 1e-5, 8).eval()
        sd = S.make_state_dict(pq.state_dict(), seed=5, regime="perturbed")
        pq.load_state_dict(sd)
        g = torch.Generator().manual_seed(200 + c)
        x = torch.randn(2, c, 32, 32, generator=g)
        z = pq.proj_in(x).permute(0, 2, 3, 1).reshape(-1, 8)
        # codebook rescaled to the latent statistics (what _init_ema would do, vq.py:76-94)
        pq.embed.copy_(S.rescale_codebook(sd["embed"], z))
        quant, idx, loss = pq(x)
        out.update({f"proj{c}_embed": _np(pq.embed), f"proj{c}_idx": _np(idx).astype(np.int16),
                    f"proj{c}_loss": _np(loss), f"proj{c}_quant_sub": _np(quant[:, :, ::4, ::4]),
                    f"proj{c}_quant_stats": stats(quant), f"proj{c}_z": _np(z),
                    f"proj{c}_gap": _np(top2_gap(z, pq.embed))})
    return out


@torch.no_grad()
def block_cases(cb_mod):
    out = {}
    conf = pre_activation_fixup(n_layers=12)
    conf.pop("_target_"), conf.pop("_recursive_"), conf.pop("in_channels"), conf.pop(
        "out_channels"), conf.pop("mode")
    for name, (cin, cout, mode, hw) in {
        "same16": (16, 16, "same", 16), "same64": (64, 64, "same", 16),
        "down8": (8, 16, "down", 16), "down32": (32, 64, "down", 8),
        "up16": (16, 8, "up", 8), "up64": (64, 32, "up", 4),
    }.items():
        blk = cb_mod.PreActFixupResBlock(in_channels=cin, out_channels=cout, mode=mode,
                                         **conf).eval()
        blk.load_state_dict(S.make_state_dict(blk.state_dict(), seed=11, regime="perturbed",
                                              n_layers=12))
        g = torch.Generator().manual_seed(300 + cin)
        x = torch.randn(2, cin, hw, hw, generator=g)
        out[f"{name}_x"] = _np(x)
        out[f"{name}_y"] = _np(blk(x))
    return out


# block cases at sizes the tcgen05 kernels tile (W % 32 == 0, H % 16 == 0): name -> (cin, cout, mode,
# hw).  x is regenerated from its seed by the tests (helpers.tc_block_input); y is stored on a
# sub-grid that contains both image borders.
TC_BLOCK_CASES = {
    "same8": (8, 8, "same", 32), "same16": (16, 16, "same", 32), "same32": (32, 32, "same", 32),
    "same64": (64, 64, "same", 32), "same128": (128, 128, "same", 32),
    "down8": (8, 16, "down", 32), "down16": (16, 32, "down", 32), "down32": (32, 64, "down", 32),
    "down64": (64, 128, "down", 32),
    "up16": (16, 8, "up", 16), "up32": (32, 16, "up", 16), "up64": (64, 32, "up", 16),
    "up128": (128, 64, "up", 16),
}


def tc_sub_index(n: int):
    return sorted(set(range(0, n, 3)) | {1, n - 2, n - 1})


def tc_block_input(name: str) -> torch.Tensor:
    cin, _, _, hw = TC_BLOCK_CASES[name]
    g = torch.Generator().manual_seed(700 + cin + {"same": 0, "down": 1, "up": 2}[TC_BLOCK_CASES[name][2]])
    return torch.randn(2, cin, hw, hw, generator=g)


@torch.no_grad()
def block_cases_tc(cb_mod):
    out = {}
    conf = pre_activation_fixup(n_layers=12)
    for k in ("_target_", "_recursive_", "in_channels", "out_channels", "mode"):
        conf.pop(k)
    for name, (cin, cout, mode, hw) in TC_BLOCK_CASES.items():
        blk = cb_mod.PreActFixupResBlock(in_channels=cin, out_channels=cout, mode=mode,
                                         **conf).eval()
        blk.load_state_dict(S.make_state_dict(blk.state_dict(), seed=13, regime="perturbed",
                                              n_layers=12))
        x = tc_block_input(name)
        y = blk(x)
        ii = torch.tensor(tc_sub_index(y.shape[-1]))
        out[f"{name}_y_sub"] = _np(y[:, :, ii][:, :, :, ii])
        out[f"{name}_y_stats"] = stats(y)
        out[f"{name}_x_stats"] = stats(x)         # guards the regenerated input
        out[f"{name}_branch_absmax"] = np.array(float((y - x).abs().max()) if mode == "same"
                                                else float(y.abs().max()))
    return out


@torch.no_grad()
def model_case(model_mod, n_down: int, regime: str, batch: int, size: int, seed: int):
    conf = compose_vqae_conf(n_down=n_down)
    conf.pop("_target_"), conf.pop("_recursive_")
    torch.manual_seed(42)
    m = model_mod.VQAE(**conf).eval()
    sd = S.make_state_dict(m.state_dict(), seed=seed, regime=regime)
    m.load_state_dict(sd)
    x = S.synthetic_patches(batch, size, seed + 1000)
    enc = m.encoder
    vq = enc.vq_layers[0]
    h = enc.pre_enc_layers[0](enc.down_layers[0](enc.in_stem(x)))
    z = vq.proj_in(h).permute(0, 2, 3, 1).reshape(-1, 8)
    out = {}
    if regime == "perturbed":
        vq.embed.copy_(S.rescale_codebook(sd["encoder.vq_layers.0.embed"], z))
        vq.embed_avg.copy_(vq.embed)
    (e,), (idx,), (loss,) = enc(x)
    recon, (loss2,) = m(x)
    assert torch.equal(loss, loss2)
    out.update(
        embed=_np(vq.embed), x_stats=stats(x), pre_vq_sub=_np(h[:, ::8, ::4, ::4]),
        pre_vq_stats=stats(h), z=_np(z), gap=_np(top2_gap(z, vq.embed)),
        idx=_np(idx).astype(np.int16), loss=_np(loss), enc_sub=_np(e[:, ::8, ::4, ::4]),
        enc_stats=stats(e), recon_sub=_np(recon[:, :, ::8, ::8]), recon_stats=stats(recon),
        in_stem_sub=_np(enc.in_stem(x)[:, :, ::16, ::16]),
        n_params=np.array(sum(p.numel() for p in m.parameters())),
        n_state=np.array(len(sd)), codes_used=np.array(idx.unique().numel()),
    )
    # decode-from-codes (embed_code -> proj_out -> decoder), scope row f-2
    q = vq.embed_code(idx).permute(0, 3, 1, 2)
    dec = m.decoder((vq.proj_out(q),))
    out["decode_codes_sub"] = _np(dec[:, :, ::8, ::8])
    return out


@torch.no_grad()
def _settle_codebooks(enc, x, sd):
    """Rescale every level's codebook to the statistics of ITS latents (what _init_ema would do on the
    first training batch).  A level's latents depend on the codebooks below it, so levels settle one
    forward at a time, lowest first; latents are read with forward hooks on the reference's proj_in."""
    zs = {}
    hooks = [vq.proj_in.register_forward_hook(
        lambda _m, _i, o, i=i: zs.__setitem__(i, o.permute(0, 2, 3, 1).reshape(-1, o.shape[1]).clone()))
        for i, vq in enumerate(enc.vq_layers)]
    for i, vq in enumerate(enc.vq_layers):                  # low-res first = dependency order
        enc(x)
        vq.embed.copy_(S.rescale_codebook(sd[f"encoder.vq_layers.{i}.embed"], zs[i]))
        vq.embed_avg.copy_(vq.embed)
    out = enc(x)
    for h in hooks:
        h.remove()
    return out, zs


@torch.no_grad()
def multilevel_cases(model_mod):
    """Scope row f-4: the multi-level hierarchy of model.py:144-187, 203-215, 274-291 run by the reference's
    own Encoder / Decoder classes on configurations written in its configuration language."""
    out = {}
    # (1) a real hierarchy: 32 ch @ 64x64 above 64 ch @ 32x32, 'up' shortcut block between them
    conf = compose_multilevel_conf(level_downs=(2, 1), n_pre_enc_layers=(2, 3), shortcut_mode="up")
    econf = dict(conf["encoder_conf"])
    econf.pop("_target_"), econf.pop("_recursive_")
    torch.manual_seed(42)
    enc = model_mod.Encoder(**econf).eval()
    sd = S.make_state_dict({"encoder." + k: v for k, v in enc.state_dict().items()}, seed=11,
                           regime="perturbed")
    enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items()})
    x = S.synthetic_patches(2, 256, 1011)
    (encs, idxs, losses), zs = _settle_codebooks(enc, x, sd)
    for i, (e, idx, loss) in enumerate(zip(encs, idxs, losses)):
        vq = enc.vq_layers[i]
        out.update({f"hier_embed{i}": _np(vq.embed), f"hier_idx{i}": _np(idx).astype(np.int16),
                    f"hier_loss{i}": _np(loss), f"hier_enc_sub{i}": _np(e[:, ::8, ::4, ::4]),
                    f"hier_enc_stats{i}": stats(e), f"hier_z{i}": _np(zs[i]),
                    f"hier_gap{i}": _np(top2_gap(zs[i], vq.embed)),
                    f"hier_codes_used{i}": np.array(idx.unique().numel())})
    out["hier_n_state"] = np.array(len(sd))
    # (2) a full VQAE with two levels of equal width (the only shape the reference's Decoder accepts)
    conf = compose_multilevel_conf(level_downs=(3, 0), n_pre_enc_layers=(2, 2), shortcut_mode="same")
    conf.pop("_target_"), conf.pop("_recursive_")
    torch.manual_seed(42)
    m = model_mod.VQAE(**conf).eval()
    sd = S.make_state_dict(m.state_dict(), seed=12, regime="perturbed")
    m.load_state_dict(sd)
    x = S.synthetic_patches(2, 256, 1012)
    (encs, idxs, losses), zs = _settle_codebooks(m.encoder, x, sd)
    recon, losses2 = m(x)
    assert all(torch.equal(a, b) for a, b in zip(losses, losses2))
    for i, (e, idx, loss) in enumerate(zip(encs, idxs, losses)):
        vq = m.encoder.vq_layers[i]
        out.update({f"flat_embed{i}": _np(vq.embed), f"flat_idx{i}": _np(idx).astype(np.int16),
                    f"flat_loss{i}": _np(loss), f"flat_enc_sub{i}": _np(e[:, ::8, ::4, ::4]),
                    f"flat_z{i}": _np(zs[i]), f"flat_gap{i}": _np(top2_gap(zs[i], vq.embed))})
    out.update(flat_recon_sub=_np(recon[:, :, ::8, ::8]), flat_recon_stats=stats(recon),
               flat_n_state=np.array(len(sd)))
    return out


def ema_training_cases(vq_mod):
    """Scope row f-4: TRAINING-mode forwards of the reference's quantisers (vq.py:47-94, 118-133): the
    first pass initialises the codebook from the batch, every pass updates the EMA buffers."""
    out = {}
    for tag, c in (("bare", 8), ("proj", 64)):
        torch.manual_seed(7)
        q = (vq_mod.EMAVectorQuantizer(256, 8, 0.25, 0.99, 1e-5) if tag == "bare"
             else vq_mod.ProjectedEMAVectorQuantizer2d(256, c, 0.25, 0.99, 1e-5, 8))
        sd = S.make_state_dict(q.state_dict(), seed=21, regime="perturbed")
        q.load_state_dict(sd)
        q.train()
        g = torch.Generator().manual_seed(300 + c)
        out[f"{tag}_embed0"] = _np(q.embed)
        for step in range(3):
            hw = 32 if tag == "bare" else 16
            # fp16-representable values, stored as fp16 (exact, half the fixture size)
            x = (torch.randn(2, c, hw, hw, generator=g) * (1.0 + 0.5 * step) + 0.3 * step).half().float()
            with torch.no_grad():
                quant, idx, loss = q(x)
            out.update({f"{tag}_x{step}": _np(x.half()), f"{tag}_idx{step}": _np(idx).astype(np.int16),
                        f"{tag}_loss{step}": _np(loss), f"{tag}_quant_stats{step}": stats(quant),
                        f"{tag}_embed_after{step}": _np(q.embed),
                        f"{tag}_embed_avg_after{step}": _np(q.embed_avg),
                        f"{tag}_cluster_size_after{step}": _np(q.cluster_size),
                        f"{tag}_first_pass_after{step}": _np(q.first_pass)})
    return out


MBCONV_BLOCK_CASES = {                   # name -> (c_in, c_out, mode, hw, batchnorm, se)
    "same16": (16, 16, "same", 16, True, True), "same8to16": (8, 16, "same", 12, True, True),
    "down16": (16, 32, "down", 16, True, True), "up32": (32, 16, "up", 8, True, True),
    "same16_plain": (16, 16, "same", 8, False, False), "down8_nose": (8, 16, "down", 8, True, False),
}


@torch.no_grad()
def mbconv_cases(model_mod, cb_mod):
    """Scope row f-4: the reference's MBConv blocks (conv_block.py:240-321) one by one, and a whole VQAE
    built from them (conf/model/{encoder,decoder}/efficientnetv2.yaml)."""
    out = {}
    for name, (cin, cout, mode, hw, use_bn, use_se) in MBCONV_BLOCK_CASES.items():
        conf = mbconv(batchnorm=use_bn, se=use_se)
        for k in ("_target_", "_recursive_", "in_channels", "out_channels", "mode"):
            conf.pop(k)
        torch.manual_seed(42)
        blk = cb_mod.MBConv(in_channels=cin, out_channels=cout, mode=mode, **conf).eval()
        blk.load_state_dict(S.make_state_dict(blk.state_dict(), seed=41, regime="perturbed"))
        gen = torch.Generator().manual_seed(500 + cin + hw)
        x = torch.randn(2, cin, hw, hw, generator=gen)
        out[f"blk_{name}_x"] = _np(x)
        out[f"blk_{name}_y"] = _np(blk(x))
    conf = compose_efficientnetv2_conf(n_down=3, n_enc_layers_trunk=3)
    conf.pop("_target_"), conf.pop("_recursive_")
    torch.manual_seed(42)
    m = model_mod.VQAE(**conf).eval()
    sd = S.make_state_dict(m.state_dict(), seed=31, regime="perturbed")
    m.load_state_dict(sd)
    x = S.synthetic_patches(2, 256, 1031)
    ((e,), (idx,), (loss,)), zs = _settle_codebooks(m.encoder, x, sd)
    recon, _ = m(x)
    vq = m.encoder.vq_layers[0]
    out.update(model_embed=_np(vq.embed), model_idx=_np(idx).astype(np.int16), model_loss=_np(loss),
               model_z=_np(zs[0]), model_gap=_np(top2_gap(zs[0], vq.embed)),
               model_enc_sub=_np(e[:, ::8, ::4, ::4]), model_recon_sub=_np(recon[:, :, ::8, ::8]),
               model_recon_stats=stats(recon), model_n_state=np.array(len(sd)),
               model_codes_used=np.array(idx.unique().numel()))
    return out


def main():
    if not ref_shim.reference_available():
        raise SystemExit("reference checkout not found; run this in the build container")
    model_mod, vq_mod, cb_mod, _ = ref_shim.load_reference()
    GOLDEN.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    if "--only-f4" in sys.argv:                      # the round-2 additions only (scope row f-4)
        np.savez_compressed(GOLDEN / "multilevel.npz", **multilevel_cases(model_mod))
        np.savez_compressed(GOLDEN / "ema_training.npz", **ema_training_cases(vq_mod))
        np.savez_compressed(GOLDEN / "mbconv.npz", **mbconv_cases(model_mod, cb_mod))
        return

    np.savez_compressed(GOLDEN / "quantizer.npz", **quantizer_cases(vq_mod))
    np.savez_compressed(GOLDEN / "blocks.npz", **block_cases(cb_mod))
    np.savez_compressed(GOLDEN / "blocks_tc.npz", **block_cases_tc(cb_mod))
    for tag, (n_down, regime, batch, size, seed) in {
        "model_nd3_perturbed": (3, "perturbed", 2, 256, 1),
        "model_nd3_fixup": (3, "fixup", 2, 256, 2),
        "model_nd4_perturbed_256": (4, "perturbed", 2, 256, 3),   # as-shipped conf: 16x16 grid
        "model_nd4_perturbed_512": (4, "perturbed", 1, 512, 4),   # 512^2 -> 32x32 grid
    }.items():
        np.savez_compressed(GOLDEN / f"{tag}.npz",
                            **model_case(model_mod, n_down, regime, batch, size, seed),
                            meta=np.array([n_down, batch, size, seed]))
        print("wrote", tag)
    np.savez_compressed(GOLDEN / "multilevel.npz", **multilevel_cases(model_mod))
    np.savez_compressed(GOLDEN / "ema_training.npz", **ema_training_cases(vq_mod))
    np.savez_compressed(GOLDEN / "mbconv.npz", **mbconv_cases(model_mod, cb_mod))
    for f in sorted(GOLDEN.glob("*.npz")):
        print(f.name, f.stat().st_size // 1024, "KiB")


if __name__ == "__main__":
    main()
