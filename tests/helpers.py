"""Shared test helpers (CPU): golden loading, deterministic weights, oracle wrappers."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

import vqae_b200
from vqae_b200 import synthetic as S
from vqae_b200.config import (compose_efficientnetv2_conf, compose_multilevel_conf, compose_vqae_conf,
                              mbconv, pre_activation_fixup)

GOLDEN = Path(__file__).resolve().parent / "golden"

# the committed model goldens: tag -> (n_down, regime, batch, size, seed)   (oracle/make_golden.py)
MODEL_CASES = {
    "model_nd3_perturbed": (3, "perturbed", 2, 256, 1),
    "model_nd3_fixup": (3, "fixup", 2, 256, 2),
    "model_nd4_perturbed_256": (4, "perturbed", 2, 256, 3),
    "model_nd4_perturbed_512": (4, "perturbed", 1, 512, 4),
}
BLOCK_CASES = {
    "same16": (16, 16, "same", 16), "same64": (64, 64, "same", 16),
    "down8": (8, 16, "down", 16), "down32": (32, 64, "down", 8),
    "up16": (16, 8, "up", 8), "up64": (64, 32, "up", 4),
}
# blocks at sizes the tcgen05 kernels tile; goldens in blocks_tc.npz (oracle/make_golden.py, same table)
TC_BLOCK_CASES = {
    "same8": (8, 8, "same", 32), "same16": (16, 16, "same", 32), "same32": (32, 32, "same", 32),
    "same64": (64, 64, "same", 32), "same128": (128, 128, "same", 32),
    "down8": (8, 16, "down", 32), "down16": (16, 32, "down", 32), "down32": (32, 64, "down", 32),
    "down64": (64, 128, "down", 32),
    "up16": (16, 8, "up", 16), "up32": (32, 16, "up", 16), "up64": (64, 32, "up", 16),
    "up128": (128, 64, "up", 16),
}
# the reference's near-tie rule: indices must match wherever the top-2 relative gap of the
# un-rooted L4 sums exceeds this (SURVEY.md section 7 hard part 2)
NEAR_TIE_REL_GAP = 16.0 * 2.0 ** -23


def golden(name: str):
    return np.load(GOLDEN / f"{name}.npz")


_models = {}


def model_and_state(tag: str):
    """(package VQAE on CPU in eval mode with the case's weights loaded, state_dict, x)."""
    n_down, regime, batch, size, seed = MODEL_CASES[tag]
    if tag not in _models:
        m = vqae_b200.build_vqae(n_down=n_down).eval()
        sd = S.make_state_dict(m.state_dict(), seed=seed, regime=regime)
        g = golden(tag)
        sd["encoder.vq_layers.0.embed"] = torch.from_numpy(g["embed"])
        sd["encoder.vq_layers.0.embed_avg"] = torch.from_numpy(g["embed"]).clone()
        m.load_state_dict(sd)
        _models[tag] = (m, sd)
    m, sd = _models[tag]
    return m, sd, S.synthetic_patches(batch, size, seed + 1000)


def make_block(name: str):
    from vqae_b200.layers.conv_block import PreActFixupResBlock
    cin, cout, mode, hw = BLOCK_CASES[name]
    conf = pre_activation_fixup(n_layers=12)
    for k in ("_target_", "_recursive_", "in_channels", "out_channels", "mode"):
        conf.pop(k)
    blk = PreActFixupResBlock(in_channels=cin, out_channels=cout, mode=mode, **conf).eval()
    blk.load_state_dict(S.make_state_dict(blk.state_dict(), seed=11, regime="perturbed",
                                          n_layers=12))
    return blk


def tc_sub_index(n: int):
    return sorted(set(range(0, n, 3)) | {1, n - 2, n - 1})


def tc_block_input(name: str) -> torch.Tensor:
    """The seeded input of a blocks_tc.npz case (checked against the golden's x_stats)."""
    cin, _, mode, hw = TC_BLOCK_CASES[name]
    g = torch.Generator().manual_seed(700 + cin + {"same": 0, "down": 1, "up": 2}[mode])
    x = torch.randn(2, cin, hw, hw, generator=g)
    want = golden("blocks_tc")[f"{name}_x_stats"]
    got = np.array([x.double().mean().item(), x.double().std().item(), x.double().abs().sum().item()])
    assert np.allclose(got, want, rtol=1e-9), "seeded input differs from the one the golden was made from"
    return x


def make_tc_block(name: str):
    from vqae_b200.layers.conv_block import PreActFixupResBlock
    cin, cout, mode, hw = TC_BLOCK_CASES[name]
    conf = pre_activation_fixup(n_layers=12)
    for k in ("_target_", "_recursive_", "in_channels", "out_channels", "mode"):
        conf.pop(k)
    blk = PreActFixupResBlock(in_channels=cin, out_channels=cout, mode=mode, **conf).eval()
    blk.load_state_dict(S.make_state_dict(blk.state_dict(), seed=13, regime="perturbed",
                                          n_layers=12))
    return blk


def sub_grid(y: torch.Tensor) -> torch.Tensor:
    ii = torch.tensor(tc_sub_index(y.shape[-1]), device=y.device)
    return y[:, :, ii][:, :, :, ii]


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def index_mismatches_outside_ties(idx, ref_idx, gap, thresh=NEAR_TIE_REL_GAP):
    idx, ref_idx, gap = (np.asarray(v).reshape(-1) for v in (idx, ref_idx, gap))
    bad = idx != ref_idx
    return int((bad & (gap >= thresh)).sum()), int(bad.sum()), int((gap < thresh).sum())


# ---- scope row f-4: multi-level hierarchies (tests/golden/multilevel.npz, oracle/make_golden.py) ----
MULTILEVEL_CASES = {
    # tag -> (compose_multilevel_conf kwargs, seed, input seed, whole VQAE?)
    "hier": (dict(level_downs=(2, 1), n_pre_enc_layers=(2, 3), shortcut_mode="up"), 11, 1011, False),
    "flat": (dict(level_downs=(3, 0), n_pre_enc_layers=(2, 2), shortcut_mode="same"), 12, 1012, True),
}


def multilevel_model_and_state(tag: str):
    """(package module on CPU in eval mode -- an Encoder for "hier", a VQAE for "flat" -- with the golden's
    weights and codebooks, its state_dict in ``encoder.`` / ``decoder.`` naming, x)."""
    kwargs, seed, xseed, whole = MULTILEVEL_CASES[tag]
    conf = compose_multilevel_conf(**kwargs)
    g = golden("multilevel")
    if whole:
        m = vqae_b200.instantiate(conf).eval()
        sd = S.make_state_dict(m.state_dict(), seed=seed, regime="perturbed")
    else:
        m = vqae_b200.instantiate(conf["encoder_conf"]).eval()
        sd = S.make_state_dict({"encoder." + k: v for k, v in m.state_dict().items()}, seed=seed,
                               regime="perturbed")
    n_levels = len(kwargs["level_downs"])
    for i in range(n_levels):
        e = torch.from_numpy(g[f"{tag}_embed{i}"])
        sd[f"encoder.vq_layers.{i}.embed"] = e
        sd[f"encoder.vq_layers.{i}.embed_avg"] = e.clone()
    m.load_state_dict(sd if whole else {k[len("encoder."):]: v for k, v in sd.items()})
    return m, sd, S.synthetic_patches(2, 256, xseed)


# ---- scope row f-4: MBConv blocks and the efficientnetv2 model (tests/golden/mbconv.npz) -------------
MBCONV_BLOCK_CASES = {                   # name -> (c_in, c_out, mode, hw, batchnorm, se)
    "same16": (16, 16, "same", 16, True, True), "same8to16": (8, 16, "same", 12, True, True),
    "down16": (16, 32, "down", 16, True, True), "up32": (32, 16, "up", 8, True, True),
    "same16_plain": (16, 16, "same", 8, False, False), "down8_nose": (8, 16, "down", 8, True, False),
}


def make_mbconv(name: str):
    from vqae_b200.layers.conv_block import MBConv
    cin, cout, mode, hw, use_bn, use_se = MBCONV_BLOCK_CASES[name]
    conf = mbconv(batchnorm=use_bn, se=use_se)
    for k in ("_target_", "_recursive_", "in_channels", "out_channels", "mode"):
        conf.pop(k)
    blk = MBConv(in_channels=cin, out_channels=cout, mode=mode, **conf).eval()
    blk.load_state_dict(S.make_state_dict(blk.state_dict(), seed=41, regime="perturbed"))
    return blk, mode


def mbconv_model_and_state():
    m = vqae_b200.instantiate(compose_efficientnetv2_conf(n_down=3, n_enc_layers_trunk=3)).eval()
    sd = S.make_state_dict(m.state_dict(), seed=31, regime="perturbed")
    e = torch.from_numpy(golden("mbconv")["model_embed"])
    sd["encoder.vq_layers.0.embed"], sd["encoder.vq_layers.0.embed_avg"] = e, e.clone()
    m.load_state_dict(sd)
    return m, sd, S.synthetic_patches(2, 256, 1031)
