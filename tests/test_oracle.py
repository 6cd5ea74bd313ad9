"""CPU: the oracle restatement (oracle/vqae_oracle.py, oracle/l4_quantize.c) against the golden
vectors written by the unmodified reference (oracle/make_golden.py)."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

import helpers as H
import vqae_oracle as O

REPO = Path(__file__).resolve().parent.parent


def test_bare_quantizer_matches_reference_golden():
    g = H.golden("quantizer")
    x, embed = torch.from_numpy(g["bare_x"]), torch.from_numpy(g["bare_embed"])
    quant, idx, loss, gap = O.ema_quantizer_forward(x, embed, 1.0)
    assert np.array_equal(idx.numpy(), g["bare_idx"].astype(np.int64))      # bit-exact indices
    assert np.array_equal(quant.numpy(), g["bare_quant"])
    assert abs(loss.item() - float(g["bare_loss"])) <= 1e-7 * abs(float(g["bare_loss"]))
    np.testing.assert_allclose(gap.reshape(-1).numpy(), g["bare_gap"], rtol=2e-3, atol=1e-6)


def test_tie_breaks_to_lowest_index():
    g = H.golden("quantizer")
    x, embed = torch.from_numpy(g["bare_x"]), torch.from_numpy(g["tie_embed"])
    _, idx, loss, _ = O.ema_quantizer_forward(x, embed, 1.0)
    assert np.array_equal(idx.numpy(), g["tie_idx"].astype(np.int64))
    assert not np.isin(idx.numpy(), (77, 200)).any()     # duplicates of row 3 never win
    assert bool(g["tie_cl_is_channels_last"]) and bool(g["tie_cl_idx_equal"])


def test_quantizer_rejects_wrong_channel_dim():
    with pytest.raises(NotImplementedError):
        O.ema_quantizer_forward(torch.zeros(1, 4, 2, 2), torch.zeros(256, 8))
    with pytest.raises(AssertionError):
        O.ema_quantizer_forward(torch.zeros(4, 8), torch.zeros(256, 8))


@pytest.mark.parametrize("c", [64, 128])
def test_projected_quantizer_matches_reference_golden(c):
    from vqae_b200 import synthetic as S
    from vqae_b200.layers.vq import ProjectedEMAVectorQuantizer2d
    g = H.golden("quantizer")
    tmpl = ProjectedEMAVectorQuantizer2d(256, c, 1.0, 0.99, 1e-5, 8).state_dict()
    sd = S.make_state_dict(tmpl, seed=5, regime="perturbed")
    sd["embed"] = torch.from_numpy(g[f"proj{c}_embed"])
    x = torch.randn(2, c, 32, 32, generator=torch.Generator().manual_seed(200 + c))
    out, idx, loss, gap, z = O.projected_quantizer_forward(x, sd, "", 1.0)
    bad, total_bad, n_ties = H.index_mismatches_outside_ties(idx, g[f"proj{c}_idx"], g[f"proj{c}_gap"])
    assert bad == 0 and total_bad <= n_ties
    np.testing.assert_allclose(out[:, :, ::4, ::4].numpy(), g[f"proj{c}_quant_sub"], rtol=1e-5, atol=1e-6)
    assert abs(loss.item() - float(g[f"proj{c}_loss"])) < 1e-6


@pytest.mark.parametrize("name", sorted(H.BLOCK_CASES))
def test_fixup_block_matches_reference_golden(name):
    g = H.golden("blocks")
    blk = H.make_block(name)
    y = O.fixup_block(torch.from_numpy(g[f"{name}_x"]), blk.state_dict())
    assert H.rel_err(y, torch.from_numpy(g[f"{name}_y"])) < 2e-6


@pytest.mark.parametrize("tag", ["model_nd3_perturbed", "model_nd3_fixup", "model_nd4_perturbed_256"])
def test_model_matches_reference_golden(tag):
    g = H.golden(tag)
    _, sd, x = H.model_and_state(tag)
    assert len(sd) == int(g["n_state"])
    with torch.no_grad():
        (enc,), (idx,), (loss,), (gap, z, h) = O.encoder_forward(x, sd, with_aux=True)
        recon = O.decoder_forward((enc,), sd)
        dec = O.decode_from_codes(idx, sd)
    np.testing.assert_allclose(O.normalize_u8(np.zeros((2, 2, 3), np.uint8))[:, 0, 0],
                               [-0.7279 / 0.2419, -0.5955 / 0.3083, -0.7762 / 0.1741], rtol=1e-6)
    bad, total_bad, n_ties = H.index_mismatches_outside_ties(idx, g["idx"], g["gap"])
    assert bad == 0 and total_bad <= n_ties, (bad, total_bad, n_ties)
    assert H.rel_err(h[:, ::8, ::4, ::4], torch.from_numpy(g["pre_vq_sub"])) < 1e-5
    assert H.rel_err(recon[:, :, ::8, ::8], torch.from_numpy(g["recon_sub"])) < 1e-4
    assert H.rel_err(dec[:, :, ::8, ::8], torch.from_numpy(g["decode_codes_sub"])) < 1e-4
    assert abs(loss.item() - float(g["loss"])) < 1e-5 * max(1.0, abs(float(g["loss"])))
    assert int(idx.unique().numel()) == int(g["codes_used"])


def _c_oracle():
    subprocess.run(["make", "-s", "-C", str(REPO / "oracle")], check=True)
    lib = ctypes.CDLL(str(REPO / "oracle" / "_build" / "liboracle_l4.so"))
    lib.oracle_l4_quantize.restype = ctypes.c_double
    return lib


def test_c_oracle_bit_exact_with_torch_cdist_and_golden():
    lib = _c_oracle()
    g = H.golden("quantizer")
    x = torch.from_numpy(g["bare_x"]).permute(0, 2, 3, 1).reshape(-1, 8).contiguous()
    embed = torch.from_numpy(g["bare_embed"]).contiguous()
    n, k, d = x.shape[0], embed.shape[0], 8
    dist = np.empty((n, k), np.float32)
    fp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    xn, en = x.numpy(), embed.numpy()
    lib.oracle_l4_cdist(fp(xn), ctypes.c_int64(n), fp(en), k, d, fp(dist))
    ref = torch.cdist(x, embed, 4, compute_mode='donot_use_mm_for_euclid_dist').numpy()
    # ATen's scalar loop and this restatement agree to the last bit on almost every entry; the
    # remainder differ by one ulp of powf
    assert np.mean(dist == ref) > 0.99 and np.max(np.abs(dist - ref) / ref) < 3e-7
    idx = np.empty(n, np.int64)
    q = np.empty((n, d), np.float32)
    gap = np.empty(n, np.float32)
    sq = lib.oracle_l4_quantize(fp(xn), ctypes.c_int64(n), fp(en), k, d, fp(idx), fp(q), fp(gap))
    assert np.array_equal(idx.reshape(4, 16, 16), g["bare_idx"].astype(np.int64))
    assert abs(sq / (n * d) - float(g["bare_loss"])) < 1e-6


def test_slide_geometry_and_stitching():
    assert O.slide_grid((50000, 50000), 256) == (195, 195)
    assert O.patch_rc(196, 195) == (1, 1)
    tiles = np.arange(6 * 4 * 4).reshape(6, 4, 4) % 200
    m = O.stitch_code_map(tiles, 2, 3)
    assert m.shape == (8, 12) and m.dtype == np.uint8
    assert np.array_equal(m[4:8, 8:12], tiles[5])
    assert O.cast_to_lowest_dtype(np.array([0, 1])).dtype == bool


# ---- scope row f-4 ---------------------------------------------------------------------------------
def test_multilevel_hierarchy_matches_reference_golden():
    """Encoder with two VQ levels and an 'up' shortcut block (model.py:144-187, 203-215)."""
    g = H.golden("multilevel")
    _, sd, x = H.multilevel_model_and_state("hier")
    assert int(g["hier_n_state"]) == len(sd)
    with torch.no_grad():
        encs, idxs, losses = O.encoder_forward_levels(x, sd)
    assert len(encs) == 2
    assert tuple(encs[0].shape) == (2, 64, 32, 32) and tuple(encs[1].shape) == (2, 32, 64, 64)
    for i in range(2):
        ref = g[f"hier_idx{i}"].astype(np.int64)
        bad = (idxs[i].numpy() != ref).reshape(-1) & (g[f"hier_gap{i}"] >= H.NEAR_TIE_REL_GAP)
        assert int(bad.sum()) == 0
        assert H.rel_err(encs[i][:, ::8, ::4, ::4], torch.from_numpy(g[f"hier_enc_sub{i}"])) < 1e-5
        assert abs(losses[i].item() - float(g[f"hier_loss{i}"])) < 1e-5 * float(g[f"hier_loss{i}"])


def test_multilevel_vqae_matches_reference_golden():
    """Whole VQAE with two equal-width levels (the shape the reference's Decoder accepts)."""
    g = H.golden("multilevel")
    _, sd, x = H.multilevel_model_and_state("flat")
    assert int(g["flat_n_state"]) == len(sd)
    with torch.no_grad():
        encs, idxs, losses = O.encoder_forward_levels(x, sd)
        recon = O.decoder_forward_levels(encs, sd)
    for i in range(2):
        assert np.array_equal(idxs[i].numpy(), g[f"flat_idx{i}"].astype(np.int64))
        assert abs(losses[i].item() - float(g[f"flat_loss{i}"])) < 1e-5 * float(g[f"flat_loss{i}"])
    assert H.rel_err(recon[:, :, ::8, ::8], torch.from_numpy(g["flat_recon_sub"])) < 1e-5


@pytest.mark.parametrize("tag", ["bare", "proj"])
def test_ema_training_updates_match_reference_golden(tag):
    """_init_ema / _update_ema restated without the one-hot matrix (vq.py:47-94) against three
    training-mode forwards of the reference's quantisers."""
    g = H.golden("ema_training")
    from vqae_b200.layers.vq import EMAVectorQuantizer, ProjectedEMAVectorQuantizer2d
    c = 8 if tag == "bare" else 64
    q = (EMAVectorQuantizer(256, 8, 0.25, 0.99, 1e-5) if tag == "bare"
         else ProjectedEMAVectorQuantizer2d(256, c, 0.25, 0.99, 1e-5, 8))
    sd = H.S.make_state_dict(q.state_dict(), seed=21, regime="perturbed")
    embed, avg, cs = sd["embed"], sd["embed_avg"], sd["cluster_size"]
    assert np.array_equal(embed.numpy(), g[f"{tag}_embed0"])
    for step in range(3):
        x = torch.from_numpy(g[f"{tag}_x{step}"]).float()
        z = x if tag == "bare" else torch.nn.functional.conv2d(x, sd["proj_in.weight"], sd["proj_in.bias"])
        flat = z.permute(0, 2, 3, 1).reshape(-1, 8)
        if step == 0:
            embed, avg, cs = O.ema_init(flat, embed, cs)
        _, idx, loss, _ = O.ema_quantizer_forward(z, embed, 0.25)
        assert np.array_equal(idx.numpy(), g[f"{tag}_idx{step}"].astype(np.int64))
        assert abs(loss.item() - float(g[f"{tag}_loss{step}"])) < 1e-5 * float(g[f"{tag}_loss{step}"])
        embed, avg, cs = O.ema_update(flat, idx, avg, cs, 0.99, 1e-5)
        np.testing.assert_allclose(cs.numpy(), g[f"{tag}_cluster_size_after{step}"], rtol=1e-6)
        np.testing.assert_allclose(avg.numpy(), g[f"{tag}_embed_avg_after{step}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(embed.numpy(), g[f"{tag}_embed_after{step}"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", sorted(H.MBCONV_BLOCK_CASES))
def test_mbconv_block_matches_reference_golden(name):
    g = H.golden("mbconv")
    blk, mode = H.make_mbconv(name)
    with torch.no_grad():
        y = O.mbconv_block(torch.from_numpy(g[f"blk_{name}_x"]), blk.state_dict(), mode)
    assert H.rel_err(y, torch.from_numpy(g[f"blk_{name}_y"])) < 1e-6


def test_mbconv_model_matches_reference_golden():
    """conf/model/{encoder,decoder}/efficientnetv2.yaml: every conv block an MBConv."""
    g = H.golden("mbconv")
    _, sd, x = H.mbconv_model_and_state()
    assert int(g["model_n_state"]) == len(sd)
    with torch.no_grad():
        (enc,), (idx,), (loss,) = O.encoder_forward_levels(x, sd)
        recon = O.decoder_forward_levels((enc,), sd)
    assert np.array_equal(idx.numpy(), g["model_idx"].astype(np.int64))
    assert abs(loss.item() - float(g["model_loss"])) < 1e-5 * float(g["model_loss"])
    assert H.rel_err(recon[:, :, ::8, ::8], torch.from_numpy(g["model_recon_sub"])) < 1e-5
