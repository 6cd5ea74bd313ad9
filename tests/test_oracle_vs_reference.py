"""Container-only pin: the oracle restatement against the UNMODIFIED reference modules, live
(through oracle/ref_shim.py).  Skipped where /root/reference does not exist (the GPU box);
there the committed goldens (tests/test_oracle.py) carry the pin."""
import numpy as np
import pytest
import torch

import helpers as H
import ref_shim
import vqae_oracle as O

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(),
                                reason="reference checkout not mounted")


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load_reference()


def test_l4_not_l2(ref):
    """The reference's metric is L4 (p = inputs.dim() = 4, vq.py:97,121-129), not Euclidean."""
    _, vq_mod, _, _ = ref
    torch.manual_seed(0)
    q = vq_mod.EMAVectorQuantizer(256, 8, 1.0, 0.99, 1e-5).eval()
    x = torch.randn(2, 8, 16, 16)
    _, idx, _ = q(x)
    flat = x.permute(0, 2, 3, 1).reshape(-1, 8)
    assert torch.equal(idx.reshape(-1), torch.cdist(flat, q.embed, 4).argmin(1))
    assert (idx.reshape(-1) != torch.cdist(flat, q.embed, 2).argmin(1)).float().mean() > 0.05
    _, o_idx, _, _ = O.ema_quantizer_forward(x, q.embed, 1.0)
    assert torch.equal(o_idx, idx)


def test_state_dict_layout_matches_reference(ref):
    import vqae_b200
    from vqae_b200.config import compose_vqae_conf
    model_mod = ref[0]
    for n_down, n_params in ((3, 4934287), (4, 19750707)):      # SURVEY.md 8c cross-check
        conf = compose_vqae_conf(n_down=n_down)
        conf.pop("_target_"), conf.pop("_recursive_")
        r = model_mod.VQAE(**conf)
        m = vqae_b200.build_vqae(n_down=n_down)
        rs, ms = r.state_dict(), m.state_dict()
        assert list(rs.keys()) == list(ms.keys())
        assert all(rs[k].shape == ms[k].shape and rs[k].dtype == ms[k].dtype for k in rs)
        assert sum(p.numel() for p in m.parameters()) == n_params
        m.load_state_dict(rs)                                     # loads unchanged


def test_fixup_init_distribution_matches_reference(ref):
    """Same Fixup initialisation rules (conv_block.py:218-237): conv3 = 0, conv1 std."""
    import vqae_b200
    m = vqae_b200.build_vqae(n_down=3)
    blk = m.encoder.pre_enc_layers[0][0]
    assert float(blk.branch_conv3.weight.abs().max()) == 0.0
    expected = np.sqrt(2 / 64) * 124 ** -0.5
    assert abs(float(blk.branch_conv1.weight.std()) - expected) / expected < 0.1


@pytest.mark.parametrize("tag", ["model_nd3_perturbed"])
def test_oracle_equals_reference_live(ref, tag):
    from vqae_b200.config import compose_vqae_conf
    model_mod = ref[0]
    n_down = H.MODEL_CASES[tag][0]
    _, sd, x = H.model_and_state(tag)
    conf = compose_vqae_conf(n_down=n_down)
    conf.pop("_target_"), conf.pop("_recursive_")
    r = model_mod.VQAE(**conf).eval()
    r.load_state_dict(sd)
    with torch.no_grad():
        (enc,), (idx,), (loss,) = r.encoder(x)
        recon, _ = r(x)
        (o_enc,), (o_idx,), (o_loss,) = O.encoder_forward(x, sd)
        o_recon = O.decoder_forward((o_enc,), sd)
    assert torch.equal(idx, o_idx)
    assert H.rel_err(o_enc, enc) < 1e-6 and H.rel_err(o_recon, recon) < 1e-5
    assert abs(loss.item() - o_loss.item()) < 1e-7
    g = H.golden(tag)                      # and the committed golden is what the reference says
    assert np.array_equal(idx.numpy().astype(np.int16), g["idx"])
