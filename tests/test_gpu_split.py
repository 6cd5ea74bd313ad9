"""GPU: the fp32-ACCURATE tensor-core kernels (precision "fp32tc": split fp16 operands hi + lo, three
tensor-core products per GEMM, exact activations -- csrc/tc_split.cu, csrc/mma_down.cu SPLIT) against
the fp32 CUDA-core path, which is pinned to the reference goldens; the goldens themselves are asserted
in test_gpu_parity.py::test_model_vs_reference_golden[fp32tc-*] and
test_gpu_golden_reduced.py::test_block_at_tensor_core_sizes_vs_reference_golden[fp32tc-*]."""
import pytest
import torch

import helpers as H
import vqae_b200
from vqae_b200 import engine as E
from vqae_b200 import synthetic as S

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _block(c_in, mode, seed):
    from vqae_b200.config import pre_activation_fixup
    from vqae_b200.layers.conv_block import PreActFixupResBlock
    conf = pre_activation_fixup(n_layers=12)
    for k in ("_target_", "_recursive_", "in_channels", "out_channels", "mode"):
        conf.pop(k)
    c_out = c_in if mode == "same" else 2 * c_in
    blk = PreActFixupResBlock(in_channels=c_in, out_channels=c_out, mode=mode, **conf).eval()
    blk.load_state_dict(S.make_state_dict(blk.state_dict(), seed=seed, regime="perturbed", n_layers=12))
    return blk.to(DEV)


@pytest.mark.parametrize("warp_mma", [True, False])
@pytest.mark.parametrize("mode,c_in,h,w,batch", [
    ("same", 8, 64, 64, 2), ("same", 8, 8, 32, 3), ("same", 16, 32, 96, 2), ("same", 32, 24, 32, 3),
    ("same", 64, 32, 32, 5), ("same", 64, 8, 64, 1),
    ("down", 8, 32, 64, 2), ("down", 16, 16, 32, 3), ("down", 32, 64, 64, 2)])
def test_split_block_vs_fp32_path(mode, c_in, h, w, batch, warp_mma, monkeypatch):
    """One launch per block, within 1e-5 of the fp32 kernels (measured ~1e-6: both are fp32-accurate,
    they differ in summation order), deterministic.  'same' blocks at C = 8, 16 have two split forms:
    warp-level MMAs (mma_same_split.cu, the one dispatched) and tcgen05 (tc_split.cu)."""
    if not warp_mma:
        if not (mode == "same" and c_in in (8, 16)):
            pytest.skip("one split form only")
        monkeypatch.setattr(E, "SPLIT_MMA", set())
    blk = _block(c_in, mode, 51)
    pk = blk.packed()
    x = torch.randn(batch, h, w, c_in, generator=torch.Generator().manual_seed(h * w + c_in)).to(DEV)
    y32 = E.fixup_forward_nhwc(pk, x, precision="fp32")
    y = E.fixup_forward_nhwc(pk, x, precision="fp32tc")          # packs on first use
    before = E.launch_count()
    y2 = E.fixup_forward_nhwc(pk, x, precision="fp32tc")
    torch.cuda.synchronize()
    assert E.launch_count() - before == 1
    assert y.shape == y32.shape and y.dtype == torch.float32
    err = float((y - y32).abs().max()) / float(y32.abs().max())
    assert err < 1e-5, err
    # the branch alone (what went through the split GEMMs), relative to its own magnitude
    if mode == "same":
        br = float((y32 - x).abs().max())
        assert float(((y - x) - (y32 - x)).abs().max()) < 2e-5 * br
    assert torch.equal(y, y2)


def test_split_block_small_and_large_weights():
    """The power-of-two pre-scaling keeps the low halves normal fp16 numbers whatever the weight
    magnitude: the same block with its convs scaled by 2^-9 / 2^7 / 2^5 (compensated in the next
    stage's bias and scale so that activations stay O(1)) is as accurate as the unscaled one."""
    blk = _block(64, "same", 52)
    x = torch.randn(2, 16, 32, 64, generator=torch.Generator().manual_seed(5)).to(DEV)
    with torch.no_grad():
        blk.branch_conv1.weight.mul_(2.0 ** -9)
        blk.bias2a.mul_(2.0 ** -9)
        blk.branch_conv3.weight.mul_(2.0 ** 7)
        blk.scale.mul_(2.0 ** -7)
    pk = E.pack_blocks([blk])[0]
    y32 = E.fixup_forward_nhwc(pk, x, precision="fp32")
    y = E.fixup_forward_nhwc(pk, x, precision="fp32tc")
    assert float((y - y32).abs().max()) / float(y32.abs().max()) < 1e-5
    assert pk.split_premul[0] == 2.0 ** 9 * pk.split_premul[1] or pk.split_premul[0] > pk.split_premul[1]


def test_fp32tc_encoder_codes_equal_fp32_codes():
    """Whole 256-model encoder: the split tensor-core path and the fp32 CUDA-core path give the same
    codes except at near-ties, with latents within 1e-5."""
    tag = "model_nd3_perturbed"
    m, sd, x = H.model_and_state(tag)
    m = m.to(DEV)
    try:
        xd = S.synthetic_patches_u8(4, 256, 77).to(DEV)
        with torch.no_grad():
            vqae_b200.set_precision(m, "fp32")
            _, i32, _, _, z32 = m.encoder.encode(xd, want_latents=True)
            vqae_b200.set_precision(m, "fp32tc")
            m.encoder.encode(xd)                                  # packs the split operands
            n0 = E.launch_count()
            _, itc, _, _, ztc = m.encoder.encode(xd, want_latents=True)
            n_tc = E.launch_count() - n0
        assert float((ztc - z32).abs().max()) < 1e-5 * float(z32.abs().max())
        assert float((itc != i32).float().mean()) < 1e-3
        # stem + 1 + 1 + 5 + 1 + 5 + 1 + 54 blocks + quantiser: one launch each, no fp32 conv kernels
        assert n_tc == 1 + 68 + 1, n_tc
    finally:
        vqae_b200.set_precision(m, None)
        m.cpu()
