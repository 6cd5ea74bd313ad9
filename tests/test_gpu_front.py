"""GPU: the fused encoder front end (csrc/mma_front.cu: in_stem + 'same' C = 8 + 'down' 8 -> 16 in one
launch) against the three separate launches it replaces and against the fp32 exact path (which is
pinned to the reference goldens); the goldens themselves cover it through the model tests of
test_gpu_golden_reduced.py (the "fp16" encoder runs this kernel)."""
import pytest
import torch

import helpers as H
import vqae_b200
from vqae_b200 import engine as E
from vqae_b200 import plan as P
from vqae_b200 import synthetic as S

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


_plans = {}


def _front(m, x, precision, fused):
    enc = m.encoder
    blocks = (P.flat_blocks(enc.down_layers) + P.flat_blocks(enc.pre_enc_layers))[:2]
    packed = _plans.setdefault(id(m), P.Plan()).get(blocks)     # packed once per model and device
    old = E.FRONT_FUSED
    E.FRONT_FUSED = fused
    try:
        h, used = E.encoder_front(x, enc.in_stem.weight, enc.in_stem.bias, None, None, packed, precision)
        h = E.run_blocks_nhwc(packed[used:], h, precision)
    finally:
        E.FRONT_FUSED = old
    return h, used


@pytest.mark.parametrize("kind,hw,batch", [
    ("u8", (256, 256), 2), ("u8", (16, 32), 3), ("u8", (32, 64), 2), ("u8", (48, 32), 1),
    ("nchw", (64, 96), 2), ("channels_last", (32, 32), 2)])
def test_front_fused_vs_separate_launches(kind, hw, batch):
    m, sd, _ = H.model_and_state("model_nd3_perturbed")
    m = m.to(DEV)
    try:
        g = torch.Generator().manual_seed(hw[0] * 7 + hw[1])
        if kind == "u8":
            x = torch.randint(0, 256, (batch, hw[0], hw[1], 3), generator=g, dtype=torch.uint8).to(DEV)
        else:
            x = torch.randn(batch, 3, hw[0], hw[1], generator=g).to(DEV)
            if kind == "channels_last":
                x = x.contiguous(memory_format=torch.channels_last)
        with torch.no_grad():
            ref16, used0 = _front(m, x, "fp16", False)
            ref32, _ = _front(m, x, "fp32", False)
            _front(m, x, "fp16", True)                     # packs
            n0 = E.launch_count()
            got, used = _front(m, x, "fp16", True)
            torch.cuda.synchronize()
        assert used0 == 0 and used == 2
        assert E.launch_count() - n0 == 1
        assert got.shape == ref16.shape == (batch, hw[0] // 2, hw[1] // 2, 16)
        scale = float(ref32.abs().max())
        # same fp16 arithmetic in the two blocks; the stem differs in summation order only (~1e-6),
        # which can move an fp16 rounding of an operand by one step here and there
        assert float((got - ref16).abs().max()) < 2e-3 * scale
        assert float((got - ref16).abs().mean()) < 2e-5 * scale
        assert float((got - ref32).abs().max()) < 1e-2 * scale
        with torch.no_grad():
            again, _ = _front(m, x, "fp16", True)
        assert torch.equal(got, again)
    finally:
        m.cpu()


def test_front_fused_in_the_fp16_encoder():
    """E.FRONT_FUSED = True swaps the encoder's first three launches for the fused kernel."""
    m, sd, _ = H.model_and_state("model_nd3_perturbed")
    m = vqae_b200.set_precision(m.to(DEV), "fp16")
    old = E.FRONT_FUSED
    E.FRONT_FUSED = True
    try:
        x = S.synthetic_patches_u8(2, 256, 11).to(DEV)
        with torch.no_grad():
            m.encoder.encode(x)
            n0 = E.launch_count()
            _, idx, _, _, _ = m.encoder.encode(x)
            n_fused = E.launch_count() - n0
            E.FRONT_FUSED = False
            try:
                n0 = E.launch_count()
                _, idx_sep, _, _, _ = m.encoder.encode(x)
                n_sep = E.launch_count() - n0
            finally:
                E.FRONT_FUSED = True
        assert n_sep - n_fused == 2
        assert float((idx != idx_sep).float().mean()) < 2e-3
    finally:
        E.FRONT_FUSED = old
        vqae_b200.set_precision(m, None)
        m.cpu()


@pytest.mark.parametrize("hw,batch,cl", [((256, 256), 2, False), ((16, 32), 3, False), ((48, 64), 2, True),
                                         ((32, 32), 1, True)])
def test_stem_out_mma_vs_exact_fp32_kernel(hw, batch, cl):
    """csrc/mma_stem.cu (out_stem on split-operand MMAs, the decoder's last launch in the "fp16" and
    "fp32tc" paths) against the exact fp32 kernel, which the model goldens pin to the reference:
    fp32-accurate (1e-5 of the output range; zero padding at the image border included)."""
    m, sd, _ = H.model_and_state("model_nd3_perturbed")
    dec = m.to(DEV).decoder
    try:
        x = torch.randn(batch, hw[0], hw[1], 8, generator=torch.Generator().manual_seed(hw[0] + batch)).to(DEV)
        ref = E.stem_out(x, dec.out_stem.weight, dec.out_stem.bias, cl, "fp32")
        n0 = E.launch_count()
        got = E.stem_out(x, dec.out_stem.weight, dec.out_stem.bias, cl, "fp16")
        torch.cuda.synchronize()
        assert E.launch_count() - n0 == 1
        assert got.shape == ref.shape == (batch, 3, hw[0], hw[1])
        assert E.is_channels_last(got) == cl or not cl
        assert float((got - ref).abs().max()) < 1e-5 * float(ref.abs().max())
        assert torch.equal(got, E.stem_out(x, dec.out_stem.weight, dec.out_stem.bias, cl, "fp32tc"))
    finally:
        m.cpu()


@pytest.mark.parametrize("kind,hw,batch", [("u8", (256, 256), 2), ("u8", (16, 32), 3), ("nchw", (48, 64), 2),
                                           ("channels_last", (32, 96), 1)])
def test_stem_in_mma_vs_exact_fp32_kernel(kind, hw, batch):
    """csrc/mma_stem.cu in_stem (nine split-operand MMAs per 16 pixels; the encoder's first launch in
    the "fp16" and "fp32tc" paths) against the exact fp32 kernel, which the model goldens pin to the
    reference (`in_stem_sub`): 1e-5 of the output range, zero padding at the border included."""
    m, sd, _ = H.model_and_state("model_nd3_perturbed")
    enc = m.to(DEV).encoder
    try:
        g = torch.Generator().manual_seed(hw[1] + batch)
        if kind == "u8":
            x = torch.randint(0, 256, (batch, hw[0], hw[1], 3), generator=g, dtype=torch.uint8).to(DEV)
        else:
            x = torch.randn(batch, 3, hw[0], hw[1], generator=g).to(DEV)
            if kind == "channels_last":
                x = x.contiguous(memory_format=torch.channels_last)
        ref = E.stem_in(x, enc.in_stem.weight, enc.in_stem.bias)
        n0 = E.launch_count()
        got = E.stem_in(x, enc.in_stem.weight, enc.in_stem.bias, precision="fp16")
        torch.cuda.synchronize()
        assert E.launch_count() - n0 == 1
        assert got.shape == ref.shape == (batch, hw[0], hw[1], 8) and got.dtype == torch.float32
        assert float((got - ref).abs().max()) < 1e-5 * float(ref.abs().max())
        assert not torch.equal(got, ref) or hw == (16, 32)       # it IS the other kernel
    finally:
        m.cpu()


@pytest.mark.parametrize("hw,batch,wscale", [((48, 96), 3, 1.0), ((16, 64), 5, 1.0), ((32, 32), 2, 300.0),
                                             ((64, 128), 2, 1e-4)])
def test_stem_in_u8_folded_normalisation_kernel(hw, batch, wscale):
    """The uint8 form of the in_stem MMA kernel (pixels exact in fp16, normalisation folded into split
    weights that are scaled by a power of two, `inside` as a fourth input channel) against the exact fp32
    kernel: interior and border tiles, batch slices (the window is read with aligned 32-bit loads),
    saturated pixels, and weights far from O(1)."""
    m, sd, _ = H.model_and_state("model_nd3_perturbed")
    enc = m.to(DEV).encoder
    try:
        g = torch.Generator().manual_seed(hw[0] + 3 * hw[1] + batch)
        big = torch.randint(0, 256, (batch + 1, hw[0], hw[1], 3), generator=g, dtype=torch.uint8)
        big[0, :, : hw[1] // 2] = 255                      # saturated and black regions
        big[-1, hw[0] // 2:] = 0
        big = big.to(DEV)
        w = (enc.in_stem.weight.detach() * wscale).contiguous()
        b = enc.in_stem.bias.detach()
        for x in (big, big[1:], big[:1]):
            ref = E.stem_in(x, w, b)
            got = E.stem_in(x, w, b, precision="fp16")
            torch.cuda.synchronize()
            assert got.shape == ref.shape
            assert float((got - ref).abs().max()) < 4e-6 * float(ref.abs().max())
        assert torch.equal(E.stem_in(big, w, b, precision="fp16"), E.stem_in(big, w, b, precision="fp16"))
    finally:
        m.cpu()
