"""GPU: config-2-size (N = 524 288) quantiser output against the plain-C oracle
(oracle/l4_quantize.c, pthreads: a few seconds on the host), and the code-extraction caller
(StreamingEncoder, get_encodings) against straightforward per-batch encodes."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

import helpers as H
import vqae_b200
from vqae_b200 import engine as E
from vqae_b200 import extract as X
from vqae_b200 import synthetic as S
from vqae_b200.layers.vq import EMAVectorQuantizer, ProjectedEMAVectorQuantizer2d

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
REPO = Path(__file__).resolve().parent.parent


def _c_oracle():
    subprocess.run(["make", "-s", "-C", str(REPO / "oracle"), "_build/liboracle_l4.so"], check=True)
    lib = ctypes.CDLL(str(REPO / "oracle" / "_build" / "liboracle_l4.so"))
    lib.oracle_l4_quantize_mt.restype = ctypes.c_double
    return lib


def _oracle_quantize(z: np.ndarray, embed: np.ndarray):
    lib = _c_oracle()
    n, d = z.shape
    idx = np.empty(n, np.int64)
    gap = np.empty(n, np.float32)
    fp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    sq = lib.oracle_l4_quantize_mt(fp(z), ctypes.c_int64(n), fp(embed), embed.shape[0], d, fp(idx),
                                   None, fp(gap), 0)
    return idx, gap, sq / (n * d)


def _ties_in_band(device_count: int, o_gap: np.ndarray) -> None:
    """The device flags (second - best) < t * second on ITS fp32 sums, the oracle reports
    (second - best) / second of its own: rows within a rounding of the threshold t may fall on
    either side, so the device count must lie between the oracle's counts at t / 2 and 2 t."""
    lo = int((o_gap < 0.5 * H.NEAR_TIE_REL_GAP).sum())
    hi = int((o_gap < 2.0 * H.NEAR_TIE_REL_GAP).sum())
    assert lo <= device_count <= hi, (lo, device_count, hi)


def test_bare_quantizer_config2_size_vs_c_oracle():
    """EMAVectorQuantizer on [512,8,32,32] (N = 524 288 vectors): every index against the C
    restatement of ATen's cdist(p=4) + argmin, bit-exact outside reported near-ties."""
    torch.manual_seed(2023)                       # the module draws its codebook from the global RNG
    q = EMAVectorQuantizer(256, 8, 1.0, 0.99, 1e-5).eval().to(DEV)
    g = torch.Generator().manual_seed(2024)
    x = torch.randn(512, 8, 32, 32, generator=g)
    quant, idx, loss = q(x.to(DEV))
    z = x.permute(0, 2, 3, 1).reshape(-1, 8).contiguous().numpy()
    o_idx, o_gap, o_loss = _oracle_quantize(z, q.embed.cpu().numpy())
    bad, total_bad, n_ties = H.index_mismatches_outside_ties(idx.cpu().numpy(), o_idx, o_gap)
    assert bad == 0, (bad, total_bad, n_ties)
    assert total_bad <= n_ties
    _ties_in_band(int(q.last_near_ties.item()), o_gap)
    assert abs(loss.item() - o_loss) < 1e-5 * o_loss
    # quantised rows are the codebook rows of the chosen codes (straight-through arithmetic: 1 ulp)
    ref_q = q.embed[idx.reshape(-1)].reshape(512, 32, 32, 8).permute(0, 3, 1, 2)
    assert H.rel_err(quant, ref_q) < 2e-7


@pytest.mark.parametrize("c", [64, 128])
def test_projected_quantizer_config2_size_vs_c_oracle(c):
    """ProjectedEMAVectorQuantizer2d on [512,C,32,32] NHWC -- the config-2 microbench call (the fused
    tcgen05 kernel where one is built): indices against the C oracle run on the kernel's own
    projected latents (bit-exact outside near-ties), the latents against proj_in in fp64."""
    pq = ProjectedEMAVectorQuantizer2d(256, c, 1.0, 0.99, 1e-5, 8).eval()
    sd = S.make_state_dict(pq.state_dict(), seed=5, regime="perturbed")
    pq.load_state_dict(sd)
    g = torch.Generator().manual_seed(77 + c)
    x = torch.randn(512, 32, 32, c, generator=g)                         # NHWC memory
    w, b = pq.proj_in.weight.detach().reshape(8, c), pq.proj_in.bias.detach()
    z64 = x.reshape(-1, c).double() @ w.double().t() + b.double()
    pq.embed.copy_(S.rescale_codebook(sd["embed"], z64.float()))
    pq = pq.to(DEV)
    xd = x.to(DEV)
    packed = pq.packed()
    out, idx, loss, ties, z = E.quantize(packed, xd.reshape(-1), True, True, 512, 1024,
                                         want_out=True, want_z=True)
    z_np = z.cpu().numpy()
    assert float(np.abs(z_np - z64.numpy()).max()) < 2e-5 * float(z64.abs().max())
    o_idx, o_gap, o_loss = _oracle_quantize(np.ascontiguousarray(z_np), pq.embed.cpu().numpy())
    bad, total_bad, n_ties = H.index_mismatches_outside_ties(idx.cpu().numpy(), o_idx, o_gap)
    assert bad == 0, (bad, total_bad, n_ties)
    assert total_bad <= n_ties
    _ties_in_band(int(ties.item()), o_gap)
    # loss = commitment_cost * mse(z, embed[idx]) in the projected space (vq.py:143)
    assert abs(loss.item() - o_loss) < 1e-5 * o_loss
    # out rows = proj_out(embed)[idx] (table gather)
    table = (pq.embed.double() @ pq.proj_out.weight.detach().reshape(c, 8).double().t()
             + pq.proj_out.bias.detach().double())
    ref_out = table[idx.reshape(-1)[:8192]].float()
    assert H.rel_err(out.view(-1, c)[:8192], ref_out) < 2e-6


def test_encode_stream_collected_equals_per_batch_encodes():
    """list(encode_stream(...)) must hold every batch's own codes (pinned staging buffers are
    recycled internally); with reuse_buffers=True the documented validity window holds."""
    tag = "model_nd3_perturbed"
    m, sd, _ = H.model_and_state(tag)
    m = vqae_b200.set_precision(m.to(DEV), "fp16")
    try:
        batches = [S.synthetic_patches_u8(3, 256, 900 + i).pin_memory() for i in range(7)]
        ref = [X.encode_patches(m.encoder, b.to(DEV)).cpu() for b in batches]
        st = X.StreamingEncoder(m.encoder, torch.device(DEV))
        got = list(st.encode_stream(batches))
        assert len(got) == len(ref)
        for a, b in zip(got, ref):
            assert torch.equal(a, b)
        # zero-copy mode: a result is still intact after two more have been yielded
        gen = st.encode_stream(batches, reuse_buffers=True)
        first = next(gen)
        keep = first.clone()
        next(gen), next(gen)
        assert torch.equal(first, keep) and torch.equal(first, ref[0])
        rest = [t.clone() for t in gen]
        assert all(torch.equal(a, b) for a, b in zip(rest, ref[3:]))
    finally:
        vqae_b200.set_precision(m, None)
        m.cpu()


def test_get_encodings_assembles_and_names_slides(tmp_path):
    """The slide-assembly loop (extract_embeddings.py:43-89,176-185): patches of two slides arrive
    interleaved across batches; each finished slide comes out once, narrowed with
    cast_to_lowest_dtype, and is saved as <ckpt>/encodings/<parent>/<stem>.npy."""
    tag = "model_nd3_perturbed"
    m, sd, _ = H.model_and_state(tag)
    m = vqae_b200.set_precision(m.to(DEV), "fp32")
    try:
        sizes = [(2, 2), (1, 3)]
        lengths = [4, 3]
        paths = ["/data/CAMELYON16/training/normal/normal_001.tif",
                 "/data/CAMELYON16/training/tumor/tumor_007.tif"]
        imgs = S.synthetic_patches_u8(7, 256, 31)
        # dataset order: slide 0 patches 0..3 then slide 1 patches 0..2; batches of 3
        meta = [(0, (0, 0)), (0, (0, 1)), (0, (1, 0)), (0, (1, 1)), (1, (0, 0)), (1, (0, 1)), (1, (0, 2))]

        def batches():
            for s in range(0, 7, 3):
                sl = slice(s, min(s + 3, 7))
                yield (imgs[sl].to(DEV), [paths[meta[i][0]] for i in range(sl.start, sl.stop)],
                       torch.tensor([meta[i][0] for i in range(sl.start, sl.stop)]),
                       torch.tensor([meta[i][1] for i in range(sl.start, sl.stop)]))

        idx = X.encode_patches(m.encoder, imgs.to(DEV)).cpu().numpy()
        out = dict(X.get_encodings(m.encoder, batches(), lengths, sizes))
        assert sorted(out) == ["normal/normal_001", "tumor/tumor_007"]
        import vqae_oracle as O
        for name, sl, grid in (("normal/normal_001", slice(0, 4), (2, 2)),
                               ("tumor/tumor_007", slice(4, 7), (1, 3))):
            ref = O.stitch_code_map(idx[sl], *grid)
            assert out[name].dtype == np.uint8 and np.array_equal(out[name], ref)
        written = list(X.extract_and_save(m.encoder, batches(), lengths, sizes, tmp_path))
        assert [p.relative_to(tmp_path).as_posix() for p in written] == [
            "encodings/normal/normal_001.npy", "encodings/tumor/tumor_007.npy"]
        assert np.array_equal(np.load(written[1]), out["tumor/tumor_007"])
    finally:
        vqae_b200.set_precision(m, None)
        m.cpu()


def test_compress_slide_rejects_wide_codebooks():
    class FakeVQ:
        num_embeddings = 512

    class FakeEnc:
        vq_layers = [FakeVQ()]

    with pytest.raises(ValueError, match="at most 256"):
        X.compress_slide(FakeEnc(), [], (1, 1), (32, 32), torch.device(DEV))
    with pytest.raises(ValueError, match="uint8"):
        E.codemap_place(torch.zeros(1, 2, 2, dtype=torch.int64, device=DEV), 0, 1,
                        torch.zeros(2, 2, dtype=torch.int64, device=DEV))
