"""GPU: CUDA-graph replay of whole steps (vqae_b200.graphs.CapturedStep) against eager launches."""
import pytest
import torch

import helpers as H
import vqae_b200
from vqae_b200 import engine as E
from vqae_b200 import extract as X
from vqae_b200 import synthetic as S
from vqae_b200.graphs import CapturedStep
from vqae_b200.plan import resolve_precision

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_captured_encode_and_round_trip_equal_eager_and_count_launches():
    m, _, _ = H.model_and_state("model_nd3_perturbed")
    m = vqae_b200.set_precision(m.to(DEV), "fp16")
    try:
        xs = [S.synthetic_patches_u8(4, 256, 70 + i).to(DEV) for i in range(2)]

        def roundtrip(x):
            _, idx, _, _, _ = m.encoder.encode(x, want_quantized=True)
            return idx, m.decode_codes(idx)

        step = CapturedStep(roundtrip, key_extra=lambda: resolve_precision(m.encoder))
        with torch.no_grad():
            ref = [tuple(t.clone() for t in roundtrip(x)) for x in xs]
            l0 = E.launch_count()
            roundtrip(xs[0])
            per_step = E.launch_count() - l0
            assert per_step > 10
            l0 = E.launch_count()
            for rep in range(4):                     # first call eager, then captures (+ replay), then replays
                for x, (ridx, rrec) in zip(xs, ref):
                    idx, rec = step(x)
                    assert torch.equal(idx, ridx) and torch.equal(rec, rrec)
            torch.cuda.synchronize()
            # every call counts as one executed step, captured or not
            assert E.launch_count() - l0 == 8 * per_step
            # another precision never replays this precision's graph
            vqae_b200.set_precision(m, "fp32")
            idx32, rec32 = step(xs[0])
            e32 = roundtrip(xs[0])
            assert torch.equal(rec32, e32[1]) and not torch.equal(rec32, ref[0][1])
    finally:
        vqae_b200.set_precision(m, None)
        m.cpu()


def test_streaming_encoder_with_and_without_graphs_agree():
    m, _, _ = H.model_and_state("model_nd3_perturbed")
    m = vqae_b200.set_precision(m.to(DEV), "fp16")
    try:
        batches = [S.synthetic_patches_u8(3, 256, 500 + i).pin_memory() for i in range(9)]
        batches.append(S.synthetic_patches_u8(2, 256, 600).pin_memory())        # short last batch
        eager = list(X.StreamingEncoder(m.encoder, torch.device(DEV), graphs=False).encode_stream(batches))
        st = X.StreamingEncoder(m.encoder, torch.device(DEV))
        for _ in range(2):                           # second pass: every full-size batch is a replay
            got = list(st.encode_stream(batches))
            assert len(got) == len(eager) and all(torch.equal(a, b) for a, b in zip(got, eager))
        dev = [t.cpu() for t in st.encode_stream(batches, to_host=False)]
        assert all(torch.equal(a, b) for a, b in zip(dev, eager))
    finally:
        vqae_b200.set_precision(m, None)
        m.cpu()


def test_graph_outlives_workspace_growth_of_later_captures():
    """A captured encode must stay valid after another capture (here a decode, whose 'up' blocks need a
    larger scratch buffer) has been taken, dropped, and the allocator's cache emptied: scratch buffers
    used inside a capture belong to the graph's own pool (engine.workspace).  Regression: a buffer
    cached across captures was freed while nodes of an earlier graph still pointed into it."""
    m, _, _ = H.model_and_state("model_nd3_perturbed")
    m = vqae_b200.set_precision(m.to(DEV), "fp16")
    try:
        x = S.synthetic_patches_u8(4, 256, 90).to(DEV)
        with torch.no_grad():
            enc_step = CapturedStep(lambda t: X.encode_patches(m.encoder, t))
            ref = X.encode_patches(m.encoder, x).clone()
            for _ in range(3):
                assert torch.equal(enc_step(x), ref)
            dec_step = CapturedStep(lambda i: m.decode_codes(i))
            rec_ref = m.decode_codes(ref).clone()
            for _ in range(3):
                assert torch.equal(dec_step(ref), rec_ref)
            del dec_step
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
            for _ in range(3):
                assert torch.equal(enc_step(x), ref)
            torch.cuda.synchronize()
    finally:
        vqae_b200.set_precision(m, None)
        m.cpu()

