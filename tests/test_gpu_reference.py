"""GPU: an instance of the UNMODIFIED reference classes (imported through oracle/ref_shim.py from
the git-ignored copy under oracle/_ref, made by `make -C oracle ref` in the build container),
bound to the B200 path with vqae_b200.accelerate(), against the same instance running the
reference's own forward through stock PyTorch/cuDNN on the same device.

Skipped when oracle/_ref is absent (a checkout that never saw the reference)."""
import copy

import numpy as np
import pytest
import torch

import helpers as H
import ref_shim
import vqae_b200
from vqae_b200 import engine as E
from vqae_b200 import synthetic as S
from vqae_b200.config import compose_vqae_conf

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_shim.reference_available(),
                                 reason="no reference copy under oracle/_ref")]
DEV = "cuda:0"


def _reference_vqae(n_down, seed, tag):
    model_mod, *_ = ref_shim.load_reference()
    conf = compose_vqae_conf(n_down=n_down)
    conf.pop("_target_"), conf.pop("_recursive_")
    m = model_mod.VQAE(**conf).eval()
    sd = S.make_state_dict(m.state_dict(), seed=seed, regime="perturbed")
    g = H.golden(tag)
    sd["encoder.vq_layers.0.embed"] = torch.from_numpy(g["embed"])
    m.load_state_dict(sd)
    assert type(m).__module__ == "vq_ae.model" and "vqae_b200" not in type(m.encoder).__module__
    return m


@pytest.fixture(autouse=True)
def _no_tf32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_accelerated_reference_model_matches_its_own_forward():
    tag = "model_nd3_perturbed"
    n_down, regime, batch, size, seed = H.MODEL_CASES[tag]
    ref = _reference_vqae(n_down, seed, tag).to(DEV)
    fast = vqae_b200.accelerate(copy.deepcopy(ref))
    assert set(fast.state_dict()) == set(ref.state_dict())
    x = S.synthetic_patches(batch, size, seed + 1000).to(DEV)
    with torch.no_grad():
        (r_enc,), (r_idx,), (r_loss,) = ref.encoder(x)            # stock torch / cuDNN, fp32
        r_recon, _ = ref(x)
        before = E.launch_count()
        (f_enc,), (f_idx,), (f_loss,) = fast.encoder(x)           # B200 plan behind the same class
        f_recon, (f_loss2,) = fast(x)
        launches = E.launch_count() - before
    assert launches > 0                                            # the library ran, not the reference
    assert f_idx.dtype == r_idx.dtype and f_idx.shape == r_idx.shape
    g = H.golden(tag)
    same = (f_idx == r_idx).cpu().numpy().reshape(-1)
    # two fp32 evaluations (cuDNN vs our kernels): codes may differ only at near-ties
    assert not (~same & (g["gap"] >= 1e-4)).any()
    assert (~same).mean() < 2e-3
    msk = torch.from_numpy(same.reshape(tuple(r_idx.shape))).to(DEV)
    assert float(((f_enc - r_enc).abs() * msk[:, None]).max() / r_enc.abs().max()) < 1e-4
    if same.all():
        assert H.rel_err(f_recon, r_recon) < 1e-4
    assert abs(f_loss.item() - r_loss.item()) < 1e-4 * abs(r_loss.item())
    # training mode / CPU tensors keep the reference's own forward
    cpu_model = vqae_b200.accelerate(_reference_vqae(n_down, seed, tag))
    with torch.no_grad():
        n0 = E.launch_count()
        cpu_model.encoder(torch.randn(1, 3, 64, 64))
        assert E.launch_count() == n0


def test_accelerated_reference_under_autocast_vs_reference_autocast():
    """extract_embeddings.py:124-125 runs the encoder under torch.autocast('cuda').  Agreement with
    the reference's fp32 codes: ours (bf16 tensor-core path) next to the reference's own fp16
    autocast run -- the reduced-precision path must not be worse than the reference is to itself
    by more than 1 %."""
    tag = "model_nd3_perturbed"
    n_down, regime, batch, size, seed = H.MODEL_CASES[tag]
    ref = _reference_vqae(n_down, seed, tag).to(DEV)
    fast = vqae_b200.accelerate(copy.deepcopy(ref))
    x = S.synthetic_patches(batch, size, seed + 1000).to(DEV)
    with torch.no_grad():
        (_,), (i32,), _ = ref.encoder(x)
        with torch.autocast("cuda"):
            enc_t, idx_t, loss_t = tuple(zip(*ref.encoder(x)))[0]      # the reference's own re-zip
            f_enc, f_idx, f_loss = tuple(zip(*fast.encoder(x)))[0]
    a_ref = float((idx_t == i32).float().mean())
    a_fast = float((f_idx == i32).float().mean())
    print(f"agreement with reference fp32 codes: reference autocast(fp16) {a_ref:.4f}, B200 bf16 path {a_fast:.4f}")
    assert a_fast >= 0.985
    assert a_fast >= a_ref - 0.01
