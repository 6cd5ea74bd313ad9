"""vqae_b200 -- B200-native (sm_100a) inference hot path of 2D-VQ-AE-2 behind the reference's
``vq_ae.model`` / ``vq_ae.layers`` nn.Module API.

    import vqae_b200
    vqae_b200.install_as_vq_ae()                  # `import vq_ae.model` now resolves here
    model = vqae_b200.build_vqae(n_down=3).cuda().eval()
    (enc,), (idx,), (loss,) = model.encoder(x)    # same tuples as the reference

There is no CPU fallback: eval-mode forwards need CUDA tensors and ``libvqae_b200.so``.
"""
from __future__ import annotations

import sys
import types

from . import _instantiate, config, engine  # noqa: F401
from ._instantiate import instantiate
from .config import compose_vqae_conf
from .layers import conv, conv_block, misc, vq  # noqa: F401
from .model import VQAE, Decoder, Encoder  # noqa: F401
from .plan import accelerate  # noqa: F401

__all__ = ["VQAE", "Encoder", "Decoder", "build_vqae", "compose_vqae_conf",
           "install_as_vq_ae", "instantiate", "engine", "set_precision", "accelerate"]


def build_vqae(n_down: int = 4, **conf_overrides) -> VQAE:
    """Random-init VQAE from the hand-composed equivalent of conf/model/vq_ae.yaml.
    ``n_down=4`` is the as-shipped 512^2 model, ``n_down=3`` the README's 256^2 -> 32x32 one."""
    conf = compose_vqae_conf(n_down=n_down, **conf_overrides)
    return instantiate(conf)


def set_precision(module, precision: str):
    """Select the arithmetic of every Encoder/Decoder below ``module``: "fp32" (exact CUDA-core
    kernels, index parity with the reference), "fp16" (tensor-core kernels with fp16 operands,
    fp32 accumulation and an fp32 residual stream -- what torch.autocast('cuda') selects) or
    "fp32tc" (tensor-core kernels with split fp16 operands: fp32-accurate, index parity with the
    reference outside near-ties at several times the speed of "fp32")."""
    if precision is not None and precision not in engine.PRECISIONS:
        raise ValueError(f"precision must be one of {engine.PRECISIONS} or None")
    for m in module.modules():
        if isinstance(m, (Encoder, Decoder)):
            m.precision = precision
        elif "_b200" in m.__dict__:                    # reference modules bound by accelerate()
            m.__dict__["_b200"].precision = precision
    return module


def install_as_vq_ae() -> None:
    """Register this package under the reference's module names so that pickled Hydra
    ``_target_`` strings and ``from vq_ae.model import VQAE`` (extract_embeddings.py:25) bind
    to the B200 path.  Refuses to shadow an already-imported reference package."""
    from . import layers, model
    existing = sys.modules.get("vq_ae")
    if existing is not None and not getattr(existing, "__vqae_b200__", False):
        raise RuntimeError("a different `vq_ae` package is already imported")
    root = types.ModuleType("vq_ae")
    root.__vqae_b200__ = True
    root.__path__ = []  # mark as package
    root.model, root.layers = model, layers
    sys.modules.update({
        "vq_ae": root, "vq_ae.model": model, "vq_ae.layers": layers,
        "vq_ae.layers.vq": layers.vq, "vq_ae.layers.conv_block": layers.conv_block,
        "vq_ae.layers.conv": layers.conv, "vq_ae.layers.misc": layers.misc,
    })
