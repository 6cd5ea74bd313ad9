"""TEST INFRASTRUCTURE ONLY -- import shim that lets the *unmodified* reference
(sara-nl/2D-VQ-AE-2, mounted read-only at /root/reference) run in this container.

The reference needs hydra / omegaconf / pytorch_lightning, none of which are
installed here.  This module registers ~70 lines of stand-in
modules in ``sys.modules`` and puts the reference checkout on ``sys.path`` so that
``import vq_ae.model`` executes the reference's own source files.  No reference
source is copied into this repository.

Users: ``oracle/make_golden.py`` and ``tests/test_oracle_vs_reference.py`` (build
container, against /root/reference); ``bench.py --impl reference`` and
``tests/test_gpu_reference.py`` (GPU box, against the unmodified copy under the
git-ignored ``oracle/_ref`` made by ``make -C oracle ref``; /root/reference itself
does not exist there and is never read at run time).  The product package never
imports this file.
"""
from __future__ import annotations

import collections
import collections.abc
import importlib
import os
import sys
import types

def _find_reference_root() -> str:
    """VQAE_REFERENCE_ROOT, else the read-only checkout (build container), else the unmodified
    copy that `make -C oracle ref` placed under oracle/_ref (travels to the GPU box)."""
    env = os.environ.get("VQAE_REFERENCE_ROOT")
    if env:
        return env
    here = os.path.dirname(os.path.abspath(__file__))
    for cand in ("/root/reference", os.path.join(here, "_ref")):
        if os.path.isfile(os.path.join(cand, "vq_ae", "model.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_reference_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "vq_ae", "model.py"))


def _locate(path: str):
    mod_name, _, attr = path.rpartition(".")
    return getattr(importlib.import_module(mod_name), attr)


_SPECIAL = ("_target_", "_recursive_", "_partial_", "_convert_", "_args_")


def _instantiate(config=None, *args, **kwargs):
    """Minimal stand-in for ``hydra.utils.instantiate`` (hydra 1.2 semantics for the
    subset the reference uses: ``_target_``, ``_recursive_``, keyword overrides)."""
    if config is None:
        return None
    if isinstance(config, (list, tuple)):
        return [_instantiate(c) for c in config]
    if not isinstance(config, dict):
        return config
    merged = {**config, **kwargs}
    if "_target_" not in merged:
        return merged
    recursive = merged.get("_recursive_", True)
    target = merged["_target_"]
    params = {k: v for k, v in merged.items() if k not in _SPECIAL}
    if recursive:
        params = {k: _instantiate_nested(v) for k, v in params.items()}
    fn = _locate(target) if isinstance(target, str) else target
    return fn(*args, **params)


def _instantiate_nested(value):
    if isinstance(value, dict):
        if "_target_" in value:
            return _instantiate(value)
        return {k: _instantiate_nested(v) for k, v in value.items()}
    if isinstance(value, (list, tuple)):
        return [_instantiate_nested(v) for v in value]
    return value


def install() -> None:
    """Idempotently install the stand-in modules and expose the reference on sys.path."""
    if "_vqae_ref_shim_installed" in sys.modules:
        return
    if not reference_available():
        raise FileNotFoundError(
            f"reference not found at {REFERENCE_ROOT} (nor a copy under oracle/_ref); run "
            "`make -C oracle ref` in the build container"
        )
    import torch
    from torch import nn

    # conv_block.py:1 does `from collections import Sequence` (removed in py3.10)
    if not hasattr(collections, "Sequence"):
        collections.Sequence = collections.abc.Sequence  # type: ignore[attr-defined]

    # ---- hydra -------------------------------------------------------------------
    hydra = types.ModuleType("hydra")
    hydra.main = lambda *a, **k: (lambda f: f)
    hydra.compose = lambda *a, **k: None
    hydra.initialize_config_dir = lambda *a, **k: None
    hydra_utils = types.ModuleType("hydra.utils")
    hydra_utils.instantiate = _instantiate
    hydra_utils.call = _instantiate
    hydra.utils = hydra_utils
    hydra_core = types.ModuleType("hydra.core")
    hydra_gh = types.ModuleType("hydra.core.global_hydra")
    hydra_gh.GlobalHydra = type("GlobalHydra", (), {"instance": staticmethod(lambda: None)})
    sys.modules.update({
        "hydra": hydra, "hydra.utils": hydra_utils,
        "hydra.core": hydra_core, "hydra.core.global_hydra": hydra_gh,
    })

    # ---- omegaconf ---------------------------------------------------------------
    oc = types.ModuleType("omegaconf")

    class DictConfig(dict):
        pass

    class ListConfig(list):
        pass

    class OmegaConf:
        @staticmethod
        def register_new_resolver(*a, **k):
            return None

        @staticmethod
        def save(*a, **k):
            return None

    oc.DictConfig, oc.ListConfig, oc.OmegaConf, oc.MISSING = DictConfig, ListConfig, OmegaConf, "???"
    sys.modules["omegaconf"] = oc

    # ---- pytorch_lightning -------------------------------------------------------
    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(nn.Module):
        def save_hyperparameters(self, *a, **k):
            return None

    pl.LightningModule = LightningModule
    pl.LightningDataModule = type("LightningDataModule", (), {})
    pl.Callback = type("Callback", (), {})
    pl.Trainer = type("Trainer", (), {})
    pl_utils = types.ModuleType("pytorch_lightning.utilities")
    pl_exc = types.ModuleType("pytorch_lightning.utilities.exceptions")
    pl_exc.MisconfigurationException = type("MisconfigurationException", (Exception,), {})
    pl_types = types.ModuleType("pytorch_lightning.utilities.types")
    pl_types.STEP_OUTPUT = object
    pl.utilities = pl_utils
    sys.modules.update({
        "pytorch_lightning": pl, "pytorch_lightning.utilities": pl_utils,
        "pytorch_lightning.utilities.exceptions": pl_exc,
        "pytorch_lightning.utilities.types": pl_types,
    })

    # torchvision is only used for make_grid (logging); stub it if it fails to import
    try:
        import torchvision.utils  # noqa: F401
    except Exception:  # pragma: no cover
        tv = types.ModuleType("torchvision")
        tvu = types.ModuleType("torchvision.utils")
        tvu.make_grid = lambda *a, **k: None
        tv.utils = tvu
        sys.modules.update({"torchvision": tv, "torchvision.utils": tvu})

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sys.modules["_vqae_ref_shim_installed"] = types.ModuleType("_vqae_ref_shim_installed")
    del torch


def load_reference():
    """Return the reference's own modules: (vq_ae.model, vq_ae.layers.vq,
    vq_ae.layers.conv_block, vq_ae.layers.conv)."""
    install()
    # the B200 package may have been aliased as `vq_ae` (vqae_b200.install_as_vq_ae); drop the
    # aliases so the import below executes the reference's own files
    root = sys.modules.get("vq_ae")
    if root is not None and getattr(root, "__vqae_b200__", False):
        for name in [n for n in sys.modules if n == "vq_ae" or n.startswith("vq_ae.")]:
            del sys.modules[name]
    model = importlib.import_module("vq_ae.model")
    vq = importlib.import_module("vq_ae.layers.vq")
    conv_block = importlib.import_module("vq_ae.layers.conv_block")
    conv = importlib.import_module("vq_ae.layers.conv")
    return model, vq, conv_block, conv
