"""Minimal ``hydra.utils.instantiate`` for the subset the reference's model tree uses.

The reference's modules receive config dicts and call ``hydra.utils.instantiate`` themselves
(``_recursive_: False``; vq_ae/model.py:141-176, layers/conv_block.py:110-126,164-189).  The
drop-in modules in this package do the same through this function, so no Hydra install is
needed.  ``_target_`` strings naming the reference's classes (``vq_ae.model.Encoder`` ...)
resolve to this package's classes.
"""
from __future__ import annotations

import importlib
from typing import Any, Callable, Dict

_SPECIAL = ("_target_", "_recursive_", "_partial_", "_convert_", "_args_")

# reference dotted path -> this package's module holding the same-named symbol
_ALIASES: Dict[str, str] = {
    "vq_ae.model": "vqae_b200.model",
    "vq_ae.layers.vq": "vqae_b200.layers.vq",
    "vq_ae.layers.conv_block": "vqae_b200.layers.conv_block",
    "vq_ae.layers.conv": "vqae_b200.layers.conv",
    "vq_ae.layers.misc": "vqae_b200.layers.misc",
    "utils.conf_helpers": "vqae_b200._instantiate",
}


def locate(path: str) -> Callable:
    mod_name, _, attr = path.rpartition(".")
    mod_name = _ALIASES.get(mod_name, mod_name)
    return getattr(importlib.import_module(mod_name), attr)


def _is_conf(v: Any) -> bool:
    return isinstance(v, dict) and "_target_" in v


def _nested(value: Any) -> Any:
    if _is_conf(value):
        return instantiate(value)
    if isinstance(value, dict):
        return {k: _nested(v) for k, v in value.items()}
    if isinstance(value, (list, tuple)):
        return [_nested(v) for v in value]
    return value


def instantiate(config: Any = None, *args: Any, **kwargs: Any) -> Any:
    if config is None:
        return None
    if isinstance(config, (list, tuple)):
        return [instantiate(c) for c in config]
    if not isinstance(config, dict):
        return config
    merged = {**config, **kwargs}
    if "_target_" not in merged:
        return merged
    recursive = merged.get("_recursive_", True)
    params = {k: v for k, v in merged.items() if k not in _SPECIAL}
    if recursive:
        params = {k: _nested(v) for k, v in params.items()}
    target = merged["_target_"]
    fn = locate(target) if isinstance(target, str) else target
    return fn(*args, **params)


# ---- utils/conf_helpers.py:76-136 equivalents (dict-of-named-items -> list) -------------------
def listify_nested_conf(conf: Any) -> Any:
    if isinstance(conf, dict):
        if "_target_" not in conf:
            return listify_nested_conf(list(conf.values()))
        return {k: listify_nested_conf(v) for k, v in conf.items()}
    if isinstance(conf, (list, tuple)):
        return [listify_nested_conf(v) for v in conf]
    return conf


def instantiate_dictified_listconf(**nested_conf: Any):
    de_nested = listify_nested_conf(nested_conf)
    if isinstance(de_nested, list):
        return [instantiate(elem) for elem in de_nested]
    return [instantiate(de_nested)]
