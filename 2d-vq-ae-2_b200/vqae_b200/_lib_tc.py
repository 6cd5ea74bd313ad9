"""ctypes signatures of the tcgen05 (bf16 tensor-core) entry points of libvqae_b200.so.

Kept apart from ``_lib.py`` so the fp32 exact path and the tensor-core path can be read
separately; both live in the same shared library and the same header.
"""
from __future__ import annotations

import ctypes as C

SIGNATURES: dict = {}
