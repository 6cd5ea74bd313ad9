"""ctypes signatures of the tcgen05 (bf16 tensor-core) entry points of libvqae_b200.so.

Kept apart from ``_lib.py`` so the fp32 exact path and the tensor-core path can be read
separately; both live in the same shared library and the same header.
"""
from __future__ import annotations

import ctypes as C

_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
_fp = C.POINTER(C.c_float)

SIGNATURES: dict = {
    "vqae_pack_same_block_f16": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    "vqae_same_block_f16": (_i, [_vp, _vp, _i, _vp, _fp, _i64, _i, _i, _i, _vp]),
    "vqae_same_block_mma_supported": (_i, [_i, _i, _i]),
    "vqae_same_block_mma_f16": (_i, [_vp, _vp, _vp, _fp, _i64, _i, _i, _i, _vp]),
    "vqae_down_block_mma_supported": (_i, [_i, _i, _i]),
    "vqae_down_block_mma_f16": (_i, [_vp, _vp, _vp, _fp, _i64, _i, _i, _i, _vp]),
    "vqae_stem_in_mma_supported": (_i, [_i, _i, _i]),
    "vqae_stem_in_mma_f32": (_i, [_vp, _i, _i, _vp, _vp, _vp, _i64, _i, _i, _i, _fp, _fp, _vp]),
    "vqae_stem_out_mma_supported": (_i, [_i, _i, _i]),
    "vqae_stem_out_mma_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i64, _i, _i, _i, _vp]),
    "vqae_front_fused_supported": (_i, [_i, _i]),
    "vqae_front_fused_f16": (_i, [_vp, _i, _i, _vp, _vp, _fp, _fp, _vp, _fp, _vp, _fp, _vp, _i64, _i, _i, _vp]),
    "vqae_same_block_mma_split_supported": (_i, [_i, _i, _i]),
    "vqae_same_block_mma_split_f16": (_i, [_vp, _vp, _vp, _vp, _fp, _fp, _i64, _i, _i, _i, _vp]),
    "vqae_same_block_split_supported": (_i, [_i, _i, _i]),
    "vqae_same_block_split_f16": (_i, [_vp, _vp, _vp, _vp, _fp, _fp, _i64, _i, _i, _i, _vp]),
    "vqae_down_block_split_f16": (_i, [_vp, _vp, _vp, _vp, _fp, _fp, _i64, _i, _i, _i, _vp]),
    "vqae_up_block_mma_supported": (_i, [_i, _i, _i]),
    "vqae_up_block_mma_scratch_bytes": (C.c_size_t, [_i64, _i, _i, _i]),
    "vqae_up_block_mma_f16": (_i, [_vp, _vp, _vp, _fp, _vp, C.c_size_t, _i64, _i, _i, _i, _vp]),
    "vqae_same_chain_flag_bytes": (C.c_size_t, [_i, _i64]),
    "vqae_same_chain_supported": (_i, [_i64, _i, _i, _i]),
    "vqae_same_chain_f16": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_size_t, _i, _i64, _i, _i, _i, _vp]),
    "vqae_trunk_resident_max_clusters": (_i, []),
    "vqae_trunk_resident_supported": (_i, [_i64, _i, _i, _i]),
    "vqae_pack_resident_block_f16": (_i, [_vp, _vp, _vp, _i, _f, _vp, _vp]),
    "vqae_trunk_resident_f16": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i64, _i, _i, _i, _vp]),
    "vqae_down_block_pack_elems": (C.c_size_t, [_i]),
    "vqae_pack_down_block_f16": (_i, [_vp, _vp, _vp, _vp, _i, _f, _vp, _vp]),
    "vqae_down_block_f16": (_i, [_vp, _vp, _i, _vp, _fp, _i64, _i, _i, _i, _vp]),
}

# include/vqae_b200_testaids.h (libvqae_b200_testaids.so): tests/ and profiles/ only
AIDS_SIGNATURES: dict = {
    "vqae_tc_mma_bench": (_i, [_i, _i, _i, _i, _vp, _vp]),
    "vqae_tc_mma_bench2": (_i, [_i, _i, _i, _i, _i, _i, _vp, _vp]),
    "vqae_tc_selftest": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "vqae_trunk_resident_set_profile": (None, [_vp]),
    "vqae_same_block_f16_profile": (_i, [_vp, _vp, _i, _vp, _fp, _i64, _i, _i, _i, _vp, _vp]),
    "vqae_quantize_tc_set_profile": (None, [_vp]),
}
