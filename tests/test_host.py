"""CPU: host-side logic -- drop-in module surface, C-ABI exports, error behaviour, sharding."""
import ctypes
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

import helpers as H
import vqae_b200
from vqae_b200 import _lib, sharding
from vqae_b200.extract import cast_to_lowest_dtype, tiles_to_map

REPO = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    header = (REPO / "include" / "vqae_b200.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(vqae_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 17
    lib = _lib.load()                      # builds with nvcc if needed; no GPU required
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/vqae_b200.h but not exported"
    from vqae_b200 import _lib_tc
    assert declared == set(_lib.SIGNATURES) | set(_lib_tc.SIGNATURES)
    # test / measurement aids live in their own library and header, outside the product ABI
    aid_header = re.sub(r"/\*.*?\*/", "", (REPO / "include" / "vqae_b200_testaids.h").read_text(),
                        flags=re.S)
    aid_declared = set(re.findall(r"\b(vqae_[a-z0-9_]+)\s*\(", aid_header))
    aids = _lib.load_testaids()
    assert aid_declared == set(_lib_tc.AIDS_SIGNATURES) and not (aid_declared & declared)
    for name in sorted(aid_declared):
        assert hasattr(aids, name) and not hasattr(lib, name), name
    assert lib.vqae_abi_version() == 4
    assert lib.vqae_error_string(3).decode().startswith("VQ dim != channel dim")


def test_state_dict_surface():
    m = vqae_b200.build_vqae(n_down=3)
    sd = m.state_dict()
    assert len(sd) == 1526 and sum(p.numel() for p in m.parameters()) == 4934287
    for key in ("encoder.in_stem.weight", "encoder.down_layers.0.layers.0.layers.1.skip_conv.weight",
                "encoder.pre_enc_layers.0.49.bias4", "encoder.vq_layers.0.embed",
                "encoder.vq_layers.0.first_pass", "encoder.vq_layers.0.proj_out.bias",
                "decoder.up_layers.0.layers.2.layers.1.branch_conv2.weight",
                "decoder.post_enc_layers.0.0.scale", "decoder.out_stem.bias"):
        assert key in sd, key
    assert sd["encoder.pre_enc_layers.0.0.bias1a"].shape == (1,)
    assert sd["encoder.vq_layers.0.embed"].shape == (256, 8)
    vq = m.encoder.vq_layers[0]
    assert vq.embedding_dim == 8 and vq.num_embeddings == 256       # vq.py:168-177
    assert m.encoder.down_layers[0].out_channels == 64
    assert len(m.encoder.pre_enc_layers[0]) == 50 and len(m.decoder.post_enc_layers[0]) == 50


def _drop_vq_ae():
    for name in [n for n in sys.modules if n == "vq_ae" or n.startswith("vq_ae.")]:
        del sys.modules[name]


def test_install_as_vq_ae_aliases():
    _drop_vq_ae()
    try:
        vqae_b200.install_as_vq_ae()
        from vq_ae.layers.vq import ProjectedEMAVectorQuantizer2d
        from vq_ae.model import VQAE
        assert VQAE is vqae_b200.VQAE
        assert ProjectedEMAVectorQuantizer2d is vqae_b200.vq.ProjectedEMAVectorQuantizer2d
    finally:
        _drop_vq_ae()


def test_no_cpu_fallback_and_reference_errors():
    m = vqae_b200.build_vqae(n_down=3).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 64, 64))
    vq = m.encoder.vq_layers[0]
    with pytest.raises(NotImplementedError, match="VQ dim != channel dim"):   # vq.py:100-104
        vq(torch.zeros(1, 5, 4, 4))
    with pytest.raises(AssertionError):                                       # vq.py:98
        vq(torch.zeros(4, 64))
    with pytest.raises(RuntimeError, match="training-mode"):
        m.train()(torch.zeros(1, 3, 64, 64))


def test_checkpoint_roundtrip(tmp_path):
    m = vqae_b200.build_vqae(n_down=3)
    conf = vqae_b200.compose_vqae_conf(n_down=3)
    hp = {k: conf[k] for k in ("optim_conf", "loss_f_conf", "encoder_conf", "decoder_conf")}
    torch.save({"state_dict": m.state_dict(), "hyper_parameters": hp}, tmp_path / "m.ckpt")
    m2 = vqae_b200.VQAE.load_from_checkpoint(str(tmp_path / "m.ckpt"))
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))


def test_shard_ranges_cover_exactly():
    for n, w in ((38025, 8), (7, 8), (256, 1), (0, 4), (1000, 3)):
        spans = [sharding.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert sharding.slide_grid((50000, 50000), 256) == (195, 195)
    assert sharding.patch_rc(389, 195) == (1, 194)
    with pytest.raises(ValueError):
        sharding.shard_range(10, 4, 4)


def test_tiles_to_map_and_dtype_narrowing():
    tiles = torch.arange(6 * 2 * 2, dtype=torch.uint8).view(6, 2, 2)
    m = tiles_to_map(tiles, (2, 3))
    assert m.shape == (4, 6) and torch.equal(m[2:4, 4:6], tiles[5])
    import vqae_oracle as O
    assert np.array_equal(m.numpy(), O.stitch_code_map(tiles.numpy(), 2, 3))
    assert cast_to_lowest_dtype(np.array([0, 255])).dtype == np.uint8
    assert cast_to_lowest_dtype(np.array([0, 1])).dtype == bool


_GLOO_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from vqae_b200 import sharding
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2],
                        rank=int(sys.argv[3]), world_size=2)
n = 7
lo, hi = sharding.shard_range(n, dist.get_rank(), 2)
tiles = (torch.arange(lo, hi, dtype=torch.uint8).view(-1, 1, 1) * torch.ones(1, 4, 4, dtype=torch.uint8))
full = sharding.gather_code_tiles(tiles, n)
assert full.shape == (n, 4, 4) and [int(t[0, 0]) for t in full] == list(range(n)), full[:, 0, 0]
dist.destroy_process_group()
print("ok")
'''


def test_code_tile_gather_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), str(REPO / "2d-vq-ae-2_b200"), port,
                               str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    for p in procs:
        out, err = p.communicate(timeout=120)
        assert p.returncode == 0 and "ok" in out, err[-2000:]


class _FakeH5:
    """Dict-backed stand-in for the h5py API surface the reference uses (File / create_group /
    create_dataset / `in` / keys / indexing); h5py itself is not in the build image."""
    store = {}

    class Group(dict):
        def create_group(self, name):
            g = _FakeH5.Group()
            self[name] = g
            return g

        def create_dataset(self, name, data):
            self[name] = np.asarray(data)

    class File:
        def __init__(self, path, mode='r'):
            self.path, self.mode = path, mode
            if 'w' in mode:
                _FakeH5.store[path] = _FakeH5.Group()
            self.root = _FakeH5.store[path]

        def __enter__(self):
            return self.root

        def __exit__(self, *a):
            return False


def test_hdf5_layout_matches_the_reference_converter(tmp_path):
    """scripts/convert_npy_embeddings_to_hdf5/convert.py:21-32 (writer) and
    datamodules/camelyon16.py:226-246 (reader): groups named by the folder tails below the common
    root, one dataset per .npy stem, images/<key> paired with masks/<key>_mask."""
    from vqae_b200 import formats as F
    from vqae_b200.extract import save_encoding
    root, tails = F.find_common_root([Path("/a/b/images"), Path("/a/b/masks")])
    assert root == Path("/a/b") and tails == ["images", "masks"]
    root, tails = F.find_common_root([Path("/a/x/images/t"), Path("/a/y/masks")])
    assert root == Path("/a") and tails == ["x/images/t", "y/masks"]
    assert F.find_common_root([Path("/a/b")]) == (Path("/a/b"), [""])
    rng = np.random.default_rng(0)
    maps = {"images/normal_001": rng.integers(0, 256, (64, 96)).astype(np.uint8),
            "images/tumor_002": rng.integers(0, 256, (32, 32)).astype(np.uint8),
            "masks/normal_001_mask": np.zeros((64, 96), bool),
            "masks/tumor_002_mask": rng.integers(0, 2, (32, 32)).astype(bool)}
    for name, arr in maps.items():                     # <ckpt>/encodings/<parent>/<stem>.npy
        save_encoding(tmp_path / "run", name, arr)
    out = F.convert_npy_to_hdf5(tmp_path / "run", h5=_FakeH5)
    assert out == (tmp_path / "run" / "encodings").with_suffix(".hdf5")
    db = _FakeH5.store[str(out)]
    assert sorted(db) == ["images", "masks"]
    assert sorted(db["images"]) == ["normal_001", "tumor_002"]
    images, masks = F.read_code_maps(out, h5=_FakeH5)
    assert np.array_equal(images[0], maps["images/normal_001"]) and images[0].dtype == np.uint8
    assert np.array_equal(masks[1], maps["masks/tumor_002_mask"])
    only_tumor, _ = F.read_code_maps(out, pattern="tumor", h5=_FakeH5)
    assert len(only_tumor) == 1 and np.array_equal(only_tumor[0], maps["images/tumor_002"])


def test_cpulist_parsing_for_numa_binding():
    from vqae_b200.sharding import _parse_cpulist
    assert _parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert _parse_cpulist("5") == {5}
