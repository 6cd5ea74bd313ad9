"""On-disk formats of the extracted code maps (scope row f-3).

The reference stores one ``.npy`` per slide under ``<ckpt>/encodings/<parent>/<stem>.npy``
(scripts/extract_embeddings/extract_embeddings.py:183-185; written here by
``vqae_b200.extract.save_encoding``) and packs those trees into ONE HDF5 file for training the
downstream classifier (scripts/convert_npy_embeddings_to_hdf5/convert.py:13-32):

    <common_root>.hdf5
        /<tail of folder 0>/<npy stem>      one dataset per .npy file of that folder
        /<tail of folder 1>/...

where the folders are all directories below ``run_path`` that contain ``.npy`` files,
``common_root`` is their longest common path prefix and a group's name is the remaining tail.  The
reader (``CAMELYON16EmbeddingsDataset``, datamodules/camelyon16.py:226-246) expects the groups
``images`` and ``masks`` and pairs ``images/<key>`` with ``masks/<key>_mask``.

``h5py`` is an optional dependency (absent from the build image): it is imported on use, and every
function takes an ``h5`` argument so that a file-like stand-in can be injected (tests/test_host.py).
"""
from __future__ import annotations

from glob import glob
from itertools import zip_longest
from pathlib import Path
from typing import Iterable, List, Sequence, Tuple

import numpy as np


def _h5py(h5=None):
    if h5 is not None:
        return h5
    try:
        import h5py
    except ImportError as e:  # pragma: no cover - depends on the environment
        raise ImportError("writing / reading the reference's HDF5 code-map files needs h5py "
                          "(pip install h5py)") from e
    return h5py


def find_npy_folders(folder: Path, patterns: Iterable[str] = ("*.npy",)) -> List[Path]:
    """Every directory below ``folder`` holding a file that matches one of ``patterns``
    (convert.py:54-72; glob follows symlinks, like the reference's)."""
    found = sorted({Path(p).parent for pat in patterns
                    for p in glob(str(Path(folder) / "**" / pat), recursive=True)})
    assert len(found) > 0, (
        f"No valid checkpoint folders were found in path {folder}, "
        f"Check that all folders contain all elements from {list(patterns)}")
    return found


def find_common_root(paths: Sequence[Path]) -> Tuple[Path, List[str]]:
    """Longest common prefix of ``paths`` and each path's remaining tail joined with '/'
    (convert.py:35-51)."""
    common_root = Path()
    parts_iterator = zip_longest(*(p.parts for p in paths))
    for parts in parts_iterator:
        if len(set(parts)) != 1:
            rest = [parts] + list(parts_iterator)
            tails = ['/'.join(filter(None, col)) for col in zip(*rest)]
            return common_root, tails
        common_root /= parts[0]
    return common_root, ['' for _ in paths]


def convert_npy_to_hdf5(run_path, h5=None) -> Path:
    """Pack every ``.npy`` code map below ``run_path`` into ``<common_root>.hdf5`` with the
    reference's group layout; returns the file written."""
    in_path = Path(run_path).resolve()
    assert in_path.is_dir(), f'{in_path} does not seem to be a valid dir'
    npy_paths = find_npy_folders(in_path)
    common_root, tails = find_common_root(npy_paths)
    out = Path(str(common_root) + '.hdf5')
    with _h5py(h5).File(str(out), 'w') as f:
        for npy_path, group_name in zip(npy_paths, tails):
            subgroup = f.create_group(name=group_name) if group_name else f
            for npy_array in sorted(npy_path.iterdir()):
                if npy_array.suffix == '.npy':
                    subgroup.create_dataset(npy_array.stem, data=np.load(str(npy_array)))
    return out


def read_code_maps(path, pattern: str = '', h5=None):
    """The reader side (camelyon16.py:226-246): ``(images, masks)`` tuples of arrays for every key
    of ``images`` containing ``pattern``, in sorted key order, masks looked up as ``<key>_mask``."""
    with _h5py(h5).File(str(path), mode='r') as db:
        assert 'images' in db and 'masks' in db
        img_db, mask_db = db['images'], db['masks']
        keys = [k for k in np.sort(list(img_db.keys())) if pattern in k]
        images = tuple(np.asarray(img_db[k]) for k in keys)
        masks = tuple(np.asarray(mask_db[k + '_mask']) for k in keys)
    return images, masks
