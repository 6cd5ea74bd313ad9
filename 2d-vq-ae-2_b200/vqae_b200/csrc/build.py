"""Builds ``libvqae_b200.so`` (the C-ABI library of include/vqae_b200.h) with nvcc for sm_100a.

In-tree, explicit ``nvcc -shared``: the .so sits next to the Python package so it travels to
the GPU box with the repository snapshot.  No torch headers are involved -- the boundary is a
plain C ABI loaded with ctypes (see ``vqae_b200/_lib.py``).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

CSRC = Path(__file__).resolve().parent
PKG = CSRC.parent
REPO = PKG.parent.parent
LIB_PATH = PKG / "libvqae_b200.so"
STAMP = PKG / ".libvqae_b200.stamp"

SOURCES = ["abi.cu", "conv_f32.cu", "stems.cu", "quantize.cu", "quantize_tc.cu", "tc_kernels.cu", "tc_down.cu", "tc_chain.cu", "tc_resident.cu", "tc_bench.cu", "up_tail.cu", "up_head.cu"]
HEADERS = ["common.cuh", "kernels.cuh", "tc_common.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise FileNotFoundError("nvcc not found; cannot build libvqae_b200.so")


def _digest() -> str:
    h = hashlib.sha256()
    files = [CSRC / s for s in SOURCES + HEADERS if (CSRC / s).exists()]
    files.append(REPO / "include" / "vqae_b200.h")
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile if sources changed since the last build; return the library path."""
    digest = _digest()
    if not force and LIB_PATH.exists() and STAMP.exists() and STAMP.read_text() == digest:
        return LIB_PATH
    srcs = [str(CSRC / s) for s in SOURCES if (CSRC / s).exists()]
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", str(REPO / "include"), "-I", str(CSRC),
           "-o", str(LIB_PATH), *srcs]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    (PKG / "build.log").write_text(" ".join(cmd) + "\n" + log)
    if proc.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libvqae_b200.so (see output above)")
    if verbose:
        print(log)
    STAMP.write_text(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
