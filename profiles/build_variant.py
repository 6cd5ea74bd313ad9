"""Build an experimental variant of libvqae_b200.so: recompile the named sources with extra -D flags and link
them with the product build's other objects.  usage: build_variant.py OUT.so FLAG[,FLAG..] file.cu [file.cu ..]
Run the variant with VQAE_B200_LIB=<path> (vqae_b200/_lib.py)."""
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "2d-vq-ae-2_b200")]
from vqae_b200.csrc import build as B  # noqa: E402

out, flags, files = sys.argv[1], sys.argv[2].split(","), sys.argv[3:]
B.build_library()
objs = []
for s in B.SOURCES:
    if s in files:
        obj = B.OBJ_DIR / (s + ".variant.o")
        cmd = [B._nvcc(), *B.NVCC_FLAGS, *[f"-D{f}" for f in flags], "-I", str(REPO / "include"), "-I",
               str(B.CSRC), "-c", str(B.CSRC / s), "-o", str(obj)]
        subprocess.run(cmd, check=True, capture_output=True)
        objs.append(str(obj))
    else:
        objs.append(str(B.OBJ_DIR / (s + ".o")))
subprocess.run([B._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out, *objs], check=True)
print("built", out)
