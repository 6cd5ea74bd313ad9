import sys
sys.path[:0] = ["/root/repo/2d-vq-ae-2_b200"]
import torch
from vqae_b200 import _lib as L
from vqae_b200 import engine as E
lib = L.load(); dev = torch.device("cuda:0")
out = torch.zeros(2, dtype=torch.int64, device=dev)
for n in (64, 128):
    for layout in (0, 2):
        for stride in (128, 8, 1, 3, 10, 35):
            for nacc in (2, 4):
                L.check(L.load_testaids().vqae_tc_mma_bench(n, layout | (nacc << 4), 2000, stride, E._ptr(out), E._stream(dev)), "bench")
                torch.cuda.synchronize()
                c, r = out.tolist()
                print(f"N={n} layout={layout} a_shift_rows={stride} nacc={nacc}: {c / r:.1f} cycles/MMA")
print("# small N (the low-channel levels issue N = 16 / 32 instructions)")
for n in (16, 32, 48, 64, 96, 128, 192, 256):
    L.check(L.load_testaids().vqae_tc_mma_bench(n, 0 | (2 << 4), 2000, 128, E._ptr(out), E._stream(dev)), "bench")
    torch.cuda.synchronize()
    c, r = out.tolist()
    print(f"N={n} layout=0: {c / r:.1f} cycles/MMA (ideal {128 * n / 256:.0f})")
