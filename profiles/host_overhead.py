"""Host enqueue time per step against device time, and CUDA-graph replay of the same step.
usage: python profiles/host_overhead.py [encode256|roundtrip512]"""
import sys
import time
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO), str(REPO / "2d-vq-ae-2_b200")]
import torch  # noqa: E402
import bench  # noqa: E402
import vqae_b200  # noqa: E402
from vqae_b200 import synthetic as S  # noqa: E402
from vqae_b200.extract import encode_patches  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "encode256"
w = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
model, _ = bench.build_model_and_state(w["n_down"])
model = vqae_b200.set_precision(model.to(dev), "fp16")
enc = model.encoder
B, P = w["batch"], w["patch"]
x = S.synthetic_patches_u8(B, P, 42).to(dev)


def step():
    if wl == "encode256":
        return encode_patches(enc, x)
    _, idx, _, _, _ = enc.encode(x, want_quantized=True)
    return model.decode_codes(idx)


with torch.no_grad():
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    n = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        step()
    t1 = time.perf_counter()
    e1.record()
    torch.cuda.synchronize()
    print(f"{wl}: host enqueue {1e3 * (t1 - t0) / n:.3f} ms/step, device {e0.elapsed_time(e1) / n:.3f} ms/step")
    # one step at a time (a synchronize between steps): what the host overhead costs when it is exposed
    t0 = time.perf_counter()
    for _ in range(n):
        step()
        torch.cuda.synchronize()
    print(f"{wl}: synchronised every step {1e3 * (time.perf_counter() - t0) / n:.3f} ms/step")
    # CUDA graph of the step
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        out = step()
    torch.cuda.synchronize()
    ref = step()
    g.replay()
    torch.cuda.synchronize()
    print("graph replay equals eager:", bool(torch.equal(out, ref)))
    t0 = time.perf_counter()
    for _ in range(n):
        g.replay()
        torch.cuda.synchronize()
    print(f"{wl}: graph replay, synchronised every step {1e3 * (time.perf_counter() - t0) / n:.3f} ms/step")
