"""Throughput of the paths outside the headline metric (CUDA events): decode (256-model) and the
512-model (n_down=4, C_lat=128) encode / round trip, in both precisions."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "2d-vq-ae-2_b200")]
import torch  # noqa: E402
import vqae_b200  # noqa: E402
from vqae_b200 import synthetic as S  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, reps=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for n_down, size, batch in ((3, 256, 64), (4, 512, 16)):
    m = vqae_b200.build_vqae(n_down=n_down).eval()
    m.load_state_dict(S.make_state_dict(m.state_dict(), seed=1, regime="perturbed"))
    m = m.to(dev)
    x = S.synthetic_patches(batch, size, 7).to(dev)
    for prec in ("fp32", "fp16"):
        vqae_b200.set_precision(m, prec)
        with torch.no_grad():
            (enc,), (idx,), _ = m.encoder(x)
            t_enc = timed(lambda: m.encoder(x))
            t_dec = timed(lambda: m.decoder((enc,)))
        print(f"n_down={n_down} {size}x{size} batch {batch} {prec}: encode {batch / t_enc * 1e3:8.1f} patches/s "
              f"({t_enc:7.2f} ms), decode {batch / t_dec * 1e3:8.1f} patches/s ({t_dec:7.2f} ms)")
