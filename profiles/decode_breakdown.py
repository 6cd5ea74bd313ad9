"""Per-call CUDA-event breakdown of one decode step (codes -> pixels).  usage:
    python profiles/decode_breakdown.py [fp32|bf16] [batch] [n_down]"""
import sys
from collections import OrderedDict
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "2d-vq-ae-2_b200"), str(REPO)]
import torch  # noqa: E402

import vqae_b200  # noqa: E402
from vqae_b200 import engine as E  # noqa: E402
from vqae_b200 import synthetic as S  # noqa: E402
from vqae_b200.model import _flat_blocks  # noqa: E402


def main():
    precision = sys.argv[1] if len(sys.argv) > 1 else "fp16"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    n_down = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    dev = torch.device("cuda:0")
    m = vqae_b200.build_vqae(n_down=n_down).eval()
    m.load_state_dict(S.make_state_dict(m.state_dict(), seed=1, regime="perturbed"))
    m = m.to(dev)
    dec = m.decoder
    c = 8 * 2 ** n_down
    enc = torch.randn(batch, 32, 32, c, device=dev)
    packed = E.pack_blocks(_flat_blocks(dec.post_enc_layers) + _flat_blocks(dec.up_layers))
    names = {0: "same", 1: "down", 2: "up"}
    chains = {}

    def run(record):
        ev = []

        def mark(label):
            if record:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                ev.append((label, e))
        mark("start")
        h = enc
        runs = dict(E._chain_runs(packed, h.shape[1], h.shape[2], h.shape[0])) if precision == "fp16" else {}
        i = 0
        while i < len(packed):
            pk = packed[i]
            hh = h.shape[1]
            if i in runs:
                j = runs[i]
                h = E.run_blocks_nhwc(packed[i:j], h, precision, chains)
                mark(f"run {j - i}x same C{pk.c_in} @{hh}")
                i = j
            else:
                h = E.fixup_forward_nhwc(pk, h, precision=precision)
                mark(f"{names[pk.mode]} C{pk.c_in}->{pk.c_out} @{hh}")
                i += 1
        E.stem_out(h, dec.out_stem.weight, dec.out_stem.bias, False, precision)
        mark("stem_out")
        return ev

    with torch.no_grad():
        for _ in range(3):
            run(False)
        torch.cuda.synchronize()
        ev = run(True)
        torch.cuda.synchronize()
    agg = OrderedDict()
    total = ev[0][1].elapsed_time(ev[-1][1])
    for (_, e0), (label, e1) in zip(ev, ev[1:]):
        a = agg.setdefault(label, [0, 0.0])
        a[0] += 1
        a[1] += e0.elapsed_time(e1)
    print(f"# decode, precision {precision}, batch {batch}, n_down {n_down}: {total:.3f} ms "
          f"({batch / total * 1e3:.0f} patches/s)")
    print(f"{'call':28s} {'n':>4s} {'ms':>9s} {'share':>7s} {'us/call':>9s}")
    for label, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{label:28s} {n:4d} {ms:9.3f} {ms / total:7.1%} {ms / n * 1e3:9.1f}")


if __name__ == "__main__":
    main()
