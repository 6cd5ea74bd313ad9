#!/usr/bin/env python
"""bench.py -- 256^2 patches/s (encode + quantise) on N B200s, with roofline and CPU baseline.

Contract (one JSON line on stdout from rank 0):
  python bench.py --gpus N --steps K --warmup W            # this framework
  python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port)
For N > 1 launch with torch.distributed.run (one rank per GPU); patches shard by rank with no
data-path collective (weak scaling: 256 patches per GPU per step).

A "step" = one pass of the hot path (stem -> down pyramid -> 50-block trunk -> quantiser) over a
batch of 256 synthetic 256x256x3 patches with the 256-model (n_down=3, C_lat=64, 32x32 codes),
random-init weights in the non-degenerate "perturbed" regime (vqae_b200/synthetic.py).
  value : patches/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e   : the same through the public API from pinned HOST uint8 tiles, H2D + D2H of the code
          indices inside the timed region
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
for _p in (REPO / "2d-vq-ae-2_b200", REPO / "oracle"):
    if str(_p) not in sys.path:
        sys.path.insert(0, str(_p))

import numpy as np  # noqa: E402
import torch  # noqa: E402

PATCH = 256
N_DOWN = 3
BATCH_PER_GPU = 256
FLOP_PER_PATCH = 6.255e9          # 3.127 GMAC encode + quantise, 256-model
WORKLOAD = "encode+quantise 256x256x3 patches, 256-model (n_down=3, C_lat=64, 32x32 codes), batch 256/GPU"


def load_peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])), smax.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# --------------------------------------------------------------------------------------------
# CPU leg: the oracle port of the reference's path, on the host cores
# --------------------------------------------------------------------------------------------
def cpu_encode_sample(sd, n_patches: int, seed: int, reps: int):
    """Times oracle normalise + encoder_forward on `n_patches` u8 tiles; returns patches/s."""
    import vqae_oracle as O
    from vqae_b200 import synthetic as S
    img = S.synthetic_patches_u8(n_patches, PATCH, seed).numpy()
    best = float("inf")
    with torch.no_grad():
        for _ in range(reps):
            t0 = time.perf_counter()
            x = torch.from_numpy(O.normalize_u8(img))
            O.encoder_forward(x, sd)
            best = min(best, time.perf_counter() - t0)
    return n_patches / best


def build_model_and_state(seed: int = 1):
    import vqae_b200
    from vqae_b200 import synthetic as S
    model = vqae_b200.build_vqae(n_down=N_DOWN).eval()
    sd = S.make_state_dict(model.state_dict(), seed=seed, regime="perturbed")
    model.load_state_dict(sd)
    return model, sd


def run_reference(args):
    """--impl reference: the reference's CPU implementation (oracle port; the reference is pure
    Python/PyTorch and /root/reference does not exist on the GPU box), all host threads."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm runs on rank 0 alone and is
    # meant to use all the host threads it can
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    if torch.get_num_threads() < avail:
        torch.set_num_threads(avail)
    _, sd = build_model_and_state()
    cores = torch.get_num_threads()
    sample = 8
    for _ in range(args.warmup):
        cpu_encode_sample(sd, sample, 0, 1)
    t0 = time.perf_counter()
    for s in range(args.steps):
        cpu_encode_sample(sd, sample, 100 + s, 1)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": "patches_per_sec_encode_quantize_256", "value": value,
        "unit": "patches/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{sample} patches per step on the host CPU"},
        "cpu_baseline": {"value": value, "unit": "patches/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} patches x {args.steps} steps, torch {torch.__version__} "
                                   f"CPU kernels, {cores} threads"},
        "e2e": {"value": value, "unit": "patches/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# GPU leg
# --------------------------------------------------------------------------------------------
def _event_time(launch, reps, dev):
    for i in range(3):
        launch(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record()
    for i in range(reps):
        launch(i)
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / reps


def time_dominant_kernel(dev, peaks, precision: str):
    """CUDA-event timing of the dominant kernel alone at the bench shape (batch 256, 32x32, C=64),
    launched through the C-ABI on the current stream, rotating over 4 buffer pairs (4 x 67 MB in
    + 4 x 67 MB out > 126 MB L2)."""
    from vqae_b200 import _lib as L
    from vqae_b200 import engine as E
    lib = L.load()
    B, H, W, C = BATCH_PER_GPU, 32, 32, 64
    nbuf = 4
    xs = [torch.randn(B, H, W, C, device=dev) for _ in range(nbuf)]
    ys = [torch.empty(B, H, W, C, device=dev) for _ in range(nbuf)]
    st = E._stream(dev)
    peak = peaks["bf16_tflops_sustained"]
    if precision == "bf16":
        # the trunk as the step runs it: ONE image-resident launch over 54 blocks (the 4 post layers
        # of the last DownBlock + the 50 pre_enc_layers)
        nblk = 54
        gen = torch.Generator().manual_seed(7)
        packs, scal = [], []
        for i in range(nblk):
            ws = [(torch.randn(C, C, k, k, generator=gen) * 0.05).to(dev) for k in (1, 3, 1)]
            pk = torch.empty(11 * C * C, dtype=torch.bfloat16, device=dev)
            L.check(lib.vqae_pack_resident_block_bf16(E._ptr(ws[0]), E._ptr(ws[1]), E._ptr(ws[2]), C,
                                                      0.2, E._ptr(pk), st), "pack")
            packs.append(pk)
            scal.append([0.01, 0.02, -0.01, 0.03, 0.02, -0.02, 0.01, 0.2])
        w_all = torch.cat(packs)
        scal_dev = torch.tensor(scal, dtype=torch.float32).to(dev)
        assert lib.vqae_trunk_resident_supported(B, H, W, C)

        def launch(i):
            L.check(lib.vqae_trunk_resident_bf16(E._ptr(xs[i % nbuf]), E._ptr(ys[i % nbuf]),
                                                 E._ptr(w_all), E._ptr(scal_dev), nblk, B, H, W, C, st),
                    "vqae_trunk_resident_bf16")
        ms = _event_time(launch, 5, dev)
        flops = 2.0 * B * H * W * C * C * 11 * nblk
        name = ("trunk_resident_tc_kernel (54 fused PreActFixupResBlocks 'same' per launch, C=64, "
                "32x32; tcgen05 bf16, fp32 residual resident in tensor memory, 4-CTA clusters)")
        note = ("whole 54-block run per launch: per 8x32-pixel tile 1x1 + 3x3 circular + 1x1 implicit "
                "GEMMs on tcgen05 (bf16 operands, fp32 TMEM accumulation); the residual stream stays in "
                "tensor memory and branch_conv3 accumulates into it, halo rows go through distributed "
                "shared memory, so HBM traffic is one read + one write of the activations per launch; "
                "every tcgen05.mma is 128x64x16 (N = C = 64), which the tensor pipe issues at 84 "
                "cycles against 32 at full rate (profiles/mma_bench_shift.py), i.e. the kernel's own "
                "ceiling is 38 % of the dense peak; timed alone with CUDA events on the launch stream, "
                f"{nbuf} rotating buffer pairs")
    else:
        w = E.pack_conv_weight(torch.randn(C, C, 3, 3, device=dev) * 0.05)

        def launch(i):
            L.check(lib.vqae_conv_f32(L.CONV_3x3_CIRC, E._ptr(xs[i % nbuf]), E._ptr(w),
                                      E._ptr(ys[i % nbuf]), None, B, H, W, C, C, 0.01, 1, 0.02,
                                      1.0, 0.0, st), "vqae_conv_f32")
        ms = _event_time(launch, 20, dev)
        flops = 2.0 * B * H * W * C * C * 9
        name = "conv_f32_kernel<3x3 circular, BN=64, BK=16> (trunk branch_conv2)"
        note = ("fp32 CUDA-core FFMA kernel measured against the bf16 tensor peak; timed alone "
                "with CUDA events on the launch stream, 4 rotating buffer pairs")
    achieved = flops / (ms * 1e-3) / 1e12
    # dram__bytes_read.sum + dram__bytes_write.sum of one launch at this shape, from the committed
    # ncu --set full capture (profiles/r1_trunk_resident_ncu_summary.txt): 72.0 MB + 16.7 MB
    traffic = 88.77e6 if precision == "bf16" else None
    return {"bound": "tensor", "kernel": name, "achieved": achieved, "peak": peak,
            "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
            "us_per_launch": ms * 1e3, "algorithmic_flops_per_launch": flops, "note": note}


def time_quantizer(dev, peaks):
    """Config 2: ProjectedEMAVectorQuantizer2d on [512,64,32,32] fp32 NHWC, N = 524288: the C-ABI
    call vqae_quantize_f32 (one fused tcgen05 kernel) on preallocated buffers, three rotating
    134 MB inputs (> L2), timed with CUDA events on the launch stream."""
    import ctypes as C
    from vqae_b200 import _lib as L
    from vqae_b200 import engine as E
    from vqae_b200.layers.vq import ProjectedEMAVectorQuantizer2d
    q = ProjectedEMAVectorQuantizer2d(256, 64, 1.0, 0.99, 1e-5, 8).eval().to(dev)
    pq = q.packed()
    lib = L.load()
    n = 512 * 1024
    xs = [torch.randn(n, 64, device=dev) for _ in range(3)]
    outs = [torch.empty(n, 64, device=dev) for _ in range(2)]
    idx = torch.empty(n, dtype=torch.int64, device=dev)
    loss = torch.empty((), device=dev)
    ties = torch.empty((), dtype=torch.int32, device=dev)
    ws = E.workspace(dev, lib.vqae_quantizer_scratch_bytes(n))
    st = E._stream(dev)
    tc = bool(lib.vqae_quantize_tc_supported(C.byref(pq.params), L.LAYOUT_NHWC, L.LAYOUT_NHWC, 1))

    def call(i):
        L.check(lib.vqae_quantize_f32(
            C.byref(pq.params), xs[i % 3].data_ptr(), L.LAYOUT_NHWC, outs[i % 2].data_ptr(),
            L.LAYOUT_NHWC, idx.data_ptr(), loss.data_ptr(), ties.data_ptr(), E.NEAR_TIE_REL_GAP,
            None, ws.data_ptr(), ws.numel(), 512, 1024, st), "vqae_quantize_f32")

    for i in range(3):
        call(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    torch.cuda.synchronize(dev)
    e0.record()
    for i in range(reps):
        call(i)
    e1.record()
    torch.cuda.synchronize(dev)
    us = e0.elapsed_time(e1) / reps * 1e3
    byts = n * 64 * 4 * 2 + n * 8
    gbs = byts / (us * 1e-6) / 1e9
    return {"workload": "ProjectedEMAVectorQuantizer2d [512,64,32,32] fp32 NHWC, K=256, D=8",
            "kernel": "quantize_tc_kernel (tcgen05 L4 filter + exact fp32 argmin + gather, fused "
                      "loss)" if tc else "quantize_kernel (CUDA-core)",
            "us_per_call": us, "algorithmic_bytes": byts, "achieved_gbs": gbs,
            "peak_gbs": peaks["hbm_gbs"], "frac": gbs / peaks["hbm_gbs"],
            "vectors_per_s": n / (us * 1e-6), "near_ties": int(ties.item())}


def run_gpu(args):
    rank, world, local = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist = None

    from vqae_b200 import engine as E
    from vqae_b200 import synthetic as S
    from vqae_b200.extract import encode_patches

    peaks, peaks_kind = load_peaks()
    model, sd = build_model_and_state()
    import vqae_b200
    model = vqae_b200.set_precision(model.to(dev), args.precision)
    enc = model.encoder
    B = BATCH_PER_GPU
    # two distinct resident batches per rank (seeds 42 + rank), uint8 tiles (50 MB each)
    host = [S.synthetic_patches_u8(B, PATCH, 42 + rank + 1000 * j).pin_memory() for j in range(2)]
    resident = [h.to(dev) for h in host]
    host_idx = torch.empty(B, 32, 32, dtype=torch.int64).pin_memory()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_resident(i):
        return encode_patches(enc, resident[i % 2])

    from vqae_b200.extract import StreamingEncoder
    streamer = StreamingEncoder(enc, dev)

    def run_e2e(n_steps):
        """Pinned uint8 tiles -> H2D (side stream, overlapped with the previous batch's encode) ->
        encode -> D2H of the int64 code indices into pinned memory; every step's copies are inside
        the timed region."""
        acc = 0
        for out in streamer.encode_stream(host[i % 2] for i in range(n_steps)):
            acc += int(out[0, 0, 0])             # touch the host result
        return acc

    def timed(fn):
        with torch.no_grad():
            for i in range(args.warmup):
                fn(i)
            barrier()
            l0 = E.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(args.steps):
                fn(i)
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1)
            launches = E.launch_count() - l0
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches = timed(step_resident)
    clocks = sampler.stop() if rank == 0 else None

    with torch.no_grad():
        run_e2e(args.warmup)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        run_e2e(args.steps)
        e1.record()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ms_e2e = max(e0.elapsed_time(e1), 0.0)
    if dist is not None:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())

    value = world * B * args.steps / (ms * 1e-3)
    e2e = world * B * args.steps / (ms_e2e * 1e-3)

    if rank == 0:
        roofline = time_dominant_kernel(dev, peaks, args.precision)
        roofline["peak_source"] = f"{peaks_kind} (MEASURED_PEAKS.json bf16_tflops_sustained)"
        quant = time_quantizer(dev, peaks)
        cpu_cores = torch.get_num_threads()
        cpu_val = cpu_encode_sample(sd, 8, 7, 3) if world == 1 else None
        line = {
            "metric": "patches_per_sec_encode_quantize_256", "value": value, "unit": "patches/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "bf16",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "precision": args.precision,
                       "parallelism": f"patch-sharded x{world}, no data-path collective",
                       "l2": "two rotating resident batches; per-step activation traffic >> 126 MB L2",
                       "whole_step_tflops": value * FLOP_PER_PATCH / 1e12 / world},
            "e2e": {"value": e2e, "unit": "patches/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(B * PATCH * PATCH * 3),
                    "d2h_bytes_per_step": int(B * 32 * 32 * 8),
                    "api": "vqae_b200.extract.StreamingEncoder(model.encoder).encode_stream(pinned "
                           "uint8 tiles): double-buffered H2D on a side stream, D2H of int64 codes",
                    "host_wall_ms_per_step": wall_ms / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "quantizer_microbench": quant,
        }
        if cpu_val is not None:
            line["cpu_baseline"] = {
                "value": cpu_val, "unit": "patches/s", "cores": cpu_cores, "kind": "port",
                "sample": f"8 patches, best of 3, oracle normalise+encoder_forward, torch "
                          f"{torch.__version__} CPU, {cpu_cores} threads"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
