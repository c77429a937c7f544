#!/usr/bin/env python
"""Per-config kernel measurements for BASELINE.json's configs C2/C3/C4 (slices that
fit comfortably in HBM), fp32 and fp64: Msamples/s, achieved GB/s on algorithmic
bytes and the fraction of the measured HBM peak.  One JSON line per case.

    python tools/bench_configs.py [--reps 5]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dsp_audio_project_b200 as pkg  # noqa: E402

GAINS = {"Sub-Bass": 6, "Bass": -3, "Low Mids": 4, "High Mids": -6, "Presence": 3, "Brilliance": -9}


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None or ms < best else best
    return best


def report(name, ms, samples, alg_bytes, extra=None):
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    line = {"case": name, "ms": round(ms, 4), "Msamples_per_s": round(samples / (ms * 1e-3) / 1e6, 1),
            "algorithmic_GB": round(alg_bytes / 1e9, 3), "achieved_GBs": round(gbs, 1),
            "frac_of_measured_hbm_peak": round(gbs / peak(), 3)}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="", help="c4: just the 2^16-point FFT case; c3t: just the wide EQ cases")
    args = ap.parse_args()
    if args.only == "c4":
        dev = torch.device("cuda", 0)
        gen = torch.Generator(device=dev).manual_seed(3)
        for tdt, ndt, es, tag, ch in ((torch.float32, np.float32, 4, "f32", 512), (torch.float64, np.float64, 8, "f64", 256)):
            x = torch.rand((ch, 1 << 20), generator=gen, device=dev, dtype=tdt) * 2 - 1
            fft = pkg.FftPlan(65536, ndt, hann=True)
            mag = fft.magnitudes(x)
            ms = timeit(lambda: fft.magnitudes(x, out=mag), args.reps)
            report(f"C4 slice FFT 2^16 {ch}x2^20 {tag}", ms, x.numel(), es * (x.numel() + mag.numel()))
        return
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(1)
    if args.only in ("", "c3t"):
        # C3 at its full channel count, a twelfth of the time axis per call (63 GB in place): the batch is wide
        # enough for the tensor-core form of the cascade (csrc/eq_mma.cu); also 18944 = 148 x 128 channels
        for ch, n in ((65536, 240_000), (18944, 480_000)):
            x = torch.empty((ch, n), device=dev, dtype=torch.float32)
            x.uniform_(-0.25, 0.25, generator=gen)
            for gname, gains in (("C1 gains", GAINS), ("all +15 dB", {k: 15 for k in GAINS})):
                eq = pkg.EqPlan.from_gains(48000, gains, np.float32)
                ms = timeit(lambda: eq.run(x, out=x), args.reps)
                report(f"C3 EQ {ch}x{n} {gname} f32", ms, x.numel(), 8 * x.numel(), {"kernel": eq.kernel_kind(ch, n)})
                x.uniform_(-0.25, 0.25, generator=gen)
            del x
            torch.cuda.empty_cache()
        # C3 in full: 65536 channels x 2.88 M samples as 12 blocks of 240000 with the state carried between the calls
        # (dspb200_eq_run_stream_f32; each block is regenerated in place, 1.51 TB of traffic in all)
        ch, n, blocks = 65536, 240_000, 12
        x = torch.empty((ch, n), device=dev, dtype=torch.float32)
        eq = pkg.EqPlan.from_gains(48000, GAINS, np.float32)
        x.uniform_(-0.25, 0.25, generator=gen)
        _, st = eq.run_stream(x, None, out=x)                   # warm-up
        torch.cuda.synchronize()
        total = 0.0
        st = None
        for b in range(blocks):
            x.uniform_(-0.25, 0.25, generator=gen)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _, st = eq.run_stream(x, st, out=x)
            e1.record()
            torch.cuda.synchronize()
            total += e0.elapsed_time(e1)
        report(f"C3 full EQ {ch}x{n * blocks} as {blocks} streamed blocks, C1 gains f32 (kernel time only)", total,
               ch * n * blocks, 8 * ch * n * blocks, {"kernel": "tensor", "ms_per_block": round(total / blocks, 3)})
        del x
        torch.cuda.empty_cache()
        if args.only == "c3t":
            return
    for tdt, ndt, es, tag in ((torch.float32, np.float32, 4, "f32"), (torch.float64, np.float64, 8, "f64")):
        # C2: SRC 160/147, 1024 ch x 441000
        x = torch.rand((1024, 441000), generator=gen, device=dev, dtype=tdt) - 0.5
        plan = pkg.SrcPlan(160, 147, ndt)
        y = plan.run(x)
        ms = timeit(lambda: plan.run(x, out=y), args.reps)
        report(f"C2 SRC 160/147 1024x441000 {tag}", ms, x.numel(), es * (x.numel() + y.numel()),
               {"kernel": plan.kernel_kind(1024, 441000)})
        del x, y
        # C3 slice: EQ six bands, 4096 ch x 2.88 M samples (of 65536 channels), both gain sets
        ch = 4096 if es == 4 else 2048
        x = (torch.rand((ch, 2_880_000), generator=gen, device=dev, dtype=tdt) - 0.5) * 0.5
        z = torch.empty_like(x)          # out of place, as sistema_ecualizador is (y_n = x_n.copy(), dsp_core.py:230)
        for gname, gains in (("C1 gains", GAINS), ("all +15 dB", {k: 15 for k in GAINS})):
            eq = pkg.EqPlan.from_gains(48000, gains, ndt)
            ms = timeit(lambda: eq.run(x, out=z), args.reps)
            report(f"C3 slice EQ {ch}x2880000 {gname} {tag}", ms, x.numel(), 2 * es * x.numel(),
                   {"kernel": eq.kernel_kind(ch, 2_880_000)})
        del x, z
        # C2-shaped EQ wave: 1024 channels x 10 s at 48 kHz (what the narrow-batch slicing of the tensor form is for)
        x = (torch.rand((1024, 480000), generator=gen, device=dev, dtype=tdt) - 0.5) * 0.5
        z = torch.empty_like(x)
        eq = pkg.EqPlan.from_gains(48000, GAINS, ndt)
        ms = timeit(lambda: eq.run(x, out=z), args.reps)
        report(f"EQ 1024x480000 C1 gains {tag}", ms, x.numel(), 2 * es * x.numel(), {"kernel": eq.kernel_kind(1024, 480000)})
        del x, z
        # C4 slice: 2^16-point frames of 2^20-sample clips, 512 channels (of 4096)
        ch = 512 if es == 4 else 256
        x = torch.rand((ch, 1 << 20), generator=gen, device=dev, dtype=tdt) * 2 - 1
        fft = pkg.FftPlan(65536, ndt, hann=True)
        mag = fft.magnitudes(x)
        ms = timeit(lambda: fft.magnitudes(x, out=mag), args.reps)
        report(f"C4 slice FFT 2^16 {ch}x2^20 {tag}", ms, x.numel(), es * (x.numel() + mag.numel()))
        # 4096-point frames of the same clips (C1/C5 frame size)
        fft2 = pkg.FftPlan(4096, ndt, hann=True)
        mag2 = fft2.magnitudes(x)
        ms = timeit(lambda: fft2.magnitudes(x, out=mag2), args.reps)
        report(f"FFT 4096 {ch}x2^20 {tag}", ms, x.numel(), es * (x.numel() + mag2.numel()))
        del x, mag, mag2
        torch.cuda.empty_cache()
        # either side of the path: int16 export and dB spectra
        z = torch.rand((1024, 480000), generator=gen, device=dev, dtype=tdt) * 2 - 1
        out16, _ = pkg.to_pcm16(z)
        ms = timeit(lambda: pkg.to_pcm16(z, out=out16), args.reps)
        report(f"pcm16 export 1024x480000 {tag} (peak pass + quantise pass)", ms, z.numel(), (2 * es + 2) * z.numel())
        fdb = pkg.FftPlan(4096, ndt, hann=True, db=True)
        mdb = fdb.magnitudes(z)
        ms = timeit(lambda: fdb.magnitudes(z, out=mdb), args.reps)
        report(f"FFT 4096 dB output 1024x480000 {tag}", ms, z.numel() // 4096 * 4096, es * (mdb.numel() * 4096 // 2049 + mdb.numel()))
        del z, out16, mdb
    # loader front end: 1024 stereo float64 clips of 10 s
    fr = torch.rand((1024, 441000, 2), generator=gen, device=dev, dtype=torch.float64) - 0.5
    mono, _ = pkg.mono_normalize(fr)
    ms = timeit(lambda: pkg.mono_normalize(fr), args.reps)
    report("loader front end 1024x441000 stereo f64 -> mono f32 (mean pass + normalise pass)", ms, mono.numel(),
           fr.numel() * 8 + 3 * mono.numel() * 4)


if __name__ == "__main__":
    main()
