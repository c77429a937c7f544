#!/usr/bin/env python
"""bench.py -- throughput of the dsp_core hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload chain|src|eq|fft] [--clips C] [--dtype f32|f64]

Default workload ("chain") is one wave of BASELINE.json's config C5 shaped like
config C2: `--clips` (default 18944 = 148 SMs x 128 channels, 88 GB of device
buffers) synthetic clips of 10 s @ 44.1 kHz per GPU, SRC 160/147 -> six-band EQ
-> non-overlapping 4096-point Hann magnitude spectra, float32.  A "step" is one
pass of that chain over the wave.  The metric is Msamples/s = input samples
consumed per second, whole job (all ranks).  `value` is timed with CUDA events,
inputs resident in HBM; `e2e` is the same chain through the host-buffer C-ABI
call (pinned host memory in, host memory out, copies inside the timed region)
on calls of `--e2e-clips` clips (default 1024: 4.7 GB of pinned memory).

One JSON line is printed by rank 0.  `--impl reference` times the CPU oracle
port of the reference's own algorithm (dense zero-stuffed convolution,
lfilter cascade, recursive FFT) on all host cores instead.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS_IN = 44100
CLIP_SAMPLES = 441000          # 10 s @ 44.1 kHz
L_UP, M_DOWN = 160, 147
N_FFT = 4096
GAINS = {"Sub-Bass": 6, "Bass": -3, "Low Mids": 4, "High Mids": -6, "Presence": 3, "Brilliance": -9}
FALLBACK_HBM_GBS = 6650.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="chain", choices=["chain", "c5job", "src", "eq", "fft"])
    ap.add_argument("--clips", type=int, default=18944, help="clips (channels) per GPU per step / per wave")
    ap.add_argument("--job-clips", type=int, default=262144, help="c5job: clips of the whole job (all GPUs)")
    ap.add_argument("--e2e-clips", type=int, default=1024, help="clips per host-buffer call of the e2e leg")
    ap.add_argument("--f64-clips", type=int, default=4736, help="clips of the float64 sub-line (N=1 only)")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-f64", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="N > 1: skip the side-stream spectra all-gather measurement")
    ap.add_argument("--gather-clips", type=int, default=1024, help="clips per rank whose spectra the gather leg exchanges")
    return ap.parse_args()


# ----------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference algorithm on the host cores
# ----------------------------------------------------------------------------
CPU_SAMPLE_SECONDS_OF_AUDIO = 0.25


def _cpu_one_clip(seed):
    import numpy as np
    from oracle import dsp_oracle as o

    n = int(FS_IN * CPU_SAMPLE_SECONDS_OF_AUDIO)
    x = np.random.default_rng(seed).uniform(-0.5, 0.5, n).astype(np.float32)
    y, fs2 = o.resample_reference_form(x, FS_IN, M_DOWN, L_UP)     # dense form, as dsp_core.py:148-170
    z = o.equalizer(y, fs2, GAINS)
    mags = o.frame_magnitudes(z, N_FFT)
    return float(np.sum(mags)) + float(z[0])


def cpu_reference_step(pool, cores):
    """One bounded sample: every core runs the reference-form chain on one
    0.25 s clip.  Returns (seconds, input samples processed)."""
    t0 = time.perf_counter()
    pool.map(_cpu_one_clip, range(cores))
    dt = time.perf_counter() - t0
    return dt, cores * int(FS_IN * CPU_SAMPLE_SECONDS_OF_AUDIO)


def make_pool():
    import multiprocessing as mp

    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    cores = max(1, min(cores, 64))
    ctx = mp.get_context("fork")
    return ctx.Pool(cores), cores


def cpu_baseline(repeats=2):
    pool, cores = make_pool()
    try:
        cpu_reference_step(pool, cores)            # warm the workers
        best = None
        for _ in range(repeats):
            dt, n = cpu_reference_step(pool, cores)
            rate = n / dt / 1e6
            best = rate if best is None or rate > best else best
    finally:
        pool.close()
        pool.join()
    return {"value": best, "unit": "Msamples/s", "cores": cores, "kind": "port",
            "sample": f"{cores} clips x {CPU_SAMPLE_SECONDS_OF_AUDIO} s @44.1 kHz, one per core: oracle port of the "
                      "reference chain (dense zero-stuffed np.convolve SRC 160/147, lfilter EQ, recursive FFT 4096)"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pool, cores = make_pool()
    try:
        for _ in range(max(args.warmup, 1)):
            cpu_reference_step(pool, cores)
        total_t, total_n = 0.0, 0
        for _ in range(args.steps):
            dt, n = cpu_reference_step(pool, cores)
            total_t += dt
            total_n += n
    finally:
        pool.close()
        pool.join()
    value = total_n / total_t / 1e6
    line = {
        "impl": "reference", "metric": "Msamples/s SRC->EQ->FFT chain", "value": value, "unit": "Msamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": "port",
                         "sample": f"each step: {cores} clips x {CPU_SAMPLE_SECONDS_OF_AUDIO} s, one per core"},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------
def workload_config(args):
    cfg = {
        "workload": "C5 chain, one wave shaped as C2: clips x 10 s @44.1 kHz -> SRC 160/147 -> 6-band EQ "
                    "(gains 6,-3,4,-6,3,-9 dB) -> 4096-pt Hann |FFT| frames",
        "selected": args.workload, "clips_per_gpu": args.clips, "clip_samples": CLIP_SAMPLES,
        "L": L_UP, "M": M_DOWN, "n_fft": N_FFT,
        "e2e_clips_per_call": getattr(args, "e2e_clips", None),
        "l2": "inputs per step exceed the 126 MB L2 (no flush needed)",
        "parallelism": f"channel-sharded x{args.gpus}, no collective",
    }
    if args.workload == "c5job":
        cfg["workload"] = ("C5 job: %d clips x 10 s @44.1 kHz in all, sharded by clip over the GPUs, processed in waves of "
                           "<= %d clips generated on the device by dspb200_generate_uniform_f32 inside the timed region "
                           "-> SRC 160/147 -> 6-band EQ -> 4096-pt Hann |FFT| frames" % (args.job_clips, args.clips))
        cfg["job_clips"] = args.job_clips
    return cfg


class ClockSampler:
    """Samples SM clock and throttle reasons while the timed region runs."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        self._nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None
        self._thr = threading.Thread(target=self._loop, daemon=True)
        self._thr.start()

    def _loop(self):
        names = {
            "hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
            "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80,
        }
        while not self._stop.is_set():
            try:
                if self._nvml:
                    p = self._nvml
                    self.samples.append(p.nvmlDeviceGetClockInfo(self._h, p.NVML_CLOCK_SM))
                    r = p.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(
                        p, "nvmlDeviceGetCurrentClocksEventReasons") else p.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
                else:
                    out = subprocess.run(
                        ["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm",
                         "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    a, b = out.strip().split(",")
                    self.samples.append(int(a))
                    self.max_mhz = int(b)
            except Exception:
                pass
            self._stop.wait(0.002)

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=2)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def traffic_from_profiles(kernel, clips):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (taken on
    `clips_per_launch` clips; the kernels stream, so traffic scales with the clip count)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            t = json.load(fh)
        v = t.get(kernel)
        return None if v is None else v * clips / t.get("clips_per_launch", 1024)
    except Exception:
        return None


FP64_TFMA_PEAK = 17.9      # measured on B200 with tools/microbench.cu (DFMA issue rate, round 1): 35.8 TFLOP/s


def oracle_spot_check(x_rows, z_rows, mag_rows):
    """float64 oracle (closed-form polyphase SRC, lfilter cascade, recursive FFT frames) of whole clips of the timed
    wave, compared with what the timed kernels left in z / mag.  Returns the worst errors in the tests' measures."""
    import numpy as np
    from oracle import dsp_oracle as o

    ez, em = 0.0, 0.0
    for x, z, m in zip(x_rows, z_rows, mag_rows):
        yo, fs2 = o.resample_closed_form(np.asarray(x, dtype=np.float64), FS_IN, M_DOWN, L_UP)
        zo = o.equalizer(yo, fs2, GAINS)
        ez = max(ez, o.full_scale_err(z, zo))
        frames = [0, 58, 116]
        for f in frames:
            mo = o.frame_magnitudes(zo[f * N_FFT:(f + 1) * N_FFT], N_FFT)[0]
            em = max(em, o.rel_err(m[f], mo))
    return ez, em


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import dsp_audio_project_b200 as pkg
    from dsp_audio_project_b200 import _lib, shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    # CPU baseline first: its worker pool is forked before any CUDA context exists
    cpu_line = None
    if not args.no_cpu_baseline and world == 1 and rank == 0:
        cpu_line = cpu_baseline()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep this rank's threads (and so its pinned host buffers, by first touch) on the CPUs /
        # NUMA node closest to its GPU: the host<->device legs of e2e share the host fabric
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        except Exception:
            pass
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    np_dt = np.float32 if args.dtype == "f32" else np.float64
    t_dt = torch.float32 if args.dtype == "f32" else torch.float64
    esize = 4 if args.dtype == "f32" else 8
    chain = pkg.Chain(L_UP, M_DOWN, FS_IN, GAINS, n_fft=N_FFT, dtype=np_dt)
    n_out = chain.out_len(CLIP_SAMPLES)
    n_frames = n_out // N_FFT
    bins = N_FFT // 2 + 1

    # ---- this rank's share of the work and its waves --------------------------------------------------------------
    if args.workload == "c5job":
        c0, c1 = shard.channel_block(args.job_clips, world, rank)
        my_clips = c1 - c0
        n_waves = max(1, -(-my_clips // args.clips))
        per = -(-my_clips // n_waves)
        per = -(-per // 128) * 128                       # whole groups of 128 channels: what the tensor-core kernels walk
        waves = []
        at = c0
        while at < c1:
            waves.append((at, min(per, c1 - at)))
            at += per
        clips = max(w[1] for w in waves) if waves else 0
        scaling = "strong"
    else:
        clips = args.clips
        waves = [(rank * clips, clips)]
        scaling = "weak"

    # persistent device buffers so every step reuses the same memory.  Config C5 keeps z and the spectra; y is never
    # materialised: the fused kernel goes from x to z, the cascade writes y into z and equalises in place
    sched = None
    if args.workload == "c5job":
        # the job's waves go through the package's wave scheduler: two x buffers, the next wave generated on a side
        # stream while the current wave's spectra are computed
        sched = pkg.WaveScheduler(chain, clips, CLIP_SAMPLES, dev)
        x, z, mag = sched.x[0], sched.z, sched.mag
    else:
        x = torch.empty((clips, CLIP_SAMPLES), dtype=t_dt, device=dev)
        z = torch.empty((clips, n_out), dtype=t_dt, device=dev)
        mag = torch.empty((clips, n_frames, bins), dtype=t_dt, device=dev)
    seed = 4 + rank                                      # seed 4 + rank

    def generate(first, count):
        pkg.generate_uniform(x[:count], seed, -0.5, 0.5, first_channel=first)

    generate(*waves[0])
    kind = chain.kernel_kind(clips, CLIP_SAMPLES)

    def step_chain():
        chain.run(x, z=z, mag=mag)

    def step_job():
        sched.run(waves, lambda xv, first, count: pkg.generate_uniform(xv, seed, -0.5, 0.5, first_channel=first))

    def step_srceq():
        if kind == "fused":
            chain.run_fused(x, out=z)
        else:
            chain.src.run(x, out=z)
            chain.eq.run(z, out=z)

    def step_src():
        chain.src.run(x, out=z)

    def step_eq():
        chain.eq.run(z, out=z)

    def step_fft():
        chain.fft.magnitudes(z, out=mag)

    step = {"chain": step_chain, "c5job": step_job, "src": step_src, "eq": step_eq, "fft": step_fft}[args.workload]
    step_chain()                      # populate z for the single-kernel workloads
    torch.cuda.synchronize()
    my_total = sum(w[1] for w in waves)
    samples_per_step = {"chain": clips * CLIP_SAMPLES, "c5job": my_total * CLIP_SAMPLES, "src": clips * CLIP_SAMPLES,
                        "eq": clips * n_out, "fft": clips * n_frames * N_FFT}[args.workload]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    # per-kernel timing (same stream, events between the launches), on the wave the buffers hold
    if args.workload == "c5job":
        generate(*waves[0])
    if kind == "fused" and args.workload in ("chain", "c5job"):
        names, fns = ["src_eq", "fft"], [step_srceq, step_fft]
    else:
        names, fns = ["src", "eq", "fft"], [step_src, step_eq, step_fft]
    per_kernel = {k: 0.0 for k in names}
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(fns) + 1)] for _ in range(args.steps)]
    for it in range(args.steps):
        evs[it][0].record()
        for j, fn in enumerate(fns):
            fn()
            evs[it][j + 1].record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    for it in range(args.steps):
        for j, k in enumerate(names):
            per_kernel[k] += evs[it][j].elapsed_time(evs[it][j + 1]) / args.steps

    t = torch.tensor([ms, float(samples_per_step)], dtype=torch.float64, device=dev)
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total_samples = float(t[1].item())
    else:
        total_samples = float(samples_per_step)
    ms_max = float(tmax[0].item())
    value = total_samples * args.steps / (ms_max * 1e-3) / 1e6

    # ---- parity of the timed wave itself: two whole clips against the float64 oracle (outside the timed region) -----
    parity = None
    if rank == 0 and not args.no_parity and args.workload in ("chain", "c5job") and args.dtype == "f32":
        step_chain()
        torch.cuda.synchronize()
        rows = [0, clips - 1]
        ez, em = oracle_spot_check([x[r].cpu().numpy() for r in rows], [z[r].cpu().numpy() for r in rows],
                                   [mag[r].cpu().numpy() for r in rows])
        parity = {"z_full_scale": ez, "mag_rel": em, "clips": rows, "frames_checked": [0, 58, 116],
                  "tolerance": {"z_full_scale": 1e-4, "mag_rel": 1e-4},
                  "against": "oracle/dsp_oracle.py (float64: closed-form polyphase SRC, lfilter cascade, recursive FFT)"}

    # ---- float64 sub-line: the reference's own arithmetic type, parity kernels, N = 1 only ---------------------------
    f64_line = None
    if (rank == 0 and world == 1 and not args.no_f64 and args.workload == "chain" and args.dtype == "f32"
            and args.f64_clips > 0):
        c64 = min(args.f64_clips, clips)
        ch64 = pkg.Chain(L_UP, M_DOWN, FS_IN, GAINS, n_fft=N_FFT, dtype=np.float64)
        x64 = torch.empty((c64, CLIP_SAMPLES), dtype=torch.float64, device=dev)
        pkg.generate_uniform(x64, seed, -0.5, 0.5)
        z64 = torch.empty((c64, n_out), dtype=torch.float64, device=dev)
        m64 = torch.empty((c64, n_frames, bins), dtype=torch.float64, device=dev)
        f_names = ["src", "eq", "fft"]
        f_fns = [lambda: ch64.src.run(x64, out=z64), lambda: ch64.eq.run(z64, out=z64),
                 lambda: ch64.fft.magnitudes(z64, out=m64)]
        for _ in range(2):
            for fn in f_fns:
                fn()
        f_steps = 3
        fe = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(f_steps)]
        for it in range(f_steps):
            fe[it][0].record()
            for j, fn in enumerate(f_fns):
                fn()
                fe[it][j + 1].record()
        torch.cuda.synchronize()
        peak, _ = measured_hbm_peak()
        f_bytes = {"src": 8 * c64 * (CLIP_SAMPLES + n_out), "eq": 8 * c64 * 2 * n_out,
                   "fft": 8 * c64 * n_frames * (N_FFT + bins)}
        # float64 FMAs the reference's arithmetic needs: 40.006 per resampler output, 9 flop per
        # sample and section of the cascade (5 mul + 4 add: 5 FMA-pipe instructions), 2.5 N log2 N flop per frame
        f_fma = {"src": 40.006 * c64 * n_out, "eq": 6 * 5.0 * c64 * n_out,
                 "fft": 1.25 * N_FFT * 12 * c64 * n_frames}
        f_k = {}
        tot = 0.0
        for j, k in enumerate(f_names):
            msk = sum(fe[it][j].elapsed_time(fe[it][j + 1]) for it in range(f_steps)) / f_steps
            tot += msk
            hbm_ms = f_bytes[k] / (peak * 1e9) * 1e3
            pipe_ms = f_fma[k] / (FP64_TFMA_PEAK * 1e12) * 1e3
            f_k[k] = {"ms": msk, "hbm_floor_ms": hbm_ms, "fp64_pipe_floor_ms": pipe_ms,
                      "bound": "hbm" if hbm_ms >= pipe_ms else "fp64_pipe", "frac_of_bound": max(hbm_ms, pipe_ms) / msk,
                      "frac_of_hbm": hbm_ms / msk}
        f64_line = {"dtype": "f64", "clips": c64, "ms_per_step": tot,
                    "value": c64 * CLIP_SAMPLES / (tot * 1e-3) / 1e6, "unit": "Msamples/s", "kernels": f_k,
                    "fp64_pipe_peak_tfma": FP64_TFMA_PEAK,
                    "note": "three float64 kernels (tiled SRC, scan EQ, FFT); floors: algorithmic bytes / measured HBM peak "
                            "and reference FMAs / measured DFMA rate; parity <= 1e-10 relative (tests)"}
        del x64, z64, m64

    # ---- e2e: host-buffer C-ABI call, copies inside the timed region -------
    e2e = None
    if not args.no_e2e and args.workload in ("chain", "c5job"):
        ec = max(1, min(args.e2e_clips, clips))
        xh = torch.empty((ec, CLIP_SAMPLES), dtype=t_dt, pin_memory=True)
        xh.copy_(x[:ec])
        zh = torch.empty((ec, n_out), dtype=t_dt, pin_memory=True)
        mh = torch.empty((ec, n_frames, bins), dtype=t_dt, pin_memory=True)
        xa, za, ma = xh.numpy(), zh.numpy(), mh.numpy()
        chain.run_host(xa, za, ma)                   # warm-up (allocations, clocks)
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 3))
        for _ in range(e2e_steps):
            chain.run_host(xa, za, ma)
        dt_host = time.perf_counter() - t0
        # the same bytes with no kernels at all: what the host<->device fabric alone allows (both directions at once)
        dxc, dzc, dmc = x[:ec], z[:ec], mag[:ec]
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            with torch.cuda.stream(s_in):
                dxc.copy_(xh, non_blocking=True)
            with torch.cuda.stream(s_out):
                zh.copy_(dzc, non_blocking=True)
                mh.copy_(dmc, non_blocking=True)
            torch.cuda.synchronize()
        dt_copy = time.perf_counter() - t0
        th = torch.tensor([dt_host, dt_copy], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(th, op=dist.ReduceOp.MAX)
        dt_host, dt_copy = float(th[0].item()), float(th[1].item())
        e2e = {"value": world * ec * CLIP_SAMPLES * e2e_steps / dt_host / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": ec * CLIP_SAMPLES * esize,
               "d2h_bytes_per_step": (ec * n_out + ec * n_frames * bins) * esize,
               "steps": e2e_steps, "ms_per_step": dt_host / e2e_steps * 1e3, "clips_per_step": ec,
               "api": "dspb200_chain_host_f32 (pinned host buffers, 3-stream slab pipeline)",
               "copy_only_ms_per_step": dt_copy / e2e_steps * 1e3, "frac_of_copy_ceiling": dt_copy / dt_host}
        if rank == 0 and not args.no_parity and args.dtype == "f32":
            # the host form's own output against the oracle (narrow slabs run the FFMA kernels, not the tensor-core ones)
            ez, em = oracle_spot_check([xa[0]], [za[0]], [ma[0]])
            e2e["parity_err"] = {"z_full_scale": ez, "mag_rel": em, "clips": [0]}
        # the export form of the same call: z leaves as the int16 signal app.py:349-354 writes to the WAV and the
        # spectra as the dB values app.py:207-210 plots -> 2/3 of the device-to-host bytes
        chain.run_host(xa, za, ma)                   # the copy-only loop above overwrote the host buffers
        ma_lin0 = ma[0].copy()
        chain_db = pkg.Chain(L_UP, M_DOWN, FS_IN, GAINS, n_fft=N_FFT, dtype=np_dt, db=True)
        qh = torch.empty((ec, n_out), dtype=torch.int16, pin_memory=True)
        ph = torch.empty((ec,), dtype=t_dt, pin_memory=True)
        qa, pa = qh.numpy(), ph.numpy()
        chain_db.run_host_pcm16(xa, qa, ma, pa)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            chain_db.run_host_pcm16(xa, qa, ma, pa)
        dt_exp = time.perf_counter() - t0
        export_parity = None
        if rank == 0 and not args.no_parity and args.dtype == "f32":
            from oracle import dsp_oracle as o
            ref_q = o.pcm16_export(za[0]).astype(np.int32)                      # app.py:349-354 on the float32 z
            ref_db = o.spectrum_db(ma_lin0.astype(np.float64))
            big = ma_lin0 > 1e-4 * ma_lin0.max()
            export_parity = {"pcm16_lsb": int(np.max(np.abs(qa[0].astype(np.int32) - ref_q))),
                             "db_abs": float(np.max(np.abs(ma[0][big] - ref_db[big]))), "clips": [0],
                             "against": "oracle pcm16_export / spectrum_db of the float32 z and |X| the plain call returned"}
        dq = torch.empty((ec, n_out), dtype=torch.int16, device=dev)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            with torch.cuda.stream(s_in):
                dxc.copy_(xh, non_blocking=True)
            with torch.cuda.stream(s_out):
                qh.copy_(dq, non_blocking=True)
                mh.copy_(dmc, non_blocking=True)
            torch.cuda.synchronize()
        dt_copy2 = time.perf_counter() - t0
        th = torch.tensor([dt_exp, dt_copy2], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(th, op=dist.ReduceOp.MAX)
        dt_exp, dt_copy2 = float(th[0].item()), float(th[1].item())
        e2e["export"] = {"value": world * ec * CLIP_SAMPLES * e2e_steps / dt_exp / 1e6, "unit": "Msamples/s",
                         "h2d_bytes_per_step": ec * CLIP_SAMPLES * esize,
                         "d2h_bytes_per_step": ec * n_out * 2 + ec * esize + ec * n_frames * bins * esize,
                         "ms_per_step": dt_exp / e2e_steps * 1e3,
                         "api": "dspb200_chain_host_pcm16_f32 (z as int16 per app.py:349-354, dB spectra per app.py:207-210)",
                         "copy_only_ms_per_step": dt_copy2 / e2e_steps * 1e3, "frac_of_copy_ceiling": dt_copy2 / dt_exp}
        if export_parity is not None:
            e2e["export"]["parity_err"] = export_parity
        del dq

    # ---- the optional exchange: all-gather of per-clip spectra over NCCL on a side stream, alone and
    # overlapped with the next waves' kernels.  Not part of `value`: the data path has no collective.
    gather_line = None
    if world > 1 and not args.no_gather and args.workload == "chain":
        gc = max(1, min(args.gather_clips, clips))
        snap = mag[:gc].clone()
        g = shard.SpectraGather(world * gc)
        for _ in range(2):
            g.start(snap).wait()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g_steps = max(1, min(args.steps, 5))
        g0.record()
        for _ in range(g_steps):
            g.start(snap).wait()
        g1.record()
        barrier()
        alone_ms = g0.elapsed_time(g1) / g_steps
        barrier()
        g0.record()
        for _ in range(g_steps):
            step_chain()
            g.wait()
            g.start(snap)
        g.wait()
        g1.record()
        barrier()
        with_ms = g0.elapsed_time(g1) / g_steps
        tg = torch.tensor([alone_ms, with_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        nbytes = snap.numel() * snap.element_size()
        gather_line = {"collective": "all_gather_into_tensor (NCCL) on a side stream, shard.SpectraGather",
                       "clips_per_rank": gc, "bytes_per_rank": nbytes, "alone_ms": float(tg[0].item()),
                       "busbw_gbs": nbytes * (world - 1) / (float(tg[0].item()) * 1e-3) / 1e9,
                       "chain_step_ms_with_gather_overlapped": float(tg[1].item()),
                       "chain_step_ms": ms_max / args.steps}
        del snap, g

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        alg_bytes = {
            "src": esize * clips * (CLIP_SAMPLES + n_out),
            "eq": esize * clips * 2 * n_out,
            "src_eq": esize * clips * (CLIP_SAMPLES + n_out),
            "fft": esize * clips * n_frames * (N_FFT + bins),
        }
        src_kind = chain.src.kernel_kind(clips, CLIP_SAMPLES)
        eq_kind = chain.eq.kernel_kind(clips, n_out)
        kernel_names = {"src": "src_mma_kernel" if src_kind == "tensor" else "src_tiled_kernel",
                        "eq": "lti_mma_kernel" if eq_kind == "tensor" else "eq_packed_kernel",
                        "src_eq": "xz_mma_kernel",
                        # float32 4096-point frames: 32 points per thread (csrc/fft_r32.cu) unless DSPB200_FFT_VAR < 64
                        "fft": "fft4096_r32_kernel" if (args.dtype == "f32" and N_FFT == 4096 and
                                                        int(os.environ.get("DSPB200_FFT_VAR", "64")) >= 64)
                        else "fft_fixed_kernel"}
        kernels = {}
        for k in names:
            gbs = alg_bytes[k] / (per_kernel[k] * 1e-3) / 1e9
            kernels[k] = {"kernel": kernel_names[k], "ms": per_kernel[k], "algorithmic_bytes": alg_bytes[k],
                          "achieved_gbs": gbs, "frac": gbs / peak, "frac_of_nominal_8tbs": gbs / 8000.0,
                          "traffic": traffic_from_profiles(kernel_names[k], clips)}
        dom = max(names, key=lambda k: per_kernel[k]) if args.workload in ("chain", "c5job") else args.workload
        line = {
            "metric": "Msamples/s SRC->EQ->FFT chain" if args.workload in ("chain", "c5job") else f"Msamples/s {args.workload}",
            "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic", "config": workload_config(args),
            "roofline": {"bound": "hbm", "kernel": kernels[dom]["kernel"], "achieved": kernels[dom]["achieved_gbs"],
                         "peak": peak, "unit": "GB/s", "frac": kernels[dom]["frac"],
                         "frac_of_nominal_8tbs": kernels[dom]["frac_of_nominal_8tbs"],   # both denominators
                         "traffic": kernels[dom]["traffic"], "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes"],
                         "ms_per_launch": kernels[dom]["ms"]},
            "kernels": kernels, "chain_kind": kind,
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": e2e,
        }
        if args.workload in ("chain", "c5job"):
            # the whole chain against SURVEY.md 8d's C5 bytes: x read once, z and the spectra written once
            chain_bytes = esize * (CLIP_SAMPLES + n_out + n_frames * bins)          # 4.643 MB per clip in float32
            step_clips = clips if args.workload == "chain" else my_total
            ms_chain = sum(per_kernel.values()) if args.workload == "chain" else ms_max / args.steps
            gbs = chain_bytes * step_clips / (ms_chain * 1e-3) / 1e9
            line["roofline"]["chain"] = {"algorithmic_bytes_per_clip": chain_bytes, "clips": step_clips, "ms": ms_chain,
                                         "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                         "frac_of_nominal_8tbs": gbs / 8000.0,
                                         "note": "z is written by the SRC->EQ kernel and read back by the FFT kernel: "
                                                 "1.41x the algorithmic bytes move; with the generator's writes in c5job"}
        if args.workload == "c5job":
            line["waves"] = [{"first_clip": w[0], "clips": w[1]} for w in waves]
        if parity is not None:
            line["parity_err"] = parity
        if f64_line is not None:
            line["f64"] = f64_line
        if cpu_line is not None:
            line["cpu_baseline"] = cpu_line
        if gather_line is not None:
            line["gather"] = gather_line
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def emit(line):
    """The one JSON line, on the real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    # libraries print to stdout too (NCCL's version banner at communicator creation): keep the real stdout
    # for the JSON line and send everything else written to fd 1 to stderr
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
