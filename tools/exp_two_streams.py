"""Experiment: does running the chain on two half-waves over two streams overlap
kernels with different bottlenecks (EQ: FMA/issue, FFT: shared memory)?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import dsp_audio_project_b200 as pkg
GAINS = {"Sub-Bass": 6, "Bass": -3, "Low Mids": 4, "High Mids": -6, "Presence": 3, "Brilliance": -9}
dev = torch.device("cuda", 0)
clips = 1024
chain = pkg.Chain(160, 147, 44100, GAINS, n_fft=4096, dtype=np.float32)
x = torch.rand((clips, 441000), device=dev) - 0.5
y = torch.empty((clips, 480000), device=dev); z = torch.empty_like(y)
mag = torch.empty((clips, 117, 2049), device=dev)

def run_serial():
    chain.src.run(x, out=y); chain.eq.run(y, out=z); chain.fft.magnitudes(z, out=mag)

def run_split(parts, streams):
    main = torch.cuda.current_stream()
    ev = torch.cuda.Event(); ev.record(main)
    per = clips // parts
    for i in range(parts):
        st = streams[i % len(streams)]
        st.wait_event(ev)
        with torch.cuda.stream(st):
            sl = slice(i * per, (i + 1) * per)
            chain.src.run(x[sl], out=y[sl]); chain.eq.run(y[sl], out=z[sl]); chain.fft.magnitudes(z[sl], out=mag[sl])
    for st in streams:
        e = torch.cuda.Event(); e.record(st); main.wait_event(e)

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

print("serial          ", round(timeit(run_serial), 3), "ms")
for parts, ns in ((2, 2), (4, 2), (4, 4), (8, 2), (8, 4)):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    print(f"parts={parts} streams={ns}", round(timeit(lambda: run_split(parts, streams)), 3), "ms")
