"""Development timing of the fused SRC->EQ kernel against the two separate tensor-core kernels on a wide wave.
python tools/xz_perf.py [clips] [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import dsp_audio_project_b200 as pkg  # noqa: E402

GAINS = {"Sub-Bass": 6, "Bass": -3, "Low Mids": 4, "High Mids": -6, "Presence": 3, "Brilliance": -9}


def timed(fn, reps):
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


def main():
    clips = int(sys.argv[1]) if len(sys.argv) > 1 else 18944
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    n_in = int(sys.argv[3]) if len(sys.argv) > 3 else 441000
    chain = pkg.Chain(160, 147, 44100, GAINS, n_fft=4096, dtype=np.float32)
    n_out = chain.out_len(n_in)
    pitch = int(os.environ.get("XZ_PITCH", n_in))      # row pitch of x in samples (>= n_in)
    x = (torch.rand((clips, pitch), device="cuda") - 0.5)[:, :n_in]
    z = torch.empty((clips, n_out), device="cuda")
    print("chain kind:", chain.kernel_kind(clips, n_in), flush=True)
    t_f = timed(lambda: chain.run_fused(x, out=z), reps)
    zf = z[:4].clone()
    def two():
        chain.src.run(x, out=z)
        chain.eq.run(z, out=z)
    t_2 = timed(two, reps) if os.environ.get("XZ_ONLY") is None else float("nan")
    d = float((zf - z[:4]).abs().max())
    gb = 4 * clips * (n_in + n_out) / 1e9
    print(f"[{clips} x {n_in}] fused {t_f:.3f} ms = {gb / t_f:.1f} GB/s of x+z ({gb / t_f / 6.5386 * 100:.1f} % of 6538.6); "
          f"src + eq separately {t_2:.3f} ms; max diff {d:.2e}")


if __name__ == "__main__":
    main()
