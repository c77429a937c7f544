"""Debug aid: the 2^16-point three-pass kernel (DSPB200_FFT_LONG32=1) against numpy, error per output row k mod 32."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DSPB200_FFT_LONG32"] = "1"
import dsp_audio_project_b200 as pkg          # noqa: E402

N = 65536
rng = np.random.default_rng(0)
hann = int(sys.argv[1]) if len(sys.argv) > 1 else 1
x = rng.uniform(-1, 1, (2, 2 * N)).astype(np.float32)
w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / (N - 1)) if hann else np.ones(N)
ref = np.abs(np.fft.rfft(x.astype(np.float64).reshape(2, 2, N) * w, axis=-1))
plan = pkg.FftPlan(N, np.float32, hann=bool(hann))
m = plan.magnitudes(torch.as_tensor(x, device="cuda")).cpu().numpy().astype(np.float64)
err = np.abs(m - ref) / ref.max()
print("max err", err.max(), "at", np.unravel_index(err.argmax(), err.shape))
e = err[0, 0]
for k1 in range(32):
    rowerr = e[k1:32768:32]
    print(k1, float(rowerr.max()), int(rowerr.argmax()), end=" | ")
    if k1 % 4 == 3:
        print()
print("k=32768:", e[32768], " first bins", m[0, 0, :4], ref[0, 0, :4])
