"""Development check of the fused SRC->EQ kernel (csrc/xz_mma.cu) against the float64 oracle and the
three-kernel cascade, on a small batch.  python tools/xz_check.py [channels] [n_in]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import dsp_audio_project_b200 as pkg  # noqa: E402
from oracle import dsp_oracle as o  # noqa: E402

GAINS = {"Sub-Bass": 6, "Bass": -3, "Low Mids": 4, "High Mids": -6, "Presence": 3, "Brilliance": -9}


def main():
    ch = int(sys.argv[1]) if len(sys.argv) > 1 else 130
    n_in = int(sys.argv[2]) if len(sys.argv) > 2 else 22052
    rng = np.random.default_rng(11)
    x = rng.uniform(-0.5, 0.5, (ch, n_in)).astype(np.float32)
    xt = torch.as_tensor(x, device="cuda")
    for name, gains in (("C1 gains", GAINS), ("all +15", {k: 15 for k in GAINS}), ("all -15", {k: -15 for k in GAINS})):
        chain = pkg.Chain(160, 147, 44100, gains, n_fft=4096, dtype=np.float32)
        t0 = time.time()
        z = chain.run_fused(xt)
        torch.cuda.synchronize()
        dt = time.time() - t0
        y = chain.src.run(xt)
        z3 = chain.eq.run(y)
        torch.cuda.synchronize()
        zc = z.cpu().numpy()
        errs = []
        for c in sorted({0, 1, 31, 32, 127, 128, ch - 1} & set(range(ch))):
            yo, fs2 = o.resample_closed_form(x[c].astype(np.float64), 44100, 147, 160)
            zo = o.equalizer(yo, fs2, gains)
            errs.append((c, float(np.max(np.abs(zc[c] - zo)))))
        d3 = float((z - z3).abs().max())
        print(f"{name:9s}: [{ch} x {n_in}] -> {tuple(z.shape)}  fused vs oracle max err per channel {errs}; "
              f"fused vs three kernels {d3:.2e}; nan {int(torch.isnan(z).sum())}; first call {dt*1e3:.1f} ms", flush=True)
        if max(e for _, e in errs) > 1e-4:
            c = max(errs, key=lambda t: t[1])[0]
            yo, fs2 = o.resample_closed_form(x[c].astype(np.float64), 44100, 147, 160)
            zo = o.equalizer(yo, fs2, gains)
            d = np.abs(zc[c] - zo)
            bad = np.nonzero(d > 1e-4)[0]
            print(f"   channel {c}: {bad.size} samples off; first {bad[:12]}, last {bad[-5:]}; z[:6] {zc[c][:6]} ref {zo[:6]}")
            print(f"   err by chunk of 80 (first 12): {[float(d[i*80:(i+1)*80].max()) for i in range(12)]}")


if __name__ == "__main__":
    main()
