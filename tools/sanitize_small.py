"""Small-shape pass over every kernel family for compute-sanitizer (memcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import dsp_audio_project_b200 as pk
from oracle import dsp_oracle as o

g = {"Sub-Bass": 6, "Bass": -3, "Low Mids": 4, "High Mids": -6, "Presence": 3, "Brilliance": -9}
for dt, tdt in ((np.float32, torch.float32), (np.float64, torch.float64)):
    for (L, M, n, ch) in [(160, 147, 1500, 3), (3, 2, 700, 130), (2, 3, 640, 2), (8, 1, 100, 1), (1, 8, 3000, 2), (5, 7, 9, 2)]:
        p = pk.SrcPlan(L, M, dt)
        x = torch.rand(ch, n, dtype=tdt, device="cuda") - 0.5
        y = p.run(x); y2 = p.run(x, force_generic=True)
        xo = torch.zeros(ch, n + 3, dtype=tdt, device="cuda"); xo[:, 1:n + 1] = x
        y3 = p.run(xo[:, 1:n + 1])
        assert float((y - y2).abs().max()) < 1e-5 and float((y - y3).abs().max()) < 1e-5
    for gains in (g, {k: -15 for k in g}, {f"b{i}": 3 for i in range(11)}):
        for n, ch in ((1, 1), (1000, 3), (1025, 5), (40000, 9)):
            e = pk.EqPlan.from_gains(48000, gains, dt)
            x = (torch.rand(ch, n, dtype=tdt, device="cuda") - 0.5) * 0.5
            z = e.run(x)
            xo = torch.zeros(ch, n + 3, dtype=tdt, device="cuda"); xo[:, 1:n + 1] = x
            z2 = e.run(xo[:, 1:n + 1])
            assert float((z - z2).abs().max()) < 1e-4
    for nf in (2, 8, 64, 1024, 2048, 4096, 16384, 65536):
        f = pk.FftPlan(nf, dt)
        x = torch.rand(2, nf * 2 + 5, dtype=tdt, device="cuda")
        m = f.magnitudes(x); m2 = f.magnitudes(x, hop=nf // 2 if nf > 2 else 1, offset=3, n_frames=2, n_valid=nf + 5)
        if nf >= 16:
            c = torch.randn(2, nf, dtype=torch.complex64 if dt == np.float32 else torch.complex128, device="cuda")
            if nf <= 65536 // (1 if dt == np.float32 else 1):
                f2 = pk.FftPlan(nf, dt, hann=False); f2.c2c(c)
    ch = pk.Chain(3, 2, 44100, g, n_fft=1024, dtype=dt)
    x = torch.rand(3, 9000, dtype=tdt, device="cuda") - 0.5
    ch.run(x, keep_y=True); ch.run_host(x.cpu().numpy())
torch.cuda.synchronize()
print("sanitize pass ok")
