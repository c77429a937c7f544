"""Quick GPU check + timing of the tensor-core EQ against the scan kernel (development aid)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import dsp_audio_project_b200 as pk
from oracle import dsp_oracle as o

BANDS = ["Sub-Bass", "Bass", "Low Mids", "High Mids", "Presence", "Brilliance"]
gd = dict(zip(BANDS, (6, -3, 4, -6, 3, -9)))
plan = pk.EqPlan.from_gains(48000, gd, np.float32)

def run(xt, mma):
    os.environ.pop("DSPB200_EQ_NO_MMA", None); os.environ.pop("DSPB200_EQ_FORCE_MMA", None)
    os.environ["DSPB200_EQ_FORCE_MMA" if mma else "DSPB200_EQ_NO_MMA"] = "1"
    z = plan.run(xt); torch.cuda.synchronize(); return z

def timeit(xt, mma, reps=5):
    run(xt, mma); run(xt, mma)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    out = torch.empty_like(xt)
    ev[0].record()
    for _ in range(reps): plan.run(xt, out=out)
    ev[1].record(); torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps

for ch, n in ((256, 112 * 3), (300, 4412), (512, 20000)):
    x = np.random.default_rng(1).uniform(-0.5, 0.5, (ch, n)).astype(np.float32)
    xt = torch.as_tensor(x, device="cuda")
    z = run(xt, True); zs = run(xt, False)
    ref = np.stack([o.equalizer(x[c].astype(np.float64), 48000, gd) for c in (0, ch - 1)])
    print(f"{ch}x{n}: mma vs f64 {np.abs(z.cpu().numpy()[[0, ch - 1]] - ref).max():.2e}  scan vs f64 "
          f"{np.abs(zs.cpu().numpy()[[0, ch - 1]] - ref).max():.2e}  mma vs scan {float((z - zs).abs().max()):.2e}", flush=True)

for ch, n in ((1024, 480000), (8192, 480000), (16384, 240000), (4096, 2880000)):
    xt = torch.rand((ch, n), device="cuda", dtype=torch.float32) - 0.5
    tm = timeit(xt, True); ts = timeit(xt, False)
    gb = 8.0 * ch * n / 1e9
    print(f"{ch}x{n}: mma {tm:.3f} ms ({gb / tm:.0f} GB/s)   scan {ts:.3f} ms ({gb / ts:.0f} GB/s)", flush=True)
    z = run(xt, True); zs = run(xt, False)
    print("   max diff", float((z - zs).abs().max()), flush=True)
    del xt, z, zs
