"""Per-phase cycle counts of the tensor-core EQ epilogue (DSPB200_LTI_PROF=1)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, ".")
import dsp_audio_project_b200 as pk

BANDS = ["Sub-Bass", "Bass", "Low Mids", "High Mids", "Presence", "Brilliance"]
plan = pk.EqPlan.from_gains(48000, dict(zip(BANDS, (6, -3, 4, -6, 3, -9))), np.float32)
os.environ["DSPB200_EQ_FORCE_MMA"] = "1"
shapes = [tuple(int(v) for v in s.split("x")) for s in (sys.argv[1:] or ["18944x120000", "37888x60000", "65536x30000"])]
for ch, n in shapes:
    xt = torch.rand((ch, n), device="cuda", dtype=torch.float32) - 0.5
    out = torch.empty_like(xt)
    os.environ.pop("DSPB200_LTI_PROF", None)
    for _ in range(2):
        plan.run(xt, out=out)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(5):
        plan.run(xt, out=out)
    ev[1].record(); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 5
    print(f"{ch}x{n}: {ms:.3f} ms  {8e-6 * ch * n / ms:.0f} GB/s", flush=True)
    os.environ["DSPB200_LTI_PROF"] = "1"
    plan.run(xt, out=out); torch.cuda.synchronize()
    del xt, out
