"""Development probe: sustained vs burst timing of the tensor-core EQ on wide and on sliced narrow batches."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import dsp_audio_project_b200 as pkg
GAINS = {"Sub-Bass": 6, "Bass": -3, "Low Mids": 4, "High Mids": -6, "Presence": 3, "Brilliance": -9}
plan = pkg.EqPlan.from_gains(48000, GAINS, np.float32)
def timed(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for ch, n in ((18944, 480000), (4096, 2880000), (8192, 1440000), (18944, 620000)):
    x = torch.rand((ch, n), device="cuda") * 0.5 - 0.25
    z = torch.empty_like(x)
    gb = 8 * ch * n / 1e9
    plan.run(x, out=z); torch.cuda.synchronize()
    for sl in (0, 4):
        if sl: os.environ["DSPB200_EQ_SLICES"] = str(sl)
        else: os.environ.pop("DSPB200_EQ_SLICES", None)
        if sl and ch > 4096: continue
        time.sleep(1.0)
        t1 = timed(lambda: plan.run(x, out=z), 1)
        t10 = timed(lambda: plan.run(x, out=z), 10)
        print(ch, n, "slices", sl, "burst", round(t1, 3), "ms", round(gb / t1 / 6.5386, 3), "| sustained x10", round(t10, 3), "ms", round(gb / t10 / 6.5386, 3), flush=True)
    os.environ.pop("DSPB200_EQ_SLICES", None)
    del x, z
