"""Timing of the EQ on batches too narrow to give every SM a group of 128 channels: tensor-core form on overlapping
time slices (default where it pays) against the FFMA scan kernel.  python tools/eq_narrow_perf.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import dsp_audio_project_b200 as pkg  # noqa: E402

GAINS = {"Sub-Bass": 6, "Bass": -3, "Low Mids": 4, "High Mids": -6, "Presence": 3, "Brilliance": -9}
PEAK = 6538.6


def timed(fn, reps=5):
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
    plan = pkg.EqPlan.from_gains(48000, GAINS, np.float32)
    print(json.dumps({"warm_chunks": plan.warm_chunks()}))
    for ch, n in ((1024, 480000), (2048, 480000), (4096, 480000), (8192, 480000), (4096, 2880000), (1024, 2880000)):
        x = torch.rand((ch, n), device="cuda") * 0.5 - 0.25
        z = torch.empty_like(x)
        gb = 8 * ch * n / 1e9
        row = {"case": f"EQ {ch} x {n} f32", "kind": plan.kernel_kind(ch, n), "algorithmic_GB": round(gb, 3)}
        t = timed(lambda: plan.run(x, out=z))
        row["default_ms"] = round(t, 4)
        row["default_frac"] = round(gb / t / PEAK * 1e3, 3)
        os.environ["DSPB200_EQ_NO_MMA"] = "1"
        t = timed(lambda: plan.run(x, out=z))
        del os.environ["DSPB200_EQ_NO_MMA"]
        row["scan_ms"] = round(t, 4)
        row["scan_frac"] = round(gb / t / PEAK * 1e3, 3)
        os.environ["DSPB200_EQ_FORCE_MMA"] = "1"
        os.environ["DSPB200_EQ_NO_OVERLAP"] = "1"
        t = timed(lambda: plan.run(x, out=z), 2)
        del os.environ["DSPB200_EQ_FORCE_MMA"], os.environ["DSPB200_EQ_NO_OVERLAP"]
        row["tensor_one_slice_ms"] = round(t, 4)
        print(json.dumps(row), flush=True)
        del x, z


if __name__ == "__main__":
    main()
