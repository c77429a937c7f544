"""C2 alone: 1024 channels x 441000 samples through SrcPlan.run (float32, 160/147), ten timed launches.
Used under ncu for per-SM balance (sm__cycles_active.avg / .max / .min)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dsp_audio_project_b200 as pkg
dev = torch.device("cuda", 0)
x = torch.rand((1024, 441000), device=dev) - 0.5
src = pkg.SrcPlan(160, 147, np.float32)
y = src.run(x)
for _ in range(3): src.run(x, out=y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): src.run(x, out=y)
e1.record(); torch.cuda.synchronize()
print("C2 SRC ms", e0.elapsed_time(e1) / 10, src.kernel_kind(1024, 441000))
