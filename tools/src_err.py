"""fp32 SRC error of the tensor-core form vs the fp64 kernel on the C2 shape (first channels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dsp_audio_project_b200 as pkg
gen = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand((1024, 441000), generator=gen, device="cuda") - 0.5
p32, p64 = pkg.SrcPlan(160, 147, np.float32), pkg.SrcPlan(160, 147, np.float64)
print("kind", p32.kernel_kind(1024, 441000))
y = p32.run(x)
yt = p32.run(x, force_tiled=True)
ref = p64.run(x[:64].double())
print("tensor  max|err| / full scale:", float((y[:64].double() - ref).abs().max()), " rms:", float((y[:64].double() - ref).pow(2).mean().sqrt()))
print("ffma    max|err| / full scale:", float((yt[:64].double() - ref).abs().max()), " rms:", float((yt[:64].double() - ref).pow(2).mean().sqrt()))
print("max|y|", float(ref.abs().max()))
