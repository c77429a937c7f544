import json, sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import dsp_audio_project_b200 as pkg
def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    best=1e9
    for _ in range(reps):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best=min(best,e0.elapsed_time(e1))
    return best
dev=torch.device("cuda",0)
for nf,tdt,ndt,ch in ((4096,torch.float64,np.float64,256),(2048,torch.float32,np.float32,512),(65536,torch.float64,np.float64,256),(16384,torch.float32,np.float32,512)):
    x=torch.rand((ch,1<<20),device=dev,dtype=tdt)*2-1
    f=pkg.FftPlan(nf,ndt,hann=True); m=f.magnitudes(x)
    ms=timeit(lambda: f.magnitudes(x,out=m))
    es=8 if tdt==torch.float64 else 4
    print(nf, ndt.__name__, round(ms,4), "ms", round(es*(x.numel()+m.numel())/ms/1e6/6538.6,3))
