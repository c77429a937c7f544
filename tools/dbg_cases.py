"""Run each kernel family in its own process and report pass/fail (debug aid)."""
import subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {
 "src_generic_f64": "p=pk.SrcPlan(3,2,np.float64); x=torch.rand(2,2000,dtype=torch.float64,device='cuda'); y=p.run(x,force_generic=True); torch.cuda.synchronize(); ref=o.resample_closed_form(x[0].cpu().numpy(),44100,2,3)[0]; print('err',o.rel_err(y[0].cpu().numpy(),ref))",
 "src_tiled_notma_f64": "p=pk.SrcPlan(3,2,np.float64); xo=torch.rand(2,2003,dtype=torch.float64,device='cuda'); x=xo[:,1:2001]; y=p.run(x); torch.cuda.synchronize(); ref=o.resample_closed_form(x[0].cpu().numpy(),44100,2,3)[0]; print('err',o.rel_err(y[0].cpu().numpy(),ref))",
 "src_tiled_tma_f64": "p=pk.SrcPlan(3,2,np.float64); x=torch.rand(2,2000,dtype=torch.float64,device='cuda'); y=p.run(x); torch.cuda.synchronize(); ref=o.resample_closed_form(x[0].cpu().numpy(),44100,2,3)[0]; print('err',o.rel_err(y[0].cpu().numpy(),ref))",
 "src_tiled_notma_f32": "p=pk.SrcPlan(160,147,np.float32); xo=torch.rand(5,4413,dtype=torch.float32,device='cuda'); x=xo[:,1:4411]; y=p.run(x); torch.cuda.synchronize(); ref=o.resample_closed_form(x[0].cpu().numpy().astype(np.float64),44100,147,160)[0]; print('err',o.full_scale_err(y[0].cpu().numpy(),ref))",
 "src_tiled_tma_f32": "p=pk.SrcPlan(160,147,np.float32); x=torch.rand(5,4412,dtype=torch.float32,device='cuda'); y=p.run(x); torch.cuda.synchronize(); ref=o.resample_closed_form(x[0].cpu().numpy().astype(np.float64),44100,147,160)[0]; print('err',o.full_scale_err(y[0].cpu().numpy(),ref))",
 "eq_f32": "g={'Sub-Bass':6,'Bass':-3,'Low Mids':4,'High Mids':-6,'Presence':3,'Brilliance':-9}; p=pk.EqPlan.from_gains(48000,g,np.float32); x=(torch.rand(3,5000,device='cuda')-0.5)*0.5; z=p.run(x); torch.cuda.synchronize(); ref=o.equalizer(x[1].cpu().numpy().astype(np.float64),48000,g); print('err',o.full_scale_err(z[1].cpu().numpy(),ref))",
 "eq_f64": "g={'Sub-Bass':6,'Bass':-3,'Low Mids':4,'High Mids':-6,'Presence':3,'Brilliance':-9}; p=pk.EqPlan.from_gains(48000,g,np.float64); x=(torch.rand(3,5000,device='cuda',dtype=torch.float64)-0.5)*0.5; z=p.run(x); torch.cuda.synchronize(); ref=o.equalizer(x[1].cpu().numpy(),48000,g); print('err',o.rel_err(z[1].cpu().numpy(),ref))",
 "eq_f64_real": "g={'Sub-Bass':-15,'Bass':-15}; p=pk.EqPlan.from_gains(48000,g,np.float64); x=(torch.rand(3,5000,device='cuda',dtype=torch.float64)-0.5)*0.5; z=p.run(x); torch.cuda.synchronize(); ref=o.equalizer(x[1].cpu().numpy(),48000,g); print('err',o.rel_err(z[1].cpu().numpy(),ref))",
 "fft_4096_f32": "p=pk.FftPlan(4096,np.float32); x=torch.rand(2,9000,device='cuda')*2-1; m=p.magnitudes(x); torch.cuda.synchronize(); ref=o.frame_magnitudes(x.cpu().numpy().astype(np.float64),4096); print('err',o.rel_err(m.cpu().numpy(),ref))",
 "fft_4096_f64": "p=pk.FftPlan(4096,np.float64); x=torch.rand(2,9000,device='cuda',dtype=torch.float64)*2-1; m=p.magnitudes(x); torch.cuda.synchronize(); ref=o.frame_magnitudes(x.cpu().numpy(),4096); print('err',o.rel_err(m.cpu().numpy(),ref))",
 "fft_64_f64": "p=pk.FftPlan(64,np.float64); x=torch.rand(2,200,device='cuda',dtype=torch.float64)*2-1; m=p.magnitudes(x); torch.cuda.synchronize(); ref=o.frame_magnitudes(x.cpu().numpy(),64); print('err',o.rel_err(m.cpu().numpy(),ref))",
 "fft_8_f64": "p=pk.FftPlan(8,np.float64); x=torch.rand(2,20,device='cuda',dtype=torch.float64)*2-1; m=p.magnitudes(x); torch.cuda.synchronize(); ref=o.frame_magnitudes(x.cpu().numpy(),8); print('err',o.rel_err(m.cpu().numpy(),ref))",
 "fft_65536_f32": "p=pk.FftPlan(65536,np.float32); x=torch.rand(2,65536,device='cuda')*2-1; m=p.magnitudes(x); torch.cuda.synchronize(); ref=np.abs(np.fft.rfft(x.cpu().numpy().astype(np.float64)*o.hann_symmetric(65536)))[:,None,:]; print('err',o.rel_err(m.cpu().numpy(),ref))",
 "fft_65536_f64": "p=pk.FftPlan(65536,np.float64); x=torch.rand(2,65536,device='cuda',dtype=torch.float64)*2-1; m=p.magnitudes(x); torch.cuda.synchronize(); ref=np.abs(np.fft.rfft(x.cpu().numpy()*o.hann_symmetric(65536)))[:,None,:]; print('err',o.rel_err(m.cpu().numpy(),ref))",
 "c2c_4096_f64": "p=pk.FftPlan(4096,np.float64,hann=False); x=torch.randn(2,4096,device='cuda',dtype=torch.complex128); X=p.c2c(x); torch.cuda.synchronize(); print('err',o.rel_err(X.cpu().numpy(),np.fft.fft(x.cpu().numpy())))",
 "c2c_65536_f64": "p=pk.FftPlan(65536,np.float64,hann=False); x=torch.randn(1,65536,device='cuda',dtype=torch.complex128); X=p.c2c(x); torch.cuda.synchronize(); print('err',o.rel_err(X.cpu().numpy(),np.fft.fft(x.cpu().numpy())))",
}
PRE = "import sys; sys.path.insert(0,%r); import numpy as np, torch; import dsp_audio_project_b200 as pk; from oracle import dsp_oracle as o; " % ROOT
sel = sys.argv[1:] or list(CASES)
for name in sel:
    r = subprocess.run([sys.executable, "-c", PRE + CASES[name]], capture_output=True, text=True, timeout=600)
    tail = (r.stdout.strip().splitlines() or [""])[-1] if r.returncode == 0 else (r.stderr.strip().splitlines() or ["?"])[-1]
    print(f"{name:24s} rc={r.returncode} {tail[:200]}", flush=True)
if len(sel) == 1:
    print(PRE + CASES[sel[0]])
