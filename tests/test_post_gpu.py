"""GPU parity for the kernels either side of the path: loader
front end (bit-exact against the real reference's output), int16 export
(bit-exact in float64, +-1 LSB in float32) and the fused dB spectrum."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import dsp_oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    import dsp_audio_project_b200 as pk
    if not torch.cuda.is_available():
        pytest.fail("a CUDA device is required")
    return torch, pk


def test_mono_normalize_bit_exact_vs_reference(env):
    torch, pk = env
    g = load_golden("loader.npz")
    for name in ("stereo", "mono", "quad", "tiny", "silence"):
        x = g[f"in_{name}"]
        for dt in (np.float64,):
            xt = torch.as_tensor(np.ascontiguousarray(x, dtype=dt)[None, ...], device="cuda")
            mono, peaks = pk.mono_normalize(xt)
            assert mono.dtype == torch.float32
            assert np.array_equal(mono[0].cpu().numpy(), g[f"out_{name}"]), name
    # batched: several clips at once, float32 input path against the oracle
    rng = np.random.default_rng(1)
    x = rng.uniform(-0.6, 0.6, (7, 4001, 2)).astype(np.float32)
    mono, peaks = pk.mono_normalize(torch.as_tensor(x, device="cuda"))
    for c in range(7):
        assert np.array_equal(mono[c].cpu().numpy(), o.load_mono_normalize(x[c]))


def test_pcm16_export(env):
    torch, pk = env
    rng = np.random.default_rng(2)
    z = rng.uniform(-0.8, 0.8, (5, 30011))
    z[1, 17] = np.nan
    z[2] = 0.0
    z[3, 5] = np.inf
    ref = np.stack([o.pcm16_export(z[c]) for c in range(5)])
    out, peaks = pk.to_pcm16(torch.as_tensor(z, device="cuda"))
    assert out.dtype == torch.int16 and np.array_equal(out.cpu().numpy(), ref)        # float64: bit exact
    z32 = z.astype(np.float32)
    out32, _ = pk.to_pcm16(torch.as_tensor(z32, device="cuda"))
    ref32 = np.stack([o.pcm16_export(z32[c].astype(np.float64)) for c in (0, 1, 2, 4)])
    got32 = out32.cpu().numpy()[[0, 1, 2, 4]].astype(np.int32)
    assert np.max(np.abs(got32 - ref32)) <= 1                                          # float32: +-1 LSB
    # a misaligned view
    zo = torch.zeros((5, 30014), dtype=torch.float64, device="cuda")
    zo[:, 1:30012] = torch.as_tensor(z, device="cuda")
    out2, _ = pk.to_pcm16(zo[:, 1:30012])
    assert np.array_equal(out2.cpu().numpy(), ref)


@pytest.mark.parametrize("n_fft", [8, 2048, 4096, 65536])
def test_db_spectrum(env, n_fft):
    torch, pk = env
    rng = np.random.default_rng(n_fft)
    x = rng.uniform(-1, 1, (2, 2 * n_fft))
    w = o.hann_symmetric(n_fft)
    mag = np.abs(np.fft.rfft(x.reshape(2, 2, n_fft) * w, axis=-1))
    ref = o.spectrum_db(mag)
    for dt, tol in ((np.float64, 1e-8), (np.float32, 2e-3)):
        plan = pk.FftPlan(n_fft, dt, hann=True, db=True)
        db = plan.magnitudes(torch.as_tensor(x.astype(dt), device="cuda")).cpu().numpy()
        big = mag > 1e-4 * mag.max()          # dB of near-zero bins amplifies rounding without bound
        assert np.max(np.abs(db[big] - ref[big])) <= tol, (dt, np.max(np.abs(db[big] - ref[big])))
