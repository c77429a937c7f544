"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle
and the golden vectors generated from the real reference.

Tolerances (north star): float64 kernels <= 1e-10 relative error
(max|a-ref| / max|ref|); float32 fast path <= 1e-5 of full scale for SRC and
FFT, <= 1e-4 of full scale for the EQ cascade.  Full scale is 1.0 for time
signals and max|X| of the reference frame for spectra.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import gains_dict
from oracle import dsp_oracle as o

pytestmark = pytest.mark.gpu

TOL_F64 = 1e-10
TOL_F32_SRC = 1e-5
TOL_F32_FFT = 1e-5
TOL_F32_EQ = 1e-4
C1_GAINS = (6, -3, 4, -6, 3, -9)


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("a CUDA device is required for -m gpu tests (no CPU fallback exists)")
    return torch


@pytest.fixture(scope="module")
def pk():
    import dsp_audio_project_b200 as pkg
    return pkg


@pytest.fixture(scope="module")
def dc():
    from modules import dsp_core
    return dsp_core


# ------------------------------------------------------------------ SRC ----
def test_src_golden_dropin(dc, golden_src):
    g = golden_src
    for idx, (L, M, N, fs_new, fs) in enumerate(g["cases"]):
        x, y = g[f"x_{idx}"], g[f"y_{idx}"]
        out, fs_out = dc.conversion_tasa_muestreo(x, int(fs), int(M), int(L))
        assert fs_out == fs_new
        if L == 1 and M == 1:
            assert out is x
            continue
        assert out.dtype == np.float64 and out.shape == y.shape
        assert o.rel_err(out, y) <= TOL_F64, (idx, L, M, N)


def test_src_golden_f32_and_generic(pk, golden_src, torch_cuda):
    torch = torch_cuda
    g = golden_src
    for idx, (L, M, N, fs_new, fs) in enumerate(g["cases"]):
        if L == 1 and M == 1:
            continue
        x, y = g[f"x_{idx}"], g[f"y_{idx}"]
        p32 = pk.SrcPlan(int(L), int(M), np.float32)
        out = p32.run_host(x.astype(np.float32))[0]
        assert o.full_scale_err(out, y) <= TOL_F32_SRC, (idx, L, M, N)
        # tiled and generic kernels agree with the reference on device tensors too
        for dt, tol in ((torch.float32, TOL_F32_SRC), (torch.float64, 1e-12)):
            plan = p32 if dt == torch.float32 else pk.SrcPlan(int(L), int(M), np.float64)
            xt = torch.as_tensor(np.stack([x, -0.5 * x]).astype(np.float64), device="cuda").to(dt)
            a = plan.run(xt).cpu().numpy()
            b = plan.run(xt, force_generic=True).cpu().numpy()
            assert o.full_scale_err(a[0], y) <= tol and o.full_scale_err(b[0], y) <= tol
            assert o.full_scale_err(a[1], -0.5 * y) <= tol


@pytest.mark.parametrize("L,M,n_in,channels", [
    (160, 147, 4410, 5), (160, 147, 4411, 130), (3, 2, 5003, 129), (2, 3, 4096, 7),
    (8, 8, 1000, 3), (1, 8, 9000, 2), (8, 1, 777, 33), (147, 160, 3000, 2), (7, 5, 2048, 64),
])
def test_src_batched_vs_oracle(pk, torch_cuda, L, M, n_in, channels):
    torch = torch_cuda
    rng = np.random.default_rng(L * 1000 + M)
    x = rng.uniform(-0.5, 0.5, (channels, n_in))
    ref = np.stack([o.resample_closed_form(x[c], 44100, M, L)[0] for c in range(min(channels, 4))])
    for dt, tol in ((np.float64, TOL_F64), (np.float32, TOL_F32_SRC)):
        plan = pk.SrcPlan(L, M, dt)
        xt = torch.as_tensor(x.astype(dt), device="cuda")
        y = plan.run(xt).cpu().numpy()
        assert y.shape == (channels, plan.out_len(n_in))
        err = o.rel_err(y[:ref.shape[0]], ref) if dt == np.float64 else o.full_scale_err(y[:ref.shape[0]], ref)
        assert err <= tol, (dt, plan.kernel_kind(channels, n_in), err)
        # last channel too (exercises the partial channel tile)
        last = o.resample_closed_form(x[-1], 44100, M, L)[0]
        assert o.full_scale_err(y[-1], last) <= max(tol, 1e-12)
        # a strided (non 16-byte aligned) view goes through the non-TMA loader
        xo = torch.zeros((channels, n_in + 3), dtype=xt.dtype, device="cuda")
        xo[:, 1:n_in + 1] = xt
        y2 = plan.run(xo[:, 1:n_in + 1]).cpu().numpy()
        assert np.max(np.abs(y2 - y)) <= 1e-6 if dt == np.float32 else np.max(np.abs(y2 - y)) <= 1e-13


def test_src_known_answers(pk, torch_cuda):
    torch = torch_cuda
    plan = pk.SrcPlan(160, 147, np.float64)
    # DC: unit gain away from the edges (sum(h) * L / L = 1)
    y = plan.run(torch.ones((2, 2000), dtype=torch.float64, device="cuda")).cpu().numpy()
    assert np.max(np.abs(y[:, 100:-100] - 1.0)) < 2e-3
    # impulse: the decimated taps h[m*M + P - i*L]
    x = np.zeros(600); x[300] = 1.0
    y = plan.run(torch.as_tensor(x[None, :], device="cuda")).cpu().numpy()[0]
    h = o.src_filter(160, 147)
    T, P, _, n_out = o.src_geometry(600, 160, 147)
    idx = np.arange(n_out) * 147 + P - 300 * 160
    exp = np.where((idx >= 0) & (idx < T), h[np.clip(idx, 0, T - 1)], 0.0)
    assert np.max(np.abs(y - exp)) <= 1e-15


def test_src_c2_full_size_properties(pk, torch_cuda):
    """Config C2 at full size: 1024 ch x 441000 -> 480000, fp32."""
    torch = torch_cuda
    ch, n_in = 1024, 441000
    plan = pk.SrcPlan(160, 147, np.float32)
    assert plan.kernel_kind(ch, n_in) == "tensor"
    gen = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand((ch, n_in), generator=gen, device="cuda", dtype=torch.float32) - 0.5
    y = plan.run(x)
    assert y.shape == (ch, 480000)
    # oracle on a few whole channels, including the first/last of channel tiles
    for c in (0, 127, 128, 1023):
        ref = o.resample_closed_form(x[c].cpu().numpy().astype(np.float64), 44100, 147, 160)[0]
        assert o.full_scale_err(y[c].cpu().numpy(), ref) <= TOL_F32_SRC
    # linearity: SRC(a*x1 + b*x2) = a*SRC(x1) + b*SRC(x2)
    x2 = torch.flip(x, dims=(0,))
    y2 = plan.run(x2)
    ymix = plan.run(0.25 * x + 0.75 * x2)
    assert float((ymix - (0.25 * y + 0.75 * y2)).abs().max()) <= 5e-6
    # DC in, DC out away from the clip edges
    ydc = plan.run(torch.full((ch, n_in), 0.5, device="cuda"))
    assert float((ydc[:, 64:-64] - 0.5).abs().max()) <= 2e-3


# ------------------------------------------------------------------- EQ ----
def test_eq_golden_dropin(dc, golden_eq):
    g = golden_eq
    for idx, row in enumerate(g["cases"]):
        fs, gs = row[0], row[1:]
        x, z = g[f"x_{idx}"], g[f"z_{idx}"]
        out = dc.sistema_ecualizador(x, fs, gains_dict(gs))
        assert (out is x) == bool(g[f"alias_{idx}"]), idx
        if out is x:
            continue
        assert out.dtype == z.dtype and out.shape == z.shape, idx
        assert o.rel_err(out, z) <= TOL_F64, (idx, row)
    x = g["x_unknown"]
    out = dc.sistema_ecualizador(x, 48000, {"Sub-Bass": 4, "Air": -7, "Brilliance": 5})
    assert o.rel_err(out, g["z_unknown"]) <= TOL_F64
    bands = ["Sub-Bass", "Bass", "Low Mids", "High Mids", "Presence", "Brilliance"]
    rev = gains_dict(C1_GAINS[::-1], bands[::-1])
    assert o.rel_err(dc.sistema_ecualizador(x, 48000, rev), g["z_reversed"]) <= TOL_F64
    b, a = g["ba_lf"][:3], g["ba_lf"][3:]
    assert o.rel_err(dc.aplicar_ecuacion_diferencias(g["x_lf"], b, a), g["y_lf"]) <= TOL_F64


def test_eq_golden_f32(pk, golden_eq):
    g = golden_eq
    for idx, row in enumerate(g["cases"]):
        fs, gs = row[0], row[1:]
        if bool(g[f"alias_{idx}"]):
            continue
        plan = pk.EqPlan.from_gains(fs, gains_dict(gs), np.float32)
        out = plan.run_host(g[f"x_{idx}"].astype(np.float32))[0]
        assert o.full_scale_err(out, g[f"z_{idx}"]) <= TOL_F32_EQ, (idx, row)


@pytest.mark.parametrize("gains", [C1_GAINS, (15,) * 6, (-15,) * 6, (-12.0412, 15, -12.5, 6, 0, -3)])
@pytest.mark.parametrize("n", [1, 31, 1024, 1025, 48000 * 3 + 7])
def test_eq_batched_vs_oracle(pk, torch_cuda, gains, n):
    torch = torch_cuda
    rng = np.random.default_rng(n)
    channels = 19
    x = rng.uniform(-0.25, 0.25, (channels, n))
    gd = gains_dict(gains)
    ref = np.stack([o.equalizer(x[c], 48000, gd) for c in range(channels)])
    for dt in (np.float64, np.float32):
        plan = pk.EqPlan.from_gains(48000, gd, dt)
        xt = torch.as_tensor(x.astype(dt), device="cuda")
        z = plan.run(xt)
        if dt == np.float64:
            assert o.rel_err(z.cpu().numpy(), ref) <= TOL_F64
        else:
            assert o.full_scale_err(z.cpu().numpy(), ref) <= TOL_F32_EQ
        # in place, and through a misaligned view (scalar loader)
        zi = xt.clone()
        plan.run(zi, out=zi)
        assert torch.equal(zi, z)
        xo = torch.zeros((channels, n + 3), dtype=xt.dtype, device="cuda")
        xo[:, 1:n + 1] = xt
        z2 = plan.run(xo[:, 1:n + 1])
        assert float((z2 - z).abs().max()) <= (1e-12 if dt == np.float64 else 1e-5)


def test_eq_long_stream_c3_shape(pk, torch_cuda):
    """C3's time axis (60 s @ 48 kHz = 2.88 M samples) on a handful of channels,
    both gain sets of SURVEY 8d, fp32 and fp64."""
    torch = torch_cuda
    n = 2_880_000
    rng = np.random.default_rng(2)
    x = rng.uniform(-0.25, 0.25, (4, n))
    for gains in (C1_GAINS, (15,) * 6):
        gd = gains_dict(gains)
        ref = np.stack([o.equalizer(x[c], 48000, gd) for c in range(4)])
        z64 = pk.EqPlan.from_gains(48000, gd, np.float64).run(torch.as_tensor(x, device="cuda")).cpu().numpy()
        assert o.rel_err(z64, ref) <= TOL_F64
        z32 = pk.EqPlan.from_gains(48000, gd, np.float32).run(
            torch.as_tensor(x.astype(np.float32), device="cuda")).cpu().numpy()
        assert o.full_scale_err(z32, ref) <= TOL_F32_EQ


def test_eq_many_sections_and_corners(pk, dc, torch_cuda):
    rng = np.random.default_rng(9)
    x = rng.uniform(-0.3, 0.3, 5000)
    # 11 active entries (unknown keys map to 1 kHz): more than one fused pass
    gains = {f"band{i}": (-1) ** i * (1 + i) for i in range(11)}
    ref = o.equalizer(x, 48000, gains)
    assert o.rel_err(dc.sistema_ecualizador(x, 48000, gains), ref) <= TOL_F64
    # |g| == 0.1 everywhere: a clipped copy in the input dtype, no filtering
    g01 = gains_dict((0.1,) * 6)
    big = (3 * x).astype(np.float32)
    out = dc.sistema_ecualizador(big, 48000, g01)
    assert out.dtype == np.float32 and np.array_equal(out, np.clip(big, -1, 1))
    # clamp: at fs = 8 kHz Presence and Brilliance both become 3600 Hz sections
    assert o.rel_err(dc.sistema_ecualizador(x, 8000, gains_dict(C1_GAINS)),
                     o.equalizer(x, 8000, gains_dict(C1_GAINS))) <= TOL_F64
    # state-space description: DC gain of every section is 1 (peaking EQ)
    plan = pk.EqPlan.from_gains(48000, gains_dict(C1_GAINS), np.float64)
    for a00, a01, a10, a11, b0, b1, c0, c1, d in plan.describe():
        A = np.array([[a00, a01], [a10, a11]])
        dc_gain = np.array([c0, c1]) @ np.linalg.solve(np.eye(2) - A, np.array([b0, b1])) + d
        assert abs(dc_gain - 1.0) < 1e-9


# ------------------------------------------------------------------ FFT ----
def test_fft_golden_dropin(dc, golden_fft):
    g = golden_fft
    for n in (1, 2, 4, 8, 16, 64, 256, 1024, 2048, 4096):
        for kind in ("r", "c"):
            x, X = g[f"x{kind}_{n}"], g[f"X{kind}_{n}"]
            out = dc.fft_diezmado_en_tiempo(x)
            if n == 1:
                assert out is x
                continue
            assert out.dtype == np.complex128 and out.shape == X.shape
            assert o.rel_err(out, X) <= TOL_F64, (n, kind)
    for bad in (3, 6, 12, 100):
        with pytest.raises(ValueError):
            dc.fft_diezmado_en_tiempo(np.zeros(bad))


def test_fft_large_c2c_and_real(pk, dc, golden_fft, torch_cuda):
    torch = torch_cuda
    rng = np.random.default_rng(5)
    for n in (8192, 16384, 65536):
        x = rng.normal(size=n) + 1j * rng.normal(size=n)
        assert o.rel_err(dc.fft_diezmado_en_tiempo(x), np.fft.fft(x)) <= TOL_F64, n
    # 2^16-point real magnitude against the golden vector from the reference FFT
    x = golden_fft["xr_65536"].astype(np.float64)
    for dt, tol in ((np.float64, TOL_F64), (np.float32, TOL_F32_FFT)):
        plan = pk.FftPlan(65536, dt, hann=False)
        mag = plan.magnitudes_host(x.astype(dt))[0, 0]
        assert o.rel_err(mag, golden_fft["Xr_65536_mag"]) <= tol


def test_spectrum_golden_dropin(dc, golden_spectrum):
    g = golden_spectrum
    for n in g["ok_lens"]:
        x = g[f"x_{n}"]
        f, m = dc.calcular_espectro_magnitud(x, 48000)
        assert f.shape == g[f"f_{n}"].shape and m.shape == g[f"m_{n}"].shape, n
        assert np.allclose(f, g[f"f_{n}"], rtol=0, atol=1e-9)
        ref = g[f"m_{n}"]
        if np.isnan(ref).any():
            assert np.isnan(m).all()
        else:
            assert np.max(np.abs(m - ref)) <= TOL_F64 * max(1.0, np.max(np.abs(ref))), n
    for n in g["valueerror_lens"]:
        with pytest.raises(ValueError):
            dc.calcular_espectro_magnitud(np.zeros(int(n)), 48000)


@pytest.mark.parametrize("n_fft", [32, 64, 512, 2048, 4096, 16384, 65536])
def test_fftmag_frames_vs_oracle(pk, torch_cuda, n_fft):
    torch = torch_cuda
    rng = np.random.default_rng(n_fft)
    channels, n_frames = 3, 3
    n = n_fft * n_frames + 17
    x = rng.uniform(-1, 1, (channels, n))
    w = o.hann_symmetric(n_fft)
    ref = np.abs(np.fft.rfft(x[:, :n_fft * n_frames].reshape(channels, n_frames, n_fft) * w, axis=-1))
    if n_fft <= 4096:   # the reference's own FFT on one frame pins np.fft as the stand-in
        one = np.abs(o.fft_dit_recursive(x[0, :n_fft] * w))[:n_fft // 2 + 1]
        assert o.rel_err(ref[0, 0], one) <= 1e-13
    for dt, tol in ((np.float64, TOL_F64), (np.float32, TOL_F32_FFT)):
        plan = pk.FftPlan(n_fft, dt, hann=True)
        xt = torch.as_tensor(x.astype(dt), device="cuda")
        mag = plan.magnitudes(xt).cpu().numpy()
        assert mag.shape == ref.shape
        for c in range(channels):
            for f in range(n_frames):
                assert o.rel_err(mag[c, f], ref[c, f]) <= tol, (dt, c, f)
        # hop/offset (odd offset = misaligned loads) and zero padding past n_valid
        m2 = plan.magnitudes(xt, hop=n_fft // 2, offset=3, n_frames=2, n_valid=n_fft + 5).cpu().numpy()
        xz = np.zeros((channels, 4 * n_fft)); xz[:, :n_fft + 5] = x[:, :n_fft + 5]
        for f in range(2):
            s = 3 + f * (n_fft // 2)
            r2 = np.abs(np.fft.rfft(xz[:, s:s + n_fft] * w, axis=-1))
            assert o.rel_err(m2[:, f], r2) <= tol


@pytest.mark.parametrize("n_fft", [2048, 4096])
def test_fftmag_many_frames_grid_stride(pk, torch_cuda, n_fft, monkeypatch):
    """More frames than resident CTAs: every CTA walks several (channel, frame) items, which the fp32
    magnitude kernel tracks by grid-stride carries instead of divisions; 41 x 47 frames make the carry
    fire at irregular steps.  Also pins the kernel variants (DSPB200_FFT_VAR: unset / 64 = 32 points per thread at
    4096 points (fft_r32.cu), 7 = the best 16-points-per-thread form, 0 = table Hann/twiddles, 15 = six CTAs per SM,
    23 = loads after the passes, 39 = register-resident real split), the dB store and the rectangular window."""
    torch = torch_cuda
    rng = np.random.default_rng(7 + n_fft)
    channels, n_frames = 41, 47
    x = rng.uniform(-1, 1, (channels, n_fft * n_frames + 5)).astype(np.float32)
    w = o.hann_symmetric(n_fft)
    ref = np.abs(np.fft.rfft(x[:, :n_fft * n_frames].astype(np.float64).reshape(channels, n_frames, n_fft) * w, axis=-1))
    xt = torch.as_tensor(x, device="cuda")
    plan = pk.FftPlan(n_fft, np.float32, hann=True)
    for var in (None, "64", "7", "0", "15", "23", "39"):
        if var is None:
            monkeypatch.delenv("DSPB200_FFT_VAR", raising=False)
        else:
            monkeypatch.setenv("DSPB200_FFT_VAR", var)
        mag = plan.magnitudes(xt).cpu().numpy()
        assert mag.shape == ref.shape
        err = np.max(np.abs(mag - ref), axis=-1) / np.max(np.abs(ref), axis=-1)
        assert err.max() <= TOL_F32_FFT, (var, float(err.max()), np.unravel_index(err.argmax(), err.shape))
    monkeypatch.delenv("DSPB200_FFT_VAR", raising=False)
    db = pk.FftPlan(n_fft, np.float32, hann=True, db=True).magnitudes(xt).cpu().numpy()
    ref_db = 20.0 * np.log10(ref + 1e-12)
    # a magnitude within TOL_F32_FFT of full scale, read at a bin 40 dB below it, is off by at most
    # 20*log10(1 + 1e-5 / 1e-2) = 8.7e-3 dB; log10f's own rounding is two orders below that
    loud = ref > 1e-2 * ref.max()
    assert np.max(np.abs(db - ref_db)[loud]) <= 20.0 * np.log10(1.0 + TOL_F32_FFT / 1e-2) + 1e-4
    # no window, frames that overlap (hop < n_fft) and start at odd offsets, a tail frame padded with zeros
    plain = pk.FftPlan(n_fft, np.float32, hann=False)
    hop, off, nf = n_fft // 2 + 2, 1, 2 * n_frames - 1
    m3 = plain.magnitudes(xt, hop=hop, offset=off, n_frames=nf).cpu().numpy()
    xp = np.zeros((channels, off + hop * nf + n_fft)); xp[:, :x.shape[1]] = x
    for f in (0, 1, nf // 2, nf - 2, nf - 1):
        r3 = np.abs(np.fft.rfft(xp[:, off + f * hop:off + f * hop + n_fft], axis=-1))
        assert o.rel_err(m3[:, f], r3) <= TOL_F32_FFT, f


def test_fft_parseval_and_linearity_full_c4_frame_count(pk, torch_cuda):
    """A slice of C4 (2^16-point frames of 2^20-sample clips): 64 channels x 16
    frames, fp32 -- Parseval against the time-domain energy of each frame."""
    torch = torch_cuda
    n_fft, channels = 65536, 64
    gen = torch.Generator(device="cuda").manual_seed(3)
    x = torch.rand((channels, 1 << 20), generator=gen, device="cuda") * 2 - 1
    plan = pk.FftPlan(n_fft, np.float32, hann=True)
    mag = plan.magnitudes(x)
    assert mag.shape == (channels, 16, n_fft // 2 + 1)
    w = torch.as_tensor(o.hann_symmetric(n_fft), device="cuda")
    frames = x.view(channels, 16, n_fft).double() * w
    energy_t = (frames ** 2).sum(-1)
    m = mag.double()
    energy_f = (2 * (m ** 2).sum(-1) - m[..., 0] ** 2 - m[..., -1] ** 2) / n_fft
    assert float(((energy_f - energy_t).abs() / energy_t).max()) <= 1e-4
    ref = torch.fft.rfft(frames[:2], dim=-1).abs()
    assert float((mag[:2].double() - ref).abs().max() / ref.max()) <= TOL_F32_FFT


@pytest.mark.parametrize("L,M", [(160, 147), (3, 2), (2, 3), (147, 160), (8, 8), (7, 5), (1, 4), (4, 1)])
def test_src_tensor_core_form_matches_oracle_and_tiled(pk, torch_cuda, L, M, monkeypatch):
    """fp32 SRC runs as a banded-Toeplitz GEMM on tcgen05 (3-term TF32 split) for long signals.
    Against the fp64 closed form, against the FFMA kernel, on ragged channel counts (not a
    multiple of the 256-channel tile), odd lengths and padded row strides."""
    torch = torch_cuda
    rng = np.random.default_rng(1000 * L + M)
    plan = pk.SrcPlan(L, M, np.float32)
    for channels, n_in in ((3, 4412), (261, 2052), (1, 30000), (2, 4411)):
        if n_in * L < o.src_geometry(n_in, L, M)[0]:
            continue
        # rows must be 16-byte aligned for the TMA path; odd lengths fall back to the FFMA kernels
        monkeypatch.setenv("DSPB200_SRC_FORCE_MMA", "1")     # small channel counts normally stay on the FFMA kernel
        assert (plan.kernel_kind(channels, n_in) == "tensor") == (n_in % 4 == 0)
        x = rng.uniform(-1, 1, (channels, n_in)).astype(np.float32)
        xt = torch.as_tensor(x, device="cuda")
        y = plan.run(xt)
        y_tiled = plan.run(xt, force_tiled=True)
        ref = np.stack([o.resample_closed_form(x[c].astype(np.float64), 48000, M, L)[0]
                        for c in (0, channels // 2, channels - 1)])
        got = y.cpu().numpy()[[0, channels // 2, channels - 1]]
        scale = max(1.0, float(np.max(np.abs(ref))))
        assert got.shape == ref.shape
        assert np.max(np.abs(got - ref)) <= TOL_F32_SRC * scale, (channels, n_in)
        assert float((y - y_tiled).abs().max()) <= TOL_F32_SRC * scale
        # a strided view (row pitch larger than the row) and an output tensor with its own pitch
        big = torch.zeros((channels, n_in + 12), device="cuda", dtype=torch.float32)
        big[:, :n_in] = xt
        y2 = plan.run(big[:, :n_in])
        assert torch.equal(y2, y)
        monkeypatch.delenv("DSPB200_SRC_FORCE_MMA")
        assert plan.kernel_kind(channels, n_in) != "tensor" or 4 * channels >= 3 * (-(-channels // 256) * 256)


@pytest.mark.parametrize("gains", [C1_GAINS, (15,) * 6, (3, 0, -2, 0, 0, 5), (0, 0, 0, 0, 0, 12), (-15,) * 6,
                                   (-12.0412, 15, -12.5, 6, 0, -3)])
def test_eq_tensor_core_form_matches_oracle_and_scan(pk, torch_cuda, gains, monkeypatch):
    """fp32 EQ on wide batches runs the cascade as one linear system, 112 samples per tcgen05 GEMM tile,
    the state carried from chunk to chunk in the registers of the thread that owns the channel
    (csrc/eq_mma.cu).  Forced onto narrow batches here; ragged channel counts, lengths that are not a
    multiple of the chunk, more channel groups than SMs (a CTA restarts from a zero state), in
    place, padded and misaligned row strides."""
    torch = torch_cuda
    gd = gains_dict(gains)
    plan = pk.EqPlan.from_gains(48000, gd, np.float32)
    rng = np.random.default_rng(77)
    for channels, n in ((300, 4412), (256, 96 * 9), (5, 20000), (128 * 150 + 5, 1124), (3, 1123), (2, 100)):
        monkeypatch.setenv("DSPB200_EQ_FORCE_MMA", "1")
        assert (plan.kernel_kind(channels, n) == "tensor") == (n % 4 == 0 and n >= 96)
        x = rng.uniform(-0.5, 0.5, (channels, n)).astype(np.float32)
        xt = torch.as_tensor(x, device="cuda")
        z = plan.run(xt)
        monkeypatch.setenv("DSPB200_EQ_NO_MMA", "1")
        assert plan.kernel_kind(channels, n) == "scan"
        z_scan = plan.run(xt)
        monkeypatch.delenv("DSPB200_EQ_NO_MMA")
        pick = sorted({0, channels // 2, channels - 1})
        ref = np.stack([o.equalizer(x[c].astype(np.float64), 48000, gd) for c in pick])
        assert o.full_scale_err(z.cpu().numpy()[pick], ref) <= TOL_F32_EQ, (channels, n)
        assert float((z - z_scan).abs().max()) <= TOL_F32_EQ, (channels, n)
        zi = xt.clone()
        plan.run(zi, out=zi)
        # in place the time axis is never cut into overlapping slices; out of place a long narrow batch is (5 x 20000
        # here), and then the two agree to float32 rounding instead of bit for bit
        if n < 96 * 16:
            assert torch.equal(zi, z)
        else:
            assert float((zi - z).abs().max()) <= 2e-6
        big = torch.zeros((channels, n + 12), device="cuda", dtype=torch.float32)
        big[:, :n] = xt
        out = torch.full((channels, n + 4), 7.0, device="cuda", dtype=torch.float32)
        plan.run(big[:, :n], out=out[:, :n])
        assert torch.equal(out[:, :n], z) and bool((out[:, n:] == 7.0).all())
        monkeypatch.delenv("DSPB200_EQ_FORCE_MMA")
    # the shape rule: wide batches take the tensor form on their own, narrow ones stay on the scan kernel
    assert plan.kernel_kind(18944, 3000) == "tensor" and plan.kernel_kind(1024, 2000) == "scan"
    x = torch.rand((18944, 3000), device="cuda", dtype=torch.float32) - 0.5
    z = plan.run(x)
    monkeypatch.setenv("DSPB200_EQ_NO_MMA", "1")
    z_scan = plan.run(x)
    monkeypatch.delenv("DSPB200_EQ_NO_MMA")
    assert float((z - z_scan).abs().max()) <= TOL_F32_EQ
    # fp64 stays on the scan kernel
    assert pk.EqPlan.from_gains(48000, gd, np.float64).kernel_kind(18944, 3000) == "scan"


@pytest.mark.parametrize("n_sections", [2, 5, 7, 8])
def test_eq_tensor_core_form_section_counts(pk, torch_cuda, n_sections, monkeypatch):
    """Every state-count variant of the tensor form (4, 8, 12 and 16 padded states): cascades of 2..8 sections
    built from explicit band lists (eight sections need both k-blocks of the free-response operand)."""
    torch = torch_cuda
    centres = [60.0, 250.0, 700.0, 1500.0, 3200.0, 6000.0, 9000.0, 14000.0][:n_sections]
    gains = [5.0, -4.0, 7.5, -6.0, 3.0, -9.0, 12.0, -2.5][:n_sections]
    plan = pk.EqPlan(48000, list(zip(centres, gains)), np.float32)
    rng = np.random.default_rng(n_sections)
    channels, n = 261, 9600 + 4
    x = rng.uniform(-0.4, 0.4, (channels, n)).astype(np.float32)
    xt = torch.as_tensor(x, device="cuda")
    monkeypatch.setenv("DSPB200_EQ_FORCE_MMA", "1")
    assert plan.kernel_kind(channels, n) == "tensor"
    z = plan.run(xt).cpu().numpy()
    monkeypatch.delenv("DSPB200_EQ_FORCE_MMA")
    for c in (0, 130, 260):
        ref = x[c].astype(np.float64)
        for fc, g in zip(centres, gains):
            ref = o.difference_equation(ref, *o.peaking_biquad(fc, 48000, g))
        assert o.full_scale_err(z[c], np.clip(ref, -1, 1)) <= TOL_F32_EQ, (n_sections, c)


@pytest.mark.parametrize("slices", [2, 5, 13])
def test_eq_tensor_core_form_time_slices(pk, torch_cuda, slices, monkeypatch):
    """With more channel groups than SMs the tensor form cuts the time axis into slices to even out the rounds;
    a later slice starts from the end state its predecessor left in memory.  Forced here on a small batch (the
    predecessor is then still running when the successor asks for its state)."""
    torch = torch_cuda
    gd = gains_dict((15,) * 6)
    plan = pk.EqPlan.from_gains(48000, gd, np.float32)
    rng = np.random.default_rng(slices)
    channels, n = 389, 96 * 61 + 40
    x = rng.uniform(-0.25, 0.25, (channels, n)).astype(np.float32)
    xt = torch.as_tensor(x, device="cuda")
    monkeypatch.setenv("DSPB200_EQ_FORCE_MMA", "1")
    z1 = plan.run(xt)
    monkeypatch.setenv("DSPB200_EQ_SLICES", str(slices))
    z = plan.run(xt)
    monkeypatch.delenv("DSPB200_EQ_SLICES")
    monkeypatch.delenv("DSPB200_EQ_FORCE_MMA")
    assert float((z - z1).abs().max()) <= 2e-6            # same arithmetic up to the fp32 round trip of the state
    pick = [0, 200, 388]
    ref = np.stack([o.equalizer(x[c].astype(np.float64), 48000, gd) for c in pick])
    assert o.full_scale_err(z.cpu().numpy()[pick], ref) <= TOL_F32_EQ


def test_eq_tensor_core_form_more_groups_than_sms(pk, torch_cuda, monkeypatch):
    """150 x 128 channels on 148 SMs: the slice count is chosen by the library; against the scan kernel."""
    torch = torch_cuda
    plan = pk.EqPlan.from_gains(48000, gains_dict(C1_GAINS), np.float32)
    x = torch.rand((150 * 128, 9600), device="cuda", dtype=torch.float32) - 0.5
    assert plan.kernel_kind(150 * 128, 9600) == "tensor"
    z = plan.run(x)
    monkeypatch.setenv("DSPB200_EQ_NO_MMA", "1")
    z_scan = plan.run(x)
    monkeypatch.delenv("DSPB200_EQ_NO_MMA")
    assert float((z - z_scan).abs().max()) <= TOL_F32_EQ


def test_eq_tensor_core_form_c3_channel_count_properties(pk, torch_cuda):
    """C3's channel count (65536 = 512 groups on 148 SMs: the library slices the time axis) through the tensor form,
    checked by properties that do not need the oracle at full size: impulse responses scale with the impulse,
    arrive where the impulse was put, and equal the cascade's impulse response; the map is linear before the clip."""
    torch = torch_cuda
    ch, n = 65536, 96 * 300
    gd = gains_dict(C1_GAINS)
    sections = pk.select_sections(48000, gd)[1]
    plan = pk.EqPlan(48000, sections, np.float32, clip=False)
    assert plan.kernel_kind(ch, n) == "tensor"
    # impulses: channel c gets amplitude a_c at time t_c
    gen = torch.Generator(device="cuda").manual_seed(5)
    amp = torch.rand(ch, device="cuda", generator=gen) + 0.5
    t_c = torch.randint(0, n - 4000, (ch,), device="cuda", generator=gen)
    x = torch.zeros((ch, n), device="cuda", dtype=torch.float32)
    x[torch.arange(ch, device="cuda"), t_c] = amp
    z = plan.run(x)
    h = np.zeros(4000)
    h[0] = 1.0
    for fc, g in sections:
        h = o.difference_equation(h, *o.peaking_biquad(fc, 48000, g))
    ht = torch.as_tensor(h, device="cuda", dtype=torch.float32)
    pick = torch.randint(0, ch, (512,), device="cuda", generator=gen)
    idx = t_c[pick][:, None] + torch.arange(4000, device="cuda")[None, :]
    got = torch.gather(z[pick], 1, idx)
    assert float((got - amp[pick][:, None] * ht[None, :]).abs().max()) <= 1e-5
    before = torch.arange(n, device="cuda")[None, :] < t_c[pick][:, None]
    assert float((z[pick] * before).abs().max()) == 0.0           # nothing before the impulse
    # linearity on noise
    x1 = torch.rand((ch, n), device="cuda", generator=gen) - 0.5
    x.uniform_(-0.5, 0.5, generator=gen)
    z1, z2 = plan.run(x1), plan.run(x)
    x1.mul_(0.75).add_(x, alpha=-1.25)
    z12 = plan.run(x1)
    assert float((z12 - (0.75 * z1 - 1.25 * z2)).abs().max()) <= 2e-5


def test_eq_streaming_blocks_reproduce_one_pass(pk, torch_cuda, monkeypatch):
    """dspb200_eq_run_stream_f32: consecutive time blocks with the state carried between the calls (how C3's
    65536 x 2.88 M samples fit through HBM) give the same samples as one pass, bit for bit, and match the oracle."""
    torch = torch_cuda
    gd = gains_dict((15, -3, 4, -14, 3, -9))
    plan = pk.EqPlan.from_gains(48000, gd, np.float32)
    chunk = plan.stream_chunk()
    assert chunk == 96
    rng = np.random.default_rng(21)
    channels, n = 300, chunk * 50 + 36          # rows must stay 16-byte aligned for the streaming form
    x = rng.uniform(-0.3, 0.3, (channels, n)).astype(np.float32)
    xt = torch.as_tensor(x, device="cuda")
    monkeypatch.setenv("DSPB200_EQ_FORCE_MMA", "1")
    whole = plan.run(xt)
    monkeypatch.delenv("DSPB200_EQ_FORCE_MMA")
    cuts = [0, chunk * 7, chunk * 27, chunk * 28, n]
    state, parts = None, []
    for a, b in zip(cuts[:-1], cuts[1:]):
        blk = xt[:, a:b].contiguous()
        z, state = plan.run_stream(blk, state)
        parts.append(z)
    got = torch.cat(parts, dim=1)
    assert torch.equal(got, whole)
    pick = [0, 150, 299]
    ref = np.stack([o.equalizer(x[c].astype(np.float64), 48000, gd) for c in pick])
    assert o.full_scale_err(got.cpu().numpy()[pick], ref) <= TOL_F32_EQ
    # in place on a view of a larger buffer, and the error paths
    buf = xt.clone()
    state = None
    for a, b in zip(cuts[:-1], cuts[1:]):
        _, state = plan.run_stream(buf[:, a:b], state, out=buf[:, a:b])
    assert torch.equal(buf, whole)
    with pytest.raises(ValueError):
        plan.run_stream(xt, torch.zeros((channels, 8), device="cuda"))
    with pytest.raises(Exception):
        pk.EqPlan(48000, [(1000.0 + 100 * i, 3.0) for i in range(9)], np.float32).run_stream(xt)


def test_eq_tensor_core_form_long_stream(pk, torch_cuda, monkeypatch):
    """C3's time axis (2.88 M samples = 25 715 chunks) through the tensor form on two channel groups."""
    torch = torch_cuda
    n = 2_880_000
    gd = gains_dict((15,) * 6)
    plan = pk.EqPlan.from_gains(48000, gd, np.float32)
    rng = np.random.default_rng(3)
    x = rng.uniform(-0.25, 0.25, (2, n)).astype(np.float32)
    xt = torch.zeros((256, n), device="cuda", dtype=torch.float32)
    xt[17], xt[255] = torch.as_tensor(x[0], device="cuda"), torch.as_tensor(x[1], device="cuda")
    monkeypatch.setenv("DSPB200_EQ_FORCE_MMA", "1")
    assert plan.kernel_kind(256, n) == "tensor"
    z = plan.run(xt)
    monkeypatch.delenv("DSPB200_EQ_FORCE_MMA")
    ref = np.stack([o.equalizer(x[c].astype(np.float64), 48000, gd) for c in range(2)])
    assert o.full_scale_err(z[[17, 255]].cpu().numpy(), ref) <= TOL_F32_EQ
    assert float(z[0].abs().max()) == 0.0


@pytest.mark.parametrize("n_fft", [32768, 65536, 131072])
def test_fft_long_split_transforms(pk, dc, torch_cuda, n_fft):
    """Transforms longer than one CTA's shared memory take a top-level radix-2..16
    split through a workspace; ragged / misaligned / many-frame input against numpy."""
    torch = torch_cuda
    rng = np.random.default_rng(n_fft + 1)
    w = o.hann_symmetric(n_fft)
    channels, n_frames = 2, 3
    x = rng.uniform(-1, 1, (channels, n_fft * n_frames + 9))
    ref = np.abs(np.fft.rfft(x[:, :n_fft * n_frames].reshape(channels, n_frames, n_fft) * w, axis=-1))
    xz = np.zeros((channels, 3 * n_fft)); xz[:, :n_fft + 7] = x[:, :n_fft + 7]
    ref2 = np.stack([np.abs(np.fft.rfft(xz[:, 3 + f * (n_fft // 2):3 + f * (n_fft // 2) + n_fft] * w, axis=-1))
                     for f in range(2)], axis=1)
    for dt, tol in ((np.float64, TOL_F64), (np.float32, TOL_F32_FFT)):
        plan = pk.FftPlan(n_fft, dt, hann=True)
        xt = torch.as_tensor(x.astype(dt), device="cuda")
        mag = plan.magnitudes(xt).cpu().numpy()
        assert o.rel_err(mag, ref) <= tol, dt
        m2 = plan.magnitudes(xt, hop=n_fft // 2, offset=3, n_frames=2, n_valid=n_fft + 7).cpu().numpy()
        assert o.rel_err(m2, ref2) <= tol, (dt, "ragged")
    if n_fft == 32768:
        # many frames: every CTA loops over several transforms
        gen = torch.Generator(device="cuda").manual_seed(11)
        xm = torch.rand((700, n_fft), generator=gen, device="cuda") * 2 - 1
        plan = pk.FftPlan(n_fft, np.float32, hann=False)
        mag = plan.magnitudes(xm)[:, 0]
        refm = torch.fft.rfft(xm.double(), dim=-1).abs()
        assert float((mag.double() - refm).abs().max() / refm.max()) <= TOL_F32_FFT
        xc = rng.normal(size=n_fft) + 1j * rng.normal(size=n_fft)
        assert o.rel_err(dc.fft_diezmado_en_tiempo(xc), np.fft.fft(xc)) <= TOL_F64


# ---------------------------------------------------------------- chain ----
def test_chain_c1_golden(dc, pk, golden_chain, torch_cuda):
    torch = torch_cuda
    g = golden_chain
    x = g["x"]
    gd = gains_dict(C1_GAINS)
    y, fs2 = dc.conversion_tasa_muestreo(x, 44100, 2, 3)
    z = dc.sistema_ecualizador(y, fs2, gd)
    assert fs2 == int(g["fs2"])
    assert o.rel_err(y[:4096], g["y_head"]) <= TOL_F64 and o.rel_err(y[-4096:], g["y_tail"]) <= TOL_F64
    assert o.rel_err(z[:4096], g["z_head"]) <= TOL_F64 and o.rel_err(z[-4096:], g["z_tail"]) <= TOL_F64
    f, m = dc.calcular_espectro_magnitud(z[:100000], fs2)
    assert o.rel_err(m, g["m_app"]) <= TOL_F64 and np.allclose(f, g["f_app"])
    mid = len(z) // 2
    frame = pk.FftPlan(4096, np.float64).magnitudes_host(z, offset=mid, n_frames=1)[0, 0]
    assert o.rel_err(frame, g["mag4096"]) <= TOL_F64
    assert o.rel_err(y[mid:mid + 4096], g["y_mid"]) <= TOL_F64 and o.rel_err(z[mid:mid + 4096], g["z_mid"]) <= TOL_F64
    assert np.allclose([np.sum(z), np.sum(np.abs(z)), np.sum(z * z)], g["z_sum"], rtol=1e-9)
    # batched chain object, fp64 and fp32, device and host forms: 30 s against the reference's own
    # y/z windows and its 4096-point frames every 2^17 samples, and against the oracle over the whole clip
    yo, zo, mo, _ = o.chain(x.astype(np.float64), 44100, 2, 3, gd, n_fft=4096, n_frames=3)
    fidx = (g["frame_starts"] // 4096).astype(int)
    for dt, ty, tz, tm in ((np.float64, TOL_F64, TOL_F64, TOL_F64), (np.float32, TOL_F32_SRC, TOL_F32_EQ, 1e-4)):
        ch = pk.Chain(3, 2, 44100, gd, n_fft=4096, dtype=dt)
        xt = torch.as_tensor(np.stack([x, x[::-1]]).astype(dt), device="cuda")
        yd, zd, md = ch.run(xt, keep_y=True)
        assert o.full_scale_err(yd[0].cpu().numpy(), yo) <= ty
        assert o.full_scale_err(zd[0].cpu().numpy(), zo) <= tz
        assert o.full_scale_err(zd[0, -4096:].cpu().numpy(), g["z_tail"]) <= tz
        assert o.rel_err(md[0, :3].cpu().numpy(), mo) <= tm
        got = md[0].cpu().numpy()[fidx]
        for k in range(len(fidx)):
            assert o.rel_err(got[k], g["frames_mag"][k]) <= tm, (dt, k)
        zh, mh = ch.run_host(np.stack([x, x[::-1]]).astype(dt))
        assert np.array_equal(zh, zd.cpu().numpy()) and np.array_equal(mh, md.cpu().numpy())
        _, z2, m2 = ch.run(xt, keep_y=False)
        assert torch.equal(z2, zd) and torch.equal(m2, md)


def test_chain_c5_shape_slice(pk, torch_cuda):
    """C5's clip shape (10 s @ 44.1 kHz -> 48 kHz, 6-band EQ, 4096-point frames)
    on 256 clips; oracle on two clips, round-trip properties on the rest."""
    torch = torch_cuda
    gd = gains_dict(C1_GAINS)
    ch = pk.Chain(160, 147, 44100, gd, n_fft=4096, dtype=np.float32)
    gen = torch.Generator(device="cuda").manual_seed(4)
    x = torch.rand((256, 441000), generator=gen, device="cuda") - 0.5
    _, z, mag = ch.run(x)
    assert z.shape == (256, 480000) and mag.shape == (256, 117, 2049)
    assert float(z.abs().max()) <= 1.0
    for c in (0, 255):
        yo, zo, mo, _ = o.chain(x[c].cpu().numpy().astype(np.float64), 44100, 147, 160, gd, n_fft=4096)
        assert o.full_scale_err(z[c].cpu().numpy(), zo) <= TOL_F32_EQ
        assert o.rel_err(mag[c].cpu().numpy(), mo) <= 1e-4
    # Parseval per frame ties the spectra back to z
    w = torch.as_tensor(o.hann_symmetric(4096), device="cuda")
    fr = z[:, :117 * 4096].view(256, 117, 4096).double() * w
    et = (fr ** 2).sum(-1)
    m = mag.double()
    ef = (2 * (m ** 2).sum(-1) - m[..., 0] ** 2 - m[..., -1] ** 2) / 4096
    assert float(((ef - et).abs() / et).max()) <= 1e-4


# ------------------------------------------------- fused SRC->EQ (xz_mma) ----
def _chain_ref(x_row, gains_or_sections, L=160, M=147, fs=44100):
    """float64 oracle of SRC -> EQ for one channel; gains dict or [(fc, gain_db)] list"""
    y, fs2 = o.resample_closed_form(x_row.astype(np.float64), fs, M, L)
    if isinstance(gains_or_sections, dict):
        return o.equalizer(y, fs2, gains_or_sections)
    for fc, g in gains_or_sections:
        y = o.difference_equation(y, *o.peaking_biquad(fc, fs2, g))
    return np.clip(y, -1, 1)


@pytest.mark.parametrize("gains", [C1_GAINS, (15,) * 6, (-15,) * 6, (0, 0, 0, 0, 0, 12), (3, 0, -2, 0, 0, 5),
                                   (-12.0412, 15, -12.5, 6, 0, -3)])
def test_chain_fused_matches_oracle(pk, torch_cuda, gains):
    """The one-kernel SRC->EQ form (fp16 three-product split on tcgen05, y never written) against the float64 oracle
    and against the three-kernel cascade: 1..6 sections (4, 8 and 12 padded states), real and complex poles, a batch
    that ends inside a channel group, an output length that ends inside a chunk and inside a 32-sample store box."""
    torch = torch_cuda
    gd = gains_dict(gains)
    ch = pk.Chain(160, 147, 44100, gd, n_fft=4096, dtype=np.float32)
    rng = np.random.default_rng(7)
    channels, n_in = 130, 22052
    x = rng.uniform(-0.5, 0.5, (channels, n_in)).astype(np.float32)
    xt = torch.as_tensor(x, device="cuda")
    z = ch.run_fused(xt)
    assert z.shape == (channels, ch.out_len(n_in)) and not bool(torch.isnan(z).any())
    y = ch.src.run(xt)
    z3 = ch.eq.run(y)
    assert float((z - z3).abs().max()) <= TOL_F32_EQ
    zc = z.cpu().numpy()
    for c in (0, 31, 32, 127, 128, 129):
        assert o.full_scale_err(zc[c], _chain_ref(x[c], gd)) <= TOL_F32_EQ, (gains, c)


@pytest.mark.parametrize("n_sections", [7, 8])
def test_chain_fused_sixteen_states(pk, torch_cuda, n_sections):
    """Cascades of 7 and 8 sections (16 padded states: the widest variant of the fused kernel)."""
    torch = torch_cuda
    centres = [60.0, 250.0, 700.0, 1500.0, 3200.0, 6000.0, 9000.0, 14000.0][:n_sections]
    gains = [5.0, -4.0, 7.5, -6.0, 3.0, -9.0, 12.0, -2.5][:n_sections]
    ch = pk.Chain(160, 147, 44100, gains_dict(C1_GAINS), n_fft=4096, dtype=np.float32)
    ch.eq = pk.EqPlan(ch.fs_out, list(zip(centres, gains)), np.float32)
    rng = np.random.default_rng(n_sections)
    x = rng.uniform(-0.4, 0.4, (64, 9000)).astype(np.float32)
    z = ch.run_fused(torch.as_tensor(x, device="cuda")).cpu().numpy()
    for c in (0, 63):
        assert o.full_scale_err(z[c], _chain_ref(x[c], list(zip(centres, gains)))) <= TOL_F32_EQ, (n_sections, c)


@pytest.mark.parametrize("n_in", [41, 147, 300, 4412, 22049])
def test_chain_fused_lengths(pk, torch_cuda, n_in):
    """Short and ragged lengths: from the shortest long-signal input (41 samples: 41 * 160 >= 6401 taps) up; one chunk,
    an odd number of chunks, outputs that are not a multiple of 4 (rows padded by the caller)."""
    torch = torch_cuda
    gd = gains_dict(C1_GAINS)
    ch = pk.Chain(160, 147, 44100, gd, n_fft=4096, dtype=np.float32)
    rng = np.random.default_rng(n_in)
    x = rng.uniform(-1.0, 1.0, (3, n_in)).astype(np.float32)
    pitch = -(-n_in // 4) * 4
    xt = torch.zeros((3, pitch), device="cuda")[:, :n_in]
    xt.copy_(torch.as_tensor(x))
    n_out = ch.out_len(n_in)
    guard = torch.full((3, -(-n_out // 4) * 4 + 8), 7.0, device="cuda")
    z = ch.run_fused(xt, out=guard[:, :n_out])
    # TMA stores clip at 16-byte granularity: the row's own padding up to a multiple of 4 samples may be written,
    # nothing beyond it
    assert bool((guard[:, -(-n_out // 4) * 4:] == 7.0).all())
    for c in range(3):
        assert o.full_scale_err(z[c].cpu().numpy(), _chain_ref(x[c], gd)) <= TOL_F32_EQ, (n_in, c)


def test_chain_fused_more_groups_than_sms_and_dispatch(pk, torch_cuda, monkeypatch):
    """19000 channels = 149 groups of 128 on 148 SMs (one CTA walks two groups: ring, window and staging state carry
    over), picked by Chain.run on its own; narrow batches and float64 stay on the three-kernel cascade; the
    environment switches; identical results through run() and run_fused()."""
    torch = torch_cuda
    gd = gains_dict(C1_GAINS)
    ch = pk.Chain(160, 147, 44100, gd, n_fft=4096, dtype=np.float32)
    assert ch.kernel_kind(19000, 7644) == "fused" and ch.kernel_kind(1024, 441000) == "cascade"
    assert ch.kernel_kind(19000, 2000) == "cascade"          # 2177 outputs: the dense z rows would not be 16-byte aligned
    assert pk.Chain(3, 2, 44100, gd, n_fft=4096, dtype=np.float32).kernel_kind(19000, 7644) == "cascade"
    assert pk.Chain(160, 147, 44100, gd, n_fft=4096, dtype=np.float64).kernel_kind(19000, 7644) == "cascade"
    gen = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand((19000, 7644), generator=gen, device="cuda") - 0.5
    _, z, mag = ch.run(x)
    assert torch.equal(z, ch.run_fused(x))
    monkeypatch.setenv("DSPB200_CHAIN_NO_FUSED", "1")
    assert ch.kernel_kind(19000, 7644) == "cascade"
    _, z3, mag3 = ch.run(x)
    monkeypatch.delenv("DSPB200_CHAIN_NO_FUSED")
    assert float((z - z3).abs().max()) <= TOL_F32_EQ
    m_ref = float(mag3.abs().max())
    assert float((mag - mag3).abs().max()) <= 1e-4 * m_ref
    for c in (0, 127, 128, 18943, 18944, 18999):
        assert o.full_scale_err(z[c].cpu().numpy(), _chain_ref(x[c].cpu().numpy(), gd)) <= TOL_F32_EQ, c
    monkeypatch.setenv("DSPB200_CHAIN_FORCE_FUSED", "1")
    assert ch.kernel_kind(8, 7644) == "fused"
    monkeypatch.delenv("DSPB200_CHAIN_FORCE_FUSED")
    # keep_y needs y: always the cascade, and its z agrees with the fused one
    y, zk, _ = ch.run(x[:300], keep_y=True)
    assert y is not None and float((zk - z[:300]).abs().max()) <= TOL_F32_EQ


def test_chain_fused_linearity_and_full_clip_length(pk, torch_cuda):
    """C5's clip length on 256 clips through the fused kernel: oracle on two clips; linearity of the unclipped cascade
    (all gains negative keeps |z| < 1): z(a x1 + b x2) = a z(x1) + b z(x2)."""
    torch = torch_cuda
    gd = gains_dict((-3, -4, -2, -6, -3, -9))
    ch = pk.Chain(160, 147, 44100, gd, n_fft=4096, dtype=np.float32)
    gen = torch.Generator(device="cuda").manual_seed(6)
    x1 = 0.4 * (torch.rand((256, 441000), generator=gen, device="cuda") - 0.5)
    x2 = 0.4 * (torch.rand((256, 441000), generator=gen, device="cuda") - 0.5)
    z1, z2 = ch.run_fused(x1), ch.run_fused(x2)
    z12 = ch.run_fused(0.75 * x1 - 1.25 * x2)
    assert float(z12.abs().max()) < 1.0
    assert float((z12 - (0.75 * z1 - 1.25 * z2)).abs().max()) <= 1e-5
    for c in (0, 255):
        assert o.full_scale_err(z1[c].cpu().numpy(), _chain_ref(x1[c].cpu().numpy(), gd)) <= TOL_F32_EQ


def test_tensor_forms_nonfinite_inputs_stay_local_in_time(pk, torch_cuda, monkeypatch):
    """Documented deviation of the tensor-core forms: a non-finite input sample poisons the whole chunk it falls in
    (96 samples in the EQ kernel, 80 outputs in the fused chain kernel) and, through the recurrence, everything after
    it -- the reference (lfilter) only the samples from it on.  Samples before the chunk are untouched in both."""
    torch = torch_cuda
    gd = gains_dict(C1_GAINS)
    plan = pk.EqPlan.from_gains(48000, gd, np.float32)
    x = torch.rand((4, 960), device="cuda") - 0.5
    x[2, 500] = float("nan")
    monkeypatch.setenv("DSPB200_EQ_FORCE_MMA", "1")
    z = plan.run(x)
    monkeypatch.delenv("DSPB200_EQ_FORCE_MMA")
    monkeypatch.setenv("DSPB200_EQ_NO_MMA", "1")
    z_scan = plan.run(x)
    monkeypatch.delenv("DSPB200_EQ_NO_MMA")
    bad = torch.isnan(z[2]).nonzero().flatten()
    bad_scan = torch.isnan(z_scan[2]).nonzero().flatten()
    assert int(bad_scan.min()) == 500                                   # sequential form: from the sample on, like lfilter
    assert int(bad.min()) == 480 and int(bad.max()) == 959              # tensor form: from the start of its 96-sample chunk
    assert not bool(torch.isnan(z[[0, 1, 3]]).any())
    assert float((z[2, :480] - z_scan[2, :480]).abs().max()) <= TOL_F32_EQ
    ch = pk.Chain(160, 147, 44100, gd, n_fft=4096, dtype=np.float32)
    xc = torch.rand((4, 4412), device="cuda") - 0.5
    xc[1, 2000] = float("inf")
    zf = ch.run_fused(xc)
    first = int((~torch.isfinite(zf[1])).nonzero().flatten().min())
    # input 2000 first reaches output ceil((2000 * 160 - 3200) / 147) = 2156; its chunk starts at 2080 (the k-step that holds
    # the sample skips the chunk's first coefficient rows, so the poison starts somewhere inside the chunk)
    assert 2080 <= first <= 2156
    assert bool(torch.isfinite(zf[[0, 2, 3]]).all())


def test_wrong_plan_handles_are_rejected(pk, torch_cuda):
    """Plans carry a type tag: an entry point handed another plan type (or a destroyed plan) answers an error code
    (ValueError through the shim) instead of reading the wrong layout."""
    import ctypes as C
    from dsp_audio_project_b200 import _lib
    torch = torch_cuda
    lib = _lib.load()
    src = pk.SrcPlan(3, 2, np.float32)
    eq = pk.EqPlan.from_gains(48000, gains_dict(C1_GAINS), np.float32)
    fft = pk.FftPlan(256, np.float32)
    x = torch.rand((2, 512), device="cuda")
    out = torch.empty((2, 1024), device="cuda")
    state = torch.zeros((2, 16), device="cuda")
    args = (x.data_ptr(), 512, out.data_ptr(), 1024, 2, 512)
    assert lib.dspb200_eq_run_f32(src._h, *args, None) == _lib.ERR_INVALID
    assert "not a live eq plan" in _lib.last_error()
    assert lib.dspb200_eq_run_stream_f32(src._h, *args, state.data_ptr(), 1, None) == _lib.ERR_INVALID
    assert lib.dspb200_src_run_f32(eq._h, *args, None) == _lib.ERR_INVALID
    assert lib.dspb200_src_run_f32(fft._h, *args, None) == _lib.ERR_INVALID
    ws = C.c_size_t()
    assert lib.dspb200_fft_workspace_bytes(eq._h, 4, C.byref(ws)) == _lib.ERR_INVALID
    k = C.c_int()
    assert lib.dspb200_chain_kernel_kind(eq._h, src._h, 4, 512, 512, C.byref(k)) == _lib.ERR_INVALID
    assert lib.dspb200_eq_plan_destroy(fft._h) == _lib.ERR_INVALID
    with pytest.raises(ValueError):
        _lib.check(lib.dspb200_fftmag_run_f32(src._h, x.data_ptr(), 512, 512, 0, 256, 2, out.data_ptr(), 129, 258, 2, None, 0, None))
    assert not hasattr(pk.SrcPlan, "run_stream")           # the resampler has no streaming form
    # plans that own device tables refuse to run with another device current
    if torch.cuda.device_count() > 1:
        x1 = torch.rand((2, 512), device="cuda:1")
        with torch.cuda.device(1):
            rc = lib.dspb200_src_run_f32(src._h, x1.data_ptr(), 512, x1.data_ptr(), 512, 2, 300, None)
        assert rc == _lib.ERR_INVALID and "built on device 0" in _lib.last_error()


def test_library_reports_launches(pk):
    from dsp_audio_project_b200 import _lib
    assert _lib.launch_count() > 0


def test_long_fft_forms_match(pk, torch_cuda, monkeypatch):
    """The forms of the 2^16-point transform give the same spectra, ragged tail included: the float32 default (three
    radix-32 passes in one launch, csrc/fft_long32.cu), the two four-step kernels (DSPB200_FFT_LONG32=0; the float64
    default) and their opt-in single-launch form (DSPB200_FFT_FUSED4=1: clusters of four CTAs keep the workspace in L2).
    Also the rectangular window, an odd offset with overlapping frames, and the dB store of the three-pass form."""
    torch = torch_cuda
    rng = np.random.default_rng(65)
    w = o.hann_symmetric(65536)
    for dt, tol in ((np.float32, TOL_F32_FFT), (np.float64, TOL_F64)):
        x = torch.as_tensor(rng.uniform(-1, 1, (5, 3 * 65536 + 1000)).astype(dt), device="cuda")
        plan = pk.FftPlan(65536, dt, hann=True)
        got = {}
        for name, env in (("default", {}), ("four_step", {"DSPB200_FFT_LONG32": "0"}),
                          ("fused4", {"DSPB200_FFT_LONG32": "0", "DSPB200_FFT_FUSED4": "1"})):
            for k in ("DSPB200_FFT_LONG32", "DSPB200_FFT_FUSED4"):
                monkeypatch.delenv(k, raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            got[name] = plan.magnitudes(x, n_frames=4, n_valid=3 * 65536 + 1000).clone()   # the 4th frame is zero padded
        for k in ("DSPB200_FFT_LONG32", "DSPB200_FFT_FUSED4"):
            monkeypatch.delenv(k, raising=False)
        ref = got["four_step"]
        bound = 2e-6 if dt == np.float32 else 1e-14
        for name in ("default", "fused4"):
            assert float((got[name] - ref).abs().max() / ref.abs().max()) <= bound, (dt, name)
        xh = x.cpu().numpy().astype(np.float64)
        for name in got:
            want = np.abs(np.fft.rfft(xh[2, 65536:2 * 65536] * w))
            assert o.rel_err(got[name][2, 1].cpu().numpy(), want) <= tol, (dt, name)
            tail = np.zeros(65536); tail[:1000] = xh[4, 3 * 65536:]
            assert o.rel_err(got[name][4, 3].cpu().numpy(), np.abs(np.fft.rfft(tail * w))) <= tol, (dt, name)
    # float32 three-pass form: no window, odd offset (misaligned loads), overlapping frames; dB spectra
    x = torch.as_tensor(rng.uniform(-1, 1, (3, 2 * 65536 + 77)).astype(np.float32), device="cuda")
    xh = x.cpu().numpy().astype(np.float64)
    m = pk.FftPlan(65536, np.float32, hann=False).magnitudes(x, hop=32768 + 5, offset=3, n_frames=3).cpu().numpy()
    xp = np.zeros((3, max(xh.shape[1], 3 + 2 * (32768 + 5) + 65536))); xp[:, :xh.shape[1]] = xh
    for f in range(3):
        s0 = 3 + f * (32768 + 5)
        assert o.rel_err(m[:, f], np.abs(np.fft.rfft(xp[:, s0:s0 + 65536], axis=-1))) <= TOL_F32_FFT, f
    db = pk.FftPlan(65536, np.float32, hann=True, db=True).magnitudes(x).cpu().numpy()
    ref = np.abs(np.fft.rfft(xh[:, :2 * 65536].reshape(3, 2, 65536) * w, axis=-1))
    loud = ref > 1e-2 * ref.max()
    assert np.max(np.abs(db - 20.0 * np.log10(ref + 1e-12))[loud]) <= 20.0 * np.log10(1.0 + TOL_F32_FFT / 1e-2) + 1e-4


def test_c5_wave_frames_parseval_and_spot_frames(pk, torch_cuda):
    """An eighth of the bench wave (2368 clips x 480000 samples: 277056 frames of 4096 points, float32) through the
    32-points-per-thread kernel: every group walks ~230 (channel, frame) items by grid-stride carries.  Parseval of every
    frame (float64 on the device), the dropped tail (480000 = 117 x 4096 + 768) and spot frames against numpy."""
    torch = torch_cuda
    ch, n, nf = 2368, 480000, 4096
    x = torch.empty((ch, n), dtype=torch.float32, device="cuda")
    pk.generate_uniform(x, 4, -1.0, 1.0)
    plan = pk.FftPlan(nf, np.float32, hann=True)
    mag = plan.magnitudes(x)
    frames = n // nf
    assert tuple(mag.shape) == (ch, frames, nf // 2 + 1)
    wh = o.hann_symmetric(nf)
    w = torch.as_tensor(wh, device="cuda", dtype=torch.float64)
    worst = 0.0
    for c0 in range(0, ch, 296):
        xf = x[c0:c0 + 296, :frames * nf].reshape(296, frames, nf).double() * w
        e_t = (xf * xf).sum(-1)
        m = mag[c0:c0 + 296].double()
        e_f = (m[..., 0] ** 2 + 2.0 * (m[..., 1:-1] ** 2).sum(-1) + m[..., -1] ** 2) / nf
        worst = max(worst, float(((e_f - e_t).abs() / e_t).max()))
        del xf, m
    assert worst <= 1e-5, worst
    for c, f in ((0, 0), (1, 116), (1183, 58), (2367, 116), (1500, 1)):
        ref = np.abs(np.fft.rfft(x[c, f * nf:(f + 1) * nf].cpu().numpy().astype(np.float64) * wh))
        assert o.rel_err(mag[c, f].cpu().numpy(), ref) <= TOL_F32_FFT, (c, f)


def test_c4_many_transforms_parseval_and_spot_frames(pk, torch_cuda):
    """A quarter of C4 (1024 channels x 2^20 samples: 16384 transforms of 2^16 points, float32): every persistent CTA of the
    three-pass kernel reuses its workspace slot and row buffers over a hundred times.  Parseval of every frame against the
    windowed time-domain energy (float64 on the device), and frames from the first, middle and last CTAs / iterations
    against numpy."""
    torch = torch_cuda
    ch, n, nf = 1024, 1 << 20, 65536
    x = torch.empty((ch, n), dtype=torch.float32, device="cuda")
    pk.generate_uniform(x, 3, -1.0, 1.0)
    mag = pk.FftPlan(nf, np.float32, hann=True).magnitudes(x)
    assert tuple(mag.shape) == (ch, n // nf, nf // 2 + 1)
    wh = o.hann_symmetric(nf)
    w = torch.as_tensor(wh, device="cuda", dtype=torch.float64)
    worst = 0.0
    for c0 in range(0, ch, 128):
        xf = x[c0:c0 + 128].view(128, n // nf, nf).double() * w
        e_t = (xf * xf).sum(-1)
        m = mag[c0:c0 + 128].double()
        e_f = (m[..., 0] ** 2 + 2.0 * (m[..., 1:-1] ** 2).sum(-1) + m[..., -1] ** 2) / nf
        worst = max(worst, float(((e_f - e_t).abs() / e_t).max()))
        del xf, m
    assert worst <= 1e-5, worst
    for c, f in ((0, 0), (511, 7), (700, 3), (1023, 15)):
        ref = np.abs(np.fft.rfft(x[c, f * nf:(f + 1) * nf].cpu().numpy().astype(np.float64) * wh))
        assert o.rel_err(mag[c, f].cpu().numpy(), ref) <= TOL_F32_FFT, (c, f)


def test_eq_tensor_form_on_narrow_batches_overlapping_slices(pk, torch_cuda, monkeypatch):
    """C2-shaped EQ (1024 channels x 10 s @ 48 kHz) fills the GPU by cutting the time axis into independent slices that
    start plan.warm_chunks() chunks early from a zero state (csrc/eq_mma.cu).  Against the float64 oracle on whole
    channels, against the exact single-slice form (bounded by float32 rounding), impulse / step responses across the
    slice boundaries, and the conditions under which the form is NOT taken (in place; short signals)."""
    torch = torch_cuda
    for gains in (C1_GAINS, (15,) * 6):
        gd = gains_dict(gains)
        plan = pk.EqPlan.from_gains(48000, gd, np.float32)
        ch, n = 1024, 480000
        assert plan.warm_chunks() > 0
        kind = plan.kernel_kind(ch, n)
        if gains == C1_GAINS:
            assert kind == "tensor"
        x = torch.empty((ch, n), dtype=torch.float32, device="cuda")
        pk.generate_uniform(x, 2, -0.25, 0.25)
        x[7, :] = 0.0
        x[7, 300000] = 1.0                                     # an impulse late in the signal
        x[8, :] = 0.0
        x[8, 100:] = 0.5                                       # a step: the slow poles ring across every boundary
        monkeypatch.setenv("DSPB200_EQ_FORCE_MMA", "1")
        z = plan.run(x)
        monkeypatch.setenv("DSPB200_EQ_NO_OVERLAP", "1")       # the same kernel, one slice per group, exact state
        z1 = plan.run(x)
        monkeypatch.delenv("DSPB200_EQ_NO_OVERLAP", raising=False)
        monkeypatch.delenv("DSPB200_EQ_FORCE_MMA", raising=False)
        # float32 rounding of two different state paths, at the scale of the signal BEFORE the clip (+15 dB on every band
        # lifts it by up to 5.6 x per band): 2e-6 with the C1 gains, 5e-6 with all +15 dB
        assert float((z - z1).abs().max()) <= 2e-5
        for c in (0, 7, 8, 1023):
            ref = o.equalizer(x[c].cpu().numpy().astype(np.float64), 48000, gd)
            assert o.full_scale_err(z[c].cpu().numpy(), ref) <= TOL_F32_EQ, (gains, c)
        # in place there is no overlap form (a slice would re-read samples its predecessor has overwritten)
        zi = x.clone()
        plan.run(zi, out=zi)
        for c in (0, 8):
            ref = o.equalizer(x[c].cpu().numpy().astype(np.float64), 48000, gd)
            assert o.full_scale_err(zi[c].cpu().numpy(), ref) <= TOL_F32_EQ
    plan = pk.EqPlan.from_gains(48000, gains_dict(C1_GAINS), np.float32)
    assert plan.kernel_kind(1024, 20000) == "scan"            # too short to pay for the warm-up
    assert plan.kernel_kind(32, 480000) == "scan"             # a quarter of one group
    assert plan.kernel_kind(4096, 2880000) == "tensor"        # C3 slice


def test_wave_scheduler_matches_wave_by_wave(pk, torch_cuda):
    """pkg.WaveScheduler (the C5 job's waves, next wave produced on a side stream into the other x buffer): every wave's
    z and spectra equal a plain Chain.run on the same clips, ragged last wave included; the producer sees each wave once."""
    torch = torch_cuda
    gd = gains_dict(C1_GAINS)
    chain = pk.Chain(3, 2, 44100, gd, n_fft=1024, dtype=np.float32)
    n_in, per = 9000, 6
    waves = [(0, 6), (6, 6), (12, 6), (18, 3)]
    ws = pk.WaveScheduler(chain, per, n_in, "cuda")
    seen, got = [], {}

    def produce(xv, first, count):
        seen.append((first, count))
        pk.generate_uniform(xv, 9, -0.5, 0.5, first_channel=first)

    def consume(zv, mv, first, count):
        got[first] = (zv.clone(), mv.clone())

    for _ in range(2):                     # twice: the buffers and events are reused across jobs
        seen.clear()
        got.clear()
        ws.run(waves, produce, consume)
        torch.cuda.synchronize()
        assert seen == waves
        for first, count in waves:
            x = torch.empty((count, n_in), dtype=torch.float32, device="cuda")
            pk.generate_uniform(x, 9, -0.5, 0.5, first_channel=first)
            _, z, mag = chain.run(x)
            assert torch.equal(got[first][0], z) and torch.equal(got[first][1], mag), first
    ref = o.chain(o.synthetic_clips(1, n_in, 9, -0.5, 0.5, first_channel=19)[0].astype(np.float64), 44100, 2, 3, gd,
                  n_fft=1024, n_frames=4)
    assert o.full_scale_err(got[18][0][1].cpu().numpy(), ref[1]) <= TOL_F32_EQ
    assert o.rel_err(got[18][1][1, :4].cpu().numpy(), ref[2]) <= 1e-4


def test_chain_c2_shaped_wave_uses_the_sliced_tensor_eq(pk, torch_cuda, monkeypatch):
    """A 1024-clip wave (C2 / C5 shape, 8 channel groups): too narrow for the fused SRC->EQ kernel, so the chain runs
    three kernels -- and gives the resampler's output a scratch buffer, so that the equaliser takes its tensor-core form
    on overlapping time slices instead of the in-place scan kernel.  Oracle on two clips; same z as the in-place route
    to float32 rounding; spectra tied to z by Parseval."""
    torch = torch_cuda
    gd = gains_dict(C1_GAINS)
    ch = pk.Chain(160, 147, 44100, gd, n_fft=4096, dtype=np.float32)
    clips, n_in = 1024, 441000
    assert ch.kernel_kind(clips, n_in) == "cascade" and ch.eq.kernel_kind(clips, 480000) == "tensor"
    need = C.c_size_t()
    from dsp_audio_project_b200 import _lib
    _lib.check(_lib.load().dspb200_chain_workspace_bytes(ch.src._h, ch.eq._h, ch.fft._h, clips, n_in, 0, C.byref(need)))
    assert need.value >= clips * 480000 * 4                   # the y scratch
    x = torch.empty((clips, n_in), dtype=torch.float32, device="cuda")
    pk.generate_uniform(x, 4, -0.5, 0.5)
    before = _lib.launch_count()
    _, z, mag = ch.run(x)
    torch.cuda.synchronize()
    assert _lib.launch_count() - before == 3
    for c in (0, 1023):
        yo, zo, mo, _ = o.chain(x[c].cpu().numpy().astype(np.float64), 44100, 147, 160, gd, n_fft=4096, n_frames=2)
        assert o.full_scale_err(z[c].cpu().numpy(), zo) <= TOL_F32_EQ
        assert o.rel_err(mag[c, :2].cpu().numpy(), mo) <= 1e-4
    monkeypatch.setenv("DSPB200_EQ_NO_OVERLAP", "1")          # in place, scan kernel
    _, z2, _ = ch.run(x)
    monkeypatch.delenv("DSPB200_EQ_NO_OVERLAP")
    assert float((z - z2).abs().max()) <= 5e-6
