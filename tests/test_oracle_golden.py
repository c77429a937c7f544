"""Pin the CPU oracle against the golden vectors produced by the real reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import gains_dict
from oracle import dsp_oracle as o


def test_design_taps(golden_design):
    g = golden_design
    i = 0
    while f"taps_{i}" in g:
        wc, n = g[f"taps_{i}_args"]
        h = o.sinc_lowpass_taps(float(wc), int(n))
        assert h.shape == g[f"taps_{i}"].shape
        assert np.max(np.abs(h - g[f"taps_{i}"])) <= 1e-15
        i += 1
    assert i >= 6


def test_design_biquads(golden_design):
    for row in golden_design["biquads"]:
        fc, fs, gdb = row[:3]
        b, a = o.peaking_biquad(fc, fs, gdb)
        assert np.max(np.abs(b - row[3:6])) <= 1e-15
        assert np.max(np.abs(a - row[6:9])) <= 1e-15


def test_src_closed_and_faithful_forms(golden_src):
    g = golden_src
    for idx, (L, M, N, fs_new, fs) in enumerate(g["cases"]):
        x, y = g[f"x_{idx}"], g[f"y_{idx}"]
        ya, fa = o.resample_reference_form(x, int(fs), int(M), int(L))
        yb, fb = o.resample_closed_form(x, int(fs), int(M), int(L))
        assert fa == fb == fs_new
        assert len(ya) == len(yb) == len(y)
        assert o.src_geometry(int(N), int(L), int(M))[3] == len(y) or (L == 1 and M == 1)
        assert np.max(np.abs(np.asarray(ya, dtype=np.float64) - y)) <= 1e-14
        assert np.max(np.abs(np.asarray(yb, dtype=np.float64) - y)) <= 1e-14
        if L == 1 and M == 1:
            assert ya is x and yb is x


def test_convolve_loop_matches_numpy():
    rng = np.random.default_rng(5)
    for na, nv in [(50, 7), (7, 50), (1, 1), (30, 30), (12, 121)]:
        a, v = rng.normal(size=na), rng.normal(size=nv)
        assert np.max(np.abs(o.convolve_same_direct(a, v) - np.convolve(a, v, "same"))) <= 1e-13
    with pytest.raises(ValueError):
        o.convolve_same_direct(np.zeros(0), np.ones(3))


def test_eq_cases(golden_eq):
    g = golden_eq
    for idx, row in enumerate(g["cases"]):
        fs, gs = row[0], row[1:]
        x, z = g[f"x_{idx}"], g[f"z_{idx}"]
        out = o.equalizer(x, fs, gains_dict(gs))
        assert (out is x) == bool(g[f"alias_{idx}"])
        assert out.dtype == z.dtype
        assert np.max(np.abs(np.asarray(out, dtype=np.float64) - z)) <= 1e-13


def test_eq_dict_order_and_unknown_band(golden_eq):
    g = golden_eq
    x = g["x_unknown"]
    out = o.equalizer(x, 48000, {"Sub-Bass": 4, "Air": -7, "Brilliance": 5})
    assert np.max(np.abs(out - g["z_unknown"])) <= 1e-13
    bands = ["Sub-Bass", "Bass", "Low Mids", "High Mids", "Presence", "Brilliance"]
    rev = gains_dict((6, -3, 4, -6, 3, -9)[::-1], bands[::-1])
    assert np.max(np.abs(o.equalizer(x, 48000, rev) - g["z_reversed"])) <= 1e-13


def test_df2t_loop_matches_reference_engine(golden_eq):
    g = golden_eq
    b, a = g["ba_lf"][:3], g["ba_lf"][3:]
    y = o.lfilter_df2t_loop(b, a, g["x_lf"])
    assert o.rel_err(y, g["y_lf"]) <= 1e-12
    assert o.rel_err(o.difference_equation(g["x_lf"], b, a), g["y_lf"]) <= 1e-15


def test_fft(golden_fft):
    g = golden_fft
    for n in (1, 2, 4, 8, 16, 64, 256, 1024, 2048, 4096):
        for kind in ("r", "c"):
            x, X = g[f"x{kind}_{n}"], g[f"X{kind}_{n}"]
            out = np.asarray(o.fft_dit_recursive(x))
            assert out.shape == X.shape and out.dtype == X.dtype
            assert np.max(np.abs(out - X)) <= 1e-12 * max(1.0, np.max(np.abs(X)))
            if n >= 2:
                assert o.rel_err(out, np.fft.fft(x)) <= 1e-14


def test_fft_65536(golden_fft):
    x = golden_fft["xr_65536"].astype(np.float64)
    mag = np.abs(o.fft_dit_recursive(x))[:32769]
    assert o.rel_err(mag, golden_fft["Xr_65536_mag"]) <= 1e-14


def test_spectrum(golden_spectrum):
    g = golden_spectrum
    for n in g["ok_lens"]:
        x = g[f"x_{n}"]
        with np.errstate(all="ignore"):
            f, m = o.magnitude_spectrum(x, 48000)
        assert f.shape == g[f"f_{n}"].shape and m.shape == g[f"m_{n}"].shape
        assert np.allclose(f, g[f"f_{n}"], rtol=0, atol=1e-9)
        assert np.allclose(m, g[f"m_{n}"], rtol=0, atol=1e-12, equal_nan=True)
    for n in g["valueerror_lens"]:
        with pytest.raises(ValueError):
            o.magnitude_spectrum(np.zeros(int(n)), 48000)
    assert len(g["valueerror_lens"]) >= 3


def test_chain_c1(golden_chain):
    g = golden_chain
    x = g["x"]
    y, z, mags, fs2 = o.chain(x, 44100, 2, 3, gains_dict((6, -3, 4, -6, 3, -9)), n_fft=4096, n_frames=2)
    assert fs2 == int(g["fs2"])
    assert np.max(np.abs(y[:4096] - g["y_head"])) <= 1e-13
    assert np.max(np.abs(y[-4096:] - g["y_tail"])) <= 1e-13
    assert np.max(np.abs(z[:4096] - g["z_head"])) <= 1e-12
    assert np.max(np.abs(z[-4096:] - g["z_tail"])) <= 1e-12
    sums = np.array([np.sum(z), np.sum(np.abs(z)), np.sum(z * z)])
    assert np.allclose(sums, g["z_sum"], rtol=1e-11)
    assert np.allclose([np.sum(y), np.sum(np.abs(y)), np.sum(y * y)], g["y_sum"], rtol=1e-11)
    mid = len(z) // 2
    assert len(x) == 30 * 44100 and len(z) == 45 * 44100          # C1 is 30 s
    assert np.max(np.abs(y[mid:mid + 4096] - g["y_mid"])) <= 1e-13
    assert np.max(np.abs(z[mid:mid + 4096] - g["z_mid"])) <= 1e-12
    frame = o.frame_magnitudes(z, 4096, offset=mid)[0]
    assert o.rel_err(frame, g["mag4096"]) <= 1e-10
    for s0, ref_mag in zip(g["frame_starts"], g["frames_mag"]):
        assert o.rel_err(o.frame_magnitudes(z, 4096, offset=int(s0), n_frames=1)[0], ref_mag) <= 1e-10
    assert o.rel_err(mags[0], g["frames_mag"][0]) <= 1e-10
    f, m = o.magnitude_spectrum(z[:100000], fs2)
    assert o.rel_err(m, g["m_app"]) <= 1e-10 and np.allclose(f, g["f_app"])


def test_known_answers():
    # DC gain of the resampler away from the edges is 1 (sum h * L / L)
    y, _ = o.resample_closed_form(np.ones(400), 48000, 2, 3)
    assert np.max(np.abs(y[100:-100] - 1.0)) < 2e-3
    # peaking biquad: unit DC gain, g dB at fc
    b, a = o.peaking_biquad(1000, 48000, 6.0)
    assert abs(b.sum() / a.sum() - 1.0) < 1e-12
    w = np.exp(-1j * 2 * np.pi * 1000 / 48000)
    H = (b[0] + b[1] * w + b[2] * w * w) / (a[0] + a[1] * w + a[2] * w * w)
    assert abs(20 * np.log10(abs(H)) - 6.0) < 1e-9
    # FFT of an impulse is all ones; Parseval
    imp = np.zeros(64); imp[0] = 1
    assert np.max(np.abs(o.fft_dit_recursive(imp) - 1)) < 1e-15
    x = np.random.default_rng(0).normal(size=256)
    X = o.fft_dit_recursive(x)
    assert abs(np.sum(np.abs(X) ** 2) / 256 - np.sum(x * x)) < 1e-9


def test_loader_front_end_matches_reference():
    from conftest import load_golden
    g = load_golden("loader.npz")
    for name in ("stereo", "mono", "quad", "tiny", "silence"):
        out = o.load_mono_normalize(g[f"in_{name}"])
        assert out.dtype == np.float32 and np.array_equal(out, g[f"out_{name}"]), name


def test_app_helpers_pinned_by_the_reference_statements(golden_app):
    """app.py:207-208 (f > 0.5 mask, 20 log10(mag + 1e-12)) and app.py:349-354 (nan_to_num, peak normalise, * 32767,
    int16) as executed from the reference's own statements by tests/golden/make_golden.py (ast extraction)."""
    g = golden_app
    assert dict(zip(g["app_line_names"].tolist(), g["app_lines"].tolist())) == {
        "mask_in": 207, "db_in": 208, "y_final_audio": 349, "peak": 350, "peak_if": 351, "int16_expr": 354}
    for tag in ("c1", "edge"):
        mask, db = o.spectrum_db_masked(g[f"f_{tag}"], g[f"m_{tag}"])
        assert np.array_equal(mask, g[f"mask_{tag}"])
        assert np.array_equal(db, g[f"db_{tag}"])
        assert np.array_equal(o.spectrum_db(g[f"m_{tag}"])[mask], g[f"db_{tag}"])
    for name in g["pcm_names"].tolist():
        with np.errstate(all="ignore"):
            out = o.pcm16_export(g[f"z_{name}"])
        assert out.dtype == np.int16 and np.array_equal(out, g[f"pcm_{name}"]), name


def test_export_helpers_known_answers():
    z = np.array([0.5, -1.0, np.nan, 0.25])
    assert np.array_equal(o.pcm16_export(z), np.array([16383, -32767, 0, 8191], dtype=np.int16))
    assert np.array_equal(o.pcm16_export(np.zeros(3)), np.zeros(3, dtype=np.int16))
    assert abs(o.spectrum_db(np.array([1.0]))[0]) < 1e-9 and abs(o.spectrum_db(np.array([0.0]))[0] + 240) < 1e-9
