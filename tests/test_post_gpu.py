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


def test_pcm16_matches_the_reference_statements(env):
    """int16 export against app.py:349-354 as executed from the reference's own statements
    (tests/golden/app_helpers.npz, made by make_golden.py's ast extraction)."""
    torch, pk = env
    g = load_golden("app_helpers.npz")
    for name in g["pcm_names"].tolist():
        z = g[f"z_{name}"]
        ref = g[f"pcm_{name}"].astype(np.int32)
        out, _ = pk.to_pcm16(torch.as_tensor(z[None, :], device="cuda"))
        got = out[0].cpu().numpy().astype(np.int32)
        if z.dtype == np.float64:
            assert np.array_equal(got, ref), name                    # float64 kernel: bit exact
        else:
            assert np.max(np.abs(got - ref)) <= 1, name              # float32 kernel: +-1 LSB
        z32 = z.astype(np.float32)
        out32, _ = pk.to_pcm16(torch.as_tensor(z32[None, :], device="cuda"))
        assert np.max(np.abs(out32[0].cpu().numpy().astype(np.int32) - ref)) <= 1, name


def test_db_and_mask_match_the_reference_statements(env):
    """The C1 spectrum the reference computed (calcular_espectro_magnitud of z[:100000]) through app.py:207-208:
    the product's fused dB store and the f > 0.5 mask as a suffix view."""
    torch, pk = env
    from conftest import c1_input, gains_dict
    from modules import dsp_core as dc
    g = load_golden("app_helpers.npz")
    gc1 = dict(load_golden("chain_c1.npz"))
    x = c1_input(gc1)
    y, fs2 = dc.conversion_tasa_muestreo(x, 44100, 2, 3)
    z = dc.sistema_ecualizador(y, fs2, gains_dict((6, -3, 4, -6, 3, -9)))[:100000]
    f, m = dc.calcular_espectro_magnitud(z, fs2)
    assert np.allclose(f, g["f_c1"]) and o.rel_err(m, g["m_c1"]) <= 1e-10
    plan = pk.FftPlan(2048, np.float64, hann=True, db=True)
    k0 = plan.first_bin_above(fs2)
    assert np.array_equal(np.arange(plan.bins) >= k0, g["mask_c1"])
    # the reference takes the window at len // 2 (dsp_core.py:76-78)
    db = plan.magnitudes(torch.as_tensor(z[None, :], device="cuda"), offset=len(z) // 2, n_frames=1)[0, 0]
    got = db.cpu().numpy()[k0:]
    assert got.shape == g["db_c1"].shape
    assert np.max(np.abs(got - g["db_c1"])) <= 1e-7
    plan32 = pk.FftPlan(2048, np.float32, hann=True, db=True)
    db32 = plan32.magnitudes(torch.as_tensor(z[None, :].astype(np.float32), device="cuda"), offset=len(z) // 2,
                             n_frames=1)[0, 0].cpu().numpy()[k0:]
    big = g["m_c1"][k0:] > 1e-3 * g["m_c1"].max()
    assert np.max(np.abs(db32[big] - g["db_c1"][big])) <= 2e-3
    # the mask rule at other rates / sizes, against the reference's comparison on rfftfreq
    for fs, n in ((48000, 2048), (8, 64), (1.0, 4), (0.9, 2), (44100, 1 << 16), (100.0, 4096)):
        p = pk.FftPlan(n, np.float32)
        fr = np.fft.rfftfreq(n, 1 / fs)
        assert np.array_equal(np.arange(p.bins) >= p.first_bin_above(fs), fr > 0.5), (fs, n)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_chain_host_export_form(env, dt):
    """dspb200_chain_host_pcm16_*: the host cascade with z leaving as app.py:349-354's int16 signal and the spectra in
    dB (app.py:207-210).  Checked against the plain host form of the same chain plus the oracle's export helpers
    (themselves pinned by the reference's statements, tests/test_oracle_golden.py)."""
    torch, pk = env
    from conftest import gains_dict
    gd = gains_dict((6, -3, 4, -6, 3, -9))
    rng = np.random.default_rng(11)
    x = rng.uniform(-0.9, 0.9, (37, 9008)).astype(dt)       # 37 clips: two 32-clip slabs, the second ragged; n_out = 13512 = 8 k: dense int16 rows
    x[5] = 0.0                                              # a silent clip: peak 0, no division
    lin = pk.Chain(3, 2, 44100, gd, n_fft=1024, dtype=dt)
    dbc = pk.Chain(3, 2, 44100, gd, n_fft=1024, dtype=dt, db=True)
    z, mag = lin.run_host(x)
    q, peaks, db = dbc.run_host_pcm16(x)
    assert q.dtype == np.int16 and q.shape == z.shape and db.shape == mag.shape
    assert np.array_equal(peaks, np.max(np.abs(z), axis=1))
    ref_q = np.stack([o.pcm16_export(z[c].astype(np.float64) if dt == np.float64 else z[c]) for c in range(len(z))])
    diff = np.abs(q.astype(np.int32) - ref_q.astype(np.int32))
    assert diff.max() <= (0 if dt == np.float64 else 1)
    assert np.all(q[5] == 0)
    ref_db = o.spectrum_db(mag.astype(np.float64))
    big = mag > 1e-4 * mag.max()
    assert np.max(np.abs(db[big] - ref_db[big])) <= (1e-8 if dt == np.float64 else 2e-3)
    # odd row lengths (n_out not a multiple of 8) take the pitched device rows
    x2 = np.ascontiguousarray(x[:3, :9000 - 7])
    z2, _ = lin.run_host(x2)
    q2, _, _ = dbc.run_host_pcm16(x2)
    ref2 = np.stack([o.pcm16_export(z2[c].astype(np.float64) if dt == np.float64 else z2[c]) for c in range(3)])
    assert np.abs(q2.astype(np.int32) - ref2.astype(np.int32)).max() <= (0 if dt == np.float64 else 1)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_generate_uniform_matches_its_numpy_twin(env, dt):
    """dspb200_generate_uniform_*: the counter-based clip generator of the throughput configurations,
    bit for bit against oracle.synthetic_clips; a rank's block equals the same rows of the whole batch."""
    torch, pk = env
    tdt = torch.float32 if dt == np.float32 else torch.float64
    for n in (1001, 1000, 7, 4096):
        x = torch.empty((6, n), dtype=tdt, device="cuda")
        pk.generate_uniform(x, 4, -0.5, 0.5)
        ref = o.synthetic_clips(6, n, 4, -0.5, 0.5, dtype=dt)
        assert np.array_equal(x.cpu().numpy(), ref), n
        part = torch.empty((2, n), dtype=tdt, device="cuda")
        pk.generate_uniform(part, 4, -0.5, 0.5, first_channel=3)
        assert np.array_equal(part.cpu().numpy(), ref[3:5]), n
    # a padded (strided) destination, other bounds
    buf = torch.full((3, 520), 9.0, dtype=tdt, device="cuda")
    pk.generate_uniform(buf[:, :513], 77, -1.0, 0.25)
    ref = o.synthetic_clips(3, 513, 77, -1.0, 0.25, dtype=dt)
    got = buf.cpu().numpy()
    assert np.array_equal(got[:, :513], ref) and np.all(got[:, 513:] == 9.0)
    big = torch.empty((4, 200000), dtype=tdt, device="cuda")
    pk.generate_uniform(big, 5, -0.5, 0.5)
    b = big.cpu().numpy().astype(np.float64)
    assert b.min() >= -0.5 and b.max() < 0.5 and abs(b.mean()) < 2e-3 and abs(b.std() - 0.2887) < 2e-3
    assert abs(np.corrcoef(b[0, :-1], b[0, 1:])[0, 1]) < 0.01          # the two halves of a hash are independent
