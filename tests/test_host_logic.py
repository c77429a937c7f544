"""CPU-only tests: the C ABI loads and exports every declared symbol, host-side
design code matches the golden vectors, drop-in bypass/argument semantics, and
the product fails loudly (no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, gains_dict
from oracle import dsp_oracle as o


@pytest.fixture(scope="module")
def lib():
    from dsp_audio_project_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        _lib.build_library()
    return _lib.load()


def test_header_symbols_exported(lib):
    from dsp_audio_project_b200 import _lib
    header = open(os.path.join(ROOT, "include", "dspb200.h")).read()
    declared = set(re.findall(r"\b(dspb200_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found in include/dspb200.h"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared but not exported"
    assert lib.dspb200_version() == 100


def test_design_taps_match_reference(lib, golden_design):
    from modules.dsp_core import generar_respuesta_impulso_sinc
    g = golden_design
    i = 0
    while f"taps_{i}" in g:
        wc, n = g[f"taps_{i}_args"]
        h = generar_respuesta_impulso_sinc(float(wc), int(n))
        assert h.shape == g[f"taps_{i}"].shape
        assert np.max(np.abs(h - g[f"taps_{i}"])) <= 1e-15
        i += 1
    cap = 6401
    buf = (C.c_double * cap)()
    n = C.c_int()
    assert lib.dspb200_design_src_filter(160, 147, buf, cap, C.byref(n)) == 0 and n.value == 6401
    assert np.max(np.abs(np.array(buf[:]) - o.src_filter(160, 147))) <= 1e-15


def test_design_biquads_match_reference(lib, golden_design):
    from modules.dsp_core import disenar_coeficientes_diferencias
    for row in golden_design["biquads"]:
        b, a = disenar_coeficientes_diferencias(*row[:3])
        assert np.max(np.abs(b - row[3:6])) <= 1e-15 and np.max(np.abs(a - row[6:9])) <= 1e-15


def test_src_geometry_matches_oracle(lib):
    from dsp_audio_project_b200 import src_geometry
    for L, M, n in [(3, 2, 500), (3, 2, 10), (160, 147, 441000), (5, 7, 1), (8, 8, 100), (1, 8, 3)]:
        T, P, _, n_out = o.src_geometry(n, L, M)
        assert src_geometry(L, M, n) == (T, P, n_out)
    assert src_geometry(160, 147, 441000)[2] == 480000


def test_section_selection_rules(lib):
    from dsp_audio_project_b200 import select_sections
    for fs in (48000, 8000, 66150, 44100):
        for gs in [(6, -3, 4, -6, 3, -9), (0.1,) * 6, (0.05, 0, -0.09, 0, 0, 0), (0.11, 0, 0, 0, 0, 0)]:
            gd = gains_dict(gs)
            bypass, secs = select_sections(fs, gd)
            assert bypass == all(abs(g) < 0.1 for g in gs)
            assert secs == o.eq_active_sections(fs, gd)
    _, secs = select_sections(48000, {"Air": 3.0, "Bass": -2})
    assert secs == [(1000.0, 3.0), (150.0, -2.0)]


def test_state_space_sections_reproduce_lfilter(lib):
    """The state-space form the EQ kernel runs (plan describe) reproduces the
    DF2T difference equation on the CPU, for complex- and real-pole sections."""
    from dsp_audio_project_b200 import EqPlan
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, 4000)
    secs = [(40.0, 15.0), (150.0, -15.0), (1000.0, -12.0412), (10000.0, 6.0), (3000.0, -13.0)]
    plan = EqPlan(48000, secs, dtype=np.float64, clip=False)
    ss = plan.describe()
    assert ss.shape == (5, 9)
    y = x.copy()
    ref = x.copy()
    for (fc, g), (a00, a01, a10, a11, b0, b1, c0, c1, d) in zip(secs, ss):
        b, a = o.peaking_biquad(fc, 48000, g)
        ref = o.difference_equation(ref, b, a)
        q0 = q1 = 0.0
        out = np.empty_like(y)
        for n in range(len(y)):
            out[n] = c0 * q0 + c1 * q1 + d * y[n]
            q0, q1 = a00 * q0 + a01 * q1 + b0 * y[n], a10 * q0 + a11 * q1 + b1 * y[n]
        y = out
    assert o.rel_err(y, ref) <= 1e-11


def _tf32_trunc(v):
    return (np.asarray(v, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def _tf32_round(v):
    return ((np.asarray(v, np.float32).view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


@pytest.mark.parametrize("gains", [(6, -3, 4, -6, 3, -9), (15,) * 6, (-15,) * 6, (-12.0412, 15, -12.5, 6, 0, -3),
                                   (0, 0, 0, 0, 0, 12)])
def test_eq_chunk_system_reproduces_the_cascade(lib, gains):
    """The tensor-core form of the EQ (csrc/eq_mma.cu) multiplies a chunk system built on the host in float64:
    z = T x + O s, s' = Phi s + K x over 96 samples.  Stepping it in numpy must reproduce the reference cascade
    (lfilter per active band, clip once, dsp_core.py:216-254); stepping it with the kernel's arithmetic (three-
    product TF32 split of [T; K] x and of O s, fp32 state update) must stay inside the fp32 EQ bound."""
    import dsp_audio_project_b200 as pk
    gd = gains_dict(gains)
    plan = pk.EqPlan.from_gains(48000, gd, np.float32)
    T, K, O, Phi = plan.chunk_system()
    L, S = T.shape[0], Phi.shape[0]
    assert L == 96 and S == 2 * plan.n_sections and np.allclose(T, np.tril(T))
    assert np.allclose(np.linalg.norm(K, axis=1), 1.0)             # balanced states
    x = np.random.default_rng(11).uniform(-0.5, 0.5, L * 40 + 17)
    n = x.size
    xp = np.zeros(-(-n // L) * L)
    xp[:n] = x
    X = xp.reshape(-1, L)
    ref = o.equalizer(x, 48000, gd)
    # float64
    s, out = np.zeros(S), []
    for xk in X:
        out.append(T @ xk + O @ s)
        s = Phi @ s + K @ xk
    assert np.max(np.abs(np.clip(np.concatenate(out)[:n], -1, 1) - ref)) <= 1e-10
    # the kernel's arithmetic
    M = np.vstack([T, K]).astype(np.float32)
    Mh = _tf32_round(M)
    Ml = _tf32_trunc(M - Mh)
    Oh = _tf32_round(O.astype(np.float32))
    Ol = _tf32_trunc(O.astype(np.float32) - Oh)
    P32 = Phi.astype(np.float32)
    s, out = np.zeros(S, np.float32), []
    for xk in X.astype(np.float32):
        xh = _tf32_trunc(xk)
        xl = _tf32_trunc(xk - xh)
        acc = Mh.astype(np.float64) @ xh + Ml.astype(np.float64) @ xh + Mh.astype(np.float64) @ xl
        s1 = _tf32_trunc(s)
        s2 = _tf32_trunc(s - s1)
        s3 = _tf32_trunc(s - s1 - s2)
        acc[:L] += Oh.astype(np.float64) @ (s1.astype(np.float64) + s2 + s3) + Ol.astype(np.float64) @ s1
        acc = acc.astype(np.float32)
        out.append(acc[:L])
        s = (P32 @ s + acc[L:]).astype(np.float32)
    assert np.max(np.abs(np.clip(np.concatenate(out)[:n], -1, 1) - ref)) <= 1e-4


def test_eq_chunk_system_section_counts(lib):
    import dsp_audio_project_b200 as pk
    centres = [60.0, 250.0, 700.0, 1500.0, 3200.0, 6000.0, 9000.0, 14000.0, 18000.0]
    gains = [5.0, -4.0, 7.5, -6.0, 3.0, -9.0, 12.0, -2.5, 4.0]
    x = np.random.default_rng(12).uniform(-0.4, 0.4, 96 * 30)
    for ns in range(0, 10):
        plan = pk.EqPlan(48000, list(zip(centres[:ns], gains[:ns])), np.float32, clip=False)
        cs = plan.chunk_system()
        if ns == 0 or ns > 8:
            assert cs is None                                      # no tensor form: clip-only copy / too many states
            continue
        T, K, O, Phi = cs
        ref = x.copy()
        for fc, g in zip(centres[:ns], gains[:ns]):
            ref = o.difference_equation(ref, *o.peaking_biquad(fc, 48000, g))
        s, out = np.zeros(2 * ns), []
        for xk in x.reshape(-1, 96):
            out.append(T @ xk + O @ s)
            s = Phi @ s + K @ xk
        assert np.max(np.abs(np.concatenate(out) - ref)) <= 1e-10 * max(1.0, np.max(np.abs(ref))), ns


def test_bypass_and_argument_semantics(lib):
    from modules import dsp_core as dc
    x = np.arange(8, dtype=np.float32)
    y, fs = dc.conversion_tasa_muestreo(x, 44100, 1, 1)
    assert y is x and fs == 44100
    assert dc.sistema_ecualizador(x, 48000, gains_dict((0.05,) * 6)) is x
    assert dc.fft_diezmado_en_tiempo(x[:1]) is not None and len(dc.fft_diezmado_en_tiempo(x[:0])) == 0
    for bad in (3, 6, 12, 100):
        with pytest.raises(ValueError):
            dc.fft_diezmado_en_tiempo(np.zeros(bad))
    for n in (2049, 3000, 4094):
        with pytest.raises(ValueError):
            dc.calcular_espectro_magnitud(np.zeros(n), 48000)
    with pytest.raises(ValueError):
        dc.conversion_tasa_muestreo(np.zeros(0), 44100, 2, 3)
    with pytest.raises(ValueError):
        dc.conversion_tasa_muestreo(np.zeros(10), 44100, 0, 3)
    x100, fs = dc.cargar_senal_audio("/nonexistent.wav")
    assert x100.shape == (100,) and x100.dtype == np.float32 and fs == 44100


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from dsp_audio_project_b200 import _lib
    from modules import dsp_core as dc
    with pytest.raises(_lib.Dspb200Error):
        dc.conversion_tasa_muestreo(np.zeros(100), 44100, 2, 3)
    with pytest.raises(_lib.Dspb200Error):
        dc.sistema_ecualizador(np.zeros(100), 48000, gains_dict((6, 0, 0, 0, 0, 0)))
    with pytest.raises(_lib.Dspb200Error):
        dc.calcular_espectro_magnitud(np.zeros(100), 48000)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "dsp_audio_project_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no CPU fallback", ""), f


@pytest.mark.parametrize("n_fft", [1024, 4096, 16384])
def test_hann_by_angle_addition_matches_the_reference_window(n_fft):
    """The fp32 magnitude kernel forms the reference's symmetric Hann (dsp_core.py:85-87) as
    1/2 + A_t cos(u D) + B_t sin(u D) for the samples n = 2(t + uQ) + {0,1} a thread owns (csrc/fft.cu,
    plan_build); the same arithmetic in float32 stays within 2e-7 of the float64 window."""
    m = n_fft // 2
    q = m // 16
    step = 2.0 * np.pi / (n_fft - 1)
    t = np.arange(q)[:, None, None]
    h = np.arange(2)[None, None, :]
    u = np.arange(16)[None, :, None]
    phi = step * (2 * t + h)
    a = (-0.5 * np.cos(phi)).astype(np.float32)
    b = (0.5 * np.sin(phi)).astype(np.float32)
    cc = np.cos(step * 2 * q * u).astype(np.float32)
    ss = np.sin(step * 2 * q * u).astype(np.float32)
    w = (a * cc + (b * ss + np.float32(0.5))).astype(np.float32)       # two fused multiply-adds in the kernel
    n = (2 * (t + u * q) + h).reshape(-1)
    ref = o.hann_symmetric(n_fft)
    assert sorted(n.tolist()) == list(range(n_fft))
    assert np.max(np.abs(w.reshape(-1).astype(np.float64) - ref[n])) <= 2e-7


@pytest.mark.parametrize("fs,gains", [(48000, (6, -3, 4, -6, 3, -9)), (48000, (15,) * 6), (48000, (-15,) * 6),
                                      (48000, (0, 0, 0, 0, 0, 12)), (96000, (6, -3, 4, -6, 3, -9)), (8000, (15,) * 6)])
def test_eq_warm_chunks_bound_holds(lib, fs, gains):
    """Narrow batches run the tensor-core EQ on independent, overlapping time slices: a slice starts
    plan.warm_chunks() chunks early from a ZERO state (csrc/eq_mma.cu, lti_warm_chunks).  Stepping the float64 chunk
    system both ways -- from the true state and from zero -- the outputs must agree to 2^-24 of max|x| after the
    warm-up, for worst-case-ish inputs too (a full-scale square wave at the slowest pole's frequency)."""
    import dsp_audio_project_b200 as pk
    plan = pk.EqPlan.from_gains(fs, gains_dict(gains), np.float32)
    w = plan.warm_chunks()
    T, K, O, Phi = plan.chunk_system()
    L, S = T.shape[0], Phi.shape[0]
    assert 1 <= w <= 4096
    rho = np.max(np.abs(np.linalg.eigvals(Phi)))
    assert rho ** w < 1e-3                                       # the slowest mode has decayed a good way
    lead = 3 * w + 20
    n = (lead + w + 12) * L
    t = np.arange(n)
    f_slow = 40.0 if gains[0] != 0 else 10000.0
    for x in (np.random.default_rng(5).uniform(-1, 1, n), np.sign(np.sin(2 * np.pi * f_slow * t / fs) + 1e-9),
              np.ones(n)):
        X = x.reshape(-1, L)
        s = np.zeros(S)
        true_out = []
        for k, xk in enumerate(X):
            if k >= lead + w:
                true_out.append(T @ xk + O @ s)
            s = Phi @ s + K @ xk
        s = np.zeros(S)
        cold = []
        for k in range(lead, len(X)):
            if k >= lead + w:
                cold.append(T @ X[k] + O @ s)
            s = Phi @ s + K @ X[k]
        err = np.max(np.abs(np.concatenate(true_out) - np.concatenate(cold)))
        assert err <= 2.0 ** -24, (fs, gains, w, err)
    # no tensor form, no bound
    assert pk.EqPlan(48000, [], np.float32).warm_chunks() == 0
