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
