"""The ctypes stub shown in INTEGRATION.md (what a reference maintainer would
paste into modules/dsp_core.py) really works against libdspb200.so."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, gains_dict
from oracle import dsp_oracle as o

pytestmark = pytest.mark.gpu


def _stub_namespace():
    from dsp_audio_project_b200 import _lib
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(.*?)```", text, re.S).group(1)
    block = block.replace('C.CDLL("libdspb200.so")', f'C.CDLL({_lib.LIB_PATH!r})')
    ns = {}
    exec(compile(block, "INTEGRATION.md", "exec"), ns)
    return ns


def test_integration_stub_matches_reference(golden_src, golden_eq, golden_spectrum):
    ns = _stub_namespace()
    g = golden_src
    for idx in (0, 1, 2, 12, 15):
        L, M, N, fs_new, fs = (int(v) for v in g["cases"][idx])
        y, fs_out = ns["conversion_tasa_muestreo"](g[f"x_{idx}"], fs, M, L)
        assert fs_out == fs_new and o.rel_err(y, g[f"y_{idx}"]) <= 1e-10
    x = np.arange(4.0)
    assert ns["conversion_tasa_muestreo"](x, 44100, 1, 1)[0] is x
    ge = golden_eq
    for idx in (0, 2, 4, 6):
        row = ge["cases"][idx]
        out = ns["sistema_ecualizador"](ge[f"x_{idx}"], row[0], gains_dict(row[1:]))
        assert o.rel_err(out, ge[f"z_{idx}"]) <= 1e-10
    assert ns["sistema_ecualizador"](x, 48000, gains_dict((0.05,) * 6)) is x
    gs = golden_spectrum
    for n in (100, 2048, 5000):
        f, m = ns["calcular_espectro_magnitud"](gs[f"x_{n}"], 48000)
        assert np.allclose(f, gs[f"f_{n}"]) and o.rel_err(m, gs[f"m_{n}"]) <= 1e-10
    with pytest.raises(ValueError):
        ns["calcular_espectro_magnitud"](np.zeros(3000), 48000)
