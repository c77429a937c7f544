#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the REAL reference.

Run in the build container only (it imports /root/reference/modules/dsp_core.py
through oracle.ref_loader, with a stub ``soundfile``):

    python tests/golden/make_golden.py

The reference ships no tests or fixtures, so these files --
inputs made from seeded numpy generators, outputs computed by the unmodified
reference functions -- are what pins the oracle and the CUDA path.  Versions at
generation time are recorded in ``manifest.json``.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.ref_loader import load_reference_dsp_core  # noqa: E402

BANDS = ["Sub-Bass", "Bass", "Low Mids", "High Mids", "Presence", "Brilliance"]


def gains_dict(values, keys=BANDS):
    return {k: v for k, v in zip(keys, values)}


def extract_app_helpers(app_path="/root/reference/app.py"):
    """The two pieces of arithmetic app.py applies to the kernels' outputs are inline statements of a Streamlit script
    that cannot be imported (streamlit/plotly absent).  They are pinned by EXECUTING the reference's own statements:
    the script is parsed with ``ast``, the assignments that compute ``mask_in``/``db_in`` (app.py:207-208) and
    ``y_final_audio``/``peak`` plus the int16 expression handed to ``write`` (app.py:349-354) are lifted out unchanged
    and compiled.  Returns (spectrum_db_masked(f, mag) -> (mask, db), pcm16(z) -> int16, {name: line})."""
    import ast

    tree = ast.parse(open(app_path, encoding="utf-8").read(), filename=app_path)
    found = {}

    def assigns_to(node, name):
        return (isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name)
                and node.targets[0].id == name)

    for node in ast.walk(tree):
        for name in ("mask_in", "db_in", "y_final_audio", "peak"):
            if assigns_to(node, name) and name not in found:
                found[name] = node
        if isinstance(node, ast.If) and isinstance(node.test, ast.Compare) and isinstance(node.test.left, ast.Name) \
                and node.test.left.id == "peak" and "peak_if" not in found:
            found["peak_if"] = node
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id == "write" \
                and len(node.args) == 3 and "int16_expr" not in found:
            found["int16_expr"] = node.args[2]
    missing = {"mask_in", "db_in", "y_final_audio", "peak", "peak_if", "int16_expr"} - set(found)
    if missing:
        raise RuntimeError(f"app.py no longer holds {sorted(missing)}")
    lines = {k: int(v.lineno) for k, v in found.items()}

    def module(*nodes):
        m = ast.Module(body=list(nodes), type_ignores=[])
        ast.fix_missing_locations(m)
        return compile(m, app_path, "exec")

    db_code = module(found["mask_in"], found["db_in"])
    pcm_code = module(found["y_final_audio"], found["peak"], found["peak_if"])
    expr = ast.Expression(body=found["int16_expr"])
    ast.fix_missing_locations(expr)
    pcm_expr = compile(expr, app_path, "eval")

    def spectrum_db_masked(f, mag):
        ns = {"np": np, "f_in": f, "mag_in": mag}
        exec(db_code, ns)
        return ns["mask_in"], ns["db_in"]

    def pcm16(z):
        ns = {"np": np, "z_final": z}
        exec(pcm_code, ns)
        return eval(pcm_expr, ns)

    return spectrum_db_masked, pcm16, lines


def main():
    import scipy

    ref = load_reference_dsp_core()
    manifest = {"numpy": np.__version__, "scipy": scipy.__version__,
                "python": sys.version.split()[0], "files": {}}

    # ---------------- design helpers (dsp_core.py:104-131, :179-203) -------
    design = {}
    for i, (wc, n) in enumerate([(1 / 3, 121), (1 / 160, 6401), (0.5, 80), (1 / 8, 321),
                                 (0.25, 1), (1 / 7, 281)]):
        design[f"taps_{i}_args"] = np.array([wc, n])
        design[f"taps_{i}"] = ref.generar_respuesta_impulso_sinc(wc, n)
    bq = []
    for fc in (40, 150, 1000, 3000, 5000, 10000, 21600.0, 3600.0):
        for fs in (48000, 44100, 66150, 8000):
            for g in (-15, -12.0412, -6, -0.5, 0.2, 3, 6, 15):
                if fc < fs / 2:
                    b, a = ref.disenar_coeficientes_diferencias(fc, fs, g)
                    bq.append(np.concatenate([[fc, fs, g], b, a]))
    design["biquads"] = np.array(bq)
    np.savez_compressed(os.path.join(HERE, "design.npz"), **design)

    # ---------------- SRC (dsp_core.py:133-173) ----------------------------
    src = {}
    cases = []
    rng = np.random.default_rng(20261018)
    idx = 0
    for (L, M, N, dt) in [
        (3, 2, 2000, "f4"), (3, 2, 2000, "f8"), (2, 3, 1501, "f8"), (1, 2, 777, "f4"),
        (2, 1, 300, "f8"), (8, 8, 400, "f8"), (1, 8, 3000, "f4"), (8, 1, 200, "f8"),
        (5, 7, 999, "f8"), (7, 5, 640, "f4"), (4, 6, 512, "f8"), (6, 4, 512, "f8"),
        (160, 147, 441, "f4"), (160, 147, 300, "f8"), (147, 160, 320, "f8"),
        (3, 2, 10, "f8"), (3, 2, 40, "f8"), (3, 2, 41, "f8"), (5, 7, 1, "f8"),
        (7, 5, 3, "f4"), (8, 3, 2, "f8"), (1, 1, 50, "f4"), (2, 8, 161, "f8"),
    ]:
        x = rng.uniform(-1, 1, N).astype(dt)
        y, fs_new = ref.conversion_tasa_muestreo(x, 44100, M, L)
        src[f"x_{idx}"] = x
        src[f"y_{idx}"] = np.asarray(y)
        cases.append([L, M, N, fs_new, 44100])
        idx += 1
    # impulse and DC probes
    for (L, M) in [(3, 2), (160, 147), (2, 5)]:
        x = np.zeros(200); x[100] = 1.0
        y, fs_new = ref.conversion_tasa_muestreo(x, 48000, M, L)
        src[f"x_{idx}"] = x; src[f"y_{idx}"] = y
        cases.append([L, M, 200, fs_new, 48000]); idx += 1
        x = np.ones(300)
        y, fs_new = ref.conversion_tasa_muestreo(x, 48000, M, L)
        src[f"x_{idx}"] = x; src[f"y_{idx}"] = y
        cases.append([L, M, 300, fs_new, 48000]); idx += 1
    src["cases"] = np.array(cases, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "src.npz"), **src)

    # ---------------- EQ (dsp_core.py:205-254) -----------------------------
    eq = {}
    eq_cases = []
    rng = np.random.default_rng(2)
    gain_sets = [
        (6, -3, 4, -6, 3, -9), (15, 15, 15, 15, 15, 15), (-15, -15, -15, -15, -15, -15),
        (-12.0412, 0, 0, 0, 0, 0), (0, -12.0412, -12.5, 0, 0, 0), (0.1, 0.1, 0.1, 0.1, 0.1, 0.1),
        (0.05, 0, -0.09, 0, 0, 0), (0, 0, 0, 0, 0, 12), (0.11, 0, 0, 0, 0, 0),
        (-13, 9, -15, 15, -1, 1), (15, 0, 0, 0, 0, 0),
    ]
    idx = 0
    for fs in (48000, 66150, 44100, 8000):
        for gs in gain_sets:
            for dt, amp in (("f8", 0.25), ("f4", 1.0)):
                if fs != 48000 and (dt == "f4" or gs not in gain_sets[:4]):
                    continue
                N = 2500
                x = (rng.uniform(-amp, amp, N)).astype(dt)
                z = ref.sistema_ecualizador(x, fs, gains_dict(gs))
                eq[f"x_{idx}"] = x
                eq[f"z_{idx}"] = np.asarray(z)
                eq[f"alias_{idx}"] = np.array(z is x)
                eq_cases.append([fs] + list(gs))
                idx += 1
    # a long low-frequency-heavy run: the 40 Hz section's time constant is ~1000 samples
    x = rng.uniform(-0.25, 0.25, 24000)
    eq[f"x_{idx}"] = x
    eq[f"z_{idx}"] = ref.sistema_ecualizador(x, 48000, gains_dict((6, -3, 4, -6, 3, -9)))
    eq[f"alias_{idx}"] = np.array(False)
    eq_cases.append([48000, 6, -3, 4, -6, 3, -9]); idx += 1
    # unknown band key and reversed key order (dict order matters, dsp_core.py:233-235)
    x = rng.uniform(-0.5, 0.5, 3000)
    g_unknown = {"Sub-Bass": 4, "Air": -7, "Brilliance": 5}
    eq["x_unknown"] = x
    eq["z_unknown"] = ref.sistema_ecualizador(x, 48000, g_unknown)
    g_rev = gains_dict((6, -3, 4, -6, 3, -9)[::-1], BANDS[::-1])
    eq["z_reversed"] = ref.sistema_ecualizador(x, 48000, g_rev)
    # single section engine
    b, a = ref.disenar_coeficientes_diferencias(40, 48000, 15)
    eq["x_lf"] = x
    eq["y_lf"] = ref.aplicar_ecuacion_diferencias(x, b, a)
    eq["ba_lf"] = np.concatenate([b, a])
    eq["cases"] = np.array(eq_cases, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "eq.npz"), **eq)

    # ---------------- FFT / spectrum (dsp_core.py:41-98) -------------------
    fft = {}
    rng = np.random.default_rng(3)
    for n in (1, 2, 4, 8, 16, 64, 256, 1024, 2048, 4096):
        xr = rng.uniform(-1, 1, n)
        xc = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
        fft[f"xr_{n}"] = xr
        fft[f"Xr_{n}"] = np.asarray(ref.fft_diezmado_en_tiempo(xr))
        fft[f"xc_{n}"] = xc
        fft[f"Xc_{n}"] = np.asarray(ref.fft_diezmado_en_tiempo(xc))
    xr = rng.uniform(-1, 1, 65536)
    fft["xr_65536"] = xr.astype(np.float32)          # stored as f32 to keep the file small
    fft["Xr_65536_mag"] = np.abs(ref.fft_diezmado_en_tiempo(fft["xr_65536"].astype(np.float64)))[:32769]
    np.savez_compressed(os.path.join(HERE, "fft.npz"), **fft)

    spec = {}
    rng = np.random.default_rng(4)
    lens = [0, 1, 2, 3, 5, 100, 1000, 2047, 2048, 4096, 4097, 5000, 20001]
    err_lens = []
    for n in lens + [2049, 3000, 4094, 4095]:
        x = rng.uniform(-1, 1, n).astype(np.float32 if n % 2 else np.float64)
        try:
            with np.errstate(all="ignore"):
                f, m = ref.calcular_espectro_magnitud(x, 48000)
        except ValueError:
            err_lens.append(n)
            continue
        spec[f"x_{n}"] = x
        spec[f"f_{n}"] = f
        spec[f"m_{n}"] = m
    spec["ok_lens"] = np.array([n for n in lens + [2049, 3000, 4094, 4095] if n not in err_lens])
    spec["valueerror_lens"] = np.array(err_lens)
    np.savez_compressed(os.path.join(HERE, "spectrum.npz"), **spec)

    # ---------------- C1 stand-in chain: 1 channel x 30 s @ 44.1 kHz ----------------
    # x is not stored (5 MB of noise): it is default_rng(20261018).uniform(-1, 1, 30 * 44100) -> float32 ->
    # peak-normalised as the loader does (dsp_core.py:26-31); x_check pins the generator's stream.
    n_c1 = 30 * 44100
    rng = np.random.default_rng(20261018)
    x = rng.uniform(-1, 1, n_c1).astype(np.float32)
    peak = np.max(np.abs(x))
    x = x / peak                                             # loader's normalisation, dsp_core.py:29-31
    y, fs2 = ref.conversion_tasa_muestreo(x, 44100, 2, 3)
    z = ref.sistema_ecualizador(y, fs2, gains_dict((6, -3, 4, -6, 3, -9)))
    mid = len(z) // 2
    frame = z[mid:mid + 4096]
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(4096) / 4095)
    mag = np.abs(ref.fft_diezmado_en_tiempo(frame * w))[:2049]
    f3, m3 = ref.calcular_espectro_magnitud(z[:100000], fs2)
    # 4096-point frames every 2^17 samples: what the batched chain's framed spectra are compared with
    frame_starts = np.arange(0, len(z) - 4096 + 1, 1 << 17)
    frames_mag = np.stack([np.abs(ref.fft_diezmado_en_tiempo(z[s0:s0 + 4096] * w))[:2049] for s0 in frame_starts])
    np.savez_compressed(os.path.join(HERE, "chain_c1.npz"), n=np.array(n_c1), seed=np.array(20261018),
                        x_check=np.array([float(np.sum(x.astype(np.float64))), float(x[0]), float(x[-1])]),
                        y_head=y[:4096], y_tail=y[-4096:], y_mid=y[mid:mid + 4096],
                        z_head=z[:4096], z_tail=z[-4096:], z_mid=z[mid:mid + 4096],
                        y_sum=np.array([np.sum(y), np.sum(np.abs(y)), np.sum(y * y)]),
                        z_sum=np.array([np.sum(z), np.sum(np.abs(z)), np.sum(z * z)]),
                        mag4096=mag, f_app=f3, m_app=m3, fs2=np.array(fs2),
                        frame_starts=frame_starts, frames_mag=frames_mag)

    # ---------------- what app.py does to the kernels' outputs (app.py:207-210, :349-354) ----------------
    # executed from the reference's own statements (extract_app_helpers), on the C1 spectrum and on edge inputs
    db_masked, pcm16, app_lines = extract_app_helpers()
    helpers = {"app_lines": np.array([app_lines[k] for k in sorted(app_lines)]),
               "app_line_names": np.array(sorted(app_lines))}
    mask, db = db_masked(f3, m3)
    helpers["f_c1"], helpers["m_c1"], helpers["mask_c1"], helpers["db_c1"] = f3, m3, mask, db
    rng = np.random.default_rng(6)
    m_edge = np.concatenate([[0.0, 1e-13, 1e-12, 1.0, 1e6], rng.uniform(0, 40, 251)])
    f_edge = np.concatenate([[0.0, 0.25, 0.5, 0.5000001, 23.4], np.arange(251) * 23.4375 + 46.875])
    mask, db = db_masked(f_edge, m_edge)
    helpers["f_edge"], helpers["m_edge"], helpers["mask_edge"], helpers["db_edge"] = f_edge, m_edge, mask, db
    zs = {
        "c1": z[:48000].copy(),
        "noise": rng.uniform(-0.8, 0.8, 30011),
        "nan_inf": np.concatenate([rng.uniform(-0.3, 0.3, 500), [np.nan, 0.7, -0.9]]),
        "zeros": np.zeros(257),
        "f32": rng.uniform(-0.5, 0.5, 4099).astype(np.float32),
        "full": np.array([1.0, -1.0, 0.5, -0.5, 0.999969482421875, 1.0 / 32767, 0.5 / 32767]),
    }
    with np.errstate(all="ignore"):
        for name, zz in zs.items():
            helpers[f"z_{name}"] = zz.copy()
            helpers[f"pcm_{name}"] = pcm16(zz.copy())
    helpers["pcm_names"] = np.array(sorted(zs))
    np.savez_compressed(os.path.join(HERE, "app_helpers.npz"), **helpers)

    # ---------------- loader front end (dsp_core.py:10-35) ------------------
    # the real cargar_senal_audio, fed through a stand-in soundfile.read
    rng = np.random.default_rng(5)
    loader = {}
    frames_in = {
        "stereo": rng.uniform(-0.7, 0.7, (3000, 2)),
        "mono": rng.uniform(-0.2, 0.2, 2500),
        "quad": rng.uniform(-1.5, 1.5, (1200, 4)),
        "tiny": rng.uniform(-1, 1, (500, 2)) * 1e-7,
        "silence": np.zeros((300, 2)),
    }
    import soundfile as sf_stub
    for name, arr in frames_in.items():
        sf_stub.read = (lambda a: (lambda _buf: (a, 44100)))(arr)
        out, fs_l = ref.cargar_senal_audio("ignored")
        assert fs_l == 44100 and out.dtype == np.float32 and out.shape == arr.shape[:1]
        loader[f"in_{name}"] = arr
        loader[f"out_{name}"] = out
    np.savez_compressed(os.path.join(HERE, "loader.npz"), **loader)

    for name in sorted(os.listdir(HERE)):
        if name.endswith(".npz"):
            manifest["files"][name] = os.path.getsize(os.path.join(HERE, name))
    with open(os.path.join(HERE, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    print(json.dumps(manifest, indent=1))


if __name__ == "__main__":
    main()
