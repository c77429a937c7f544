"""Drop-in GPU replacement of the reference's ``modules/dsp_core.py``.

Same eight public names, positional signatures, return types and error
behaviour as ``/root/reference/modules/dsp_core.py`` (cited per function), so

    from modules.dsp_core import (cargar_senal_audio, conversion_tasa_muestreo,
                                  sistema_ecualizador, calcular_espectro_magnitud)

in ``app.py:13-18`` keeps working.  The three numeric kernels run as float64
sm_100a CUDA kernels behind the C ABI (``include/dspb200.h``); 1-D numpy in,
fresh float64 numpy out (a batch of one channel).  There is no CPU fallback:
without the library or a B200 these functions raise.

Documented deviations from the reference ("semantic contract"):
* ``fft_diezmado_en_tiempo`` raises ``ValueError`` for EVERY non-power-of-two
  length; the reference raises for most (6, 12, 100, ...) but silently returns
  a wrong-length array for a few (N=3 -> length 4).
* ``aplicar_ecuacion_diferencias`` accepts sections up to second order (the
  only kind the reference ever builds); longer b/a raise NotImplementedError.
* Inputs are float32/float64 real (complex input is accepted by the FFT only).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import F64, check
from .plans import (BAND_CENTRES_HZ, EqPlan, cached_fft_plan, cached_src_plan, select_sections,
                    src_geometry)

__all__ = [
    "cargar_senal_audio", "fft_diezmado_en_tiempo", "calcular_espectro_magnitud",
    "generar_respuesta_impulso_sinc", "conversion_tasa_muestreo",
    "disenar_coeficientes_diferencias", "aplicar_ecuacion_diferencias", "sistema_ecualizador",
]

_N_VENTANA = 2048  # dsp_core.py:74


def cargar_senal_audio(buffer_archivo):
    """File loader (dsp_core.py:10-35) -- I/O, outside the GPU hot path; kept so
    the module is a complete drop-in: soundfile read, stereo->mono mean,
    float32, peak normalisation, and the reference's catch-all fallback."""
    try:
        import soundfile as sf
        x_n, fs = sf.read(buffer_archivo)
        if len(x_n.shape) > 1:
            x_n = x_n.mean(axis=1)
        x_n = x_n.astype(np.float32)
        peak = np.max(np.abs(x_n))
        if peak > 1e-6:
            x_n = x_n / peak
        return x_n, fs
    except Exception:
        return np.zeros(100, dtype=np.float32), 44100


def _is_pow2(n: int) -> bool:
    return n >= 1 and (n & (n - 1)) == 0


def fft_diezmado_en_tiempo(x):
    """Radix-2 DIT FFT (dsp_core.py:41-66): natural-order complex128 spectrum;
    length <= 1 returns the argument itself (:52)."""
    n = len(x)
    if n <= 1:
        return x
    if not _is_pow2(n):
        raise ValueError(f"fft_diezmado_en_tiempo needs a power-of-two length (got {n})")
    a = np.ascontiguousarray(x, dtype=np.complex128)
    if a.ndim != 1:
        raise ValueError("fft_diezmado_en_tiempo expects a 1-D sequence")
    plan = cached_fft_plan(n, F64, False)
    return plan.c2c_host_f64(a[None, :])[0]


def calcular_espectro_magnitud(x_n, fs, n_fft=None):
    """Hann-windowed magnitude spectrum (dsp_core.py:68-98).  Window rule of
    :74-82: a 2048-sample slice from len//2 when the signal is longer than
    2048, else zero-padding to the next power of two.  ``n_fft`` (extension,
    default 2048 = the reference's constant) changes the window length."""
    n_win = _N_VENTANA if n_fft is None else int(n_fft)
    x = np.asarray(x_n)
    if x.ndim != 1:
        raise ValueError("calcular_espectro_magnitud expects a 1-D signal")
    length = len(x)
    if length > n_win:
        offset = length // 2
        n = min(n_win, length - offset)          # the slice x[mid:mid+n_win] may come up short
    else:
        offset = 0
        n = 1 << (length - 1).bit_length()       # len 0 -> 2, as the reference
    if not _is_pow2(n):
        raise ValueError(f"window of {n} samples is not a power of two (signal length {length})")
    plan = cached_fft_plan(n, F64, True)
    xd = np.ascontiguousarray(x, dtype=np.float64)
    mag = plan.magnitudes_host(xd[None, :], hop=n, offset=offset, n_frames=1)[0, 0]
    freqs = np.fft.rfftfreq(n, d=1 / fs)
    keep = n // 2 + 1
    return freqs[:keep], mag[:keep]


def generar_respuesta_impulso_sinc(w_c_norm, L_taps):
    """Blackman-windowed sinc low-pass (dsp_core.py:104-131), float64."""
    n = int(L_taps)
    cap = n + 1
    h = np.empty(cap)
    got = C.c_int()
    check(_lib.load().dspb200_design_sinc_taps(float(w_c_norm), n, h.ctypes.data_as(C.POINTER(C.c_double)),
                                               cap, C.byref(got)))
    return h[:got.value].copy()


def conversion_tasa_muestreo(x_n, fs_original, M, L):
    """L/M sample-rate converter (dsp_core.py:133-173).  Positional order is
    (x, fs, M, L) as in app.py:164.  L == M == 1 returns the input object."""
    if M == 1 and L == 1:
        return x_n, fs_original
    L_i, M_i = int(L), int(M)
    if L_i != L or M_i != M or L_i < 1 or M_i < 1:
        raise ValueError(f"L and M must be positive integers (got L={L}, M={M})")
    x = np.asarray(x_n)
    if x.ndim != 1:
        raise ValueError("conversion_tasa_muestreo expects a 1-D signal")
    if len(x) == 0:
        raise ValueError("v cannot be empty")  # numpy.convolve's message for the empty expansion
    xd = np.ascontiguousarray(x, dtype=np.float64)
    plan = cached_src_plan(L_i, M_i, F64)
    y = plan.run_host(xd[None, :])[0]
    return y, int(fs_original * L_i / M_i)


def disenar_coeficientes_diferencias(fc, fs, ganancia_db):
    """RBJ peaking biquad coefficients (dsp_core.py:179-203) -> (b[3], a[3])."""
    b = np.empty(3)
    a = np.empty(3)
    pd = C.POINTER(C.c_double)
    check(_lib.load().dspb200_design_peaking_biquad(float(fc), float(fs), float(ganancia_db),
                                                    b.ctypes.data_as(pd), a.ctypes.data_as(pd)))
    return b, a


def aplicar_ecuacion_diferencias(x_n, b, a):
    """Difference-equation engine (dsp_core.py:205-214 = scipy.signal.lfilter,
    zero initial state) for sections up to second order."""
    b = np.atleast_1d(np.asarray(b, dtype=np.float64))
    a = np.atleast_1d(np.asarray(a, dtype=np.float64))
    if len(b) > 3 or len(a) > 3:
        raise NotImplementedError("only sections up to second order are supported")
    if len(a) == 0 or a[0] == 0:
        raise ValueError("a[0] must be non-zero")
    ba = np.zeros(6)
    ba[:len(b)] = b
    ba[3:3 + len(a)] = a
    x = np.asarray(x_n)
    if x.ndim != 1:
        raise ValueError("aplicar_ecuacion_diferencias expects a 1-D signal")
    plan = EqPlan(1.0, (), dtype=np.float64, clip=False, raw_ba=ba)
    return plan.run_host(np.ascontiguousarray(x, dtype=np.float64)[None, :])[0]


def sistema_ecualizador(x_n, fs, ganancias_bandas):
    """Six-band peaking-EQ cascade (dsp_core.py:216-254): bypass returns the
    input object when every |g| < 0.1; bands apply in dict order with |g| > 0.1,
    centre clamped to 0.9*fs/2, skipped at <= 10 Hz; one clip to [-1, 1]."""
    bypass, sections = select_sections(fs, ganancias_bandas)
    if bypass:
        return x_n
    x = np.asarray(x_n)
    if x.ndim != 1:
        raise ValueError("sistema_ecualizador expects a 1-D signal")
    if not sections:
        # |g| == 0.1 corner: a clipped copy in the input's own float type
        dt = np.float32 if x.dtype == np.float32 else np.float64
        plan = EqPlan(fs, (), dtype=dt, clip=True)
        return plan.run_host(np.ascontiguousarray(x, dtype=dt)[None, :])[0]
    plan = EqPlan(fs, sections, dtype=np.float64, clip=True)
    return plan.run_host(np.ascontiguousarray(x, dtype=np.float64)[None, :])[0]
