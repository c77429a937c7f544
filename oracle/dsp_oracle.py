"""CPU oracle: float64 restatement of the reference's numeric hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Each function cites the
lines of ``/root/reference/modules/dsp_core.py`` whose behaviour it restates.
The third-party arithmetic the reference leans on is

* ``numpy.convolve``  (dense direct convolution, call site dsp_core.py:166) and
* ``scipy.signal.lfilter`` (direct-form-II-transposed recurrence, dsp_core.py:214);

``requirements.txt:1-6`` pins no versions; this container has numpy 2.3.5 and
scipy 1.18.1.  Both algorithms are restated here in plain loops
(``convolve_same_direct``, ``lfilter_df2t_loop``) and the fast library calls are
only used after being checked against those loops (tests/test_oracle_golden.py).

Parity pinning: outputs of the real reference on seeded inputs are committed
under ``tests/golden`` (generator: ``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import math

import numpy as np

try:  # scipy is the reference's own recurrence engine (dsp_core.py:3, :214)
    from scipy.signal import lfilter as _scipy_lfilter
except Exception:  # pragma: no cover - scipy is in the image
    _scipy_lfilter = None

# Reference constants (dsp_core.py:74, :158, :225-228, :222, :234, :240, :249)
SPECTRUM_WINDOW = 2048
TAPS_PER_UNIT = 40
BAND_CENTRES_HZ = {
    "Sub-Bass": 40, "Bass": 150, "Low Mids": 1000,
    "High Mids": 3000, "Presence": 5000, "Brilliance": 10000,
}
BAND_ORDER = ("Sub-Bass", "Bass", "Low Mids", "High Mids", "Presence", "Brilliance")
UNKNOWN_BAND_HZ = 1000
GAIN_BYPASS_DB = 0.1
NYQUIST_SAFETY = 0.90
MIN_CENTRE_HZ = 10


# --------------------------------------------------------------------------
# FIR design  (dsp_core.py:104-131)
# --------------------------------------------------------------------------
def sinc_lowpass_taps(w_c_norm: float, n_taps: int) -> np.ndarray:
    """Blackman-windowed sinc low-pass, unit DC gain.

    dsp_core.py:114 forces an odd length, :116 centres the index grid, :120 is
    ``sinc(wc*n)`` with numpy's normalised sinc, :123 multiplies by
    ``np.blackman`` and :127-129 divides by the sum when it is non-zero.
    """
    n_taps = int(n_taps)
    if n_taps % 2 == 0:
        n_taps += 1
    half = n_taps // 2
    n = np.arange(-half, half + 1)
    arg = np.pi * (w_c_norm * n).astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        core = np.where(arg == 0.0, 1.0, np.sin(arg) / arg)
    if n_taps == 1:
        win = np.ones(1)
    else:
        k = np.arange(n_taps)
        win = (0.42 - 0.5 * np.cos(2.0 * np.pi * k / (n_taps - 1))
               + 0.08 * np.cos(4.0 * np.pi * k / (n_taps - 1)))
    h = core * win
    total = h.sum()
    if total != 0:
        h = h / total
    return h


def src_filter(L: int, M: int) -> np.ndarray:
    """The resampler's filter: cutoff 1/max(L,M) (dsp_core.py:155), length
    40*max(L,M)+1 (:158), scaled by L (:162)."""
    big = max(int(L), int(M))
    return sinc_lowpass_taps(1.0 / big, TAPS_PER_UNIT * big + 1) * int(L)


# --------------------------------------------------------------------------
# Sample-rate conversion  (dsp_core.py:133-173)
# --------------------------------------------------------------------------
def convolve_same_direct(a: np.ndarray, v: np.ndarray) -> np.ndarray:
    """Plain-loop restatement of ``np.convolve(a, v, 'same')`` (numpy's
    published definition: the centred slice, of length max(len a, len v), of
    the full linear convolution).  O(len a * len v): small cases only."""
    a = np.asarray(a, dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    if a.size == 0 or v.size == 0:
        raise ValueError("convolve: inputs cannot be empty")
    full = np.zeros(a.size + v.size - 1)
    for i, ai in enumerate(a):
        if ai != 0.0:
            full[i:i + v.size] += ai * v
    n_out = max(a.size, v.size)
    start = (min(a.size, v.size) - 1) // 2
    return full[start:start + n_out]


def src_geometry(n_in: int, L: int, M: int):
    """(taps T, centre offset P, filtered length, output length) of the
    resampler for an input of n_in samples.  With the expanded signal of
    length n_in*L convolved in 'same' mode against T taps (dsp_core.py:166)
    numpy keeps max(n_in*L, T) samples starting (min(n_in*L, T)-1)//2 into the
    full convolution; the decimator (:170) keeps every M-th of them."""
    big = max(L, M)
    T = TAPS_PER_UNIT * big + 1
    n_exp = n_in * L
    P = (min(n_exp, T) - 1) // 2
    n_filt = max(n_exp, T)
    n_out = -(-n_filt // M)
    return T, P, n_filt, n_out


def resample_reference_form(x, fs, M, L, *, use_numpy_convolve=True):
    """Faithful restatement of dsp_core.py:133-173: bypass for L==M==1
    (:144-145, returns the input object), zero-stuffing (:148-150), dense
    'same' convolution with the L-scaled filter (:159-166), stride-M pick
    (:170), truncated rate (:172).  Note the positional order (x, fs, M, L)."""
    if M == 1 and L == 1:
        return x, fs
    x = np.asarray(x)
    n = x.shape[0]
    stuffed = np.zeros(n * L, dtype=x.dtype)
    stuffed[::L] = x
    h = src_filter(L, M)
    if use_numpy_convolve:
        filt = np.convolve(stuffed, h, mode="same")
    else:
        filt = convolve_same_direct(stuffed, h)
    return filt[::M], int(fs * L / M)


def resample_closed_form(x, fs, M, L, *, block=65536):
    """Polyphase closed form of the same map:
    ``y[m] = sum_i h[m*M + P - i*L] * x[i]``; float64 accumulation.  Used for
    sizes where the dense form is too slow; pinned against the faithful form
    and the golden vectors in tests/test_oracle_golden.py."""
    if M == 1 and L == 1:
        return x, fs
    x = np.asarray(x)
    n = x.shape[0]
    if n == 0:
        raise ValueError("convolve: inputs cannot be empty")
    xd = x.astype(np.float64)
    h = src_filter(L, M)
    T, P, _, n_out = src_geometry(n, L, M)
    jmax = -(-T // L)
    hp = np.concatenate([h, np.zeros(jmax * L + L - T)])
    y = np.empty(n_out)
    j = np.arange(jmax)
    for m0 in range(0, n_out, block):
        m = np.arange(m0, min(n_out, m0 + block))
        q = m * M + P
        i0 = q // L
        ph = q % L
        ti = ph[:, None] + j[None, :] * L          # tap index
        xi = i0[:, None] - j[None, :]              # input index
        ok = (xi >= 0) & (xi < n) & (ti < T)
        taps = hp[np.minimum(ti, hp.size - 1)]
        xs = xd[np.clip(xi, 0, n - 1)]
        y[m] = np.where(ok, taps * xs, 0.0).sum(axis=1)
    return y, int(fs * L / M)


# --------------------------------------------------------------------------
# Equaliser  (dsp_core.py:179-254)
# --------------------------------------------------------------------------
def peaking_biquad(fc: float, fs: float, gain_db: float):
    """RBJ peaking section with Q fixed by alpha=sin(w0)/2 (dsp_core.py:187-189),
    coefficients :192-197, normalised by a0 (:200-201).  Returns (b[3], a[3])."""
    w0 = 2.0 * np.pi * fc / fs
    alpha = np.sin(w0) / 2.0
    A = 10.0 ** (gain_db / 40.0)
    cw = np.cos(w0)
    a0 = 1.0 + alpha / A
    b = np.array([1.0 + alpha * A, -2.0 * cw, 1.0 - alpha * A]) / a0
    a = np.array([a0, -2.0 * cw, 1.0 - alpha / A]) / a0
    return b, a


def lfilter_df2t_loop(b, a, x) -> np.ndarray:
    """Plain-loop restatement of ``scipy.signal.lfilter(b, a, x)`` for a
    second-order section: direct form II transposed, zero initial state
    (the engine behind dsp_core.py:214).  Python loop: small cases only."""
    b = np.asarray(b, dtype=np.float64) / a[0]
    a = np.asarray(a, dtype=np.float64) / a[0]
    x = np.asarray(x, dtype=np.float64)
    y = np.empty_like(x)
    z0 = z1 = 0.0
    b0, b1, b2 = b
    _, a1, a2 = a
    for n in range(x.size):
        xn = x[n]
        yn = b0 * xn + z0
        z0 = b1 * xn - a1 * yn + z1
        z1 = b2 * xn - a2 * yn
        y[n] = yn
    return y


def difference_equation(x, b, a):
    """dsp_core.py:205-214: the LTI difference-equation engine."""
    if _scipy_lfilter is not None:
        return _scipy_lfilter(b, a, x)
    return lfilter_df2t_loop(b, a, x)


def eq_active_sections(fs: float, gains: dict):
    """Which bands the cascade applies and at which centre (dsp_core.py:233-251):
    dict order, |g| > 0.1, unknown key -> 1000 Hz, clamp to 0.9*fs/2, skip when
    the effective centre is <= 10 Hz.  Returns [(fc_eff, gain_db), ...]."""
    out = []
    nyq = fs / 2.0
    for name, g in gains.items():
        if abs(g) > GAIN_BYPASS_DB:
            fc = BAND_CENTRES_HZ.get(name, UNKNOWN_BAND_HZ)
            ceiling = nyq * NYQUIST_SAFETY
            fc_eff = ceiling if fc >= ceiling else fc
            if fc_eff > MIN_CENTRE_HZ:
                out.append((fc_eff, g))
    return out


def equalizer(x, fs, gains: dict):
    """dsp_core.py:216-254: bypass (returns the input object) when every
    |g| < 0.1 (:222-223); otherwise copy, run the active sections in series and
    clip once to [-1, 1] at the end (:254)."""
    if all(abs(g) < GAIN_BYPASS_DB for g in gains.values()):
        return x
    y = np.array(x, copy=True)
    for fc_eff, g in eq_active_sections(fs, gains):
        b, a = peaking_biquad(fc_eff, fs, g)
        y = difference_equation(y, b, a)
    return np.clip(y, -1.0, 1.0)


# --------------------------------------------------------------------------
# FFT and magnitude spectrum  (dsp_core.py:41-98)
# --------------------------------------------------------------------------
def fft_dit_recursive(x):
    """dsp_core.py:41-66: recursive radix-2 decimation in time.  Length <= 1
    returns the argument itself (:52); each level transforms the even and odd
    samples (:55-56), multiplies the odd half by exp(-2j*pi*k/N) (:59-60) and
    emits [E+t, E-t] (:63-64).  Power-of-two lengths only."""
    n = len(x)
    if n <= 1:
        return x
    even = fft_dit_recursive(x[0::2])
    odd = fft_dit_recursive(x[1::2])
    k = np.arange(n // 2)
    tw = np.exp(-2j * np.pi * k / n) * odd
    return np.concatenate([even + tw, even - tw])


def hann_symmetric(n: int) -> np.ndarray:
    """dsp_core.py:85-87: 0.5 - 0.5*cos(2*pi*k/(n-1)); n == 1 divides 0 by 0
    and yields NaN exactly as the reference does."""
    k = np.arange(n)
    with np.errstate(invalid="ignore", divide="ignore"):
        return 0.5 - 0.5 * np.cos(2 * np.pi * k / (n - 1))


def spectrum_segment(x):
    """dsp_core.py:74-82: a 2048-sample window starting at len//2 when the
    signal is longer than 2048, otherwise zero-padding to the next power of
    two (len 0 -> 2)."""
    x = np.asarray(x)
    n = len(x)
    if n > SPECTRUM_WINDOW:
        mid = n // 2
        return x[mid:mid + SPECTRUM_WINDOW]
    target = 1 << (n - 1).bit_length()
    return np.pad(x, (0, target - n))


def magnitude_spectrum(x, fs):
    """dsp_core.py:68-98: window selection, Hann, manual FFT, abs, rfftfreq
    axis, first N/2+1 bins.  Raises ValueError (from the FFT's concatenation)
    for 2049 <= len <= 4094 unless the slice happens to be a power of two."""
    seg = spectrum_segment(x)
    n = len(seg)
    spec = fft_dit_recursive(seg * hann_symmetric(n))
    mag = np.abs(spec)
    freqs = np.fft.rfftfreq(n, d=1 / fs)
    keep = n // 2 + 1
    return freqs[:keep], mag[:keep]


def frame_magnitudes(x, n_fft: int, hop: int | None = None, offset: int = 0, n_frames: int | None = None):
    """Framing rule fixed in SURVEY.md 8d for the throughput configs: frames of
    n_fft samples every ``hop`` (default n_fft) starting at ``offset``, tail
    dropped; each frame is |FFT(frame*hann)|[:n_fft/2+1] with the reference's
    FFT and window.  x may be [time] or [channels, time]."""
    x = np.asarray(x, dtype=np.float64)
    hop = n_fft if hop is None else hop
    one_d = x.ndim == 1
    x2 = x[None, :] if one_d else x
    n = x2.shape[1]
    fit = 0 if n - offset < n_fft else (n - offset - n_fft) // hop + 1
    n_frames = fit if n_frames is None else min(int(n_frames), fit)
    w = hann_symmetric(n_fft)
    out = np.empty((x2.shape[0], n_frames, n_fft // 2 + 1))
    for c in range(x2.shape[0]):
        for f in range(n_frames):
            s = offset + f * hop
            out[c, f] = np.abs(fft_dit_recursive(x2[c, s:s + n_fft] * w))[:n_fft // 2 + 1]
    return out[0] if one_d else out


# --------------------------------------------------------------------------
# The app's cascade (app.py:161-167, :202-205) on one signal
# --------------------------------------------------------------------------
def chain(x, fs, M, L, gains, n_fft=4096, n_frames=None):
    """SRC -> EQ -> framed magnitude spectra of z; n_frames limits the spectra to the
    first frames (the recursive FFT costs ~60 ms per 4096-point frame)."""
    y, fs2 = resample_closed_form(x, fs, M, L)
    z = equalizer(y, fs2, gains)
    return y, z, frame_magnitudes(z, n_fft, n_frames=n_frames), fs2


# --------------------------------------------------------------------------
# Either side of the path
# --------------------------------------------------------------------------
def load_mono_normalize(frames):
    """dsp_core.py:23-31 after the file read: stereo -> mono mean (:23-24),
    float32 (:26), division by the peak when it exceeds 1e-6 (:29-31).
    `frames` is what soundfile returns: [n] or [n, channels]."""
    x = np.asarray(frames)
    if x.ndim > 1:
        x = x.mean(axis=1)
    x = x.astype(np.float32)
    peak = np.max(np.abs(x)) if x.size else np.float32(0)
    if peak > 1e-6:
        x = x / peak
    return x


def spectrum_db(mag):
    """app.py:207-210: 20*log10(mag + 1e-12)."""
    return 20 * np.log10(np.asarray(mag) + 1e-12)


def spectrum_db_masked(f, mag):
    """app.py:207-208: mask = f > 0.5 (drops the DC bin of any practical frame), dB of the bins that pass."""
    f = np.asarray(f)
    mask = f > 0.5
    return mask, 20 * np.log10(np.asarray(mag)[mask] + 1e-12)


def pcm16_export(z):
    """app.py:349-354: nan_to_num, divide by the peak when it is > 0, * 32767,
    truncate to int16.  Inline code of the Streamlit script, which cannot be
    imported here; pinned by tests/golden/app_helpers.npz, which make_golden.py
    produces by executing those statements lifted from app.py with ``ast``."""
    y = np.nan_to_num(np.asarray(z))
    peak = np.max(np.abs(y)) if y.size else 0
    if peak > 0:
        y = y / peak
    return (y * 32767).astype(np.int16)


def synthetic_clips(channels: int, n: int, seed: int, lo: float = -0.5, hi: float = 0.5, first_channel: int = 0,
                          dtype=np.float32):
    """numpy restatement of the library's on-device clip generator (dspb200_generate_uniform_*, csrc/post.cu: splitmix64
    of a per-pair counter, 24 bits per sample): not reference behaviour, only how the throughput configurations of
    SURVEY.md 8d make their inputs, so that the tests can reproduce a device-generated wave on the host."""
    pairs = (n + 1) // 2
    idx = (np.arange(first_channel, first_channel + channels, dtype=np.uint64)[:, None] * np.uint64(pairs)
           + np.arange(1, pairs + 1, dtype=np.uint64)[None, :])
    with np.errstate(over="ignore"):
        z = np.uint64(int(seed) & (2 ** 64 - 1)) + np.uint64(0x9E3779B97F4A7C15) * idx
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    dt = np.dtype(dtype)
    bits = np.empty((channels, 2 * pairs), dtype=np.uint64)
    bits[:, 0::2] = z >> np.uint64(40)                                   # even samples: top 24 bits
    bits[:, 1::2] = (z & np.uint64(0xFFFFFFFF)) >> np.uint64(8)          # odd samples: top 24 bits of the low word
    u = bits[:, :n].astype(dt) * dt.type(1.0 / 16777216.0)
    return (dt.type(lo) + dt.type(hi - lo) * u).astype(dt)


def rel_err(a, ref) -> float:
    """max|a-ref| / max|ref| (the north star's relative error)."""
    a = np.asarray(a)
    ref = np.asarray(ref)
    if ref.size == 0:
        return 0.0
    den = float(np.max(np.abs(ref)))
    return float(np.max(np.abs(a - ref))) / (den if den > 0 else 1.0)


def full_scale_err(a, ref, full_scale: float = 1.0) -> float:
    """max|a-ref| / full_scale (fp32 tolerances are quoted against full scale)."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if ref.size == 0:
        return 0.0
    return float(np.max(np.abs(a - ref))) / full_scale


def next_pow2(n: int) -> int:
    return 1 << max(0, (int(n) - 1).bit_length())


def _self_check():  # pragma: no cover - manual sanity run
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, 200)
    a, _ = resample_reference_form(x, 44100, 2, 3)
    b, _ = resample_closed_form(x, 44100, 2, 3)
    print("src", np.max(np.abs(a - b)))
    print("fft", np.max(np.abs(fft_dit_recursive(x[:128].astype(complex)) - np.fft.fft(x[:128]))))
    print(math.pi)


if __name__ == "__main__":  # pragma: no cover
    _self_check()
