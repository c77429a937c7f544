"""Numerical model of the tensor-core EQ (csrc/eq_mma.cu): the six-section cascade as one
12-state linear system, advanced 96 samples at a time,
    y_k = T x_k + O s_k ,   s_{k+1} = Phi s_k + K x_k ,
with [T; K] x_k evaluated as the three-product TF32 split the tcgen05 kernel uses.
Run on the CPU to see the error of the formulation against float64 lfilter before
spending GPU time.  Not part of the product or the tests."""
import sys

import numpy as np
from scipy.signal import lfilter

sys.path.insert(0, ".")
from oracle import dsp_oracle as o  # noqa: E402

L = 96


def section_ss(b, a):
    """same forms as csrc/design.cu section_from_ba (complex poles only here)"""
    b = np.asarray(b, float) / a[0]
    a = np.asarray(a, float) / a[0]
    B0, B1 = b[1] - a[1] * b[0], b[2] - a[2] * b[0]
    disc = a[1] ** 2 - 4 * a[2]
    if disc < 0:
        sg, om = -a[1] / 2, np.sqrt(-disc) / 2
        A = np.array([[sg, om], [-om, sg]])
        Bv = np.array([1.0, 0.0])
        C = np.array([B0, -(sg * B0 + B1) / om])
    else:
        A = np.array([[-a[1], 1.0], [-a[2], 0.0]])
        Bv = np.array([B0, B1])
        C = np.array([1.0, 0.0])
    return A, Bv, C, b[0]


def cascade_ss(secs):
    n = 2 * len(secs)
    A = np.zeros((n, n)); B = np.zeros(n); C = np.zeros(n); D = 1.0
    for i, (Ai, Bi, Ci, Di) in enumerate(secs):
        r = slice(2 * i, 2 * i + 2)
        A[r, :] = np.outer(Bi, C)          # input of section i = output so far
        A[r, r] = Ai
        B[r] = Bi * D
        C = Di * C
        C[r] = Ci
        D = Di * D
    return A, B, C, D


def tf32_trunc(v):
    u = np.asarray(v, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)
    return u.view(np.float32)


def tf32_round(v):
    u = (np.asarray(v, np.float32).view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)
    return u.view(np.float32)


def build(secs, scale_states=True):
    A, B, C, D = cascade_ss(secs)
    n = len(B)
    pw = [np.eye(n)]
    for _ in range(L):
        pw.append(A @ pw[-1])
    h = np.array([D] + [C @ pw[i] @ B for i in range(L - 1)])
    T = np.zeros((L, L))
    for r in range(L):
        T[r, : r + 1] = h[: r + 1][::-1]
    O = np.array([C @ pw[r] for r in range(L)])
    K = np.array([pw[L - 1 - k] @ B for k in range(L)]).T
    Phi = pw[L]
    if scale_states:   # balance: unit row norms of K
        sc = 1.0 / np.maximum(np.linalg.norm(K, axis=1), 1e-300)
        K = K * sc[:, None]; O = O / sc[None, :]; Phi = (Phi * sc[:, None]) / sc[None, :]
    return T, O, K, Phi


def run_model(x, secs, emulate=True):
    T, O, K, Phi = build(secs)
    n = len(x)
    nch = -(-n // L)
    xp = np.zeros(nch * L, np.float32); xp[:n] = x
    X = xp.reshape(nch, L).T                      # [L, chunks]
    M = np.vstack([T, K]).astype(np.float32)
    if emulate:
        Mh = tf32_round(M); Ml = tf32_trunc(M - Mh)
        Xh = tf32_trunc(X); Xl = tf32_trunc(X - Xh)
        acc = (Mh.astype(np.float64) @ Xh + Ml.astype(np.float64) @ Xh + Mh.astype(np.float64) @ Xl).astype(np.float32)
    else:
        acc = (M.astype(np.float64) @ X).astype(np.float32)
    Y0, U = acc[:L], acc[L:]
    s = np.zeros(len(Phi), np.float32)
    Phi32, O32 = Phi.astype(np.float32), O.astype(np.float32)
    out = np.empty((nch, L), np.float32)
    for k in range(nch):
        y = Y0[:, k].copy()
        for j in range(len(s)):                   # fp32 FMA chain like the epilogue
            y = y + O32[:, j] * s[j]
        out[k] = y
        s = (Phi32 @ s + U[:, k]).astype(np.float32)
    return np.clip(out.reshape(-1)[:n], -1, 1)


def main():
    fs = 48000.0
    rng = np.random.default_rng(5)
    cases = {
        "C1 gains": dict(zip(("Sub-Bass", "Bass", "Low Mids", "High Mids", "Presence", "Brilliance"), (6, -3, 4, -6, 3, -9))),
        "all +15": {k: 15 for k in ("Sub-Bass", "Bass", "Low Mids", "High Mids", "Presence", "Brilliance")},
        "all -15": {k: -15 for k in ("Sub-Bass", "Bass", "Low Mids", "High Mids", "Presence", "Brilliance")},
        "all -12": {k: -12 for k in ("Sub-Bass", "Bass", "Low Mids", "High Mids", "Presence", "Brilliance")},
    }
    for name, gains in cases.items():
        for amp in (0.25, 1.0):
            x = rng.uniform(-amp, amp, 480000).astype(np.float32)
            secs = [section_ss(*o.peaking_biquad(fc, fs, g)) for fc, g in o.eq_active_sections(fs, gains)]
            ref = o.equalizer(x.astype(np.float64), fs, gains)
            for em in (False, True):
                z = run_model(x, secs, em)
                print(f"{name:9s} amp {amp}: emulate_tf32x3={em}: err vs f64 = {np.max(np.abs(z - ref)):.2e} "
                      f"(unclipped max |ref| ~ {np.max(np.abs(ref)):.2f})")
        T, O, K, Phi = build(secs)
        print(f"   |T|max {np.abs(T).max():.2f} |O|max {np.abs(O).max():.2e} |K|max {np.abs(K).max():.2e} "
              f"|Phi|max {np.abs(Phi).max():.2e} rho {np.max(np.abs(np.linalg.eigvals(Phi))):.4f}")
    # low-frequency sine: the worst case for state magnitude
    t = np.arange(480000) / fs
    x = (0.9 * np.sin(2 * np.pi * 40 * t)).astype(np.float32)
    gains = cases["all +15"]
    secs = [section_ss(*o.peaking_biquad(fc, fs, g)) for fc, g in o.eq_active_sections(fs, gains)]
    A, B, C, D = cascade_ss(secs)
    y = x.astype(np.float64)
    for fc, g in o.eq_active_sections(fs, gains):
        b, a = o.peaking_biquad(fc, fs, g)
        y = lfilter(b, a, y)
    z = run_model(x, secs, True)
    print("40 Hz sine, all +15: err vs f64 (before clip compare on clipped) =", np.max(np.abs(z - np.clip(y, -1, 1))))


if __name__ == "__main__":
    main()
