"""Numerical model of the fused SRC->EQ kernel (csrc/xz_mma.cu): per chunk of C outputs
    [z; u] = G_ph x_win + O s ,   s' = Phi s + u ,   G_ph = [T; K] A_ph
with A_ph the resampler's banded tap matrix of the chunk's phase and (T, K, O, Phi) the chunk system
of the biquad cascade (tools/eq_mma_model.py).  G_ph x is evaluated as the three-product fp16 split
the tcgen05 kernel uses (x_hi G_hi + x_hi G_lo + x_lo G_hi, fp32 accumulation), the free response as
[s1 | s2 | s3 | s1] . [O_hi | O_hi | O_hi | O_lo].  Run on the CPU to see the error of the formulation
against the float64 oracle before spending GPU time.  Not part of the product or the tests."""
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import dsp_oracle as o  # noqa: E402
from tools.eq_mma_model import cascade_ss, section_ss  # noqa: E402

C = 80
XS, SS = 6, 6          # x and the states are scaled by 2^6 before the fp16 split


def f16(v):
    return np.asarray(v, np.float32).astype(np.float16)


def split16(v):
    """v (float32) -> hi, lo in fp16 with hi + lo ~ v"""
    v = np.asarray(v, np.float32)
    hi = f16(v)
    lo = f16(v - hi.astype(np.float32))
    return hi, lo


def chunk_system(secs, rows):
    A, B, Cv, D = cascade_ss(secs)
    n = len(B)
    pw = [np.eye(n)]
    for _ in range(rows):
        pw.append(A @ pw[-1])
    h = np.array([D] + [Cv @ pw[i] @ B for i in range(rows - 1)])
    T = np.zeros((rows, rows))
    for r in range(rows):
        T[r, : r + 1] = h[: r + 1][::-1]
    O = np.array([Cv @ pw[r] for r in range(rows)])
    K = np.array([pw[rows - 1 - k] @ B for k in range(rows)]).T
    Phi = pw[rows]
    sc = 1.0 / np.maximum(np.linalg.norm(K, axis=1), 1e-300)
    return T, O / sc[None, :], K * sc[:, None], (Phi * sc[:, None]) / sc[None, :]


def geometry(L, M, T):
    P = (T - 1) // 2
    g = np.gcd(L, C * M)
    n_ph = L // g
    starts, ends = [], []
    for k in range(n_ph):
        m0 = k * C
        starts.append(-((-(m0 * M + P - T + 1)) // L))
        ends.append(((m0 + C - 1) * M + P) // L)
    return P, n_ph, starts, ends


def run(x, L, M, fs_out, gains, emulate=True):
    h = o.src_filter(L, M)
    T = len(h)
    P, n_ph, starts, ends = geometry(L, M, T)
    W = max(e - s + 1 for s, e in zip(starts, ends))
    adv = n_ph * C * M // L
    secs = [section_ss(*o.peaking_biquad(fc, fs_out, g)) for fc, g in o.eq_active_sections(fs_out, gains)]
    Te, Oe, Ke, Phi = chunk_system(secs, C)
    ns = Ke.shape[0]
    G = []
    for ph in range(n_ph):
        A = np.zeros((C, W))
        for r in range(C):
            for w in range(W):
                t = (ph * C + r) * M + P - (starts[ph] + w) * L
                if 0 <= t < T:
                    A[r, w] = h[t]
        G.append(np.vstack([Te, Ke]) @ A)
    gmax = max(np.abs(g).max() for g in G)
    ge = int(np.floor(np.log2(8192.0 / gmax)))
    omax = np.abs(Oe).max()
    n_in = len(x)
    n_out = -(-n_in * L // M)
    nch = -(-n_out // C)
    z = np.zeros(nch * C, np.float32)
    s = np.zeros(ns, np.float32)
    Phi32 = Phi.astype(np.float32)
    Gs = [split16((g * 2.0 ** ge).astype(np.float32)) for g in G]
    Os = split16((Oe * 2.0 ** (ge + XS - SS)).astype(np.float32))
    unscale = np.float32(2.0 ** -(ge + XS))
    for k in range(nch):
        ph = k % n_ph
        s0 = starts[ph] + (k // n_ph) * adv
        idx = np.arange(s0, s0 + W)
        xw = np.where((idx >= 0) & (idx < n_in), x[np.clip(idx, 0, n_in - 1)], 0).astype(np.float32)
        if emulate:
            xh, xl = split16(xw * np.float32(2.0 ** XS))
            gh, gl = Gs[ph]
            d = (gh.astype(np.float64) @ xh.astype(np.float64) + gl.astype(np.float64) @ xh.astype(np.float64)
                 + gh.astype(np.float64) @ xl.astype(np.float64))
            sv = s * np.float32(2.0 ** SS)
            s1 = f16(sv); r1 = sv - s1.astype(np.float32)
            s2 = f16(r1); s3 = f16(r1 - s2.astype(np.float32))
            oh, ol = Os
            d[:C] += oh.astype(np.float64) @ (s1.astype(np.float64) + s2.astype(np.float64) + s3.astype(np.float64)) \
                + ol.astype(np.float64) @ s1.astype(np.float64)
            d = d.astype(np.float32) * unscale
        else:
            d = (G[ph] @ xw.astype(np.float64))
            d[:C] += Oe @ s.astype(np.float64)
            d = d.astype(np.float32)
        z[k * C:(k + 1) * C] = d[:C]
        s = (Phi32 @ s + d[C:]).astype(np.float32)
    info = dict(W=W, n_ph=n_ph, starts=starts, ends=ends, gmax=gmax, ge=ge, omax=omax,
                smax=None)
    return np.clip(z[:n_out], -1, 1), info


def main():
    L, M, fs_in = 160, 147, 44100
    fs_out = fs_in * L // M
    rng = np.random.default_rng(7)
    names = ("Sub-Bass", "Bass", "Low Mids", "High Mids", "Presence", "Brilliance")
    cases = {
        "C1 gains": dict(zip(names, (6, -3, 4, -6, 3, -9))),
        "all +15": {k: 15 for k in names},
        "all -15": {k: -15 for k in names},
    }
    n_in = 44100
    for name, gains in cases.items():
        for amp in (0.5, 1.0):
            x = rng.uniform(-amp, amp, n_in).astype(np.float32)
            yo, _ = o.resample_closed_form(x.astype(np.float64), fs_in, M, L)
            zo = o.equalizer(yo, fs_out, gains)
            for em in (False, True):
                z, info = run(x, L, M, fs_out, gains, em)
                print(f"{name:9s} amp {amp}: fp16x3={em}: err vs f64 = {np.max(np.abs(z - zo)):.2e}  "
                      f"(W {info['W']}, phases {info['n_ph']}, |G|max {info['gmax']:.3f} 2^{info['ge']}, |O|max {info['omax']:.3g})")
    t = np.arange(n_in) / fs_in
    x = (0.9 * np.sin(2 * np.pi * 40 * t)).astype(np.float32)
    gains = cases["all +15"]
    yo, _ = o.resample_closed_form(x.astype(np.float64), fs_in, M, L)
    zo = o.equalizer(yo, fs_out, gains)
    z, _ = run(x, L, M, fs_out, gains, True)
    print("40 Hz sine, all +15: err =", np.max(np.abs(z - zo)))


if __name__ == "__main__":
    main()
