// K3 -- batched radix-2 FFT and windowed magnitude spectrum.
//
// Replaces fft_diezmado_en_tiempo (dsp_core.py:41-66) and the arithmetic of
// calcular_espectro_magnitud (dsp_core.py:85-98: symmetric Hann, FFT, abs,
// first N/2+1 bins).  The reference recurses a radix-2 decimation-in-time split
// down to length 1; here the same radix-2 butterflies are grouped 4 levels at a
// time into register-resident 16-point blocks (radix 2/4/8 for the remainder)
// and the levels are chained Stockham-style through shared memory, so the
// result is the same DFT in natural order without a bit-reversal pass.
//
// Real input uses the conjugate-symmetry trick: N real samples are read as N/2
// complex points z[j] = x[2j] + i x[2j+1] (one 8/16-byte load per point, Hann
// applied on the fly), transformed with an N/2-point complex FFT and split
//     Xe = (Z[k] + conj Z[n-k]) / 2,  Xo = -i (Z[k] - conj Z[n-k]) / 2,
//     |X[k]| = |Xe + W_N^k Xo|,  |X[n-k]| = |Xe - W_N^k Xo|,   n = N/2,
// so only N/2+1 magnitudes are ever written.
//
// One CTA transforms one frame entirely on chip when n <= kSubMax (the 2048- and
// 4096-point sizes of the app and of config C5).  Larger transforms (2^16 of
// config C4) use a top-level radix-R decimation-in-time split: R interleaved
// sub-sequences are transformed by R CTAs into a workspace (sized to stay
// L2-resident) and a second kernel does the last radix-R level fused with the
// real split and the magnitude.
//
// Twiddles are float64-accurate tables rounded to the working type (the
// reference evaluates np.exp per level, dsp_core.py:59-60), never recurrences;
// the fp32 magnitude kernel and the four-step form derive some powers from
// table entries by product trees at most 3-4 roundings deep.
//
// Roofline (per frame of N real samples): N*sizeof(T) bytes read +
// (N/2+1)*sizeof(T) written; ~2.5 N log2 N flop.
#include <cmath>
#include <cstdlib>
#include <new>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "cpx.cuh"
#include "internal.cuh"

namespace dspb200 {

template <typename T> struct FftCfg;
template <> struct FftCfg<float> { static constexpr int kSubMax = 8192; };
template <> struct FftCfg<double> { static constexpr int kSubMax = 4096; };

constexpr int kPadShift = 4;  // one pad element per 16: conflict-free strided stores
__host__ __device__ __forceinline__ int padded(int i) { return i + (i >> kPadShift); }

// cos/sin(2 pi k / 16), k = 0..7
__device__ constexpr double kCos16[8] = {1.0, 0.92387953251128673848, 0.70710678118654752440,
                                         0.38268343236508977173, 0.0, -0.38268343236508977173,
                                         -0.70710678118654752440, -0.92387953251128673848};
__device__ constexpr double kSin16[8] = {0.0, 0.38268343236508977173, 0.70710678118654752440,
                                         0.92387953251128673848, 1.0, 0.92387953251128673848,
                                         0.70710678118654752440, 0.38268343236508977173};

// R-point DFT in registers, natural order in and out: the reference's radix-2
// decimation-in-time recursion (even/odd split, twiddle, butterfly) unrolled.
template <typename T, int R> struct Dft {
  typedef typename Cpx<T>::type C;
  static __device__ __forceinline__ void run(C* v) {
    C e[R / 2], o[R / 2];
#pragma unroll
    for (int k = 0; k < R / 2; ++k) { e[k] = v[2 * k]; o[k] = v[2 * k + 1]; }
    Dft<T, R / 2>::run(e);
    Dft<T, R / 2>::run(o);
#pragma unroll
    for (int k = 0; k < R / 2; ++k) {
      if (k == 0) {
        v[k] = cadd(e[k], o[k]);
        v[k + R / 2] = csub(e[k], o[k]);
      } else if (4 * k == R) {
        const C t = mul_neg_i(o[k]);
        v[k] = cadd(e[k], t);
        v[k + R / 2] = csub(e[k], t);
      } else {
        // e + w o as FMAs onto e, e - w o = 2 e - (e + w o): 6 operations per butterfly instead of 8
        C w;
        w.x = static_cast<T>(kCos16[k * (16 / R)]);
        w.y = static_cast<T>(-kSin16[k * (16 / R)]);
        const C lo = cmadd(w, o[k], e[k]);
        v[k] = lo;
        v[k + R / 2] = twice_minus(e[k], lo);
      }
    }
  }
};
template <typename T> struct Dft<T, 1> {
  typedef typename Cpx<T>::type C;
  static __device__ __forceinline__ void run(C*) {}
};

template <typename T> struct FftArgs {
  typedef typename Cpx<T>::type C;
  // input
  const T* x;              // real frames: [channels, time]; c2c: interleaved complex
  long long x_stride;
  long long n_valid, offset, hop, n_frames;
  const T* window;         // Hann [2*nc] or nullptr
  // Hann by angle addition (fixed-size kernel): per-thread (-cos/2, sin/2) of samples 2t, 2t+1 and the
  // step rotations cos/sin(u * 2 pi 2Q/(N-1)), duplicated into both halves of a packed operand
  const T* hann_ab;        // [Q][4] = (-cos(phi_2t)/2, -cos(phi_2t+1)/2, sin(phi_2t)/2, sin(phi_2t+1)/2)
  C hann_cc[16], hann_ss[16];
  int db;                  // write 20*log10(mag + 1e-12) instead of mag (app.py:207-210)
  // geometry
  int nc;                  // complex points of the whole transform
  int m;                   // complex points of the per-CTA sub-transform (nc = m * r_top)
  int r_top;
  int n_pass;
  int radix[8];
  const C* tw_pass;        // per-pass twiddle rows, pass p at tw_offset[p]: [ns_p][17], row[r] = W^(r k)
  int tw_offset[8];
  int tw_total;
  int tw_in_smem;
  const C* tw_tree;        // radix-16 passes only: [ns_p][kTreePitch] = W^k, W^2k, W^4k, W^8k (product-tree form)
  int tw_tree_offset[8];
  int tw_tree_total;
  const C* tw_top;         // W_nc[k]
  const C* tw_post;        // W_(2nc)[k], k = 0..nc/2
  // output
  T* mag; long long mag_frame_stride, mag_channel_stride;
  C* out;                  // c2c result or workspace
  long long n_items;       // transforms * r_top (of this launch)
  long long first;         // index of the launch's first transform (split transforms run in L2-sized chunks)
};

// Thread t owns the 16 points  t + s*(m/16), s = 0..15, of a pass's input (a
// coalesced / conflict-free access for consecutive t).  For radix R that is
// B = 16/R butterflies; butterfly b takes its r-th operand from slot b + r*B.
template <typename T, int R>
__device__ __forceinline__ void gather_butterflies(typename Cpx<T>::type* v, const typename Cpx<T>::type* tmp) {
  constexpr int B = 16 / R;
#pragma unroll
  for (int b = 0; b < B; ++b)
#pragma unroll
    for (int r = 0; r < R; ++r) v[b * R + r] = tmp[b + r * B];
}

constexpr int kTwPitch = 17;   // per-pass twiddle rows [k][17]: 16-lane 8-byte reads hit distinct banks
constexpr int kTreePitch = 5;  // compact rows [k][5]: (10 k + 2 c) mod 32 is distinct over 16 consecutive k

template <typename T, int R>
__device__ __forceinline__ void compute_store_pass(const typename Cpx<T>::type* tmp, typename Cpx<T>::type* s,
                                                   const typename Cpx<T>::type* __restrict__ tw, int t,
                                                   int m, int ns) {
  typedef typename Cpx<T>::type C;
  constexpr int B = 16 / R;
  C v[16];
  gather_butterflies<T, R>(v, tmp);
#pragma unroll
  for (int b = 0; b < B; ++b) {
    const int j = t + b * (m / 16);
    const int k = j & (ns - 1);
    C* w = v + b * R;
    if (ns > 1) {
      const C* row = tw + k * kTwPitch;      // row[r] = exp(-2 pi i r k / (ns R))
      C wv[R];
#pragma unroll
      for (int r = 1; r < R; ++r) wv[r] = row[r];
#pragma unroll
      for (int r = 1; r < R; ++r) w[r] = cmul(w[r], wv[r]);
    }
    Dft<T, R>::run(w);
    const int base = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) s[padded(base + r * ns)] = w[r];
  }
}

template <typename T>
__device__ __forceinline__ void run_pass(int R, const typename Cpx<T>::type* tmp, typename Cpx<T>::type* s,
                                         const typename Cpx<T>::type* tw, int t, int m, int ns) {
  switch (R) {
    case 16: compute_store_pass<T, 16>(tmp, s, tw, t, m, ns); break;
    case 8: compute_store_pass<T, 8>(tmp, s, tw, t, m, ns); break;
    case 4: compute_store_pass<T, 4>(tmp, s, tw, t, m, ns); break;
    default: compute_store_pass<T, 2>(tmp, s, tw, t, m, ns); break;
  }
}

// MODE 0: real frames -> magnitudes (r_top == 1)      MODE 1: real frames -> workspace
// MODE 2: complex -> complex (r_top == 1)              MODE 3: complex -> workspace
template <typename T, int MODE, int MAXT>
__global__ void __launch_bounds__(MAXT, (sizeof(T) == 4 ? (MAXT <= 128 ? 5 : (MAXT <= 256 ? 2 : 1)) : (MAXT <= 128 ? 2 : 1)))
fft_stockham_kernel(const FftArgs<T> a) {
  typedef typename Cpx<T>::type C;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* s = reinterpret_cast<C*>(smem_raw);
  const int t = threadIdx.x;
  const int m = a.m;
  const int q16 = m / 16;
  const bool active = t < q16;
  constexpr bool kReal = (MODE == 0 || MODE == 1);

  // per-pass twiddle tables: shared memory when they fit, else straight from global (L1/L2)
  const C* tw = a.tw_pass;
  if (a.tw_in_smem) {
    C* stw = s + padded(m) + 1;
    for (int i = t; i < a.tw_total; i += blockDim.x) stw[i] = a.tw_pass[i];
    tw = stw;
    __syncthreads();
  }

  for (long long item = blockIdx.x; item < a.n_items; item += gridDim.x) {
    const long long fl = item / a.r_top;
    const int rho = static_cast<int>(item - fl * a.r_top);
    const long long f = a.first + fl;
    C tmp[16];
    // ---- first pass: operands come straight from global memory, all 16 loads in flight ----
    if (active) {
      if constexpr (kReal) {
        const long long c = f / a.n_frames;
        const long long fr = f - c * a.n_frames;
        const T* xrow = a.x + c * a.x_stride;
        const long long fstart = a.offset + fr * a.hop;
        const bool fast = (fstart + 2LL * a.nc <= a.n_valid) &&
                          (((reinterpret_cast<uintptr_t>(xrow + fstart)) % (2 * sizeof(T))) == 0);
        if (fast) {
          const C* xp = reinterpret_cast<const C*>(xrow + fstart);
#pragma unroll
          for (int u = 0; u < 16; ++u) tmp[u] = xp[static_cast<long long>(a.r_top) * (t + u * q16) + rho];
        } else {
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const long long s0 = fstart + 2 * (static_cast<long long>(a.r_top) * (t + u * q16) + rho);
            tmp[u].x = s0 < a.n_valid ? xrow[s0] : T(0);
            tmp[u].y = s0 + 1 < a.n_valid ? xrow[s0 + 1] : T(0);
          }
        }
        if (a.window) {
          const C* wp = reinterpret_cast<const C*>(a.window);
          C wv[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) wv[u] = wp[static_cast<long long>(a.r_top) * (t + u * q16) + rho];
#pragma unroll
          for (int u = 0; u < 16; ++u) { tmp[u].x *= wv[u].x; tmp[u].y *= wv[u].y; }
        }
      } else {
        const C* xp = reinterpret_cast<const C*>(a.x) + f * a.nc;
#pragma unroll
        for (int u = 0; u < 16; ++u) tmp[u] = xp[static_cast<long long>(a.r_top) * (t + u * q16) + rho];
      }
    }
    __syncthreads();   // previous item's readers are done with s
    int ns = 1;
    for (int pass = 0; pass < a.n_pass; ++pass) {
      const int R = a.radix[pass];
      if (pass > 0) {
        if (active) {
#pragma unroll
          for (int u = 0; u < 16; ++u) tmp[u] = s[padded(t + u * q16)];
        }
        __syncthreads();
      }
      if (active) run_pass<T>(R, tmp, s, tw + a.tw_offset[pass], t, m, ns);
      __syncthreads();
      ns *= R;
    }
    // ---- epilogue ----------------------------------------------------------
    if constexpr (MODE == 0) {
      const long long c = f / a.n_frames;
      const long long fr = f - c * a.n_frames;
      T* mg = a.mag + c * a.mag_channel_stride + fr * a.mag_frame_stride;
      const int nc = a.nc;
      for (int k = t; k <= nc / 2; k += blockDim.x) {
        const C A = s[padded(k)];
        const C Bc = cconj(s[padded((nc - k) & (nc - 1))]);
        C xe, xo;
        xe.x = T(0.5) * (A.x + Bc.x); xe.y = T(0.5) * (A.y + Bc.y);
        // -i/2 * (A - B)
        xo.x = T(0.5) * (A.y - Bc.y); xo.y = T(-0.5) * (A.x - Bc.x);
        const C tt = cmul(a.tw_post[k], xo);
        const C p = cadd(xe, tt), q = csub(xe, tt);
        mg[k] = finish_mag(p.x * p.x + p.y * p.y, a.db);
        mg[nc - k] = finish_mag(q.x * q.x + q.y * q.y, a.db);
      }
    } else if constexpr (MODE == 2) {
      C* o = a.out + f * a.nc;
      for (int k = t; k < m; k += blockDim.x) o[k] = s[padded(k)];
    } else {
      C* o = a.out + item * m;
      for (int k = t; k < m; k += blockDim.x) o[k] = s[padded(k)];
    }
    // the __syncthreads at the top of the next item protects s
  }
}

// ---- compile-time sized variant ---------------------------------------------
// For the sizes that matter (n_fft 1024..16384 on chip, 2^16 split 4 x 8192) the
// radix plan, strides and padded shared-memory offsets are compile-time, so the
// address arithmetic folds into immediates.
template <int M> struct CtPlan;
template <> struct CtPlan<128>  { static constexpr int NP = 2; static constexpr int R[4] = {8, 16, 1, 1}; };
template <> struct CtPlan<256>  { static constexpr int NP = 2; static constexpr int R[4] = {16, 16, 1, 1}; };
template <> struct CtPlan<512>  { static constexpr int NP = 3; static constexpr int R[4] = {2, 16, 16, 1}; };
template <> struct CtPlan<1024> { static constexpr int NP = 3; static constexpr int R[4] = {4, 16, 16, 1}; };
template <> struct CtPlan<2048> { static constexpr int NP = 3; static constexpr int R[4] = {8, 16, 16, 1}; };
template <> struct CtPlan<4096> { static constexpr int NP = 3; static constexpr int R[4] = {16, 16, 16, 1}; };
template <> struct CtPlan<8192> { static constexpr int NP = 4; static constexpr int R[4] = {2, 16, 16, 16}; };

// VAR bit 0: radix-16 twiddles W^r, r = 1..15, from W^1, W^2, W^4, W^8 by a product tree (<= 3
//            roundings deep) instead of 15 shared-memory reads; bit 1: Hann by angle addition instead
//            of a table read per sample; bit 2: the last pass stores without padding (its stores and
//            the epilogue's mirrored reads are unit-stride, the pad only costs them a 2-way conflict).
struct NoHook { __device__ __forceinline__ void operator()() const {} };

// `after_gather` runs in the LAST pass once the operands have left `tmp` for the butterfly registers: the caller's
// chance to put the next item's loads in flight (into `tmp`) a whole pass before they are needed.
template <typename T, int M, int R, int NS, int VAR, typename Hook>
__device__ __forceinline__ void ct_pass(const typename Cpx<T>::type* tmp, typename Cpx<T>::type* s,
                                        const typename Cpx<T>::type* __restrict__ tw, const int t, Hook&& after_gather,
                                        typename Cpx<T>::type* keep) {
  typedef typename Cpx<T>::type C;
  constexpr int Q = M / 16, B = 16 / R;
  C v[16];
  gather_butterflies<T, R>(v, tmp);
  if constexpr (NS * R == M) after_gather();
#pragma unroll
  for (int b = 0; b < B; ++b) {
    const int j = t + b * Q;
    const int k = j & (NS - 1);
    C* w = v + b * R;
    if constexpr (NS > 1 && R == 16 && (VAR & 1) != 0) {
      const C* row = tw + k * kTreePitch;
      const C w1 = row[0], w2 = row[1], w4 = row[2], w8 = row[3];
      const C w3 = cmul(w1, w2);
      w[1] = cmul(w[1], w1); w[2] = cmul(w[2], w2); w[3] = cmul(w[3], w3); w[4] = cmul(w[4], w4);
      const C w5 = cmul(w1, w4), w6 = cmul(w2, w4), w7 = cmul(w3, w4);
      w[5] = cmul(w[5], w5); w[6] = cmul(w[6], w6); w[7] = cmul(w[7], w7); w[8] = cmul(w[8], w8);
      w[9] = cmul(w[9], cmul(w1, w8)); w[10] = cmul(w[10], cmul(w2, w8));
      w[11] = cmul(w[11], cmul(w3, w8)); w[12] = cmul(w[12], cmul(w4, w8));
      w[13] = cmul(w[13], cmul(w5, w8)); w[14] = cmul(w[14], cmul(w6, w8));
      w[15] = cmul(w[15], cmul(w7, w8));
    } else if constexpr (NS > 1) {
      const C* row = tw + k * kTwPitch;
      C wv[R];
#pragma unroll
      for (int r = 1; r < R; ++r) wv[r] = row[r];
#pragma unroll
      for (int r = 1; r < R; ++r) w[r] = cmul(w[r], wv[r]);
    }
    Dft<T, R>::run(w);
    const int base = (j - k) * R + k;
    if constexpr ((VAR & 32) != 0 && (VAR & 4) != 0 && NS * R == M && B == 1 && R == 16) {
      // VAR bit 5: the real split pairs Z[t + i Q], i < 8 (this thread's own first eight outputs) with
      // Z[M - t - i Q] (the LAST eight outputs of thread Q - t): only r >= 8 is published, r < 8 stays in registers
      // (thread 0 is its own partner and also needs its r = 0 as Z[M] = Z[0])
      C* sp = s + base;
#pragma unroll
      for (int r = 8; r < R; ++r) sp[r * NS] = w[r];
      if (t == 0) sp[0] = w[0];
#pragma unroll
      for (int r = 0; r < 8; ++r) keep[r] = w[r];
    } else if constexpr ((VAR & 4) != 0 && NS * R == M && B == 1) {
      C* sp = s + base;                                 // last pass: base == t, unit stride across the warp
#pragma unroll
      for (int r = 0; r < R; ++r) sp[r * NS] = w[r];
    } else {
      C* sp = s + base + (base >> kPadShift);           // padded(base + r*NS) = this + static offset
#pragma unroll
      for (int r = 0; r < R; ++r) sp[r * NS + ((r * NS) >> kPadShift)] = w[r];
    }
  }
}

template <typename T, int M, int P, int NS, int VAR = 0, typename Hook = NoHook>
__device__ __forceinline__ void ct_passes(typename Cpx<T>::type* tmp, typename Cpx<T>::type* s,
                                          const typename Cpx<T>::type* tw, const int* tw_offset, const int t,
                                          Hook&& after_gather = Hook(), typename Cpx<T>::type* keep = nullptr) {
  typedef typename Cpx<T>::type C;
  if constexpr (P < CtPlan<M>::NP) {
    constexpr int R = CtPlan<M>::R[P];
    constexpr int Q = M / 16;
    if constexpr (P > 0) {
      if constexpr (Q % 16 == 0) {
        const C* lp = s + t + (t >> kPadShift);
#pragma unroll
        for (int u = 0; u < 16; ++u) tmp[u] = lp[u * (Q + Q / 16)];
      } else {
#pragma unroll
        for (int u = 0; u < 16; ++u) tmp[u] = s[padded(t + u * Q)];
      }
      __syncthreads();
    }
    ct_pass<T, M, R, NS, VAR>(tmp, s, tw + tw_offset[P], t, after_gather, keep);   // tw_offset: the table's own offsets (full or tree)
    __syncthreads();
    ct_passes<T, M, P + 1, NS * R, VAR>(tmp, s, tw, tw_offset, t, after_gather, keep);
  }
}

template <typename T, int M, int VAR = 0> struct CtBounds {
  static constexpr int Q = M / 16;
  static constexpr int kThreads = (VAR & 8) ? 768 : 640;   // resident threads per SM aimed at
  // float64: the product-tree form (VAR bit 0) holds 4 instead of 15 twiddles per pass in registers, so three CTAs fit
  static constexpr int kThreads64 = (VAR & 1) ? 384 : 256;
  static constexpr int kMinBlocks = sizeof(T) == 4 ? (kThreads / Q > 0 ? (kThreads / Q > 16 ? 16 : kThreads / Q) : 1)
                                                   : (kThreads64 / Q > 0 ? (kThreads64 / Q > 8 ? 8 : kThreads64 / Q) : 1);
};

template <typename T> __device__ __forceinline__ T mag_sqrt(T v) { return sqrt(v); }
template <> __device__ __forceinline__ float mag_sqrt<float>(float v) {   // MUFU.SQRT: ~1 ulp, no slow path
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

template <typename T, int MODE, int M, bool kTwSmem, int VAR = 0>
__global__ void __launch_bounds__(M / 16, (CtBounds<T, M, VAR>::kMinBlocks))
fft_fixed_kernel(const FftArgs<T> a) {
  typedef typename Cpx<T>::type C;
  constexpr int Q = M / 16;
  constexpr bool kReal = (MODE == 0 || MODE == 1);
  constexpr bool kSplit = (MODE == 1 || MODE == 3);   // part of a top-level radix split (r_top > 1)
  constexpr bool kHannFly = (VAR & 2) != 0 && MODE == 0 && sizeof(T) == 4;
  constexpr bool kFlatLast = (VAR & 4) != 0 && CtPlan<M>::R[CtPlan<M>::NP - 1] == 16;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* s = reinterpret_cast<C*>(smem_raw);
  const int t = threadIdx.x;
  C* stw = s + padded(M) + 1;
  constexpr bool kTree = (VAR & 1) != 0;
  const C* tw_src = kTree ? a.tw_tree : a.tw_pass;
  const int tw_count = kTree ? a.tw_tree_total : a.tw_total;
  const int* tw_offset = kTree ? a.tw_tree_offset : a.tw_offset;
  if constexpr (kTwSmem) {
    for (int i = t; i < tw_count; i += Q) stw[i] = tw_src[i];
    __syncthreads();
  }
  const C* tw = kTwSmem ? stw : tw_src;
  // MODE 0 keeps the real-split twiddles W_2M^k, k <= M/2, in shared memory as well
  const C* twp = a.tw_post;
  if constexpr (MODE == 0 && kTwSmem) {
    C* sp = stw + tw_count;
    for (int i = t; i <= M / 2; i += Q) sp[i] = a.tw_post[i];
    twp = sp;
    __syncthreads();
  }

  // raw operands of one item, all 16 loads in flight (no window yet)
  auto fetch = [&](long long item, C* tmp, long long c_item, long long fr_item) {
    long long fl = item;
    int rho = 0;
    if constexpr (kSplit) {
      fl = item / a.r_top;
      rho = static_cast<int>(item - fl * a.r_top);
    }
    const long long f = a.first + fl;
    if constexpr (kReal) {
      long long c = c_item, fr = fr_item;
      if constexpr (kSplit) {
        c = f / a.n_frames;
        fr = f - c * a.n_frames;
      }
      const T* xrow = a.x + c * a.x_stride;
      const long long fstart = a.offset + fr * a.hop;
      const bool fast = (fstart + 2LL * a.nc <= a.n_valid) &&
                        (((reinterpret_cast<uintptr_t>(xrow + fstart)) % (2 * sizeof(T))) == 0);
      if (fast) {
        if constexpr (!kSplit) {
          const C* xp = reinterpret_cast<const C*>(xrow + fstart) + t;
#pragma unroll
          for (int u = 0; u < 16; ++u) tmp[u] = xp[u * Q];
        } else {
          const C* xp = reinterpret_cast<const C*>(xrow + fstart) + static_cast<long long>(a.r_top) * t + rho;
#pragma unroll
          for (int u = 0; u < 16; ++u) tmp[u] = xp[static_cast<long long>(a.r_top) * (u * Q)];
        }
      } else {
        // tail / misaligned frame: guarded scalar loads, 32-bit index arithmetic relative to the frame
        const long long left = a.n_valid - fstart;
        const int rem = left > 0x7fffffff ? 0x7fffffff : (left < 0 ? 0 : static_cast<int>(left));
        const T* xf = xrow + fstart;
        const int e0 = 2 * (a.r_top * t + rho), de = 2 * a.r_top * Q;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int e = e0 + u * de;
          tmp[u].x = e < rem ? xf[e] : T(0);
          tmp[u].y = e + 1 < rem ? xf[e + 1] : T(0);
        }
      }
    } else {
      if constexpr (!kSplit) {
        const C* xp = reinterpret_cast<const C*>(a.x) + f * a.nc + t;
#pragma unroll
        for (int u = 0; u < 16; ++u) tmp[u] = xp[u * Q];
      } else {
        const C* xp = reinterpret_cast<const C*>(a.x) + f * a.nc + static_cast<long long>(a.r_top) * t + rho;
#pragma unroll
        for (int u = 0; u < 16; ++u) tmp[u] = xp[static_cast<long long>(a.r_top) * (u * Q)];
      }
    }
  };

  C hann_a = {T(0), T(0)}, hann_b = {T(0), T(0)};
  if constexpr (kHannFly) {
    if (a.window) {
      hann_a = reinterpret_cast<const C*>(a.hann_ab)[2 * t];
      hann_b = reinterpret_cast<const C*>(a.hann_ab)[2 * t + 1];
    }
  }
  C tmp[16];
  // (channel, frame) of the item in flight, advanced by the grid stride without a 64-bit division per frame
  long long c_next = 0, fr_next = 0, dc = 0, dfr = 0;
  if constexpr (kReal && !kSplit) {
    const long long f0 = a.first + blockIdx.x;
    c_next = f0 / a.n_frames;
    fr_next = f0 - c_next * a.n_frames;
    dc = static_cast<long long>(gridDim.x) / a.n_frames;
    dfr = static_cast<long long>(gridDim.x) - dc * a.n_frames;
  }
  if (blockIdx.x < a.n_items) fetch(blockIdx.x, tmp, c_next, fr_next);
  for (long long item = blockIdx.x; item < a.n_items; item += gridDim.x) {
    long long fl = item;
    int rho = 0;
    if constexpr (kSplit) {
      fl = item / a.r_top;
      rho = static_cast<int>(item - fl * a.r_top);
    }
    const long long f = a.first + fl;
    long long c = c_next, fr = fr_next;
    if constexpr (kReal) {
      if constexpr (kSplit) {
        c = f / a.n_frames;
        fr = f - c * a.n_frames;
      }
      if constexpr (kHannFly) {
        if (a.window) {
          const C half = {T(0.5), T(0.5)};
#pragma unroll
          for (int u = 0; u < 16; ++u)
            tmp[u] = pmul(tmp[u], fma2(hann_a, a.hann_cc[u], fma2(hann_b, a.hann_ss[u], half)));
        }
      } else if (a.window) {
        C wv[16];
        if constexpr (!kSplit) {
          const C* wp = reinterpret_cast<const C*>(a.window) + t;
#pragma unroll
          for (int u = 0; u < 16; ++u) wv[u] = wp[u * Q];
        } else {
          const C* wp = reinterpret_cast<const C*>(a.window) + static_cast<long long>(a.r_top) * t + rho;
#pragma unroll
          for (int u = 0; u < 16; ++u) wv[u] = wp[static_cast<long long>(a.r_top) * (u * Q)];
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) tmp[u] = pmul(tmp[u], wv[u]);
      }
    }
    __syncthreads();   // previous item's readers are done with s
    // the next item's loads are put in flight inside the last pass, as soon as its operands have left `tmp`: they
    // have that pass's butterflies and the whole epilogue to land (issued after the passes they were still the
    // hottest stall of the kernel: 12 % of its samples waiting on the first use of a frame)
    auto prefetch = [&]() {
      if (item + gridDim.x < a.n_items) {
        if constexpr (kReal && !kSplit) {
          fr_next += dfr;
          c_next += dc;
          if (fr_next >= a.n_frames) { fr_next -= a.n_frames; ++c_next; }
        }
        fetch(item + gridDim.x, tmp, c_next, fr_next);
      }
    };
    constexpr bool kKeep = (VAR & 32) != 0 && kFlatLast && MODE == 0;
    C keep[kKeep ? 8 : 1];
    if constexpr ((VAR & 16) != 0) {
      ct_passes<T, M, 0, 1, VAR>(tmp, s, tw, tw_offset, t);
      prefetch();
    } else {
      ct_passes<T, M, 0, 1, VAR>(tmp, s, tw, tw_offset, t, prefetch, keep);
    }
    if constexpr (MODE == 0) {
      T* mg = a.mag + c * a.mag_channel_stride + fr * a.mag_frame_stride;
      constexpr int NC = M;   // r_top == 1
      // k = t + i*Q covers 0 .. NC/2 - 1; its partner NC - k (0 for k == 0) is read and written through
      // pointers based at NC - t, so every access of the loop is base + compile-time offset
      const C* lp = kFlatLast ? s + t : s + t + (t >> kPadShift);
      const C* lb = kFlatLast ? s + (NC - t) : s + (NC - t) + ((NC - t) >> kPadShift);
      const C* lb0 = (t == 0) ? s : lb;
      T* mlo = mg + t;
      T* mhi = mg + (NC - t);
      auto split = [&](auto db_tag) {
        constexpr bool kDb = decltype(db_tag)::value;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          C A;
          if constexpr (kKeep) A = keep[i]; else A = lp[kFlatLast ? i * Q : i * (Q + Q / 16)];
          const C Bc = cconj(i == 0 ? lb0[0] : lb[kFlatLast ? -(i * Q) : -(i * (Q + Q / 16))]);
          const C xe = pscale(cadd(A, Bc), T(0.5));              // (A + conj B) / 2
          const C xo = pscale(mul_neg_i(csub(A, Bc)), T(0.5));   // -i (A - conj B) / 2
          const C tt = cmul(xo, twp[t + i * Q]);
          const C pp = cadd(xe, tt), qq = csub(xe, tt);
          mlo[i * Q] = finish_mag(pp.x * pp.x + pp.y * pp.y, kDb ? 1 : 0);
          mhi[-(i * Q)] = finish_mag(qq.x * qq.x + qq.y * qq.y, kDb ? 1 : 0);
        }
        if (t == 0) {                          // k = NC/2: both magnitudes coincide
          const C A = s[kFlatLast ? NC / 2 : padded(NC / 2)];
          mg[NC / 2] = finish_mag(A.x * A.x + A.y * A.y, kDb ? 1 : 0);
        }
      };
      if (a.db) split(std::true_type{}); else split(std::false_type{});
    } else if constexpr (MODE == 2) {
      C* o = a.out + f * a.nc;
      const C* lp = kFlatLast ? s + t : s + t + (t >> kPadShift);
#pragma unroll
      for (int u = 0; u < 16; ++u) o[t + u * Q] = lp[kFlatLast ? u * Q : u * (Q + Q / 16)];
    } else {
      C* o = a.out + item * M;
      const C* lp = kFlatLast ? s + t : s + t + (t >> kPadShift);
#pragma unroll
      for (int u = 0; u < 16; ++u) o[t + u * Q] = lp[kFlatLast ? u * Q : u * (Q + Q / 16)];
    }
  }
}

// Last radix-R level of a split transform, fused with the real split + |.|
// (kReal) or writing the complex result.  ws: [transform][rho][m].
template <typename T, int R, bool kReal>
__global__ void __launch_bounds__(256)
fft_combine_kernel(const FftArgs<T> a, const typename Cpx<T>::type* __restrict__ ws, long long n_transforms) {
  typedef typename Cpx<T>::type C;
  const int m = a.m, nc = a.nc;
  const int per = kReal ? (m / 2 + 1) : m;
  const long long total = n_transforms * per;
  for (long long id = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; id < total;
       id += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long fl = id / per;
    const int k = static_cast<int>(id - fl * per);
    const C* w = ws + fl * nc;
    const long long f = a.first + fl;
    C za[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      za[r] = w[static_cast<long long>(r) * m + k];
      if (r > 0) za[r] = cmul(za[r], a.tw_top[r * k]);
    }
    Dft<T, R>::run(za);
    if constexpr (!kReal) {
      C* o = a.out + f * nc;
#pragma unroll
      for (int s = 0; s < R; ++s) o[k + s * m] = za[s];
    } else {
      const int k2 = (m - k) & (m - 1);
      C zb[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        zb[r] = w[static_cast<long long>(r) * m + k2];
        if (r > 0) zb[r] = cmul(zb[r], a.tw_top[r * k2]);
      }
      Dft<T, R>::run(zb);
      const long long c = f / a.n_frames;
      const long long fr = f - c * a.n_frames;
      T* mg = a.mag + c * a.mag_channel_stride + fr * a.mag_frame_stride;
      auto emit = [&](const C A, const C Bz, int idx) {
        if (idx > nc / 2) return;
        const C Bc = cconj(Bz);
        C xe, xo;
        xe.x = T(0.5) * (A.x + Bc.x); xe.y = T(0.5) * (A.y + Bc.y);
        xo.x = T(0.5) * (A.y - Bc.y); xo.y = T(-0.5) * (A.x - Bc.x);
        const C tt = cmul(a.tw_post[idx], xo);
        const C p = cadd(xe, tt), q = csub(xe, tt);
        mg[idx] = finish_mag(p.x * p.x + p.y * p.y, a.db);
        mg[nc - idx] = finish_mag(q.x * q.x + q.y * q.y, a.db);
      };
#pragma unroll
      for (int s = 0; s < R; ++s) {
        const int sp = (k == 0) ? ((R - s) & (R - 1)) : (R - 1 - s);
        emit(za[s], zb[sp], k + s * m);
        if (k != 0 && k != k2) emit(zb[s], za[sp], k2 + s * m);
      }
    }
  }
}

// ---- four-step form of long real transforms (2^15 .. 2^18 points, config C4) ----
// nc = N1 x 128 complex points, z[n1*128 + n2].  Kernel 1 transforms 16 adjacent
// columns per CTA (N1-point transforms over n1; a warp reads 2 rows x 128 B, so
// every sector is used once), multiplies by W_nc^(n2 k1) and writes ws[k1][n2].
// Kernel 2 transforms 32 rows per CTA (128-point transforms over n2): 16 rows k1
// and their mirror rows N1-k1, which is what the real split needs to turn
// Z[k1 + N1 k2] / Z[nc - k1 - N1 k2] into |X| with 64-byte coalesced stores.
constexpr int kFsCols = 128;     // N2
constexpr int kFsTile = 16;      // columns per CTA (kernel 1), primary rows per CTA (kernel 2)

template <typename T> struct FourStep {
  typedef typename Cpx<T>::type C;
  const C* tw_cols; int tw_cols_offset[8]; int tw_cols_total;   // N1-point passes
  const C* tw_rows; int tw_rows_offset[8]; int tw_rows_total;   // 128-point passes
  const C* tw_hi;                            // W_nc^(256 q), q < nc/256
  const C* tw_lo;                            // W_nc^r, r < 256
  C* ws;
};

template <int M> __host__ __device__ constexpr int fs_pitch() { return (M + (M >> kPadShift)) | 1; }
// Row slots of kernel 2: a pitch that is a multiple of 16 elements plus a skew 0,8,1,9,...,7,15.
// Adjacent rows (one half-warp during the passes) sit 8 elements apart modulo 16, so a run of 8
// consecutive elements per row is conflict-free; the 16 rows a half-warp reads at one column in
// the real split all start at different offsets modulo 16.
constexpr int kFsRowPitch = 160;
__host__ __device__ __forceinline__ constexpr int fs_row_base(int slot) {
  return slot * kFsRowPitch + ((slot & 1) << 3) + ((slot & 15) >> 1);
}

// Workspace accesses of the fused form go to L2 (.cg): the slot of a cluster is rewritten for every transform by
// other SMs, so a line left in this SM's L1 by the previous transform would be stale.
template <bool kCg, typename C> __device__ __forceinline__ C ws_load(const C* p) {
  if constexpr (kCg) return __ldcg(p); else return *p;
}
template <bool kCg, typename C> __device__ __forceinline__ void ws_store(C* p, C v) {
  if constexpr (kCg) __stcg(p, v); else *p = v;
}

// Columns [16 tile, 16 tile + 16) of transform f: N1-point transforms along n1, times W_nc^(n2 k1), into wsf[k1][n2].
// N1 threads.  s: kFsTile * PITCH points (the caller keeps other users of it away until the call returns and
// synchronises before the next use).
template <typename T, int N1, bool kCg>
__device__ __forceinline__ void fs_cols_tile(const FftArgs<T>& a, const FourStep<T>& fs, typename Cpx<T>::type* s,
                                             const typename Cpx<T>::type* s_hi, const typename Cpx<T>::type* s_lo,
                                             const typename Cpx<T>::type* tw_cols, const long long f, const int tile,
                                             typename Cpx<T>::type* wsf, const int t) {
  typedef typename Cpx<T>::type C;
  constexpr int Q = N1 / 16, NC = N1 * kFsCols, PITCH = fs_pitch<N1>();
  const int c = t % kFsTile, tp = t / kFsTile;
  const int n2 = tile * kFsTile + c;
  const long long ch = f / a.n_frames;
  const long long fr = f - ch * a.n_frames;
  const long long fstart = a.offset + fr * a.hop;
  const T* xf = a.x + ch * a.x_stride + fstart;
  const long long left = a.n_valid - fstart;
  C tmp[16];
  if (left >= 2LL * NC && (reinterpret_cast<uintptr_t>(xf) % (2 * sizeof(T))) == 0) {
    const C* xp = reinterpret_cast<const C*>(xf) + tp * kFsCols + n2;
#pragma unroll
    for (int u = 0; u < 16; ++u) tmp[u] = xp[u * Q * kFsCols];
  } else {
    const int rem = left > 0x7fffffff ? 0x7fffffff : (left < 0 ? 0 : static_cast<int>(left));
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int e = 2 * ((tp + u * Q) * kFsCols + n2);
      tmp[u].x = e < rem ? xf[e] : T(0);
      tmp[u].y = e + 1 < rem ? xf[e + 1] : T(0);
    }
  }
  if (a.window) {
    const C* wp = reinterpret_cast<const C*>(a.window) + tp * kFsCols + n2;
#pragma unroll
    for (int u = 0; u < 16; ++u) tmp[u] = pmul(tmp[u], wp[u * Q * kFsCols]);
  }
  C* sc = s + c * PITCH;
  ct_passes<T, N1, 0, 1>(tmp, sc, tw_cols, fs.tw_cols_offset, tp);
  // W_nc^(n2 k1), k1 = tp + u Q:  W^(n2 tp) * (W^(n2 Q))^u.  Two table products give the base and
  // the step; the powers of the step come from a product tree (<= 4 roundings deep) instead of
  // sixteen scattered table reads per thread, which were the busiest user of the load/store pipe.
  auto lookup = [&](int q) { return cmul(s_hi[q >> 8], s_lo[q & 255]); };
  const C wb = lookup(n2 * tp);
  C pw[16];
  pw[1] = lookup(n2 * Q);
  pw[2] = cmul(pw[1], pw[1]); pw[3] = cmul(pw[2], pw[1]); pw[4] = cmul(pw[2], pw[2]);
  pw[5] = cmul(pw[4], pw[1]); pw[6] = cmul(pw[4], pw[2]); pw[7] = cmul(pw[4], pw[3]);
  pw[8] = cmul(pw[4], pw[4]);
#pragma unroll
  for (int u = 9; u < 16; ++u) pw[u] = cmul(pw[8], pw[u - 8]);
  C* o = wsf + n2;
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    const int k1 = tp + u * Q;
    const C w = u == 0 ? wb : cmul(wb, pw[u]);
    ws_store<kCg>(o + static_cast<long long>(k1) * kFsCols, cmul(sc[padded(k1)], w));
  }
}

// Rows of group g of transform f (256 threads, 32 row slots): the 16 primary rows k1 = 16 g + 1 .. 16 g + 16
// and their mirrors N1 - k1, 128-point transforms along n2, real split, |X| out.  Row N1/2 (the last primary row
// of the last group) is its own mirror, so its mirror slot carries row 0 -- which pairs with itself too -- instead.
template <typename T, int N1, bool kCg>
__device__ __forceinline__ void fs_rows_group(const FftArgs<T>& a, const FourStep<T>& fs, typename Cpx<T>::type* s,
                                              const typename Cpx<T>::type* s_tw, const long long f, const int g,
                                              const typename Cpx<T>::type* wsf, const int t) {
  typedef typename Cpx<T>::type C;
  constexpr int M = kFsCols, Q = M / 16, NC = N1 * M;
  constexpr int G = N1 / 2 / kFsTile;     // groups of 16 primary rows k1 in [1, N1/2]
  const int tp = t % Q, slot = t / Q;     // 32 row slots: 0..15 primary rows, 16..31 their mirrors (ascending)
  const bool last = g == G - 1;
  int row = slot < kFsTile ? (kFsTile * g + 1 + slot) : (N1 - kFsTile * g - 2 * kFsTile + slot);
  if (last && slot == kFsTile) row = 0;
  C tmp[16];
  {
    const C* wp = wsf + static_cast<long long>(row) * M + tp;
#pragma unroll
    for (int u = 0; u < 16; ++u) tmp[u] = ws_load<kCg>(wp + u * Q);
  }
  // the real-split twiddles of this thread's bins, requested before the passes
  C twp[M / 16];
  const int k1 = kFsTile * g + 1 + (t % kFsTile);
#pragma unroll
  for (int i = 0; i < M / 16; ++i) {
    const int idx = k1 + N1 * (t / kFsTile + 16 * i);
    twp[i] = a.tw_post[idx <= NC / 2 ? idx : NC - idx];
  }
  ct_passes<T, M, 0, 1>(tmp, s + fs_row_base(slot), s_tw, fs.tw_rows_offset, tp);

  const long long ch = f / a.n_frames;
  const long long fr = f - ch * a.n_frames;
  T* mg = a.mag + ch * a.mag_channel_stride + fr * a.mag_frame_stride;
  auto emit = [&](const C A, const C Bz, int idx, const C wpost) {   // A = Z[idx], Bz = Z[nc - idx], idx <= nc/2
    const C Bc = cconj(Bz);
    C xe, xo;
    xe.x = T(0.5) * (A.x + Bc.x); xe.y = T(0.5) * (A.y + Bc.y);
    xo.x = T(0.5) * (A.y - Bc.y); xo.y = T(-0.5) * (A.x - Bc.x);
    const C tt = cmul(wpost, xo);
    const C p = cadd(xe, tt), q = csub(xe, tt);
    mg[idx] = finish_mag(p.x * p.x + p.y * p.y, a.db);
    mg[NC - idx] = finish_mag(q.x * q.x + q.y * q.y, a.db);
  };
  {
    const int j = t % kFsTile;
    const bool self = 2 * k1 == N1;                          // row N1/2: Z[nc - idx] is in the same row
    const C* pa = s + fs_row_base(j);
    const C* pb = s + fs_row_base(self ? j : 2 * kFsTile - 1 - j);  // row N1 - k1
#pragma unroll
    for (int i = 0; i < M / 16; ++i) {
      const int k2 = t / kFsTile + 16 * i;
      const C A = pa[padded(k2)];
      const C Bz = pb[padded(M - 1 - k2)];
      const int idx = k1 + N1 * k2;
      if (idx <= NC / 2) {
        if (!self || k2 < M / 2) emit(A, Bz, idx, twp[i]);   // a self-paired row: take each pair once
      } else if (!self) {
        emit(Bz, A, NC - idx, twp[i]);
      }
    }
  }
  if (last && t <= M / 2) {                                  // row 0 pairs with itself, k2 <-> (M - k2) mod M
    const C* r0 = s + fs_row_base(kFsTile);
    emit(r0[padded(t)], r0[padded((M - t) & (M - 1))], N1 * t, a.tw_post[N1 * t]);
  }
}

template <typename T, int N1>
__global__ void __launch_bounds__(N1)
fft4_cols_kernel(const FftArgs<T> a, const FourStep<T> fs) {
  typedef typename Cpx<T>::type C;
  constexpr int NC = N1 * kFsCols, PITCH = fs_pitch<N1>();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* s = reinterpret_cast<C*>(smem_raw);
  C* s_hi = s + kFsTile * PITCH;
  C* s_lo = s_hi + NC / 256;
  // measured on C4: the pass twiddles are faster from shared memory in fp64 (1.94 vs 2.18 ms)
  // and from global/L1 in fp32 (1.65 vs 1.80 ms)
  constexpr bool kTwSmem = sizeof(T) == 8;
  C* s_tw = s_lo + 256;
  const int t = threadIdx.x;
  for (int i = t; i < NC / 256; i += N1) s_hi[i] = fs.tw_hi[i];
  for (int i = t; i < 256; i += N1) s_lo[i] = fs.tw_lo[i];
  if constexpr (kTwSmem) {
    for (int i = t; i < fs.tw_cols_total; i += N1) s_tw[i] = fs.tw_cols[i];
    __syncthreads();
  }
  constexpr int TILES = kFsCols / kFsTile;
  const long long fl = blockIdx.x / TILES;
  const int tile = static_cast<int>(blockIdx.x - fl * TILES);
  fs_cols_tile<T, N1, false>(a, fs, s, s_hi, s_lo, kTwSmem ? s_tw : fs.tw_cols, a.first + fl, tile, fs.ws + fl * NC, t);
}

template <typename T, int N1>
__global__ void __launch_bounds__(256)
fft4_rows_kernel(const FftArgs<T> a, const FourStep<T> fs) {
  typedef typename Cpx<T>::type C;
  constexpr int NC = N1 * kFsCols;
  constexpr int G = N1 / 2 / kFsTile;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* s = reinterpret_cast<C*>(smem_raw);
  C* s_tw = s + 2 * kFsTile * kFsRowPitch;
  const int t = threadIdx.x;
  for (int i = t; i < fs.tw_rows_total; i += 256) s_tw[i] = fs.tw_rows[i];
  __syncthreads();               // pass twiddles are in place
  const long long fl = blockIdx.x / G;
  const int g = static_cast<int>(blockIdx.x - fl * G);
  fs_rows_group<T, N1, false>(a, fs, s, s_tw, a.first + fl, g, fs.ws + fl * NC, t);
}

// The two steps in ONE launch with the workspace resident in L2 (C4: 7.4 GB of DRAM traffic for 3.2 GB algorithmic
// with the two kernels above, whose workspace of a whole launch goes out to HBM and comes back).  A cluster of
// kFsCluster CTAs of 256 threads walks transforms cluster-stride: each CTA transforms its share of the column tiles
// into the cluster's own workspace slot (nc points, rewritten for every transform, so its lines stay dirty in L2 and
// never need to reach DRAM), the cluster synchronises, each CTA transforms its share of the row groups, and the
// cluster synchronises again before the slot is overwritten.  N1 = 256 only (n_fft = 2^16: 256 threads serve both
// the N1-thread column step and the 256-thread row step).
constexpr int kFsCluster = 4;
template <typename T, int N1>
__global__ void __launch_bounds__(256, sizeof(T) == 4 ? 3 : 2)
fft4_fused_kernel(const FftArgs<T> a, const FourStep<T> fs, const long long count) {
  typedef typename Cpx<T>::type C;
  static_assert(N1 == 256, "the fused four-step form is built for N1 = 256");
  constexpr int NC = N1 * kFsCols, PITCH = fs_pitch<N1>();
  constexpr int TILES = kFsCols / kFsTile, G = N1 / 2 / kFsTile;
  constexpr int kWork = kFsTile * PITCH > 2 * kFsTile * kFsRowPitch ? kFsTile * PITCH : 2 * kFsTile * kFsRowPitch;
  constexpr bool kTwSmem = sizeof(T) == 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* s = reinterpret_cast<C*>(smem_raw);
  C* s_hi = s + kWork;
  C* s_lo = s_hi + NC / 256;
  C* s_twr = s_lo + 256;
  C* s_twc = s_twr + fs.tw_rows_total;
  const int t = threadIdx.x;
  for (int i = t; i < NC / 256; i += 256) s_hi[i] = fs.tw_hi[i];
  for (int i = t; i < 256; i += 256) s_lo[i] = fs.tw_lo[i];
  for (int i = t; i < fs.tw_rows_total; i += 256) s_twr[i] = fs.tw_rows[i];
  if constexpr (kTwSmem)
    for (int i = t; i < fs.tw_cols_total; i += 256) s_twc[i] = fs.tw_cols[i];
  __syncthreads();
  unsigned rank, n_clusters_x, cluster_x;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(n_clusters_x));
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(cluster_x));
  C* wsf = fs.ws + static_cast<long long>(cluster_x) * NC;
  for (long long fl = cluster_x; fl < count; fl += n_clusters_x) {
    const long long f = a.first + fl;
    for (int tile = static_cast<int>(rank); tile < TILES; tile += kFsCluster) {
      fs_cols_tile<T, N1, true>(a, fs, s, s_hi, s_lo, kTwSmem ? s_twc : fs.tw_cols, f, tile, wsf, t);
      __syncthreads();           // the tile's last reads of s are done before the next user writes it
    }
    // columns of every CTA of the cluster are in the slot (release / acquire at cluster scope)
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    for (int g = static_cast<int>(rank); g < G; g += kFsCluster) {
      fs_rows_group<T, N1, true>(a, fs, s, s_twr, f, g, wsf, t);
      __syncthreads();
    }
    // every CTA has read its rows: the slot may be overwritten
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}

// Direct DFT for tiny transforms (nc < 16): one thread per output bin.
template <typename T, bool kReal>
__global__ void __launch_bounds__(128)
fft_small_kernel(const FftArgs<T> a, const typename Cpx<T>::type* __restrict__ tw_full, int n_fft,
                 long long n_transforms) {
  typedef typename Cpx<T>::type C;
  const int per = kReal ? (n_fft / 2 + 1) : n_fft;
  const long long total = n_transforms * per;
  for (long long id = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; id < total;
       id += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long f = id / per;
    const int k = static_cast<int>(id - f * per);
    C acc; acc.x = T(0); acc.y = T(0);
    if constexpr (kReal) {
      const long long c = f / a.n_frames;
      const long long fr = f - c * a.n_frames;
      const T* xrow = a.x + c * a.x_stride;
      const long long fstart = a.offset + fr * a.hop;
      for (int t = 0; t < n_fft; ++t) {
        T xv = fstart + t < a.n_valid ? xrow[fstart + t] : T(0);
        if (a.window) xv *= a.window[t];
        const C w = tw_full[(t * k) & (n_fft - 1)];
        acc.x += xv * w.x;
        acc.y += xv * w.y;
      }
      a.mag[c * a.mag_channel_stride + fr * a.mag_frame_stride + k] = finish_mag(acc.x * acc.x + acc.y * acc.y, a.db);
    } else {
      const C* in = reinterpret_cast<const C*>(a.x) + f * n_fft;
      for (int t = 0; t < n_fft; ++t) acc = cadd(acc, cmul(in[t], tw_full[(t * k) & (n_fft - 1)]));
      a.out[f * n_fft + k] = acc;
    }
  }
}

struct FftSide {
  int nc = 0, m = 0, r_top = 1, n_pass = 0;
  int radix[8] = {0};
  void* d_tw_pass = nullptr;
  int tw_offset[8] = {0};
  int tw_total = 0;
  void* d_tw_tree = nullptr;
  int tw_tree_offset[8] = {0};
  int tw_tree_total = 0;
  void* d_tw_top = nullptr;
  void* d_tw_post = nullptr;
};

}  // namespace dspb200

struct dspb200_fft_plan {
  uint32_t magic = dspb200::kMagicFft;   // first member: checked by every entry point
  int n_fft, hann, db, dtype, device;
  dspb200::FftSide real_side, c2c_side;
  dspb200::FftSide fs_cols, fs_rows;   // four-step form of a long real transform: N1-point columns, 128-point rows
  int fs_n1 = 0;                        // 0: not available for this size / dtype
  void* d_fs_hi = nullptr;
  void* d_fs_lo = nullptr;
  void* d_window = nullptr;
  void* d_hann_ab = nullptr;   // on-the-fly Hann of the fixed-size kernel: per-thread (-cos/2, sin/2) pairs
  double hann_cos[16] = {0}, hann_sin[16] = {0};   // cos/sin(u * 2 pi 2Q/(N-1))
  void* d_tw_full = nullptr;   // W_N, for the direct small-size kernel
  dspb200::FftR32Plan r32;     // 4096-point fp32 frames: the 32-points-per-thread kernel (fft_r32.cu)
  dspb200::FftLong32Plan l32;  // 2^16-point fp32 frames: three radix-32 passes in one launch (fft_long32.cu)
};

namespace dspb200 {

static const long double kPiL = 3.14159265358979323846264338327950288L;

template <typename T>
static int upload_twiddles(int n, int count, void** dptr) {   // W_n[k], k = 0..count-1
  typedef typename Cpx<T>::type C;
  *dptr = nullptr;
  if (count <= 0) return DSPB200_OK;
  std::vector<C> h(static_cast<size_t>(count));
  for (int k = 0; k < count; ++k) {
    // exact octant reduction keeps every entry correctly rounded
    const long double ang = -2.0L * kPiL * static_cast<long double>(k) / static_cast<long double>(n);
    h[static_cast<size_t>(k)].x = static_cast<T>(cosl(ang));
    h[static_cast<size_t>(k)].y = static_cast<T>(sinl(ang));
  }
  DSP_CUDA(cudaMalloc(dptr, h.size() * sizeof(C)));
  DSP_CUDA(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(C), cudaMemcpyHostToDevice));
  return DSPB200_OK;
}

template <typename T>
static int build_side(FftSide& s, int nc, bool real) {
  s.nc = nc;
  if (nc < 16) return DSPB200_OK;   // served by the direct kernel
  const int sub_max = FftCfg<T>::kSubMax;
  s.m = nc <= sub_max ? nc : sub_max;
  s.r_top = nc / s.m;
  if (s.r_top > 16) return fail(DSPB200_ERR_UNSUPPORTED, "transform too long (complex length %d)", nc);
  int rest = s.m, n16 = 0;
  while (rest % 16 == 0 && rest > 1) { rest /= 16; ++n16; }
  s.n_pass = 0;
  if (rest > 1) s.radix[s.n_pass++] = rest;
  for (int i = 0; i < n16; ++i) s.radix[s.n_pass++] = 16;
  {
    typedef typename Cpx<T>::type C;
    std::vector<C> h;
    int ns = 1;
    for (int p = 0; p < s.n_pass; ++p) {
      const int R = s.radix[p];
      s.tw_offset[p] = static_cast<int>(h.size());
      if (ns > 1) {
        for (int k = 0; k < ns; ++k)
          for (int r = 0; r < kTwPitch; ++r) {
            C w; w.x = T(1); w.y = T(0);
            if (r >= 1 && r < R) {
              const long double ang = -2.0L * kPiL * static_cast<long double>(r) * static_cast<long double>(k) /
                                      (static_cast<long double>(ns) * static_cast<long double>(R));
              w.x = static_cast<T>(cosl(ang));
              w.y = static_cast<T>(sinl(ang));
            }
            h.push_back(w);
          }
      }
      ns *= R;
    }
    s.tw_total = static_cast<int>(h.size());
    if (!h.empty()) {
      DSP_CUDA(cudaMalloc(&s.d_tw_pass, h.size() * sizeof(C)));
      DSP_CUDA(cudaMemcpy(s.d_tw_pass, h.data(), h.size() * sizeof(C), cudaMemcpyHostToDevice));
    }
    // compact form for the product tree: W^(k), W^(2k), W^(4k), W^(8k) of every radix-16 pass
    std::vector<C> g;
    ns = 1;
    for (int p = 0; p < s.n_pass; ++p) {
      const int R = s.radix[p];
      s.tw_tree_offset[p] = static_cast<int>(g.size());
      if (ns > 1 && R == 16) {
        for (int k = 0; k < ns; ++k)
          for (int c = 0; c < kTreePitch; ++c) {
            C w; w.x = T(1); w.y = T(0);
            if (c < 4) {
              const long double ang = -2.0L * kPiL * static_cast<long double>(1 << c) * static_cast<long double>(k) /
                                      (static_cast<long double>(ns) * static_cast<long double>(R));
              w.x = static_cast<T>(cosl(ang));
              w.y = static_cast<T>(sinl(ang));
            }
            g.push_back(w);
          }
      }
      ns *= R;
    }
    s.tw_tree_total = static_cast<int>(g.size());
    if (!g.empty()) {
      DSP_CUDA(cudaMalloc(&s.d_tw_tree, g.size() * sizeof(C)));
      DSP_CUDA(cudaMemcpy(s.d_tw_tree, g.data(), g.size() * sizeof(C), cudaMemcpyHostToDevice));
    }
  }
  if (s.r_top > 1) DSP_TRY(upload_twiddles<T>(nc, nc, &s.d_tw_top));
  if (real) DSP_TRY(upload_twiddles<T>(2 * nc, nc / 2 + 1, &s.d_tw_post));
  return DSPB200_OK;
}

template <typename T>
static int plan_build(dspb200_fft_plan* p) {
  const int N = p->n_fft;
  DSP_TRY(build_side<T>(p->real_side, N / 2, true));
  {
    // a complex side that is too long only disables the complex entry points
    const int rc = build_side<T>(p->c2c_side, N, false);
    if (rc == DSPB200_ERR_UNSUPPORTED) p->c2c_side.nc = -1;
    else if (rc != DSPB200_OK) return rc;
  }
  if (N < 32) DSP_TRY(upload_twiddles<T>(N, N, &p->d_tw_full));
  if constexpr (sizeof(T) == 4) {
    DSP_TRY(fft_r32_build(N, p->r32));
    DSP_TRY(fft_long32_build(N, p->l32));
  }
  {
    const int nc = N / 2;
    const int n1 = nc / kFsCols;
    const int n1_max = sizeof(T) == 4 ? 1024 : 512;      // 16 columns of N1 points must fit shared memory
    if (p->real_side.r_top > 1 && n1 >= 128 && n1 <= n1_max) {
      typedef typename Cpx<T>::type C;
      DSP_TRY(build_side<T>(p->fs_cols, n1, false));
      DSP_TRY(build_side<T>(p->fs_rows, kFsCols, false));
      std::vector<C> hi(static_cast<size_t>(nc / 256)), lo(256);
      for (int q = 0; q < nc / 256; ++q) {
        const long double ang = -2.0L * kPiL * static_cast<long double>(q) * 256.0L / static_cast<long double>(nc);
        hi[static_cast<size_t>(q)].x = static_cast<T>(cosl(ang));
        hi[static_cast<size_t>(q)].y = static_cast<T>(sinl(ang));
      }
      for (int r = 0; r < 256; ++r) {
        const long double ang = -2.0L * kPiL * static_cast<long double>(r) / static_cast<long double>(nc);
        lo[static_cast<size_t>(r)].x = static_cast<T>(cosl(ang));
        lo[static_cast<size_t>(r)].y = static_cast<T>(sinl(ang));
      }
      DSP_CUDA(cudaMalloc(&p->d_fs_hi, hi.size() * sizeof(C)));
      DSP_CUDA(cudaMemcpy(p->d_fs_hi, hi.data(), hi.size() * sizeof(C), cudaMemcpyHostToDevice));
      DSP_CUDA(cudaMalloc(&p->d_fs_lo, lo.size() * sizeof(C)));
      DSP_CUDA(cudaMemcpy(p->d_fs_lo, lo.data(), lo.size() * sizeof(C), cudaMemcpyHostToDevice));
      p->fs_n1 = n1;
    }
  }
  if (p->hann) {
    std::vector<T> w(static_cast<size_t>(N));
    for (int k = 0; k < N; ++k) {
      // 0.5 - 0.5 cos(2 pi k / (N-1)); N == 1 gives 0/0 = NaN as in dsp_core.py:87
      const double ratio = static_cast<double>(k) / static_cast<double>(N - 1);
      w[static_cast<size_t>(k)] = static_cast<T>(0.5 - 0.5 * std::cos(2.0 * 3.14159265358979323846 * ratio));
    }
    DSP_CUDA(cudaMalloc(&p->d_window, w.size() * sizeof(T)));
    DSP_CUDA(cudaMemcpy(p->d_window, w.data(), w.size() * sizeof(T), cudaMemcpyHostToDevice));
    if (p->real_side.r_top == 1 && p->real_side.m >= 16 && N > 1) {
      // sample n = 2(t + uQ) + {0,1}:  w = 1/2 - cos(phi_t + u D)/2 = 1/2 + A cos(uD) + B sin(uD),
      // A = -cos(phi_t)/2, B = sin(phi_t)/2, phi_n = 2 pi n/(N-1), D = 2 pi 2Q/(N-1)
      const int Q = p->real_side.m / 16;
      const long double step = 2.0L * kPiL / static_cast<long double>(N - 1);
      std::vector<T> ab(static_cast<size_t>(Q) * 4);
      for (int t = 0; t < Q; ++t)
        for (int h = 0; h < 2; ++h) {
          const long double phi = step * static_cast<long double>(2 * t + h);
          ab[static_cast<size_t>(t) * 4 + h] = static_cast<T>(-0.5L * cosl(phi));
          ab[static_cast<size_t>(t) * 4 + 2 + h] = static_cast<T>(0.5L * sinl(phi));
        }
      for (int u = 0; u < 16; ++u) {
        const long double d = step * static_cast<long double>(2 * Q) * static_cast<long double>(u);
        p->hann_cos[u] = static_cast<double>(cosl(d));
        p->hann_sin[u] = static_cast<double>(sinl(d));
      }
      DSP_CUDA(cudaMalloc(&p->d_hann_ab, ab.size() * sizeof(T)));
      DSP_CUDA(cudaMemcpy(p->d_hann_ab, ab.data(), ab.size() * sizeof(T), cudaMemcpyHostToDevice));
    }
  }
  return DSPB200_OK;
}

// Per-chunk workspace cap of split transforms.  (An L2-sized cap, 32 MB, was measured SLOWER on C4:
// 3.8 ms vs 2.8 ms for 8192 frames -- the small launches cost more than the saved DRAM traffic.)
constexpr size_t kSplitWorkspaceBytes = size_t(2) << 30;
static int64_t split_chunk(const FftSide& s, size_t csize) {
  const int64_t c = static_cast<int64_t>(kSplitWorkspaceBytes / (static_cast<size_t>(s.nc) * csize));
  return c < 1 ? 1 : c;
}
static size_t side_workspace(const FftSide& s, int64_t n_transforms, size_t csize) {
  if (s.nc < 16 || s.r_top == 1) return 0;
  const int64_t chunk = split_chunk(s, csize);
  return static_cast<size_t>(n_transforms < chunk ? n_transforms : chunk) * s.nc * csize;
}

template <typename T>
static void fill_args(FftArgs<T>& a, const FftSide& s) {
  typedef typename Cpx<T>::type C;
  a.nc = s.nc; a.m = s.m; a.r_top = s.r_top; a.n_pass = s.n_pass;
  for (int i = 0; i < 8; ++i) a.radix[i] = s.radix[i];
  a.tw_pass = static_cast<const C*>(s.d_tw_pass);
  for (int i = 0; i < 8; ++i) a.tw_offset[i] = s.tw_offset[i];
  a.tw_total = s.tw_total;
  a.tw_tree = static_cast<const C*>(s.d_tw_tree);
  for (int i = 0; i < 8; ++i) a.tw_tree_offset[i] = s.tw_tree_offset[i];
  a.tw_tree_total = s.tw_tree_total;
  a.tw_in_smem = 0;
  a.tw_top = static_cast<const C*>(s.d_tw_top);
  a.tw_post = static_cast<const C*>(s.d_tw_post);
}

template <typename T, int MODE, int MAXT>
static int launch_stockham_t(FftArgs<T> a, cudaStream_t stream) {
  typedef typename Cpx<T>::type C;
  size_t smem = static_cast<size_t>(padded(a.m) + 1) * sizeof(C);
  const size_t tw_bytes = static_cast<size_t>(a.tw_total) * sizeof(C);
  a.tw_in_smem = (tw_bytes > 0 && tw_bytes <= 24 * 1024) ? 1 : 0;   // small tables ride along in shared memory
  if (a.tw_in_smem) smem += tw_bytes;
  auto kern = fft_stockham_kernel<T, MODE, MAXT>;
  DSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int threads = a.m / 16;
  if (threads < 32) threads = 32;
  int per_sm = 1;
  DSP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
  if (per_sm < 1) per_sm = 1;
  const long long cap = static_cast<long long>(sm_count()) * per_sm;
  const int grid = static_cast<int>(a.n_items < cap ? a.n_items : cap);
  kern<<<grid, threads, smem, stream>>>(a);
  return after_launch("fft_stockham_kernel");
}

// Variant of the fixed-size kernel per (type, mode): fp32 magnitude frames take the product-tree twiddles
// (compact tables: 31 KB instead of 44 KB of shared memory per CTA at 4096 points, which leaves the L1 its
// share of the SM's 256 KB), the angle-addition Hann and the unpadded last pass.  Measured on 4736 clips x
// 480000 samples, 4096-point frames: 4.31 ms -> 3.37 ms.  DSPB200_FFT_VAR overrides it for the 4096-point
// size (experiments: 0, 6, 7, 15, 23).
template <typename T, int MODE> struct FixedVar { static constexpr int value = MODE == 0 ? (sizeof(T) == 4 ? 7 : 5) : 0; };

template <typename T, int MODE, int M>
static int launch_fixed(FftArgs<T> a, cudaStream_t stream) {
  typedef typename Cpx<T>::type C;
  constexpr int kVar = FixedVar<T, MODE>::value;
  int var = kVar;
  if constexpr (kVar != 0) {
    if (M == 2048 && sizeof(T) == 4) {   // experiments of the float32 4096-point kernel only
      const char* ev = getenv("DSPB200_FFT_VAR");
      if (ev) var = atoi(ev);
      if (!(var == 0 || var == 6 || var == 7 || var == 15 || var == 23 || var == 39)) var = kVar;
    }
    if (a.window != nullptr && a.hann_ab == nullptr) var = 0;
  }
  const bool tree = (var & 1) != 0;
  size_t smem = static_cast<size_t>(padded(M) + 1) * sizeof(C);
  const size_t tw_bytes = static_cast<size_t>(tree ? a.tw_tree_total : a.tw_total) * sizeof(C);
  a.tw_in_smem = (tw_bytes > 0 && tw_bytes <= 24 * 1024) ? 1 : 0;
  if (a.tw_in_smem) smem += tw_bytes + (MODE == 0 ? static_cast<size_t>(M / 2 + 1) * sizeof(C) : 0);
  void (*kern)(const FftArgs<T>) = a.tw_in_smem ? fft_fixed_kernel<T, MODE, M, true> : fft_fixed_kernel<T, MODE, M, false>;
  if constexpr (kVar != 0) {
    if (var == kVar)
      kern = a.tw_in_smem ? fft_fixed_kernel<T, MODE, M, true, kVar> : fft_fixed_kernel<T, MODE, M, false, kVar>;
    if constexpr (M == 2048 && sizeof(T) == 4) {
      if (a.tw_in_smem && var == 6) kern = fft_fixed_kernel<T, MODE, M, true, 6>;
      if (a.tw_in_smem && var == 15) kern = fft_fixed_kernel<T, MODE, M, true, 15>;
      if (a.tw_in_smem && var == 23) kern = fft_fixed_kernel<T, MODE, M, true, 23>;
      if (a.tw_in_smem && var == 39) kern = fft_fixed_kernel<T, MODE, M, true, 39>;   // 7 with the first eight outputs of a thread kept in registers through the real split   // 7 with the next frame's loads issued after the passes
    }
  }
  DSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  if (const char* cv = getenv("DSPB200_FFT_CARVEOUT"))
    DSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(cv)));
  int per_sm = 1;
  DSP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, M / 16, smem));
  if (per_sm < 1) per_sm = 1;
  if (const char* cv = getenv("DSPB200_FFT_MAX_CTAS")) {
    const int lim = atoi(cv);
    if (lim >= 1 && lim < per_sm) per_sm = lim;
  }
  if (getenv("DSPB200_FFT_TRACE")) fprintf(stderr, "fft_fixed M=%d var=%d smem=%zu per_sm=%d\n", M, var, smem, per_sm);
  const long long cap = static_cast<long long>(sm_count()) * per_sm;
  const int grid = static_cast<int>(a.n_items < cap ? a.n_items : cap);
  kern<<<grid, M / 16, smem, stream>>>(a);
  return after_launch("fft_fixed_kernel");
}

template <typename T, int MODE>
static int launch_stockham(const FftArgs<T>& a, cudaStream_t stream) {
  if (getenv("DSPB200_FFT_GENERIC") == nullptr) {
    switch (a.m) {
      case 512: return launch_fixed<T, MODE, 512>(a, stream);
      case 1024: return launch_fixed<T, MODE, 1024>(a, stream);
      case 2048: return launch_fixed<T, MODE, 2048>(a, stream);
      case 4096: return launch_fixed<T, MODE, 4096>(a, stream);
      case 8192: return launch_fixed<T, MODE, 8192>(a, stream);
      default: break;
    }
  }
  const int threads = a.m / 16;
  if (threads <= 128) return launch_stockham_t<T, MODE, 128>(a, stream);
  if (threads <= 256) return launch_stockham_t<T, MODE, 256>(a, stream);
  return launch_stockham_t<T, MODE, 512>(a, stream);
}

template <typename T, bool kReal>
static int launch_combine(const FftArgs<T>& a, const typename Cpx<T>::type* ws, long long n_tr, cudaStream_t stream) {
  const int per = kReal ? (a.m / 2 + 1) : a.m;
  const long long total = n_tr * per;
  const int threads = 256;
  long long blocks = ceil_div(total, threads);
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  switch (a.r_top) {
    case 2: fft_combine_kernel<T, 2, kReal><<<static_cast<int>(blocks), threads, 0, stream>>>(a, ws, n_tr); break;
    case 4: fft_combine_kernel<T, 4, kReal><<<static_cast<int>(blocks), threads, 0, stream>>>(a, ws, n_tr); break;
    case 8: fft_combine_kernel<T, 8, kReal><<<static_cast<int>(blocks), threads, 0, stream>>>(a, ws, n_tr); break;
    case 16: fft_combine_kernel<T, 16, kReal><<<static_cast<int>(blocks), threads, 0, stream>>>(a, ws, n_tr); break;
    default: return fail(DSPB200_ERR_UNSUPPORTED, "internal: top radix %d", a.r_top);
  }
  return after_launch("fft_combine_kernel");
}

template <typename T, int N1>
static int launch_four_step(const FftArgs<T>& a, const FourStep<T>& fs, long long cnt, cudaStream_t stream) {
  typedef typename Cpx<T>::type C;
  constexpr int NC = N1 * kFsCols;
  if constexpr (N1 == 256) {
    // Opt-in (DSPB200_FFT_FUSED4=1).  Measured on C4 (512 x 2^20 samples, fp32): DRAM traffic 3.38 GB instead of 7.4 GB
    // (1.05 x algorithmic), but 1.86 ms against 1.57 ms for the two kernels: with three CTAs of 256 threads per SM and
    // two cluster barriers per transform the SMs issue 37 % of the time and the shared-memory pipe is 64 % busy --
    // the transforms are bound by the exchanges of the 16-points-per-thread passes, not by HBM.
    const char* fe = getenv("DSPB200_FFT_FUSED4");
    if (fe != nullptr && atoi(fe) != 0) {
      constexpr int kWork = kFsTile * fs_pitch<N1>() > 2 * kFsTile * kFsRowPitch ? kFsTile * fs_pitch<N1>() : 2 * kFsTile * kFsRowPitch;
      const size_t smem = (static_cast<size_t>(kWork) + NC / 256 + 256 + fs.tw_rows_total +
                           (sizeof(T) == 8 ? fs.tw_cols_total : 0)) * sizeof(C);
      auto kf = fft4_fused_kernel<T, N1>;
      DSP_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      cudaLaunchConfig_t cfg = {};
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = kFsCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.blockDim = dim3(256, 1, 1);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = stream;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      cfg.gridDim = dim3(kFsCluster, 1, 1);
      int max_clusters = 0;
      DSP_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kf, &cfg));
      if (max_clusters < 1) max_clusters = 1;
      if (const char* ev = getenv("DSPB200_FFT_FUSED4_CLUSTERS")) {
        const int lim = atoi(ev);
        if (lim >= 1 && lim < max_clusters) max_clusters = lim;
      }
      const long long n_clusters = cnt < max_clusters ? cnt : max_clusters;
      cfg.gridDim = dim3(static_cast<unsigned>(n_clusters * kFsCluster), 1, 1);
      if (getenv("DSPB200_FFT_TRACE")) fprintf(stderr, "fft4_fused N1=%d smem=%zu clusters=%lld (max %d)\n", N1, smem, n_clusters, max_clusters);
      DSP_CUDA(cudaLaunchKernelEx(&cfg, kf, a, fs, cnt));
      return after_launch("fft4_fused_kernel");
    }
  }
  const size_t smem1 = (static_cast<size_t>(kFsTile) * fs_pitch<N1>() + NC / 256 + 256 +
                        (sizeof(T) == 8 ? fs.tw_cols_total : 0)) * sizeof(C);
  const size_t smem2 = (static_cast<size_t>(2 * kFsTile) * kFsRowPitch + fs.tw_rows_total) * sizeof(C);
  auto k1 = fft4_cols_kernel<T, N1>;
  auto k2 = fft4_rows_kernel<T, N1>;
  DSP_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem1)));
  DSP_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem2)));
  const long long b1 = cnt * (kFsCols / kFsTile), b2 = cnt * (N1 / 2 / kFsTile);
  DSP_CHECK(b1 < (1LL << 31) && b2 < (1LL << 31), "too many transforms in one launch");
  k1<<<static_cast<unsigned>(b1), N1, smem1, stream>>>(a, fs);
  DSP_TRY(after_launch("fft4_cols_kernel"));
  k2<<<static_cast<unsigned>(b2), 256, smem2, stream>>>(a, fs);
  return after_launch("fft4_rows_kernel");
}

// 4096-point fp32 magnitude frames run the 32-points-per-thread kernel (fft_r32.cu); DSPB200_FFT_VAR < 64 selects
// the 16-points-per-thread variants of fft_fixed_kernel instead (7 = the best of them)
static bool fft_r32_selected() {
  const char* ev = getenv("DSPB200_FFT_VAR");
  return ev == nullptr || atoi(ev) >= 64;
}

template <typename T>
int fftmag_run(const dspb200_fft_plan* p, const T* x, int64_t xs, int64_t n_valid, int64_t offset, int64_t hop,
               int64_t n_frames, T* mag, int64_t mfs, int64_t mcs, int64_t channels, void* ws, size_t ws_bytes,
               cudaStream_t stream) {
  typedef typename Cpx<T>::type C;
  DSP_PLAN(p, kMagicFft, "fft");
  DSP_CHECK(p->dtype == DType<T>::id, "plan dtype %d does not match the entry point", p->dtype);
  DSP_CHECK(channels >= 0 && n_frames >= 0 && n_valid >= 0 && offset >= 0 && hop >= 0, "negative argument");
  if (channels == 0 || n_frames == 0) return DSPB200_OK;
  DSP_CHECK(x != nullptr && mag != nullptr, "NULL buffer");
  DSP_CHECK(mfs >= p->n_fft / 2 + 1, "mag_frame_stride smaller than n_fft/2+1");
  DSP_TRY(ensure_device());
  DSP_TRY(check_plan_device(p->device, "fft"));
  const int64_t n_tr = channels * n_frames;
  FftArgs<T> a{};
  a.x = x; a.x_stride = xs; a.n_valid = n_valid; a.offset = offset; a.hop = hop; a.n_frames = n_frames;
  a.window = static_cast<const T*>(p->d_window);
  a.hann_ab = static_cast<const T*>(p->d_hann_ab);
  for (int u = 0; u < 16; ++u) {
    a.hann_cc[u].x = a.hann_cc[u].y = static_cast<T>(p->hann_cos[u]);
    a.hann_ss[u].x = a.hann_ss[u].y = static_cast<T>(p->hann_sin[u]);
  }
  a.db = p->db;
  a.mag = mag; a.mag_frame_stride = mfs; a.mag_channel_stride = mcs;
  const FftSide& s = p->real_side;
  if (p->n_fft < 32) {
    const int per = p->n_fft / 2 + 1;
    const int threads = 128;
    long long blocks = ceil_div(n_tr * per, threads);
    if (blocks > 65535) blocks = 65535;
    fft_small_kernel<T, true><<<static_cast<int>(blocks), threads, 0, stream>>>(
        a, static_cast<const C*>(p->d_tw_full), p->n_fft, n_tr);
    return after_launch("fft_small_kernel");
  }
  fill_args(a, s);
  a.n_items = n_tr * s.r_top;
  if constexpr (sizeof(T) == 4) {
    if (p->r32.ok && fft_r32_selected())
      return fft_r32_run(p->r32, x, xs, n_valid, offset, hop, n_frames, mag, mfs, mcs, channels, p->hann, p->db, stream);
  }
  if constexpr (sizeof(T) == 4) {
    // 2^16-point frames run the three-pass form (fft_long32.cu); DSPB200_FFT_LONG32=0 selects the four-step kernels
    const char* ev = getenv("DSPB200_FFT_LONG32");
    if (p->l32.ok && (ev == nullptr || atoi(ev) != 0))
      return fft_long32_run(p->l32, x, xs, n_valid, offset, hop, n_frames, mag, mfs, mcs, channels, p->hann, p->db, ws,
                            ws_bytes, stream);
  }
  if (s.r_top == 1) return launch_stockham<T, 0>(a, stream);
  const size_t need = side_workspace(s, n_tr, sizeof(C));
  DSP_CHECK(ws != nullptr && ws_bytes >= need, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
  // run in chunks whose workspace stays L2-resident between the two kernels
  const int64_t chunk = split_chunk(s, sizeof(C));
  a.out = static_cast<C*>(ws);
  const bool four_step = p->fs_n1 > 0 && getenv("DSPB200_FFT_NO_FOUR_STEP") == nullptr;
  FourStep<T> fs{};
  if (four_step) {
    fs.tw_cols = static_cast<const C*>(p->fs_cols.d_tw_pass);
    fs.tw_rows = static_cast<const C*>(p->fs_rows.d_tw_pass);
    for (int i = 0; i < 8; ++i) {
      fs.tw_cols_offset[i] = p->fs_cols.tw_offset[i];
      fs.tw_rows_offset[i] = p->fs_rows.tw_offset[i];
    }
    fs.tw_cols_total = p->fs_cols.tw_total;
    fs.tw_rows_total = p->fs_rows.tw_total;
    fs.tw_hi = static_cast<const C*>(p->d_fs_hi);
    fs.tw_lo = static_cast<const C*>(p->d_fs_lo);
    fs.ws = static_cast<C*>(ws);
  }
  for (int64_t f0 = 0; f0 < n_tr; f0 += chunk) {
    const int64_t cnt = (n_tr - f0) < chunk ? (n_tr - f0) : chunk;
    a.first = f0;
    a.n_items = cnt * s.r_top;
    if (four_step) {
      switch (p->fs_n1) {
        case 128: DSP_TRY((launch_four_step<T, 128>(a, fs, cnt, stream))); break;
        case 256: DSP_TRY((launch_four_step<T, 256>(a, fs, cnt, stream))); break;
        case 512: DSP_TRY((launch_four_step<T, 512>(a, fs, cnt, stream))); break;
        case 1024:
          if constexpr (sizeof(T) == 4) { DSP_TRY((launch_four_step<T, 1024>(a, fs, cnt, stream))); break; }
        default: return fail(DSPB200_ERR_UNSUPPORTED, "internal: four-step N1 %d", p->fs_n1);
      }
      continue;
    }
    DSP_TRY((launch_stockham<T, 1>(a, stream)));
    DSP_TRY((launch_combine<T, true>(a, static_cast<const C*>(ws), cnt, stream)));
  }
  return DSPB200_OK;
}

template <typename T>
int fft_c2c_run(const dspb200_fft_plan* p, const T* in, T* out, int64_t batch, void* ws, size_t ws_bytes,
                cudaStream_t stream) {
  typedef typename Cpx<T>::type C;
  DSP_PLAN(p, kMagicFft, "fft");
  DSP_CHECK(p->dtype == DType<T>::id, "plan dtype %d does not match the entry point", p->dtype);
  DSP_CHECK(batch >= 0, "negative batch");
  if (batch == 0) return DSPB200_OK;
  DSP_CHECK(in != nullptr && out != nullptr, "NULL buffer");
  DSP_CHECK(in != out, "in-place complex transform is not supported");
  DSP_TRY(ensure_device());
  DSP_TRY(check_plan_device(p->device, "fft"));
  FftArgs<T> a{};
  a.x = in; a.out = reinterpret_cast<C*>(out); a.n_frames = 1;
  const FftSide& s = p->c2c_side;
  if (p->n_fft < 16) {
    const int threads = 128;
    long long blocks = ceil_div(batch * p->n_fft, threads);
    if (blocks > 65535) blocks = 65535;
    fft_small_kernel<T, false><<<static_cast<int>(blocks), threads, 0, stream>>>(
        a, static_cast<const C*>(p->d_tw_full), p->n_fft, batch);
    return after_launch("fft_small_kernel");
  }
  fill_args(a, s);
  a.n_items = batch * s.r_top;
  if (s.r_top == 1) return launch_stockham<T, 2>(a, stream);
  const size_t need = side_workspace(s, batch, sizeof(C));
  DSP_CHECK(ws != nullptr && ws_bytes >= need, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
  C* final_out = a.out;
  const int64_t chunk = split_chunk(s, sizeof(C));
  for (int64_t f0 = 0; f0 < batch; f0 += chunk) {
    const int64_t cnt = (batch - f0) < chunk ? (batch - f0) : chunk;
    a.first = f0;
    a.n_items = cnt * s.r_top;
    a.out = static_cast<C*>(ws);
    DSP_TRY((launch_stockham<T, 3>(a, stream)));
    a.out = final_out;
    DSP_TRY((launch_combine<T, false>(a, static_cast<const C*>(ws), cnt, stream)));
  }
  return DSPB200_OK;
}

template int fftmag_run<float>(const dspb200_fft_plan*, const float*, int64_t, int64_t, int64_t, int64_t, int64_t, float*, int64_t, int64_t, int64_t, void*, size_t, cudaStream_t);
template int fftmag_run<double>(const dspb200_fft_plan*, const double*, int64_t, int64_t, int64_t, int64_t, int64_t, double*, int64_t, int64_t, int64_t, void*, size_t, cudaStream_t);

int fft_plan_info(const dspb200_fft_plan* plan, int* n_fft, int* dtype) {
  DSP_PLAN(plan, kMagicFft, "fft");
  *n_fft = plan->n_fft; *dtype = plan->dtype;
  return DSPB200_OK;
}

template <typename T>
static int fftmag_host(const dspb200_fft_plan* p, const T* x, int64_t channels, int64_t n_samples, int64_t offset,
                       int64_t hop, int64_t n_frames, T* mag) {
  typedef typename Cpx<T>::type C;
  DSP_PLAN(p, kMagicFft, "fft");
  DSP_CHECK(channels >= 0 && n_samples >= 0 && n_frames >= 0, "negative shape");
  if (channels == 0 || n_frames == 0) return DSPB200_OK;
  DSP_CHECK(mag != nullptr && (x != nullptr || n_samples == 0), "NULL buffer");
  DSP_TRY(ensure_device());
  const int bins = p->n_fft / 2 + 1;
  const int64_t pitch = round_up(n_samples > 0 ? n_samples : 1, 16 / static_cast<int64_t>(sizeof(T)));
  T *dx = nullptr, *dm = nullptr;
  void* ws = nullptr;
  const size_t need = side_workspace(p->real_side, channels * n_frames, sizeof(C));
  cudaError_t e = cudaMalloc(&dx, static_cast<size_t>(channels) * pitch * sizeof(T));
  if (e == cudaSuccess) e = cudaMalloc(&dm, static_cast<size_t>(channels) * n_frames * bins * sizeof(T));
  if (e == cudaSuccess && need) e = cudaMalloc(&ws, need);
  int rc = DSPB200_OK;
  if (e == cudaSuccess && n_samples > 0)
    e = cudaMemcpy2DAsync(dx, pitch * sizeof(T), x, n_samples * sizeof(T), n_samples * sizeof(T), channels,
                          cudaMemcpyHostToDevice, 0);
  if (e == cudaSuccess)
    rc = fftmag_run<T>(p, dx, pitch, n_samples, offset, hop, n_frames, dm, bins, n_frames * bins, channels, ws, need, nullptr);
  if (e == cudaSuccess && rc == DSPB200_OK)
    e = cudaMemcpyAsync(mag, dm, static_cast<size_t>(channels) * n_frames * bins * sizeof(T), cudaMemcpyDeviceToHost, 0);
  if (e == cudaSuccess) e = cudaStreamSynchronize(0);
  cudaFree(dx); cudaFree(dm); cudaFree(ws);
  if (e != cudaSuccess) return fail(DSPB200_ERR_CUDA, "fftmag host path: %s", cudaGetErrorString(e));
  return rc;
}

}  // namespace dspb200

using namespace dspb200;

extern "C" {

int dspb200_fft_plan_create(int n_fft, int hann, int dtype, dspb200_fft_plan** plan) {
  DSP_CHECK(plan != nullptr, "plan output pointer is NULL");
  DSP_CHECK(n_fft >= 1 && (n_fft & (n_fft - 1)) == 0, "n_fft must be a power of two (got %d)", n_fft);
  DSP_CHECK(n_fft <= DSPB200_FFT_MAX, "n_fft above %d is not supported", DSPB200_FFT_MAX);
  DSP_CHECK(dtype == DSPB200_F32 || dtype == DSPB200_F64, "dtype must be 0 (f32) or 1 (f64)");
  DSP_TRY(ensure_device());
  dspb200_fft_plan* p = new (std::nothrow) dspb200_fft_plan();
  if (!p) return fail(DSPB200_ERR_ALLOC, "out of host memory");
  p->n_fft = n_fft; p->hann = (hann & DSPB200_FFT_HANN) ? 1 : 0; p->db = (hann & DSPB200_FFT_DB) ? 1 : 0; p->dtype = dtype;
  cudaGetDevice(&p->device);
  int rc = dtype == DSPB200_F32 ? plan_build<float>(p) : plan_build<double>(p);
  if (rc != DSPB200_OK) {
    dspb200_fft_plan_destroy(p);
    return rc;
  }
  *plan = p;
  return DSPB200_OK;
}

int dspb200_fft_plan_destroy(dspb200_fft_plan* p) {
  if (!p) return DSPB200_OK;
  DSP_PLAN(p, kMagicFft, "fft");
  p->magic = 0;
  cudaFree(p->d_fs_hi);
  cudaFree(p->d_fs_lo);
  FftSide* sides[4] = {&p->real_side, &p->c2c_side, &p->fs_cols, &p->fs_rows};
  for (FftSide* s : sides) {
    cudaFree(s->d_tw_pass);
    cudaFree(s->d_tw_tree);
    cudaFree(s->d_tw_top);
    cudaFree(s->d_tw_post);
  }
  cudaFree(p->d_window);
  cudaFree(p->d_hann_ab);
  fft_r32_free(p->r32);
  fft_long32_free(p->l32);
  cudaFree(p->d_tw_full);
  delete p;
  return DSPB200_OK;
}

int dspb200_fft_workspace_bytes(const dspb200_fft_plan* p, int64_t n_transforms, size_t* bytes) {
  DSP_PLAN(p, kMagicFft, "fft");
  DSP_CHECK(bytes != nullptr, "NULL argument");
  DSP_CHECK(n_transforms >= 0, "negative transform count");
  const size_t csize = p->dtype == DSPB200_F32 ? 8 : 16;
  const size_t r = side_workspace(p->real_side, n_transforms, csize);
  const size_t c = p->c2c_side.nc > 0 ? side_workspace(p->c2c_side, n_transforms, csize) : 0;
  *bytes = r > c ? r : c;
  return DSPB200_OK;
}

int dspb200_fftmag_run_f32(const dspb200_fft_plan* p, const float* x, int64_t xs, int64_t n_valid, int64_t offset,
                           int64_t hop, int64_t n_frames, float* mag, int64_t mfs, int64_t mcs, int64_t channels,
                           void* ws, size_t ws_bytes, void* stream) {
  return fftmag_run<float>(p, x, xs, n_valid, offset, hop, n_frames, mag, mfs, mcs, channels, ws, ws_bytes,
                           static_cast<cudaStream_t>(stream));
}
int dspb200_fftmag_run_f64(const dspb200_fft_plan* p, const double* x, int64_t xs, int64_t n_valid, int64_t offset,
                           int64_t hop, int64_t n_frames, double* mag, int64_t mfs, int64_t mcs, int64_t channels,
                           void* ws, size_t ws_bytes, void* stream) {
  return fftmag_run<double>(p, x, xs, n_valid, offset, hop, n_frames, mag, mfs, mcs, channels, ws, ws_bytes,
                            static_cast<cudaStream_t>(stream));
}
int dspb200_fft_c2c_run_f32(const dspb200_fft_plan* p, const float* in, float* out, int64_t batch, void* ws,
                            size_t ws_bytes, void* stream) {
  DSP_PLAN(p, kMagicFft, "fft");
  DSP_CHECK(p->c2c_side.nc > 0, "plan is too long for the complex transform");
  return fft_c2c_run<float>(p, in, out, batch, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}
int dspb200_fft_c2c_run_f64(const dspb200_fft_plan* p, const double* in, double* out, int64_t batch, void* ws,
                            size_t ws_bytes, void* stream) {
  DSP_PLAN(p, kMagicFft, "fft");
  DSP_CHECK(p->c2c_side.nc > 0, "plan is too long for the complex transform");
  return fft_c2c_run<double>(p, in, out, batch, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}
int dspb200_fftmag_host_f32(const dspb200_fft_plan* p, const float* x, int64_t channels, int64_t n_samples,
                            int64_t offset, int64_t hop, int64_t n_frames, float* mag) {
  DSP_PLAN(p, kMagicFft, "fft");
  DSP_CHECK(p->dtype == DSPB200_F32, "plan is not float32");
  return fftmag_host<float>(p, x, channels, n_samples, offset, hop, n_frames, mag);
}
int dspb200_fftmag_host_f64(const dspb200_fft_plan* p, const double* x, int64_t channels, int64_t n_samples,
                            int64_t offset, int64_t hop, int64_t n_frames, double* mag) {
  DSP_PLAN(p, kMagicFft, "fft");
  DSP_CHECK(p->dtype == DSPB200_F64, "plan is not float64");
  return fftmag_host<double>(p, x, channels, n_samples, offset, hop, n_frames, mag);
}
int dspb200_fft_c2c_host_f64(const dspb200_fft_plan* p, const double* in, double* out, int64_t batch) {
  DSP_PLAN(p, kMagicFft, "fft");
  DSP_CHECK(p->dtype == DSPB200_F64 && p->c2c_side.nc > 0, "plan is not float64 or too long");
  DSP_CHECK(batch >= 0, "negative batch");
  if (batch == 0) return DSPB200_OK;
  DSP_CHECK(in != nullptr && out != nullptr, "NULL buffer");
  DSP_TRY(ensure_device());
  const size_t bytes = static_cast<size_t>(batch) * p->n_fft * 16;
  const size_t need = side_workspace(p->c2c_side, batch, 16);
  double *di = nullptr, *dout = nullptr;
  void* ws = nullptr;
  cudaError_t e = cudaMalloc(&di, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&dout, bytes);
  if (e == cudaSuccess && need) e = cudaMalloc(&ws, need);
  int rc = DSPB200_OK;
  if (e == cudaSuccess) e = cudaMemcpyAsync(di, in, bytes, cudaMemcpyHostToDevice, 0);
  if (e == cudaSuccess) rc = fft_c2c_run<double>(p, di, dout, batch, ws, need, nullptr);
  if (e == cudaSuccess && rc == DSPB200_OK) e = cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, 0);
  if (e == cudaSuccess) e = cudaStreamSynchronize(0);
  cudaFree(di); cudaFree(dout); cudaFree(ws);
  if (e != cudaSuccess) return fail(DSPB200_ERR_CUDA, "fft c2c host path: %s", cudaGetErrorString(e));
  return rc;
}

}  // extern "C"
