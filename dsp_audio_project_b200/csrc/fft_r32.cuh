// Register-resident 32-point DFT and the packed-fp32 butterfly helpers shared by the 32-points-per-thread FFT
// kernels (fft_r32.cu: 4096-point frames; fft_long32.cu: 2^16-point frames).
#pragma once
#include "common.cuh"
#include "cpx.cuh"

namespace dspb200 {
namespace r32 {

constexpr int kTw1Pitch = 11;     // W_1024 table [32][11]: w^1, w^2, w^3, w^4, w^8, ..., w^28, pad  (w = W_1024^row)
constexpr int kTw1 = 32 * kTw1Pitch;

// cos/sin(2 pi k / 32), k = 0..15
__device__ constexpr double kCos32[16] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
                                          0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173,
                                          0.19509032201612826785, 0.0, -0.19509032201612826785, -0.38268343236508977173,
                                          -0.55557023301960222474, -0.70710678118654752440, -0.83146961230254523708,
                                          -0.92387953251128675613, -0.98078528040323044913};
__device__ constexpr double kSin32[16] = {0.0, 0.19509032201612826785, 0.38268343236508977173, 0.55557023301960222474,
                                          0.70710678118654752440, 0.83146961230254523708, 0.92387953251128675613,
                                          0.98078528040323044913, 1.0, 0.98078528040323044913, 0.92387953251128675613,
                                          0.83146961230254523708, 0.70710678118654752440, 0.55557023301960222474,
                                          0.38268343236508977173, 0.19509032201612826785};

// W_64^i, i = 0..7
__device__ constexpr double kW64r[8] = {1.0, 0.99518472667219688624, 0.98078528040323044913, 0.95694033573220886494,
                                        0.92387953251128675613, 0.88192126434835502971, 0.83146961230254523708,
                                        0.77301045336273696081};
__device__ constexpr double kW64i[8] = {0.0, -0.09801714032956060199, -0.19509032201612826785, -0.29028467725446236764,
                                        -0.38268343236508977173, -0.47139673682599764856, -0.55557023301960222474,
                                        -0.63439328416364549822};

// R-point DFT in registers, natural order in and out: the reference's even/odd recursion (dsp_core.py:52-66) unrolled.
template <int R> struct Dft32 {
  static __device__ __forceinline__ void run(float2* v) {
    float2 e[R / 2], o[R / 2];
#pragma unroll
    for (int k = 0; k < R / 2; ++k) { e[k] = v[2 * k]; o[k] = v[2 * k + 1]; }
    Dft32<R / 2>::run(e);
    Dft32<R / 2>::run(o);
#pragma unroll
    for (int k = 0; k < R / 2; ++k) {
      if (k == 0) {
        v[k] = cadd(e[k], o[k]);
        v[k + R / 2] = csub(e[k], o[k]);
      } else if (4 * k == R) {
        const float2 t = mul_neg_i(o[k]);
        v[k] = cadd(e[k], t);
        v[k + R / 2] = csub(e[k], t);
      } else {
        // e + w o as two packed FMAs, e - w o = 2 e - (e + w o) as a third: 6 lane operations instead of 8
        const float wr = static_cast<float>(kCos32[k * (32 / R)]), wi = static_cast<float>(-kSin32[k * (32 / R)]);
        const float2 os = make_float2(-o[k].y, o[k].x);
        const float2 lo = ffma2s(os, wi, ffma2s(o[k], wr, e[k]));
        v[k] = lo;
        v[k + R / 2] = ffma2s(e[k], 2.0f, make_float2(-lo.x, -lo.y));   // = twice_minus(e, lo), declared below
      }
    }
  }
};
template <> struct Dft32<1> {
  static __device__ __forceinline__ void run(float2*) {}
};

__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

template <bool kDb> __device__ __forceinline__ float mag_of(float2 p) {
  return finish_mag(fmaf(p.x, p.x, p.y * p.y), kDb ? 1 : 0);
}

// v[n] *= w^n, n = 1..31, with w^(a + 4 b) = w^a * w^(4 b) from ten table entries (row of the W_1024-style table):
// one rounding deep
__device__ __forceinline__ void twiddle_powers(float2* v, const float2* __restrict__ row) {
  float2 wa[4], wb[8];
#pragma unroll
  for (int i = 1; i < 4; ++i) wa[i] = row[i - 1];
#pragma unroll
  for (int i = 1; i < 8; ++i) wb[i] = row[2 + i];
#pragma unroll
  for (int n = 1; n < 32; ++n) {
    const int lo = n & 3, hi = n >> 2;
    float2 w;
    if (hi == 0) w = wa[lo];
    else if (lo == 0) w = wb[hi];
    else w = cmul(wa[lo], wb[hi]);
    v[n] = cmul(v[n], w);
  }
}

// The same 32-point DFT with its first butterfly level already done by the caller: P[i] = x_i + x_(i+16),
// M[i] = x_i - x_(i+16), i < 16 (the leaves of the even/odd recursion pair inputs 16 apart).  Lets the caller fold a
// per-input factor (window, twiddle) into that level: x_i f_i + x_j f_j is one multiply and one FMA, the difference
// 2 x_i f_i - that sum a third instruction.
template <int R> struct Dft32Pre {
  static __device__ __forceinline__ void run(const float2* P, const float2* M, float2* v) {
    float2 Pe[R / 4], Po[R / 4], Me[R / 4], Mo[R / 4], e[R / 2], o[R / 2];
#pragma unroll
    for (int k = 0; k < R / 4; ++k) { Pe[k] = P[2 * k]; Po[k] = P[2 * k + 1]; Me[k] = M[2 * k]; Mo[k] = M[2 * k + 1]; }
    Dft32Pre<R / 2>::run(Pe, Me, e);
    Dft32Pre<R / 2>::run(Po, Mo, o);
#pragma unroll
    for (int k = 0; k < R / 2; ++k) {
      if (k == 0) {
        v[k] = cadd(e[k], o[k]);
        v[k + R / 2] = csub(e[k], o[k]);
      } else if (4 * k == R) {
        const float2 t = mul_neg_i(o[k]);
        v[k] = cadd(e[k], t);
        v[k + R / 2] = csub(e[k], t);
      } else {
        const float wr = static_cast<float>(kCos32[k * (32 / R)]), wi = static_cast<float>(-kSin32[k * (32 / R)]);
        const float2 os = make_float2(-o[k].y, o[k].x);
        const float2 lo = ffma2s(os, wi, ffma2s(o[k], wr, e[k]));
        v[k] = lo;
        v[k + R / 2] = twice_minus(e[k], lo);
      }
    }
  }
};
template <> struct Dft32Pre<2> {
  static __device__ __forceinline__ void run(const float2* P, const float2* M, float2* v) { v[0] = P[0]; v[1] = M[0]; }
};

// v <- DFT32(v[n] f[n]) with real per-component factors f[n] = quarter + a cc[n] + b ss[n] (the Hann window by angle
// addition: a, b per thread, cc / ss the step rotations duplicated into both halves)
__device__ __forceinline__ void dft32_windowed(float2* v, float2 ha, float2 hb, const float2* cc, const float2* ss) {
  const float2 quarter = make_float2(0.25f, 0.25f);
  float2 P[16], M[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 wa = fma2(ha, cc[i], fma2(hb, ss[i], quarter));
    const float2 wb = fma2(ha, cc[i + 16], fma2(hb, ss[i + 16], quarter));
    const float2 t = pmul(v[i], wa);
    P[i] = fma2(v[i + 16], wb, t);
    M[i] = fma2(v[i + 16], make_float2(-wb.x, -wb.y), t);
  }
  Dft32Pre<32>::run(P, M, v);
}

// v <- DFT32(v[n] w^n) with the powers of w from a ten-entry row as in twiddle_powers
__device__ __forceinline__ void dft32_twiddled(float2* v, const float2* __restrict__ row) {
  float2 wa[4], wb[8];
#pragma unroll
  for (int i = 1; i < 4; ++i) wa[i] = row[i - 1];
#pragma unroll
  for (int i = 1; i < 8; ++i) wb[i] = row[2 + i];
  float2 P[16], M[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int lo = i & 3, hi = i >> 2;
    const float2 wj = lo == 0 ? wb[hi + 4] : cmul(wa[lo], wb[hi + 4]);   // w^(i + 16)
    float2 x = v[i];
    if (i > 0) x = cmul(x, hi == 0 ? wa[lo] : (lo == 0 ? wb[hi] : cmul(wa[lo], wb[hi])));
    P[i] = cmadd(wj, v[i + 16], x);
    M[i] = twice_minus(x, P[i]);
  }
  Dft32Pre<32>::run(P, M, v);
}

// host: the ten entries of row `num/den` (w = exp(-2 pi i num/den))
inline void fill_twiddle_row(float2* row, long double num, long double den) {
  const long double pi = 3.14159265358979323846264338327950288L;
  auto w = [&](long double m) {
    const long double ang = -2.0L * pi * num * m / den;
    return make_float2(static_cast<float>(cosl(ang)), static_cast<float>(sinl(ang)));
  };
  for (int i = 1; i < 4; ++i) row[i - 1] = w(i);
  for (int i = 1; i < 8; ++i) row[2 + i] = w(4 * i);
}

}  // namespace r32
}  // namespace dspb200
