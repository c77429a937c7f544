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

// c + a b with a's parts as broadcast scalars (two packed FMAs), and 2 e - lo (the other output of a butterfly whose
// first output lo = e + t is known)
__device__ __forceinline__ float2 cmadd(float2 a, float2 b, float2 c) {
  return ffma2s(make_float2(-b.y, b.x), a.y, ffma2s(b, a.x, c));
}
__device__ __forceinline__ float2 twice_minus(float2 e, float2 lo) { return ffma2s(e, 2.0f, make_float2(-lo.x, -lo.y)); }
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
