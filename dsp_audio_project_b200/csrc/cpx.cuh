// Complex arithmetic of the FFT kernels (fft.cu, fft_r32.cu): one fp32 complex number is one packed pair
// (FADD2 / FMUL2 / FFMA2), fp64 stays scalar.
#pragma once
#include "common.cuh"

namespace dspb200 {

template <typename T> struct Cpx;
template <> struct Cpx<float> { typedef float2 type; };
template <> struct Cpx<double> { typedef double2 type; };

template <typename C> __device__ __forceinline__ C cadd(C a, C b) { C r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { C r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
  C r;
  r.x = a.x * b.x - a.y * b.y;
  r.y = a.x * b.y + a.y * b.x;
  return r;
}
// float2 overloads: one complex number = one packed fp32 pair (FADD2 / FMUL2 / FFMA2);
// ptxas folds the half swaps and sign flips of -i*z and of the complex product
// into the instructions' operand modifiers.
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {   // a * b = a.x * (b.x, b.y) + a.y * (-b.y, b.x)
  const float2 bs = make_float2(-b.y, b.x);
  return ffma2s(bs, a.y, fmul2s(b, a.x));
}
__device__ __forceinline__ float2 pmul(float2 a, float2 b) {   // elementwise (a.x*b.x, a.y*b.y)
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ double2 pmul(double2 a, double2 b) { return make_double2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ float2 pscale(float2 a, float s) { return fmul2s(a, s); }
__device__ __forceinline__ double2 pscale(double2 a, double s) { return make_double2(a.x * s, a.y * s); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {     // elementwise a*b + c
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
      "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ double2 fma2(double2 a, double2 b, double2 c) { return make_double2(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y)); }
// c + a b, and 2 e - lo (the second output of a butterfly whose first output lo = e + t is known)
__device__ __forceinline__ float2 cmadd(float2 a, float2 b, float2 c) {   // a's parts as broadcast scalars: two packed FMAs
  return ffma2s(make_float2(-b.y, b.x), a.y, ffma2s(b, a.x, c));
}
__device__ __forceinline__ double2 cmadd(double2 a, double2 b, double2 c) {
  return make_double2(fma(-a.y, b.y, fma(a.x, b.x, c.x)), fma(a.y, b.x, fma(a.x, b.y, c.y)));
}
__device__ __forceinline__ float2 twice_minus(float2 e, float2 lo) { return ffma2s(e, 2.0f, make_float2(-lo.x, -lo.y)); }
__device__ __forceinline__ double2 twice_minus(double2 e, double2 lo) { return make_double2(fma(2.0, e.x, -lo.x), fma(2.0, e.y, -lo.y)); }
template <typename C> __device__ __forceinline__ C cconj(C a) { a.y = -a.y; return a; }
template <typename C> __device__ __forceinline__ C mul_neg_i(C a) { C r; r.x = a.y; r.y = -a.x; return r; }
// magnitude (or its dB value) from |X|^2
__device__ __forceinline__ double finish_mag(double v2, int db) {
  const double m = sqrt(v2);
  return db ? 20.0 * log10(m + 1e-12) : m;
}
__device__ __forceinline__ float finish_mag(float v2, int db) {
  float m;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(m) : "f"(v2));  // bare MUFU.SQRT (~1 ulp); |X|^2 < 1.2e-38 reads as 0
  return db ? 20.0f * log10f(m + 1e-12f) : m;
}

}  // namespace dspb200
