// K2 -- biquad equaliser cascade as a chunked parallel linear-recurrence scan.
//
// Replaces sistema_ecualizador / aplicar_ecuacion_diferencias
// (dsp_core.py:205-254): up to 6 (here: up to 16) second-order sections in
// series, zero initial state, one clip at the end (:254).
//
// Layout and mapping.  x, z: [channels, time], time fastest.  One WARP owns one
// channel and marches over it in tiles of 32*LC samples (LC = 32 fp32 / 16
// fp64).  A tile is staged global->shared with 16-byte cp.async (coalesced,
// double buffered); lane l then owns the LC consecutive samples
// [l*LC, (l+1)*LC) of the tile in registers.  For each section, in order:
//   1. zero-state pass over the lane's chunk in the 2x2 state-space form
//      (design.cu: rotation-scaling for complex poles, Schur form for real);
//   2. warp-shuffle Kogge-Stone prefix composition of the chunk end states:
//      f_l <- f_l + A^(LC*2^d) f_(l-2^d); the tile's carry-in state is folded
//      into lane 0 first, so the inclusive result is the true state at the end
//      of every chunk and lane 31's becomes the next tile's carry;
//   3. correction y[i] += (c A^i) . q_in for the lane's true initial state.
// All sections are applied to the register-resident chunk before it is
// written back: the cascade costs ONE read and ONE write of HBM per sample.
//
// Roofline: 2*sizeof(T) algorithmic bytes per sample; ~8 FMA per section per
// sample (6 zero-state + 2 correction), i.e. ~50 FMA/sample for six sections.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "design.cuh"
#include "internal.cuh"

namespace dspb200 {

constexpr int kEqPassSections = 8;   // sections fused per kernel launch
constexpr int kEqWarpsPerCta = 8;

template <typename T> struct EqCfg;
template <> struct EqCfg<float> { static constexpr int LC = 32; static constexpr int PAD = 4; };
template <> struct EqCfg<double> { static constexpr int LC = 16; static constexpr int PAD = 2; };

template <typename T, int NS> struct EqKernelParams {
  static constexpr int LC = EqCfg<T>::LC;
  T a[NS][4];
  T b[NS][2];
  T c[NS][2];
  T dd[NS];
  T g[NS][LC][2];   // c * A^i, i = 0..LC-1
  T pw[NS][5][4];   // A^(LC * 2^d), d = 0..4
  T ph[NS][4];      // A^(LC / 2): joins a lane's two half chunks (packed fp32 path)
  T gain;
  int clip;
};

template <typename T> struct EqKernelParams<T, 0> {
  T gain;
  int clip;
};

template <typename T> __device__ __forceinline__ T fma_t(T a, T b, T c);
template <> __device__ __forceinline__ float fma_t<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> __device__ __forceinline__ double fma_t<double>(double a, double b, double c) { return fma(a, b, c); }

// kPlain: every section has b = [1, 0] and its direct gain folded into `gain`.
template <typename T, int NS, bool kPlain>
__global__ void __launch_bounds__(kEqWarpsPerCta * 32)
eq_scan_kernel(const __grid_constant__ EqKernelParams<T, NS> p, const T* __restrict__ x,
               long long x_stride, T* __restrict__ z, long long z_stride, long long channels,
               long long n, int aligned) {
  constexpr int LC = EqCfg<T>::LC;
  constexpr int PITCH = LC + EqCfg<T>::PAD;        // chunk pitch: conflict-free 16-byte lane access
  constexpr int VEC = 16 / static_cast<int>(sizeof(T));   // elements per 16 bytes
  constexpr int TILE = 32 * LC;
  constexpr int PIECES = TILE / VEC / 32;          // 16-byte pieces per lane per tile
  constexpr int STAGE = 32 * PITCH;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  T* buf = reinterpret_cast<T*>(smem_raw) + static_cast<size_t>(warp) * 2 * STAGE;
  const long long warps_total = static_cast<long long>(gridDim.x) * kEqWarpsPerCta;
  const long long n_tiles = (n + TILE - 1) / TILE;

  for (long long ch = static_cast<long long>(blockIdx.x) * kEqWarpsPerCta + warp; ch < channels;
       ch += warps_total) {
    const T* xc = x + ch * x_stride;
    T* zc = z + ch * z_stride;
    T carry[NS > 0 ? NS : 1][2];
#pragma unroll
    for (int s = 0; s < (NS > 0 ? NS : 1); ++s) carry[s][0] = carry[s][1] = T(0);

    auto stage_in = [&](long long t, int st) {
      T* dst = buf + st * STAGE;
      const long long base = t * TILE;
      if (aligned) {
#pragma unroll
        for (int k = 0; k < PIECES; ++k) {
          const int q = lane + 32 * k;                  // piece index in the tile
          const int e = q * VEC;                        // first element of the piece
          const long long gi = base + e;
          long long rem = (n - gi) * static_cast<long long>(sizeof(T));
          const int nbytes = rem >= 16 ? 16 : (rem > 0 ? static_cast<int>(rem) : 0);
          const T* src = nbytes > 0 ? xc + gi : xc;
          cp_async16(dst + (e / LC) * PITCH + (e % LC), src, nbytes);
        }
      } else {
        for (int e = lane; e < TILE; e += 32) {
          const long long gi = base + e;
          dst[(e / LC) * PITCH + (e % LC)] = gi < n ? xc[gi] : T(0);
        }
      }
      cp_async_commit();
    };

    if (n_tiles > 0) stage_in(0, 0);
    for (long long t = 0; t < n_tiles; ++t) {
      const int st = static_cast<int>(t & 1);
      if (t + 1 < n_tiles) stage_in(t + 1, st ^ 1);
      else cp_async_commit();   // empty group keeps the wait depth uniform
      cp_async_wait<1>();
      __syncwarp();
      T* cur = buf + st * STAGE + lane * PITCH;
      T v[LC];
#pragma unroll
      for (int i = 0; i < LC; i += VEC) {
        if constexpr (sizeof(T) == 4) {
          const float4 q = *reinterpret_cast<const float4*>(cur + i);
          v[i] = q.x; v[i + 1] = q.y; v[i + 2] = q.z; v[i + 3] = q.w;
        } else {
          const double2 q = *reinterpret_cast<const double2*>(cur + i);
          v[i] = q.x; v[i + 1] = q.y;
        }
      }

      if constexpr (NS > 0) {
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const T a00 = p.a[s][0], a01 = p.a[s][1], a10 = p.a[s][2], a11 = p.a[s][3];
        const T c0 = p.c[s][0], c1 = p.c[s][1];
        T q0 = T(0), q1 = T(0);
        // 1. zero-state pass
#pragma unroll
        for (int i = 0; i < LC; ++i) {
          const T xi = v[i];
          T y, n0, n1;
          if constexpr (kPlain) {
            y = fma_t(c0, q0, fma_t(c1, q1, xi));
            n0 = fma_t(a00, q0, fma_t(a01, q1, xi));
            n1 = fma_t(a10, q0, a11 * q1);
          } else {
            y = fma_t(c0, q0, fma_t(c1, q1, p.dd[s] * xi));
            n0 = fma_t(a00, q0, fma_t(a01, q1, p.b[s][0] * xi));
            n1 = fma_t(a10, q0, fma_t(a11, q1, p.b[s][1] * xi));
          }
          v[i] = y;
          q0 = n0;
          q1 = n1;
        }
        // 2. prefix composition of chunk end states across the warp
        if (lane == 0) {
          const T s0 = carry[s][0], s1 = carry[s][1];
          q0 += fma_t(p.pw[s][0][0], s0, p.pw[s][0][1] * s1);
          q1 += fma_t(p.pw[s][0][2], s0, p.pw[s][0][3] * s1);
        }
#pragma unroll
        for (int d = 0; d < 5; ++d) {
          const T u0 = __shfl_up_sync(0xffffffffu, q0, 1 << d);
          const T u1 = __shfl_up_sync(0xffffffffu, q1, 1 << d);
          if (lane >= (1 << d)) {
            q0 += fma_t(p.pw[s][d][0], u0, p.pw[s][d][1] * u1);
            q1 += fma_t(p.pw[s][d][2], u0, p.pw[s][d][3] * u1);
          }
        }
        T e0 = __shfl_up_sync(0xffffffffu, q0, 1);
        T e1 = __shfl_up_sync(0xffffffffu, q1, 1);
        if (lane == 0) { e0 = carry[s][0]; e1 = carry[s][1]; }
        carry[s][0] = __shfl_sync(0xffffffffu, q0, 31);
        carry[s][1] = __shfl_sync(0xffffffffu, q1, 31);
        // 3. correction with the lane's true initial state
#pragma unroll
        for (int i = 0; i < LC; ++i) v[i] = fma_t(p.g[s][i][0], e0, fma_t(p.g[s][i][1], e1, v[i]));
      }
      }  // NS > 0

      // overall gain, clip (dsp_core.py:254), back through shared memory
#pragma unroll
      for (int i = 0; i < LC; ++i) {
        T y = v[i] * p.gain;
        if (p.clip) y = y < T(-1) ? T(-1) : (y > T(1) ? T(1) : y);   // NaN passes through like np.clip
        v[i] = y;
      }
#pragma unroll
      for (int i = 0; i < LC; i += VEC) {
        if constexpr (sizeof(T) == 4)
          *reinterpret_cast<float4*>(cur + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        else
          *reinterpret_cast<double2*>(cur + i) = make_double2(v[i], v[i + 1]);
      }
      __syncwarp();
      const T* out = buf + st * STAGE;
      const long long base = t * TILE;
      if (aligned && base + TILE <= n) {
#pragma unroll
        for (int k = 0; k < PIECES; ++k) {
          const int e = (lane + 32 * k) * VEC;
          const T* sp = out + (e / LC) * PITCH + (e % LC);
          if constexpr (sizeof(T) == 4)
            *reinterpret_cast<float4*>(zc + base + e) = *reinterpret_cast<const float4*>(sp);
          else
            *reinterpret_cast<double2*>(zc + base + e) = *reinterpret_cast<const double2*>(sp);
        }
      } else {
        for (int e = lane; e < TILE; e += 32) {
          const long long gi = base + e;
          if (gi < n) zc[gi] = out[(e / LC) * PITCH + (e % LC)];
        }
      }
      __syncwarp();   // the stage is refilled two iterations from now
    }
  }
}

// ---------------------------------------------------------------------------
// fp32 fast path: packed arithmetic + several warps per channel.
//
// * Packed: the lane's 32-sample chunk runs as two 16-sample halves side by side
//   in the two lanes of FFMA2/FMUL2 (section coefficients as the broadcast scalar
//   operand); the 2x2 state algebra of the scan is packed too (a matrix-vector
//   product is two FFMA2 on column pairs).  Pairs are carried as opaque 64-bit
//   registers so they stay in aligned register pairs for their whole life.
// * The section loop is NOT unrolled (runtime section count, one ~450-instruction
//   body that stays in the instruction cache); per-section constants come from
//   the kernel-parameter constant bank with a runtime index.
// * W warps per channel: warp wi takes tiles wi, wi+W, ... of the channel.
//   Everything except the carry fold is independent of the previous tile, so the
//   warps run as a wavefront: lane 31 publishes the tile's end state per section
//   through shared memory (+ an mbarrier when W > 1), the next tile's warp picks
//   it up just before its own fold.  One channel per CTA keeps the CTAs small
//   enough to balance 1024 long channels over 148 SMs.
constexpr int kEqPkLC = 32;
constexpr int kEqPkPitch = kEqPkLC + 4;
constexpr int kEqPkMaxNs = 8;

struct EqPackedSection {
  float a[4];                   // a00 a01 a10 a11
  float c[2];                   // c / d (direct gain folded into `gain`)
  float pad[2];
  float g[kEqPkLC / 2][2];      // (c/d) A^i, i < 16
  float pwc[5][4];              // A^(32*2^d) as column pairs (p00, p10, p01, p11)
  float phc[4];                 // A^16, column pairs
};
struct EqPackedParams {
  EqPackedSection sec[kEqPkMaxNs];
  float gain;
  int clip;
  int ns;
};

typedef unsigned long long pk2;
__device__ __forceinline__ pk2 pk_make(float lo, float hi) {
  pk2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float pk_lo(pk2 v) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return lo;
}
__device__ __forceinline__ float pk_hi(pk2 v) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return hi;
}
// a * s + c   with the scalar s broadcast to both halves
__device__ __forceinline__ pk2 pk_fma_s(pk2 a, float s, pk2 c) {
  pk2 r;
  asm("{\n.reg .b64 sv;\nmov.b64 sv, {%2, %2};\nfma.rn.f32x2 %0, %1, sv, %3;\n}" : "=l"(r) : "l"(a), "f"(s), "l"(c));
  return r;
}
__device__ __forceinline__ pk2 pk_mul_s(pk2 a, float s) {
  pk2 r;
  asm("{\n.reg .b64 sv;\nmov.b64 sv, {%2, %2};\nmul.rn.f32x2 %0, %1, sv;\n}" : "=l"(r) : "l"(a), "f"(s));
  return r;
}
// acc += col * s   (in place; call under a predicate for masked updates)
__device__ __forceinline__ void pk_acc(pk2& acc, pk2 col, float s) {
  asm("{\n.reg .b64 sv;\nmov.b64 sv, {%2, %2};\nfma.rn.f32x2 %0, %1, sv, %0;\n}" : "+l"(acc) : "l"(col), "f"(s));
}
__device__ __forceinline__ float clip_unit_nan(float y) {   // clip to [-1, 1], NaN passes through like np.clip
  float r;
  asm("max.NaN.f32 %0, %1, 0fBF800000;\n\tmin.NaN.f32 %0, %0, 0f3F800000;" : "=f"(r) : "f"(y));
  return r;
}

template <int NS, int W>
__global__ void __launch_bounds__(W * 32, 24 / W)
eq_packed_kernel(const __grid_constant__ EqPackedParams p, const float* __restrict__ x, long long x_stride,
                 float* __restrict__ z, long long z_stride, long long channels, long long n, int aligned) {
  constexpr int LC = kEqPkLC, H = LC / 2, PITCH = kEqPkPitch;
  constexpr int TILE = 32 * LC, PIECES = TILE / 4 / 32, STAGE = 32 * PITCH;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int wi = threadIdx.x >> 5;
  float* buf = reinterpret_cast<float*>(smem_raw) + static_cast<size_t>(wi) * 2 * STAGE;
  float2* cval = reinterpret_cast<float2*>(reinterpret_cast<float*>(smem_raw) + static_cast<size_t>(W) * 2 * STAGE);
  uint64_t* cbar = reinterpret_cast<uint64_t*>(cval + W * kEqPkMaxNs);   // [warp][section]
  if constexpr (W > 1) {
    if (threadIdx.x == 0) {
      for (int i = 0; i < W * kEqPkMaxNs; ++i) mbar_init(&cbar[i], 1);
      fence_mbar_init();
    }
    __syncthreads();
  }
  const long long n_tiles = (n + TILE - 1) / TILE;
  const int prev = (wi + W - 1) % W;                                   // warp that owns tile t-1
  const long long prev_tiles = n_tiles > prev ? (n_tiles - prev + W - 1) / W : 0;   // carries it publishes per channel
  uint64_t* wait_bar = cbar + prev * kEqPkMaxNs;
  const volatile float* wait_val = reinterpret_cast<const volatile float*>(cval + prev * kEqPkMaxNs);
  uint64_t* pub_bar = cbar + wi * kEqPkMaxNs;
  volatile float* pub_val = reinterpret_cast<volatile float*>(cval + wi * kEqPkMaxNs);
  // lane's 16-byte pieces of a tile: element 4*lane + 128*k -> chunk lane/8 + 4k, offset (4*lane)%32
  const int sm_off = (lane >> 3) * PITCH + ((4 * lane) & 31);

  long long ch_iter = 0;
  for (long long ch = blockIdx.x; ch < channels; ch += gridDim.x, ++ch_iter) {
    const float* xc = x + ch * x_stride;
    float* zc = z + ch * z_stride;
    if constexpr (W > 1) __syncthreads();   // all carries of the previous channel are consumed

    auto stage_in = [&](long long t, int st) {
      float* dst = buf + st * STAGE;
      const long long base = t * TILE;
      if (aligned && base + TILE <= n) {
        const float* src = xc + base + 4 * lane;
#pragma unroll
        for (int k = 0; k < PIECES; ++k) cp_async16(dst + sm_off + k * 4 * PITCH, src + 128 * k, 16);
      } else if (aligned) {
#pragma unroll 1
        for (int k = 0; k < PIECES; ++k) {
          const long long gi = base + 4 * lane + 128 * k;
          const long long rem = (n - gi) * 4;
          const int nbytes = rem >= 16 ? 16 : (rem > 0 ? static_cast<int>(rem) : 0);
          cp_async16(dst + sm_off + k * 4 * PITCH, nbytes > 0 ? xc + gi : xc, nbytes);
        }
      } else {
        for (int e = lane; e < TILE; e += 32) {
          const long long gi = base + e;
          dst[(e / LC) * PITCH + (e % LC)] = gi < n ? xc[gi] : 0.f;
        }
      }
      cp_async_commit();
    };

    if (wi < n_tiles) stage_in(wi, 0);
    int it = 0;
    for (long long t = wi; t < n_tiles; t += W, ++it) {
      const int st = it & 1;
      if (t + W < n_tiles) stage_in(t + W, st ^ 1);
      else cp_async_commit();
      cp_async_wait<1>();
      __syncwarp();
      float* cur = buf + st * STAGE + lane * PITCH;
      pk2 xp[H];
#pragma unroll
      for (int i = 0; i < H; i += 4) {
        const float4 lo = *reinterpret_cast<const float4*>(cur + i);
        const float4 hi = *reinterpret_cast<const float4*>(cur + H + i);
        xp[i] = pk_make(lo.x, hi.x); xp[i + 1] = pk_make(lo.y, hi.y);
        xp[i + 2] = pk_make(lo.z, hi.z); xp[i + 3] = pk_make(lo.w, hi.w);
      }
      // parity of the mbarrier phase that carries tile t-1's end states
      const uint32_t need = static_cast<uint32_t>((ch_iter * prev_tiles + (t - 1 - prev) / W) & 1);
      const bool first_tile = (t == 0);

#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const EqPackedSection& q = p.sec[s];
        const float a00 = q.a[0], a01 = q.a[1], a10 = q.a[2], a11 = q.a[3];
        const float c0 = q.c[0], c1 = q.c[1];
        pk2 q0 = 0ull, q1 = 0ull;
        // 1. zero-state pass over both half chunks
#pragma unroll
        for (int i = 0; i < H; ++i) {
          const pk2 xi = xp[i];
          const pk2 ty = pk_fma_s(q1, c1, xi);
          const pk2 t0 = pk_fma_s(q1, a01, xi);
          const pk2 y = pk_fma_s(q0, c0, ty);
          const pk2 n0 = pk_fma_s(q0, a00, t0);
          const pk2 n1 = pk_fma_s(q0, a10, pk_mul_s(q1, a11));
          xp[i] = y;
          q0 = n0;
          q1 = n1;
        }
        const pk2 ph0 = pk_make(q.phc[0], q.phc[1]), ph1 = pk_make(q.phc[2], q.phc[3]);
        const float fA0 = pk_lo(q0), fA1 = pk_lo(q1);
        pk2 F = pk_make(pk_hi(q0), pk_hi(q1));               // fB
        pk_acc(F, ph0, fA0);                                 // lane chunk end state from zero: fB + A^16 fA
        pk_acc(F, ph1, fA1);
        // 2. carry-in of the tile (end state of tile t-1), folded into lane 0
        float S0 = 0.f, S1 = 0.f;
        if (!first_tile) {                                   // warp-uniform
          if constexpr (W > 1) mbar_spin(&wait_bar[s], need);
          S0 = wait_val[2 * s];
          S1 = wait_val[2 * s + 1];
        }
        if (lane == 0) {
          pk_acc(F, pk_make(q.pwc[0][0], q.pwc[0][1]), S0);
          pk_acc(F, pk_make(q.pwc[0][2], q.pwc[0][3]), S1);
        }
        // 3. Kogge-Stone prefix composition across the warp
#pragma unroll
        for (int d = 0; d < 5; ++d) {
          const float ux = __shfl_up_sync(0xffffffffu, pk_lo(F), 1 << d);
          const float uy = __shfl_up_sync(0xffffffffu, pk_hi(F), 1 << d);
          if (lane >= (1 << d)) {
            pk_acc(F, pk_make(q.pwc[d][0], q.pwc[d][1]), ux);
            pk_acc(F, pk_make(q.pwc[d][2], q.pwc[d][3]), uy);
          }
        }
        float E0s = __shfl_up_sync(0xffffffffu, pk_lo(F), 1);
        float E1s = __shfl_up_sync(0xffffffffu, pk_hi(F), 1);
        if (lane == 0) { E0s = S0; E1s = S1; }
        __syncwarp();                                        // every lane has read the old carry slot
        if (lane == 31) {
          pub_val[2 * s] = pk_lo(F);
          pub_val[2 * s + 1] = pk_hi(F);
          if constexpr (W > 1) mbar_arrive(&pub_bar[s]);     // release: the next tile's warp may fold
        }
        __syncwarp();
        // 4. correction with the true initial states (second half chunk: A^16 E + fA)
        pk2 EB = pk_make(fA0, fA1);
        pk_acc(EB, ph0, E0s);
        pk_acc(EB, ph1, E1s);
        const pk2 E0 = pk_make(E0s, pk_lo(EB)), E1 = pk_make(E1s, pk_hi(EB));
#pragma unroll
        for (int i = 0; i < H; ++i) {
          pk_acc(xp[i], E1, q.g[i][1]);
          pk_acc(xp[i], E0, q.g[i][0]);
        }
      }

      // gain, clip (dsp_core.py:254), back through shared memory for coalesced stores
      float ox[H], oy[H];
#pragma unroll
      for (int i = 0; i < H; ++i) {
        const pk2 v = pk_mul_s(xp[i], p.gain);
        ox[i] = pk_lo(v);
        oy[i] = pk_hi(v);
        if (p.clip) { ox[i] = clip_unit_nan(ox[i]); oy[i] = clip_unit_nan(oy[i]); }
      }
#pragma unroll
      for (int i = 0; i < H; i += 4) {
        *reinterpret_cast<float4*>(cur + i) = make_float4(ox[i], ox[i + 1], ox[i + 2], ox[i + 3]);
        *reinterpret_cast<float4*>(cur + H + i) = make_float4(oy[i], oy[i + 1], oy[i + 2], oy[i + 3]);
      }
      __syncwarp();
      const float* out = buf + st * STAGE;
      const long long base = t * TILE;
      if (aligned && base + TILE <= n) {
        float* dstg = zc + base + 4 * lane;
#pragma unroll
        for (int k = 0; k < PIECES; ++k)
          *reinterpret_cast<float4*>(dstg + 128 * k) = *reinterpret_cast<const float4*>(out + sm_off + k * 4 * PITCH);
      } else {
        for (int e = lane; e < TILE; e += 32) {
          const long long gi = base + e;
          if (gi < n) zc[gi] = out[(e / LC) * PITCH + (e % LC)];
        }
      }
      __syncwarp();
    }
  }
}

// ---- plan ------------------------------------------------------------------
}  // namespace dspb200

struct dspb200_eq_plan {
  uint32_t magic = dspb200::kMagicEq;   // first member: checked by every entry point
  int dtype;
  int clip;
  int device;
  std::vector<dspb200::Section> sections;
  // tensor-core form (eq_mma.cu): tables built on first use, on the device current at that time
  std::mutex mma_mu;
  int mma_state = 0;   // 0 not tried, 1 built, -1 not available
  int mma_device = -1;
  dspb200::LtiMmaPlan mma;
  // fused SRC->EQ form (xz_mma.cu) for the resampler ratio it was last asked for; same lazy construction
  int xz_state = 0, xz_L = 0, xz_M = 0;
  dspb200::XzPlan xz;
};

namespace dspb200 {

template <typename T, int NS>
static void fill_params(EqKernelParams<T, NS>& kp, const Section* sec, bool plain, bool clip) {
  constexpr int LC = EqCfg<T>::LC;
  double gain = 1.0;
  for (int s = 0; s < NS; ++s) {
    const Section& S = sec[s];
    double c0 = S.c[0], c1 = S.c[1], dd = S.d;
    if (plain) {  // fold the direct gain: y/d = (c/d) . q + x
      gain *= S.d;
      c0 /= S.d;
      c1 /= S.d;
      dd = 1.0;
    }
    for (int i = 0; i < 4; ++i) kp.a[s][i] = static_cast<T>(S.a[i]);
    kp.b[s][0] = static_cast<T>(S.b0);
    kp.b[s][1] = static_cast<T>(S.b1);
    kp.c[s][0] = static_cast<T>(c0);
    kp.c[s][1] = static_cast<T>(c1);
    kp.dd[s] = static_cast<T>(dd);
    for (int i = 0; i < LC; ++i) {
      double m[4];
      mat2_power(S.a, i, m);
      kp.g[s][i][0] = static_cast<T>(c0 * m[0] + c1 * m[2]);
      kp.g[s][i][1] = static_cast<T>(c0 * m[1] + c1 * m[3]);
    }
    for (int d = 0; d < 5; ++d) {
      double m[4];
      mat2_power(S.a, static_cast<long long>(LC) << d, m);
      for (int i = 0; i < 4; ++i) kp.pw[s][d][i] = static_cast<T>(m[i]);
    }
    {
      double m[4];
      mat2_power(S.a, LC / 2, m);
      for (int i = 0; i < 4; ++i) kp.ph[s][i] = static_cast<T>(m[i]);
    }
  }
  kp.gain = static_cast<T>(gain);
  kp.clip = clip ? 1 : 0;
}

template <typename T, int NS, bool kPlain>
static int launch_pass(const Section* sec, bool clip, const T* x, int64_t xs, T* z, int64_t zs,
                       int64_t channels, int64_t n, cudaStream_t stream) {
  constexpr int LC = EqCfg<T>::LC;
  constexpr int PITCH = LC + EqCfg<T>::PAD;
  EqKernelParams<T, NS> kp;
  if constexpr (NS > 0) {
    fill_params<T, NS>(kp, sec, kPlain, clip);
  } else {
    kp.gain = T(1);
    kp.clip = clip ? 1 : 0;
  }
  const size_t smem = static_cast<size_t>(kEqWarpsPerCta) * 2 * 32 * PITCH * sizeof(T);
  auto kern = eq_scan_kernel<T, NS, kPlain>;
  DSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int vec = 16 / static_cast<int>(sizeof(T));
  const bool aligned = (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(z) % 16 == 0) &&
                       (xs % vec == 0) && (zs % vec == 0);
  const int64_t ctas_needed = ceil_div(channels, kEqWarpsPerCta);
  const int64_t max_ctas = static_cast<int64_t>(sm_count()) * 2;
  const int grid = static_cast<int>(ctas_needed < max_ctas ? ctas_needed : max_ctas);
  kern<<<grid, kEqWarpsPerCta * 32, smem, stream>>>(kp, x, xs, z, zs, channels, n, aligned ? 1 : 0);
  return after_launch("eq_scan_kernel");
}

template <int NS, int W>
static int launch_packed(const EqPackedParams& kp, const float* x, int64_t xs, float* z, int64_t zs,
                         int64_t channels, int64_t n, cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(W) * 2 * 32 * kEqPkPitch * sizeof(float) +
                      static_cast<size_t>(W) * kEqPkMaxNs * (sizeof(float2) + sizeof(uint64_t));
  auto kern = eq_packed_kernel<NS, W>;
  DSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const bool aligned = (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(z) % 16 == 0) &&
                       (xs % 4 == 0) && (zs % 4 == 0);
  const int64_t max_ctas = static_cast<int64_t>(sm_count()) * (24 / W);
  const int grid = static_cast<int>(channels < max_ctas ? channels : max_ctas);
  kern<<<grid, W * 32, smem, stream>>>(kp, x, xs, z, zs, channels, n, aligned ? 1 : 0);
  return after_launch("eq_packed_kernel");
}

static int dispatch_packed(int ns, const Section* sec, bool clip, const float* x, int64_t xs, float* z, int64_t zs,
                           int64_t channels, int64_t n, cudaStream_t stream) {
  EqPackedParams kp;
  memset(&kp, 0, sizeof(kp));
  double gain = 1.0;
  for (int s = 0; s < ns; ++s) {
    const Section& S = sec[s];
    EqPackedSection& q = kp.sec[s];
    gain *= S.d;
    const double c0 = S.c[0] / S.d, c1 = S.c[1] / S.d;
    for (int i = 0; i < 4; ++i) q.a[i] = static_cast<float>(S.a[i]);
    q.c[0] = static_cast<float>(c0);
    q.c[1] = static_cast<float>(c1);
    for (int i = 0; i < kEqPkLC / 2; ++i) {
      double m[4];
      mat2_power(S.a, i, m);
      q.g[i][0] = static_cast<float>(c0 * m[0] + c1 * m[2]);
      q.g[i][1] = static_cast<float>(c0 * m[1] + c1 * m[3]);
    }
    auto cols = [](const double m[4], float out[4]) {
      out[0] = static_cast<float>(m[0]); out[1] = static_cast<float>(m[2]);
      out[2] = static_cast<float>(m[1]); out[3] = static_cast<float>(m[3]);
    };
    for (int d = 0; d < 5; ++d) {
      double m[4];
      mat2_power(S.a, static_cast<long long>(kEqPkLC) << d, m);
      cols(m, q.pwc[d]);
    }
    double m[4];
    mat2_power(S.a, kEqPkLC / 2, m);
    cols(m, q.phc);
  }
  kp.gain = static_cast<float>(gain);
  kp.clip = clip ? 1 : 0;
  kp.ns = ns;
  // warps per channel: as few as still fill the machine (24 warps per SM)
  const int64_t resident = static_cast<int64_t>(sm_count()) * 24;
  int w = 1;
  if (const char* e = getenv("DSPB200_EQ_WARPS_PER_CHANNEL")) w = atoi(e);
  else if (n > 4 * 1024) {
    // largest W that (nearly) fits one wave of resident warps; a 20 % overshoot
    // measured faster than dropping to the next smaller W
    w = static_cast<int>((resident * 6 / 5) / (channels > 0 ? channels : 1));
    w = w < 1 ? 1 : (w > 4 ? 4 : w);
  }
#define DSP_EQ_PK(NSV)                                                                 \
  case NSV:                                                                            \
    if (w >= 4) return launch_packed<NSV, 4>(kp, x, xs, z, zs, channels, n, stream);   \
    if (w == 3) return launch_packed<NSV, 3>(kp, x, xs, z, zs, channels, n, stream);   \
    if (w == 2) return launch_packed<NSV, 2>(kp, x, xs, z, zs, channels, n, stream);   \
    return launch_packed<NSV, 1>(kp, x, xs, z, zs, channels, n, stream);
  switch (ns) {
    DSP_EQ_PK(1) DSP_EQ_PK(2) DSP_EQ_PK(3) DSP_EQ_PK(4) DSP_EQ_PK(5) DSP_EQ_PK(6) DSP_EQ_PK(7) DSP_EQ_PK(8)
    default: return fail(DSPB200_ERR_INVALID, "internal: bad section count %d", ns);
  }
#undef DSP_EQ_PK
}

template <typename T, bool kPlain>
static int dispatch_ns(int ns, const Section* sec, bool clip, const T* x, int64_t xs, T* z, int64_t zs,
                       int64_t channels, int64_t n, cudaStream_t stream) {
  switch (ns) {
    case 1: return launch_pass<T, 1, kPlain>(sec, clip, x, xs, z, zs, channels, n, stream);
    case 2: return launch_pass<T, 2, kPlain>(sec, clip, x, xs, z, zs, channels, n, stream);
    case 3: return launch_pass<T, 3, kPlain>(sec, clip, x, xs, z, zs, channels, n, stream);
    case 4: return launch_pass<T, 4, kPlain>(sec, clip, x, xs, z, zs, channels, n, stream);
    case 5: return launch_pass<T, 5, kPlain>(sec, clip, x, xs, z, zs, channels, n, stream);
    case 6: return launch_pass<T, 6, kPlain>(sec, clip, x, xs, z, zs, channels, n, stream);
    case 7: return launch_pass<T, 7, kPlain>(sec, clip, x, xs, z, zs, channels, n, stream);
    case 8: return launch_pass<T, 8, kPlain>(sec, clip, x, xs, z, zs, channels, n, stream);
    default: return fail(DSPB200_ERR_INVALID, "internal: bad section count %d", ns);
  }
}

// Would a run of this shape use the tensor-core form (eq_mma.cu)?  Builds the plan's tables on first use.
static int eq_mma_ready(const dspb200_eq_plan* plan, const float* x, int64_t xs, const float* z, int64_t zs,
                        int64_t channels, int64_t n, bool& tensor) {
  tensor = false;
  const int total = static_cast<int>(plan->sections.size());
  if (plan->dtype != DSPB200_F32 || total < 1 || total > kLtiMaxStates / 2 || getenv("DSPB200_EQ_NO_MMA") != nullptr)
    return DSPB200_OK;
  int dev = 0;
  DSP_CUDA(cudaGetDevice(&dev));
  dspb200_eq_plan* mp = const_cast<dspb200_eq_plan*>(plan);
  {
    std::lock_guard<std::mutex> lk(mp->mma_mu);
    if (mp->mma_state == 0) {
      DSP_TRY(lti_mma_build_eq(mp->sections.data(), total, mp->mma));
      mp->mma_state = mp->mma.ok ? 1 : -1;
      mp->mma_device = dev;
    }
    if (mp->mma_state != 1 || mp->mma_device != dev) return DSPB200_OK;
  }
  tensor = lti_mma_usable(plan->mma, x, xs, z, zs, channels, n);
  return DSPB200_OK;
}

// Does a float32 batch of this shape run the tensor-core form ONLY out of place?  (Narrow batches: the overlapping
// time slices re-read inputs before their own range, so in place they fall back to the scan kernel.)  The chain then
// gives the resampler's output a scratch buffer instead of equalising in place.
int eq_prefers_out_of_place(const dspb200_eq_plan* plan, int64_t channels, int64_t n, int64_t stride, bool* prefers) {
  *prefers = false;
  if (plan == nullptr || plan->dtype != DSPB200_F32) return DSPB200_OK;
  bool oop = false, inp = false;
  DSP_TRY(eq_mma_ready(plan, nullptr, stride, nullptr, stride, channels, n, oop));
  if (!oop) return DSPB200_OK;
  const float* same = reinterpret_cast<const float*>(static_cast<uintptr_t>(256));   // any aligned pointer, used as x and z
  DSP_TRY(eq_mma_ready(plan, same, stride, same, stride, channels, n, inp));
  *prefers = !inp;
  return DSPB200_OK;
}

template <typename T>
int eq_run(const dspb200_eq_plan* plan, const T* x, int64_t xs, T* z, int64_t zs, int64_t channels,
           int64_t n, cudaStream_t stream) {
  DSP_PLAN(plan, kMagicEq, "eq");
  DSP_CHECK(plan->dtype == DType<T>::id, "plan dtype %d does not match the entry point", plan->dtype);
  DSP_CHECK(channels >= 0 && n >= 0, "negative shape");
  if (channels == 0 || n == 0) return DSPB200_OK;
  DSP_CHECK(x != nullptr && z != nullptr, "NULL buffer");
  DSP_CHECK(xs >= n && zs >= n, "channel stride smaller than n");
  DSP_TRY(ensure_device());
  const int total = static_cast<int>(plan->sections.size());
  if (total == 0)
    return launch_pass<T, 0, true>(nullptr, plan->clip != 0, x, xs, z, zs, channels, n, stream);
  if (sizeof(T) == 4) {
    bool tensor = false;
    DSP_TRY(eq_mma_ready(plan, reinterpret_cast<const float*>(x), xs, reinterpret_cast<const float*>(z), zs, channels, n, tensor));
    if (tensor) {
      return lti_mma_run(plan->mma, reinterpret_cast<const float*>(x), xs, reinterpret_cast<float*>(z), zs,
                         channels, n, n, plan->clip != 0, nullptr, false, stream);
    }
  }
  const T* src = x;
  int64_t src_stride = xs;
  for (int first = 0; first < total; first += kEqPassSections) {
    const int ns = (total - first) < kEqPassSections ? (total - first) : kEqPassSections;
    const bool last = first + ns >= total;
    const Section* sec = plan->sections.data() + first;
    bool plain = true;
    for (int s = 0; s < ns; ++s)
      if (!sec[s].complex_poles || sec[s].d == 0.0 || !std::isfinite(1.0 / sec[s].d)) plain = false;
    const bool clip = last && plan->clip != 0;
    if (plain && sizeof(T) == 4 && getenv("DSPB200_EQ_NO_PACKED") == nullptr)
      DSP_TRY(dispatch_packed(ns, sec, clip, reinterpret_cast<const float*>(src), src_stride,
                              reinterpret_cast<float*>(z), zs, channels, n, stream));
    else if (plain)
      DSP_TRY((dispatch_ns<T, true>(ns, sec, clip, src, src_stride, z, zs, channels, n, stream)));
    else
      DSP_TRY((dispatch_ns<T, false>(ns, sec, clip, src, src_stride, z, zs, channels, n, stream)));
    src = z;  // later passes run in place on the output
    src_stride = zs;
  }
  return DSPB200_OK;
}

template int eq_run<float>(const dspb200_eq_plan*, const float*, int64_t, float*, int64_t, int64_t, int64_t, cudaStream_t);
template int eq_run<double>(const dspb200_eq_plan*, const double*, int64_t, double*, int64_t, int64_t, int64_t, cudaStream_t);

int eq_plan_clip(const dspb200_eq_plan* plan) { return (plan && plan->magic == kMagicEq) ? plan->clip : 0; }

int eq_plan_xz(const dspb200_eq_plan* eq, const dspb200_src_plan* src, const XzPlan** xp) {
  *xp = nullptr;
  DSP_PLAN(eq, kMagicEq, "eq");
  const std::vector<double>* taps = src_plan_taps(src);
  DSP_CHECK(taps != nullptr, "handle is not a live src plan");
  int L = 0, M = 0, dt = 0;
  DSP_TRY(src_plan_ratio(src, &L, &M, &dt));
  const int total = static_cast<int>(eq->sections.size());
  if (eq->dtype != DSPB200_F32 || dt != DSPB200_F32 || total < 1 || total > kLtiMaxStates / 2) return DSPB200_OK;
  int dev = 0;
  DSP_CUDA(cudaGetDevice(&dev));
  dspb200_eq_plan* mp = const_cast<dspb200_eq_plan*>(eq);
  std::lock_guard<std::mutex> lk(mp->mma_mu);
  if (mp->xz_state != 0 && (mp->xz_L != L || mp->xz_M != M || (mp->xz_state == 1 && mp->xz.device != dev))) {
    if (mp->xz_state == 1) xz_free(mp->xz);
    mp->xz_state = 0;
  }
  if (mp->xz_state == 0) {
    DSP_TRY(xz_build(*taps, L, M, mp->sections.data(), total, mp->xz));
    mp->xz_state = mp->xz.ok ? 1 : -1;
    mp->xz_L = L;
    mp->xz_M = M;
  }
  if (mp->xz_state == 1) *xp = &mp->xz;
  return DSPB200_OK;
}

int eq_plan_dtype(const dspb200_eq_plan* plan) { return (plan && plan->magic == kMagicEq) ? plan->dtype : -1; }

template <typename T>
static int eq_host(const dspb200_eq_plan* plan, const T* x, T* z, int64_t channels, int64_t n) {
  DSP_PLAN(plan, kMagicEq, "eq");
  DSP_CHECK(channels >= 0 && n >= 0, "negative shape");
  if (channels == 0 || n == 0) return DSPB200_OK;
  DSP_CHECK(x != nullptr && z != nullptr, "NULL buffer");
  DSP_TRY(ensure_device());
  const int vec = 16 / static_cast<int>(sizeof(T));
  const int64_t pitch = round_up(n, vec);
  T* d = nullptr;
  DSP_CUDA(cudaMalloc(&d, static_cast<size_t>(channels) * pitch * sizeof(T)));
  cudaStream_t st = nullptr;
  int rc = DSPB200_OK;
  cudaError_t e = cudaMemcpy2DAsync(d, pitch * sizeof(T), x, n * sizeof(T), n * sizeof(T), channels,
                                    cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) rc = eq_run<T>(plan, d, pitch, d, pitch, channels, n, st);
  if (e == cudaSuccess && rc == DSPB200_OK)
    e = cudaMemcpy2DAsync(z, n * sizeof(T), d, pitch * sizeof(T), n * sizeof(T), channels,
                          cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d);
  if (e != cudaSuccess) return fail(DSPB200_ERR_CUDA, "eq host path: %s", cudaGetErrorString(e));
  return rc;
}

static int make_plan(std::vector<Section>&& secs, int clip, int dtype, dspb200_eq_plan** out) {
  DSP_CHECK(out != nullptr, "plan output pointer is NULL");
  DSP_CHECK(dtype == DSPB200_F32 || dtype == DSPB200_F64, "dtype must be 0 (f32) or 1 (f64)");
  DSP_CHECK(secs.size() <= DSPB200_MAX_SECTIONS, "at most %d sections per plan", DSPB200_MAX_SECTIONS);
  dspb200_eq_plan* p = new (std::nothrow) dspb200_eq_plan();
  if (!p) return fail(DSPB200_ERR_ALLOC, "out of host memory");
  p->dtype = dtype;
  p->clip = clip ? 1 : 0;
  p->device = -1;
  p->sections = std::move(secs);
  *out = p;
  return DSPB200_OK;
}

}  // namespace dspb200

using namespace dspb200;

extern "C" {

int dspb200_eq_plan_create(double fs, const double* fc_eff, const double* gains_db, int n_sections,
                           int clip, int dtype, dspb200_eq_plan** plan) {
  DSP_CHECK(n_sections >= 0, "n_sections must be >= 0");
  DSP_CHECK(fs > 0.0, "fs must be positive");
  DSP_CHECK(n_sections == 0 || (fc_eff && gains_db), "NULL section arrays");
  std::vector<Section> secs;
  for (int i = 0; i < n_sections; ++i) {
    DSP_CHECK(fc_eff[i] > 0.0 && fc_eff[i] < fs / 2.0, "section %d: centre %g Hz outside (0, fs/2)", i, fc_eff[i]);
    double b[3], a[3];
    peaking_biquad(fc_eff[i], fs, gains_db[i], b, a);
    secs.push_back(section_from_ba(b, a));
  }
  return make_plan(std::move(secs), clip, dtype, plan);
}

int dspb200_eq_plan_create_raw(const double* ba, int n_sections, int clip, int dtype,
                               dspb200_eq_plan** plan) {
  DSP_CHECK(n_sections >= 0, "n_sections must be >= 0");
  DSP_CHECK(n_sections == 0 || ba != nullptr, "ba is NULL");
  std::vector<Section> secs;
  for (int i = 0; i < n_sections; ++i) {
    const double* r = ba + 6 * i;
    DSP_CHECK(r[3] != 0.0, "section %d: a0 must be non-zero", i);
    secs.push_back(section_from_ba(r, r + 3));
  }
  return make_plan(std::move(secs), clip, dtype, plan);
}

int dspb200_eq_plan_create_bands(double fs, const double gains_db[DSPB200_EQ_BANDS], int dtype,
                                 dspb200_eq_plan** plan, int* bypass) {
  DSP_CHECK(gains_db != nullptr && bypass != nullptr, "NULL argument");
  static const double centres[DSPB200_EQ_BANDS] = {40, 150, 1000, 3000, 5000, 10000};  // :225-228
  double fc[DSPB200_EQ_BANDS], g[DSPB200_EQ_BANDS];
  int n = 0;
  DSP_TRY(dspb200_eq_select_sections(fs, centres, gains_db, DSPB200_EQ_BANDS, fc, g, &n, bypass));
  if (*bypass) return make_plan({}, 0, dtype, plan);
  return dspb200_eq_plan_create(fs, fc, g, n, 1, dtype, plan);
}

int dspb200_eq_plan_destroy(dspb200_eq_plan* plan) {
  if (!plan) return DSPB200_OK;
  DSP_PLAN(plan, kMagicEq, "eq");
  plan->magic = 0;
  if (plan->mma_state == 1) lti_mma_free(plan->mma);
  if (plan->xz_state == 1) xz_free(plan->xz);
  delete plan;
  return DSPB200_OK;
}

int dspb200_eq_plan_kernel_kind(const dspb200_eq_plan* plan, int64_t channels, int64_t n, int64_t x_stride,
                                int* kind) {
  DSP_PLAN(plan, kMagicEq, "eq");
  DSP_CHECK(kind != nullptr, "NULL argument");
  DSP_TRY(ensure_device());
  bool tensor = false;
  DSP_TRY(eq_mma_ready(plan, nullptr, x_stride, nullptr, x_stride, channels, n, tensor));
  *kind = tensor ? 1 : 0;
  return DSPB200_OK;
}

int dspb200_eq_plan_warm_chunks(const dspb200_eq_plan* plan, int* chunks) {
  DSP_PLAN(plan, kMagicEq, "eq");
  DSP_CHECK(chunks != nullptr, "NULL argument");
  LtiChunkSystem cs;
  DSP_TRY(lti_chunk_system(plan->sections.data(), static_cast<int>(plan->sections.size()), cs));
  *chunks = cs.rows > 0 ? lti_warm_chunks(cs) : 0;
  return DSPB200_OK;
}

int dspb200_eq_plan_chunk_system(const dspb200_eq_plan* plan, int* rows, int* states, double* tk, double* o,
                                 double* phi) {
  DSP_PLAN(plan, kMagicEq, "eq");
  DSP_CHECK(rows != nullptr && states != nullptr, "NULL argument");
  LtiChunkSystem cs;
  DSP_TRY(lti_chunk_system(plan->sections.data(), static_cast<int>(plan->sections.size()), cs));
  *rows = cs.rows;
  *states = cs.states;
  if (cs.rows == 0) return DSPB200_OK;
  if (tk) memcpy(tk, cs.tk.data(), cs.tk.size() * sizeof(double));
  if (o) memcpy(o, cs.o.data(), cs.o.size() * sizeof(double));
  if (phi) memcpy(phi, cs.phi.data(), cs.phi.size() * sizeof(double));
  return DSPB200_OK;
}

int dspb200_eq_plan_describe(const dspb200_eq_plan* plan, int* n_sections, double* ss, int capacity) {
  DSP_PLAN(plan, kMagicEq, "eq");
  DSP_CHECK(n_sections != nullptr, "NULL argument");
  *n_sections = static_cast<int>(plan->sections.size());
  for (int i = 0; i < *n_sections && i < capacity && ss; ++i) {
    const Section& s = plan->sections[static_cast<size_t>(i)];
    double* o = ss + 9 * i;
    o[0] = s.a[0]; o[1] = s.a[1]; o[2] = s.a[2]; o[3] = s.a[3];
    o[4] = s.b0; o[5] = s.b1; o[6] = s.c[0]; o[7] = s.c[1]; o[8] = s.d;
  }
  return DSPB200_OK;
}

int dspb200_eq_run_f32(const dspb200_eq_plan* plan, const float* x, int64_t xs, float* z, int64_t zs,
                       int64_t channels, int64_t n, void* stream) {
  return eq_run<float>(plan, x, xs, z, zs, channels, n, static_cast<cudaStream_t>(stream));
}
int dspb200_eq_run_f64(const dspb200_eq_plan* plan, const double* x, int64_t xs, double* z, int64_t zs,
                       int64_t channels, int64_t n, void* stream) {
  return eq_run<double>(plan, x, xs, z, zs, channels, n, static_cast<cudaStream_t>(stream));
}
int dspb200_eq_stream_chunk(void) { return lti_mma_chunk(); }

int dspb200_eq_run_stream_f32(const dspb200_eq_plan* plan, const float* x, int64_t xs, float* z, int64_t zs,
                              int64_t channels, int64_t n, float* state, int first, void* stream) {
  DSP_PLAN(plan, kMagicEq, "eq");
  DSP_CHECK(state != nullptr, "NULL argument");
  DSP_CHECK(plan->dtype == DSPB200_F32, "the streaming form is float32 only");
  DSP_CHECK(channels >= 0 && n >= 0, "negative shape");
  if (channels == 0 || n == 0) return DSPB200_OK;
  DSP_CHECK(x != nullptr && z != nullptr, "NULL buffer");
  DSP_CHECK(xs >= n && zs >= n, "channel stride smaller than n");
  DSP_TRY(ensure_device());
  bool tensor = false;
  DSP_TRY(eq_mma_ready(plan, nullptr, 4, nullptr, 4, 0, 0, tensor));   // builds the plan's tables
  if (plan->mma_state != 1 || !lti_mma_possible(plan->mma, x, xs, z, zs))
    return fail(DSPB200_ERR_UNSUPPORTED,
                "the streaming form needs 1..8 sections and 16-byte aligned rows (x, z and both strides)");
  return lti_mma_run(plan->mma, x, xs, z, zs, channels, n, n, plan->clip != 0, state, first == 0,
                     static_cast<cudaStream_t>(stream));
}

int dspb200_eq_host_f32(const dspb200_eq_plan* plan, const float* x, float* z, int64_t channels, int64_t n) {
  DSP_PLAN(plan, kMagicEq, "eq");
  DSP_CHECK(plan->dtype == DSPB200_F32, "plan is not float32");
  return eq_host<float>(plan, x, z, channels, n);
}
int dspb200_eq_host_f64(const dspb200_eq_plan* plan, const double* x, double* z, int64_t channels, int64_t n) {
  DSP_PLAN(plan, kMagicEq, "eq");
  DSP_CHECK(plan->dtype == DSPB200_F64, "plan is not float64");
  return eq_host<double>(plan, x, z, channels, n);
}

}  // extern "C"
