// K3, 4096-point Hann magnitude frames (config C5, app.py:202-205) -- 32 points per thread.
//
// Same DFT as fft_fixed_kernel (dsp_core.py:41-98: symmetric Hann, radix-2 decimation in time, |.| of the first
// N/2+1 bins), regrouped so that a frame crosses shared memory TWICE instead of three times.  fft_fixed_kernel
// (16 points per thread, 2048 = 8 x 16 x 16) is bound by the SM's L1/shared data pipe: 1127 wavefronts per frame
// (ncu, profiles/r2a_ncu_full_fft.md) against the ~890 cycles an SM may spend per frame at the HBM roofline.  Here
//
//   n = n0 + 2 n1 + 64 n2,  k = k2 + 32 k1 + 1024 k0      (n0, k0 < 2; n1, n2, k1, k2 < 32)
//
//   pass 0   thread (n0, n1) owns z[n0 + 2 n1 + 64 n2], n2 = 0..31 (coalesced 8-byte loads), applies the window
//            and a 32-point DFT over n2 -> A[n0, n1, k2]
//   exchange through shared memory (the only one between passes)
//   pass 1   thread (n0, k2) multiplies by W_1024^(n1 k2) and transforms over n1 -> the two 1024-point spectra
//            E (n0 = 0: even points) and O (n0 = 1: odd points) at k' = k2 + 32 k1
//   split    E and O go to shared memory once more; a thread reads E[q], O[q], E[1024-q], O[1024-q] and forms the
//            last radix-2 level (Z[q], Z[q+1024], Z[1024-q], Z[2048-q]) together with the real split: four
//            magnitudes per quad, |X[q]|, |X[2048-q]|, |X[1024-q]|, |X[1024+q]|, stored in coalesced runs.
//
// 64 threads (two warps) per frame; a CTA holds G independent groups that synchronise on named barriers.  All
// twiddles of the split come from W_4096^t (two registers per thread) times compile-time constants; the pass-1
// twiddles from ten table entries per thread (w^1..w^3, w^4, w^8, .., w^28: one rounding deep).  The window is
// applied by angle addition as in fft_fixed_kernel, pre-scaled by 1/2 so the real split needs no halving.
#include <cmath>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "cpx.cuh"
#include "fft_r32.cuh"
#include "internal.cuh"

namespace dspb200 {

namespace {

using namespace r32;

constexpr int kN = 4096;          // real samples per frame
constexpr int kM = 2048;          // complex points
constexpr int kPitch = 33;        // exchange rows [64 threads][33]
constexpr int kBuf = 64 * kPitch; // complex entries per group buffer (>= 2048)
constexpr int kTabVt = kTw1;              // W_4096^t, t < 64
constexpr int kTabHann = kTabVt + 64;     // per thread: (A(2t), A(2t+1)), (B(2t), B(2t+1))
constexpr int kTabTotal = kTabHann + 128;

struct R32Args {
  const float* x;
  long long x_stride, n_valid, offset, hop, n_frames;
  float* mag;
  long long mfs, mcs;
  long long n_items;
  const float2* tables;
  int db, hann;
  float2 cc[32], ss[32];   // cos/sin(s * 2 pi 128/(N-1)), both halves alike
};

__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

template <int G, int MINB, bool kPrefetch, bool kFold = true>
__global__ void __launch_bounds__(64 * G, MINB) fft4096_r32_kernel(const R32Args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tw1 = reinterpret_cast<float2*>(smem_raw);
  const int g = threadIdx.x >> 6, t = threadIdx.x & 63, warp = t >> 5, lane = t & 31;
  float2* xb = tw1 + kTw1 + g * kBuf;
  for (int i = threadIdx.x; i < kTw1; i += 64 * G) tw1[i] = a.tables[i];
  const float2 vt = a.tables[kTabVt + t];                       // W_4096^t
  const float2 hann_a = a.tables[kTabHann + 2 * t], hann_b = a.tables[kTabHann + 2 * t + 1];
  __syncthreads();
  const int bar = g + 1;
  const float2* twr = tw1 + lane * kTw1Pitch;

  const long long stride = static_cast<long long>(gridDim.x) * G;
  long long item = static_cast<long long>(blockIdx.x) * G + g;
  long long c = item / a.n_frames, fr = item - c * a.n_frames;
  const long long dc = stride / a.n_frames, dfr = stride - dc * a.n_frames;

  for (; item < a.n_items; item += stride) {
    float2 v[32];
    long long cn = c + dc, frn = fr + dfr;                      // the group's next frame
    if (frn >= a.n_frames) { frn -= a.n_frames; ++cn; }
    if (kPrefetch && t == 0 && item + stride < a.n_items) {
      // bring it into L2 while this one is transformed: the loads at the top of the next iteration are this kernel's
      // longest stall (no registers to land them in early)
      const long long fstart = a.offset + frn * a.hop;
      if (fstart + kN <= a.n_valid) {
        const uintptr_t p0 = reinterpret_cast<uintptr_t>(a.x + cn * a.x_stride + fstart);
        // whole 16-byte units INSIDE the frame only (a misaligned frame loses its first and last few bytes)
        const uintptr_t p1 = (p0 + 15) & ~static_cast<uintptr_t>(15);
        prefetch_l2_bulk(reinterpret_cast<const void*>(p1), static_cast<uint32_t>((p0 + kN * 4 - p1) & ~static_cast<uintptr_t>(15)));
      }
    }
    {
      const float* xrow = a.x + c * a.x_stride;
      const long long fstart = a.offset + fr * a.hop;
      const bool fast = (fstart + kN <= a.n_valid) && ((reinterpret_cast<uintptr_t>(xrow + fstart) & 7) == 0);
      if (fast) {
        const float2* xp = reinterpret_cast<const float2*>(xrow + fstart) + t;
#pragma unroll
        for (int s = 0; s < 32; ++s) v[s] = xp[s * 64];
      } else {
        const long long left = a.n_valid - fstart;
        const int rem = left > kN ? kN : (left < 0 ? 0 : static_cast<int>(left));
        const float* xf = xrow + fstart;
#pragma unroll
        for (int s = 0; s < 32; ++s) {
          const int e = 2 * (t + 64 * s);
          v[s].x = e < rem ? xf[e] : 0.f;
          v[s].y = e + 1 < rem ? xf[e + 1] : 0.f;
        }
      }
    }
    if (a.hann && kFold) {
      dft32_windowed(v, hann_a, hann_b, a.cc, a.ss);
    } else if (a.hann) {
      const float2 quarter = make_float2(0.25f, 0.25f);
#pragma unroll
      for (int s = 0; s < 32; ++s) v[s] = pmul(v[s], fma2(hann_a, a.cc[s], fma2(hann_b, a.ss[s], quarter)));
      Dft32<32>::run(v);
    } else {
#pragma unroll
      for (int s = 0; s < 32; ++s) v[s] = pscale(v[s], 0.5f);
      Dft32<32>::run(v);
    }
    {
      float2* wp = xb + t * kPitch;
#pragma unroll
      for (int k2 = 0; k2 < 32; ++k2) wp[k2] = v[k2];
    }
    group_sync(bar);
    {
      const float2* rp = xb + warp * kPitch + lane;
#pragma unroll
      for (int n1 = 0; n1 < 32; ++n1) v[n1] = rp[n1 * 2 * kPitch];
    }
    if constexpr (kFold) {
      dft32_twiddled(v, twr);   // W_1024^(n1 k2) folded into the first butterfly level
    } else {
      twiddle_powers(v, twr);
      Dft32<32>::run(v);
    }
    {
      // in place: thread (n0, k2) puts its k1-th output where it read its n1 = k1 input, so no other thread's
      // operands are overwritten and no barrier is needed between the reads above and these stores.
      // E[k2 + 32 k1] sits at row 2 k1, O[...] at row 2 k1 + 1, column k2.
      float2* wp = xb + warp * kPitch + lane;
#pragma unroll
      for (int k1 = 0; k1 < 32; ++k1) wp[k1 * 2 * kPitch] = v[k1];
    }
    group_sync(bar);
    {
      float* mg = a.mag + c * a.mcs + fr * a.mfs;
      // quad q = t + 64 i = k2 + 32 k1 with k2 = lane, k1 = warp + 2 i; its mirror 1024 - q has k2' = (32 - lane) & 31 and
      // k1' = 31 - k1 (lane > 0) or (32 - k1) & 31 (lane == 0)
      const float2* Eq0 = xb + 2 * kPitch * warp + lane;
      const float2* Ej0 = lane ? xb + 2 * kPitch * (31 - warp) + 32 - lane : xb + 2 * kPitch * (32 - warp);
      const float2* Ej00 = t == 0 ? xb : Ej0;                 // q = 0 is its own mirror
      auto quad = [&](float* mq, float* mr, float2 vq, const float2* pe, const float2* pj, auto db_tag) {   // mq = mg + q, mr = mg - q
        constexpr bool kDb = decltype(db_tag)::value;
        const float2 Eq = pe[0], Oq = pe[kPitch], Ej = pj[0], Oj = pj[kPitch];
        const float2 w = cmul(vq, vq);                        // W_2048^q
        const float2 A = cmadd(w, Oq, Eq), B2 = twice_minus(Eq, A);           // Z[q], Z[1024 + q]
        const float2 B = cmadd(cconj(w), Oj, Ej), A2 = twice_minus(Ej, B);    // Z[2048 - q], Z[1024 - q]
        {
          const float2 Bc = cconj(B);
          const float2 S = cadd(A, Bc), D = csub(A, Bc);
          const float2 X1 = cmadd(make_float2(vq.y, -vq.x), D, S);   // S - i W_4096^q (A - conj B)
          mq[0] = mag_of<kDb>(X1);
          mr[kM] = mag_of<kDb>(twice_minus(S, X1));
        }
        {
          const float2 Bc = cconj(B2);
          const float2 S = cadd(A2, Bc), D = csub(A2, Bc);
          const float2 X1 = cmadd(make_float2(-vq.x, vq.y), D, S);   // -i W_4096^(1024-q) = -conj(W_4096^q)
          mr[1024] = mag_of<kDb>(X1);
          mq[1024] = mag_of<kDb>(twice_minus(S, X1));
        }
      };
      float* mlo = mg + t;
      float* mhi = mg - t;
      auto split = [&](auto db_tag) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 ci2 = make_float2(static_cast<float>(kW64r[i]), static_cast<float>(kW64i[i]));   // W_4096^(64 i)
          const float2 vq = i == 0 ? vt : cmul(vt, ci2);
          quad(mlo + 64 * i, mhi - 64 * i, vq, Eq0 + 4 * kPitch * i, (i == 0 ? Ej00 : Ej0) - 4 * kPitch * i, db_tag);
        }
        if (t == 0)
          quad(mg + 512, mg - 512, make_float2(0.70710678118654752440f, -0.70710678118654752440f), xb + 32 * kPitch, xb + 32 * kPitch, db_tag);
      };
      if (a.db) split(std::true_type{}); else split(std::false_type{});
    }
    group_sync(bar);   // the split's readers are done with xb before the next frame's pass 0 writes it
    c = cn;
    fr = frn;
  }
}

template <int G, int MINB, bool kPrefetch = true, bool kFold = true>
int launch_r32(const R32Args& a, cudaStream_t stream) {
  auto kern = fft4096_r32_kernel<G, MINB, kPrefetch, kFold>;
  const size_t smem = static_cast<size_t>(kTw1 + G * kBuf) * sizeof(float2);
  DSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int per_sm = 1;
  DSP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 64 * G, smem));
  if (per_sm < 1) per_sm = 1;
  if (const char* cv = getenv("DSPB200_FFT_MAX_CTAS")) {
    const int lim = atoi(cv);
    if (lim >= 1 && lim < per_sm) per_sm = lim;
  }
  if (getenv("DSPB200_FFT_TRACE")) fprintf(stderr, "fft4096_r32 G=%d smem=%zu per_sm=%d\n", G, smem, per_sm);
  const long long cap = static_cast<long long>(sm_count()) * per_sm;
  const long long ctas = ceil_div(a.n_items, G);
  const int grid = static_cast<int>(ctas < cap ? ctas : cap);
  kern<<<grid, 64 * G, smem, stream>>>(a);
  return after_launch("fft4096_r32_kernel");
}

}  // namespace

int fft_r32_build(int n_fft, FftR32Plan& rp) {
  rp.ok = 0;
  if (n_fft != kN) return DSPB200_OK;
  const long double pi = 3.14159265358979323846264338327950288L;
  std::vector<float2> h(static_cast<size_t>(kTabTotal), make_float2(1.f, 0.f));
  auto w = [&](long double num, long double den) {
    const long double ang = -2.0L * pi * num / den;
    return make_float2(static_cast<float>(cosl(ang)), static_cast<float>(sinl(ang)));
  };
  for (int k2 = 0; k2 < 32; ++k2) fill_twiddle_row(&h[static_cast<size_t>(k2 * kTw1Pitch)], k2, 1024.0L);
  for (int t = 0; t < 64; ++t) h[static_cast<size_t>(kTabVt + t)] = w(t, 4096.0L);
  // Hann sample n = 2(t + 64 s) + {0, 1}: w/2 = 1/4 + A cos(s D) + B sin(s D), A = -cos(phi_n0)/4, B = sin(phi_n0)/4,
  // phi_n = 2 pi n/(N-1), D = 2 pi 128/(N-1)  (dsp_core.py:87 with the real split's 1/2 folded in)
  const long double step = 2.0L * pi / static_cast<long double>(kN - 1);
  for (int t = 0; t < 64; ++t) {
    const long double p0 = step * (2 * t), p1 = step * (2 * t + 1);
    h[static_cast<size_t>(kTabHann + 2 * t)] = make_float2(static_cast<float>(-0.25L * cosl(p0)), static_cast<float>(-0.25L * cosl(p1)));
    h[static_cast<size_t>(kTabHann + 2 * t + 1)] = make_float2(static_cast<float>(0.25L * sinl(p0)), static_cast<float>(0.25L * sinl(p1)));
  }
  for (int s = 0; s < 32; ++s) {
    rp.hann_cos[s] = static_cast<float>(cosl(step * 128.0L * s));
    rp.hann_sin[s] = static_cast<float>(sinl(step * 128.0L * s));
  }
  DSP_CUDA(cudaMalloc(&rp.d_tables, h.size() * sizeof(float2)));
  DSP_CUDA(cudaMemcpy(rp.d_tables, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice));
  rp.ok = 1;
  return DSPB200_OK;
}

void fft_r32_free(FftR32Plan& rp) {
  if (rp.d_tables) cudaFree(rp.d_tables);
  rp.d_tables = nullptr;
  rp.ok = 0;
}

int fft_r32_run(const FftR32Plan& rp, const float* x, int64_t xs, int64_t n_valid, int64_t offset, int64_t hop,
                int64_t n_frames, float* mag, int64_t mfs, int64_t mcs, int64_t channels, int hann, int db,
                cudaStream_t stream) {
  DSP_CHECK(rp.ok, "internal: no 32-points-per-thread tables for this plan");
  R32Args a{};
  a.x = x; a.x_stride = xs; a.n_valid = n_valid; a.offset = offset; a.hop = hop; a.n_frames = n_frames;
  a.mag = mag; a.mfs = mfs; a.mcs = mcs;
  a.n_items = channels * n_frames;
  a.tables = static_cast<const float2*>(rp.d_tables);
  a.db = db; a.hann = hann;
  for (int s = 0; s < 32; ++s) {
    a.cc[s] = make_float2(rp.hann_cos[s], rp.hann_cos[s]);
    a.ss[s] = make_float2(rp.hann_sin[s], rp.hann_sin[s]);
  }
  int cfg = 81;   // G = 8 groups per CTA, 1 CTA per SM
  if (const char* ev = getenv("DSPB200_FFT_R32_CFG")) cfg = atoi(ev);
  switch (cfg) {
    case 24: return launch_r32<2, 4>(a, stream);
    case 42: return launch_r32<4, 2>(a, stream);
    case 61: return launch_r32<6, 1>(a, stream);
    case 32: return launch_r32<3, 2>(a, stream);
    case 33: return launch_r32<3, 3>(a, stream);
    case 52: return launch_r32<5, 2>(a, stream);
    case 810: return launch_r32<8, 1, false>(a, stream);
    case 811: return launch_r32<8, 1, true, false>(a, stream);
    case 71: return launch_r32<7, 1>(a, stream);
    case 91: return launch_r32<9, 1>(a, stream);
    case 101: return launch_r32<10, 1>(a, stream);
    default: return launch_r32<8, 1>(a, stream);
  }
}

}  // namespace dspb200
