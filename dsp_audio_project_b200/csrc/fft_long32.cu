// K3, 2^16-point Hann magnitude frames (config C4: 2^20-sample clips x 4096 channels) -- three radix-32 passes.
//
// Same DFT as the four-step kernels of fft.cu (dsp_core.py:41-98), regrouped as 32768 = 32 x 32 x 32 complex points so
// that a transform crosses shared memory twice (one exchange inside the 1024-point rows, one for the real split) and
// its workspace once, instead of four exchanges (256 = 16 x 16 columns, 128 = 8 x 16 rows).  The four-step form is
// bound by those exchanges, not by HBM (profiles/r2c_ncu_full_fft4_fused.md).
//
//   n = 1024 n1 + n2,  k = k1 + 32 k2        (n1, k1 < 32; n2, k2 < 1024)
//
//   step 1   thread n2 loads z[1024 n1 + n2], n1 = 0..31 (a warp reads 256 contiguous bytes per n1), applies the
//            window, transforms over n1 in registers, multiplies by W_32768^(n2 k1) and writes Y[k1][n2] into the
//            CTA's own workspace slot (256 KB, rewritten for every transform: it stays in L2)
//   step 2   a warp owns a row k1: 1024 points = 32 x 32, two register passes with one exchange through the warp's
//            own shared-memory buffer (no CTA barrier), result left there in place
//   split    rows k1 and 32 - k1 are transformed in the same round (16 rows per round, two rounds), so
//            Z[k1 + 32 k2] and Z[32768 - k1 - 32 k2] = row 32 - k1, element 1023 - k2, are both at hand: a thread
//            forms |X[k]| and |X[32768 - k]|; lanes run over eight adjacent rows, so the stores are 32-byte runs.
//
// One persistent CTA of 512 threads per SM.  fp32 only; the window is applied by angle addition (pre-scaled by 1/2
// for the real split) as in fft_r32.cu.
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

constexpr int kN = 65536;           // real samples per frame
constexpr int kNc = 32768;          // complex points
constexpr int kRow = 1024;          // points per row (n2 / k2)
constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kPairs = kWarps / 2;
// row buffer: 32 x 33 entries + entry 1056 = a copy of element 0 (so that "element 1024" of a row reads element 0);
// pitch 1 mod 16: the real split reads 8 row pairs x 4 consecutive elements per warp instruction without conflicts
constexpr int kRowPitch = 1057;
constexpr int kWsRows = 16;         // rows of a transform that go through the workspace (the other 16 stay in shared memory)
// table blob (float2 entries)
constexpr int kTabLo = kTw1;                 // W_65536^k1, k1 < 32
constexpr int kTabHi = kTabLo + 32;          // W_65536^(32 k2) = W_2048^k2, k2 < 1024
constexpr int kTabSmem = kTabHi + kRow;      // everything up to here is copied to shared memory
constexpr int kTabCol = kTabSmem;            // [1024][10]: powers of W_32768^n2 (fft_r32.cuh layout)
constexpr int kTabHann = kTabCol + kRow * 10;   // [1024][2]: (A(2 n2), A(2 n2 + 1)), (B(2 n2), B(2 n2 + 1))
constexpr int kTabTotal = kTabHann + kRow * 2;

struct L32Args {
  const float* x;
  long long x_stride, n_valid, offset, hop, n_frames;
  float* mag;
  long long mfs, mcs;
  long long n_items;
  const float2* tables;
  float2* ws;              // [gridDim.x][16][1024]: the round-1 rows of the transform in flight
  int db, hann;
  float2 cc[32], ss[32];   // cos/sin(n1 * 2 pi 2048/(N-1)), both halves alike
};

// L2 eviction policies: the workspace slot is rewritten for every transform and must stay in L2 (evict_last); the input
// and the spectra stream through once (evict_first), so they do not push the slots out
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float2 ld_ws(const float2* p, uint64_t pol) {   // L2 only: other rounds rewrite the slot
  float2 v;
  asm volatile("ld.global.cg.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol) : "memory");
  return v;
}
__device__ __forceinline__ void st_ws(float2* p, float2 v, uint64_t pol) {
  asm volatile("st.global.cg.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ float2 ld_stream(const float2* p, uint64_t pol) {
  float2 v;
  asm volatile("ld.global.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_stream(float* p, float v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}

__device__ __forceinline__ int row_addr(int idx) { return (idx >> 5) * 33 + (idx & 31); }

// slot of row k1 inside its round (round 0: rows 0, 16, 1, 31, 2, 30, .., 7, 25; round 1: rows 8, 24, 9, 23, .., 15, 17)
__host__ __device__ constexpr int row_round(int k1) { return (k1 <= 7 || k1 == 16 || k1 >= 25) ? 0 : 1; }
__host__ __device__ constexpr int row_slot(int k1) {
  return k1 == 0 ? 0 : (k1 == 16 ? 1 : (k1 <= 7 ? 2 * k1 : (k1 >= 25 ? 2 * (32 - k1) + 1 : (k1 <= 15 ? 2 * (k1 - 8) : 2 * (24 - k1) + 1))));
}

__global__ void __launch_bounds__(kThreads, 1) fft65536_l32_kernel(const L32Args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tw1 = reinterpret_cast<float2*>(smem_raw);
  float2* tlo = tw1 + kTabLo;
  float2* thi = tw1 + kTabHi;
  float2* rows = tw1 + kTabSmem;
  const int T = threadIdx.x, warp = T >> 5, lane = T & 31;
  for (int i = T; i < kTabSmem; i += kThreads) tw1[i] = a.tables[i];
  __syncthreads();
  float2* wsb = a.ws + static_cast<size_t>(blockIdx.x) * (kWsRows * kRow);
  float2* rb = rows + warp * kRowPitch;
  const float2* twr = tw1 + lane * kTw1Pitch;
  const uint64_t pol_ws = policy_evict_last(), pol_io = policy_evict_first();

  for (long long item = blockIdx.x; item < a.n_items; item += gridDim.x) {
    const long long c = item / a.n_frames, fr = item - c * a.n_frames;
    const float* xrow = a.x + c * a.x_stride;
    const long long fstart = a.offset + fr * a.hop;
    const bool fast = (fstart + kN <= a.n_valid) && ((reinterpret_cast<uintptr_t>(xrow + fstart) & 7) == 0);
    // ---- step 1: columns.  Round-0 rows go straight into their row buffers, round-1 rows into the workspace ----
#pragma unroll 1
    for (int r = 0; r < kRow / kThreads; ++r) {
      const int n2 = r * kThreads + T;
      float2 v[32];
      if (fast) {
        const float2* xp = reinterpret_cast<const float2*>(xrow + fstart) + n2;
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) v[n1] = ld_stream(xp + n1 * kRow, pol_io);
      } else {
        const long long left = a.n_valid - fstart;
        const int rem = left > kN ? kN : (left < 0 ? 0 : static_cast<int>(left));
        const float* xf = xrow + fstart;
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
          const int e = 2 * (n2 + kRow * n1);
          v[n1].x = e < rem ? xf[e] : 0.f;
          v[n1].y = e + 1 < rem ? xf[e + 1] : 0.f;
        }
      }
      if (a.hann) {
        dft32_windowed(v, a.tables[kTabHann + 2 * n2], a.tables[kTabHann + 2 * n2 + 1], a.cc, a.ss);
      } else {
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) v[n1] = pscale(v[n1], 0.5f);
        Dft32<32>::run(v);
      }
      twiddle_powers(v, a.tables + kTabCol + n2 * 10);   // W_32768^(n2 k1)
      float2* wp = wsb + n2;
      float2* sp = rows + row_addr(n2);
#pragma unroll
      for (int k1 = 0; k1 < 32; ++k1) {
        if (row_round(k1) == 0) sp[row_slot(k1) * kRowPitch] = v[k1];
        else st_ws(wp + row_slot(k1) * kRow, v[k1], pol_ws);
      }
    }
    __syncthreads();
    if (T < 16 && item + gridDim.x < a.n_items) {
      // the next transform's input into L2 while this one's rows are transformed
      const long long itn = item + gridDim.x;
      const long long cn = itn / a.n_frames, frn = itn - cn * a.n_frames;
      const long long fs2 = a.offset + frn * a.hop;
      if (fs2 + kN <= a.n_valid) {
        // sixteen pieces of whole 16-byte units INSIDE the frame only (the last piece of a misaligned frame is shorter)
        const uintptr_t f0 = reinterpret_cast<uintptr_t>(a.x + cn * a.x_stride + fs2);
        const uintptr_t p0 = ((f0 + 15) & ~static_cast<uintptr_t>(15)) + static_cast<uintptr_t>(T) * (kN * 4 / 16);
        const uintptr_t end = (f0 + kN * 4) & ~static_cast<uintptr_t>(15);
        const uintptr_t len = p0 + kN * 4 / 16 <= end ? kN * 4 / 16 : (end > p0 ? end - p0 : 0);
        if (len) prefetch_l2_bulk(reinterpret_cast<const void*>(p0), static_cast<uint32_t>(len));
      }
    }
    // ---- step 2: rows, 16 per round (rows k1 and 32 - k1 in the same round; warp w owns slot w) ----
    float* mg = a.mag + c * a.mcs + fr * a.mfs;
#pragma unroll 1
    for (int r = 0; r < 2; ++r) {
      {
        float2 v[32];
        float2* cp = rb + lane;
        if (r == 0) {
#pragma unroll
          for (int s = 0; s < 32; ++s) v[s] = cp[s * 33];        // Y[k1][lane + 32 s], left here by step 1
          __syncwarp();
        } else {
          const float2* rp = wsb + warp * kRow + lane;
#pragma unroll
          for (int s = 0; s < 32; ++s) v[s] = ld_ws(rp + 32 * s, pol_ws);
        }
        Dft32<32>::run(v);
        float2* wp = rb + lane * 33;
#pragma unroll
        for (int ka = 0; ka < 32; ++ka) wp[ka] = v[ka];
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = cp[j * 33];
        dft32_twiddled(v, twr);   // W_1024^(j ka) folded into the first butterfly level
#pragma unroll
        for (int kb = 0; kb < 32; ++kb) cp[kb * 33] = v[kb];   // in place: Z_row[ka + 32 kb]
        if (lane == 0) rb[32 * 33] = v[0];                      // "element 1024" = element 0
      }
      __syncthreads();
      {
        // thread (p, kk): row pair p of the round, elements k2 = kk + 64 i.  General pair (ra, 32 - ra):
        // A = Z_ra[k2], B = Z_(32-ra)[1023 - k2], k = ra + 32 k2.  Pair 0 of round 0 is rows 0 and 16, each its own
        // mirror: i < 8 -> row 0: A = Z_0[k2], B = Z_0[1024 - k2], k = 32 k2;  i >= 8 -> row 16 with q = k2 - 512:
        // A = Z_16[q], B = Z_16[1023 - q], k = 16 + 32 q.  Everything is base + compile-time offset per half.
        const int p = T & (kPairs - 1), kk = T / kPairs;   // kk < 64
        const bool special = (r == 0 && p == 0);
        const int ra = r * kPairs + p;
        const float2* sa = rows + (2 * p) * kRowPitch;
        const float2* sb = sa + kRowPitch;
        const float2* pA[2];
        const float2* pB[2];
        int kb0[2], hq[2];
        float2 wl[2];
        pA[0] = sa + row_addr(kk);
        pB[0] = special ? sa + row_addr(kRow - kk) : sb + row_addr(kRow - 1 - kk);
        kb0[0] = special ? 32 * kk : ra + 32 * kk;
        hq[0] = kk;
        wl[0] = tlo[special ? 0 : ra];
        pA[1] = special ? sb + row_addr(kk) : pA[0] + 66 * 8;
        pB[1] = special ? sb + row_addr(kRow - 1 - kk) : pB[0] - 66 * 8;
        kb0[1] = special ? 16 + 32 * kk : kb0[0] + 32 * 512;
        hq[1] = special ? kk : kk + 512;
        wl[1] = tlo[special ? 16 : ra];
        auto split_one = [&](float* mk, float* mm, float2 A, float2 B, float2 w, auto db_tag) {
          constexpr bool kDb = decltype(db_tag)::value;
          const float2 Bc = cconj(B);
          const float2 S = cadd(A, Bc), D = csub(A, Bc);
          const float2 X1 = cmadd(make_float2(w.y, -w.x), D, S);   // S - i W_65536^k (A - conj B)
          st_stream(mk, mag_of<kDb>(X1), pol_io);
          st_stream(mm, mag_of<kDb>(twice_minus(S, X1)), pol_io);
        };
        auto split = [&](auto db_tag) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float* mk = mg + kb0[h];
            float* mm = mg + (kNc - kb0[h]);
            const float2* th = thi + hq[h];
#pragma unroll
            for (int i = 0; i < 8; ++i)
              split_one(mk + 2048 * i, mm - 2048 * i, pA[h][66 * i], pB[h][-66 * i], cmul(wl[h], th[64 * i]), db_tag);
          }
          if (r == 0 && T == 0) {   // k = 16384: row 0, element 512 pairs with itself;  W_65536^16384 = -i
            const float2 A = rows[row_addr(512)];
            split_one(mg + kNc / 2, mg + kNc / 2, A, A, make_float2(0.f, -1.f), db_tag);
          }
        };
        if (a.db) split(std::true_type{}); else split(std::false_type{});
      }
      __syncthreads();
    }
  }
}

}  // namespace

int fft_long32_build(int n_fft, FftLong32Plan& lp) {
  lp.ok = 0;
  if (n_fft != kN) return DSPB200_OK;
  const long double pi = 3.14159265358979323846264338327950288L;
  std::vector<float2> h(static_cast<size_t>(kTabTotal), make_float2(1.f, 0.f));
  auto w = [&](long double num, long double den) {
    const long double ang = -2.0L * pi * num / den;
    return make_float2(static_cast<float>(cosl(ang)), static_cast<float>(sinl(ang)));
  };
  for (int k = 0; k < 32; ++k) fill_twiddle_row(&h[static_cast<size_t>(k * kTw1Pitch)], k, 1024.0L);
  for (int k1 = 0; k1 < 32; ++k1) h[static_cast<size_t>(kTabLo + k1)] = w(k1, 65536.0L);
  for (int k2 = 0; k2 < kRow; ++k2) h[static_cast<size_t>(kTabHi + k2)] = w(k2, 2048.0L);
  for (int n2 = 0; n2 < kRow; ++n2) fill_twiddle_row(&h[static_cast<size_t>(kTabCol + n2 * 10)], n2, 32768.0L);
  // Hann sample m = 2048 n1 + (2 n2 + c): w/2 = 1/4 + A cos(n1 D) + B sin(n1 D), A = -cos(phi_(2 n2 + c))/4,
  // B = sin(phi_(2 n2 + c))/4, phi_m = 2 pi m/(N-1), D = 2 pi 2048/(N-1)   (dsp_core.py:87, real split's 1/2 folded in)
  const long double step = 2.0L * pi / static_cast<long double>(kN - 1);
  for (int n2 = 0; n2 < kRow; ++n2) {
    const long double p0 = step * (2 * n2), p1 = step * (2 * n2 + 1);
    h[static_cast<size_t>(kTabHann + 2 * n2)] = make_float2(static_cast<float>(-0.25L * cosl(p0)), static_cast<float>(-0.25L * cosl(p1)));
    h[static_cast<size_t>(kTabHann + 2 * n2 + 1)] = make_float2(static_cast<float>(0.25L * sinl(p0)), static_cast<float>(0.25L * sinl(p1)));
  }
  for (int n1 = 0; n1 < 32; ++n1) {
    lp.hann_cos[n1] = static_cast<float>(cosl(step * 2048.0L * n1));
    lp.hann_sin[n1] = static_cast<float>(sinl(step * 2048.0L * n1));
  }
  DSP_CUDA(cudaMalloc(&lp.d_tables, h.size() * sizeof(float2)));
  DSP_CUDA(cudaMemcpy(lp.d_tables, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice));
  lp.ok = 1;
  return DSPB200_OK;
}

void fft_long32_free(FftLong32Plan& lp) {
  if (lp.d_tables) cudaFree(lp.d_tables);
  lp.d_tables = nullptr;
  lp.ok = 0;
}

size_t fft_long32_workspace(int64_t n_transforms) {
  const int64_t ctas = n_transforms < sm_count() ? n_transforms : sm_count();   // one CTA per SM, one slot each
  return static_cast<size_t>(ctas) * kWsRows * kRow * sizeof(float2);
}

int fft_long32_run(const FftLong32Plan& lp, const float* x, int64_t xs, int64_t n_valid, int64_t offset, int64_t hop,
                   int64_t n_frames, float* mag, int64_t mfs, int64_t mcs, int64_t channels, int hann, int db, void* ws,
                   size_t ws_bytes, cudaStream_t stream) {
  DSP_CHECK(lp.ok, "internal: no three-pass tables for this plan");
  L32Args a{};
  a.x = x; a.x_stride = xs; a.n_valid = n_valid; a.offset = offset; a.hop = hop; a.n_frames = n_frames;
  a.mag = mag; a.mfs = mfs; a.mcs = mcs;
  a.n_items = channels * n_frames;
  a.tables = static_cast<const float2*>(lp.d_tables);
  a.ws = static_cast<float2*>(ws);
  a.db = db; a.hann = hann;
  for (int s = 0; s < 32; ++s) {
    a.cc[s] = make_float2(lp.hann_cos[s], lp.hann_cos[s]);
    a.ss[s] = make_float2(lp.hann_sin[s], lp.hann_sin[s]);
  }
  const size_t smem = static_cast<size_t>(kTabSmem + kWarps * kRowPitch) * sizeof(float2);
  auto kern = fft65536_l32_kernel;
  DSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int grid = static_cast<int>(a.n_items < sm_count() ? a.n_items : sm_count());
  const size_t need = fft_long32_workspace(a.n_items);
  DSP_CHECK(ws != nullptr && ws_bytes >= need, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
  if (getenv("DSPB200_FFT_TRACE")) fprintf(stderr, "fft65536_l32 smem=%zu grid=%d\n", smem, grid);
  kern<<<grid, kThreads, smem, stream>>>(a);
  return after_launch("fft65536_l32_kernel");
}

}  // namespace dspb200
