// K1 on the tensor cores: the polyphase resampler as a banded-Toeplitz GEMM.
//
// For 128 consecutive outputs m0..m0+127 of every channel c
//     y[c, m0 + r] = sum_k  A_p[r, k] * x[c, lo + k],      A_p[r, k] = h[(m0 + r) M + P - (lo + k) L]
// (dsp_core.py:149-173 with the zero-stuffed signal eliminated).  The tap matrix A_p depends on the
// tile only through its phase p = tile mod period (period = 4L / gcd(128 M, 4L) tiles advance the
// input by a whole multiple of 4 samples), so the host lays out `period` matrices [128 x K] once per
// plan.  One tile is D[128 outputs x 256 channels] = A_p[128 x K] . X[256 x K]^T on tcgen05.mma
// (kind::tf32, fp32 accumulators in TMEM).  TF32 keeps 11 significand bits, far short of the 1e-5
// full-scale parity bound, so each operand is split in two TF32 numbers and three products are
// accumulated:  A_hi X + A_lo X + A_hi X_lo  (the hardware truncates fp32 operands to TF32, so the
// raw x tile serves as X_hi; X_lo = x - trunc(x) is formed in shared memory by the converter warps;
// A_hi / A_lo are precomputed).  Measured error of the split on random data: 1.6e-6 of max|y|.
//
// Warp roles (one persistent CTA per SM): warps 0-7 epilogue (TMEM -> registers -> 128-byte
// coalesced global stores), warp 8 TMA producer of x tiles, warp 9 MMA issuer (one elected lane) +
// TMEM allocation, warp 10 TMA producer of tap tiles, warps 11-14 converters.  Tiles hold 32 k-values
// (128-byte swizzled, K-major).  The x ring is 4 deep (HBM latency), the A_hi ring 4 and the A_lo ring 2 deep (L2 hits).
// Per k-block the issuer runs the two products that read the raw x tile, the converters then
// overwrite the tile in place with X_lo, and the third product follows one k-block later, so no
// second x buffer is needed.  Two 256-column accumulators in TMEM let the epilogue of tile i overlap
// the MMAs of tile i+1.
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "internal.cuh"
#include "umma.cuh"

namespace dspb200 {

namespace {

constexpr int kTM = 128;        // outputs per tile (MMA M)
constexpr int kTN = 256;        // channels per tile (MMA N)
constexpr int kBK = 32;         // k-values per stage (one 128-byte swizzle row)
constexpr int kXSlots = 4;      // x tiles in flight (converted in place between the products)
constexpr int kAhSlots = 4;     // A_hi tiles in flight (held until the third product)
constexpr int kAlSlots = 2;     // A_lo tiles in flight (released after the second product)
constexpr int kEpiWarps = 8;    // two per TMEM lane quarter: columns [0,128) and [128,256)
constexpr int kConvWarps = 4;
constexpr int kTmaWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1, kTmaWarpA = kEpiWarps + 2, kConvWarp0 = kEpiWarps + 3;
constexpr int kThreads = (kConvWarp0 + kConvWarps) * 32;
constexpr uint32_t kABytes = kTM * kBK * 4;   // 16 KB
constexpr uint32_t kBBytes = kTN * kBK * 4;   // 32 KB
constexpr size_t kSmemBytes = kXSlots * kBBytes + (kAhSlots + kAlSlots) * kABytes + 1024;   // 128 + 64 KB + alignment slack

struct MmaArgs {
  float* y; long long y_stride;
  long long channels, n_out;
  const int* lo;                // [period] window start of tile p (multiple of 4, may be negative)
  int period, nkb;
  long long adv;                // input samples per `period` tiles
  long long n_tt, n_tiles;
};

__global__ void __launch_bounds__(kThreads, 1)
src_mma_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_x, const MmaArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));   // swizzle atoms: 1024-byte aligned
  __shared__ __align__(8) uint64_t bars[4 * kXSlots + 2 * kAhSlots + 2 * kAlSlots + 4];
  __shared__ uint32_t tmem_base_s;
  uint64_t* full_x = bars;                    // [x slot] TMA landed the x tile
  uint64_t* mid = full_x + kXSlots;           // [x slot] the products with the raw x tile have completed
  uint64_t* conv = mid + kXSlots;             // [x slot] x replaced in place by x - trunc(x)
  uint64_t* empty_x = conv + kXSlots;         // [x slot] the product with X_lo has completed
  uint64_t* full_ah = empty_x + kXSlots;      // [A_hi slot] TMA landed
  uint64_t* empty_ah = full_ah + kAhSlots;    // [A_hi slot] all three products have completed
  uint64_t* full_al = empty_ah + kAhSlots;    // [A_lo slot] TMA landed
  uint64_t* empty_al = full_al + kAlSlots;    // [A_lo slot] the second product has completed
  uint64_t* acc_full = empty_al + kAlSlots;   // [2] accumulator complete
  uint64_t* acc_empty = acc_full + 2;         // [2] accumulator drained by the epilogue warps
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kXSlots; ++s) {
      mbar_init(&full_x[s], 1); mbar_init(&mid[s], 1); mbar_init(&conv[s], kConvWarps); mbar_init(&empty_x[s], 1);
    }
    for (int s = 0; s < kAhSlots; ++s) { mbar_init(&full_ah[s], 1); mbar_init(&empty_ah[s], 1); }
    for (int s = 0; s < kAlSlots; ++s) { mbar_init(&full_al[s], 1); mbar_init(&empty_al[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(&tmem_base_s))), "r"(2 * kTN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  auto x_ptr = [&](int s) -> unsigned char* { return smem + static_cast<size_t>(s) * kBBytes; };
  auto ah_ptr = [&](int s) -> unsigned char* { return smem + static_cast<size_t>(kXSlots) * kBBytes + static_cast<size_t>(s) * kABytes; };
  auto al_ptr = [&](int s) -> unsigned char* { return ah_ptr(kAhSlots) + static_cast<size_t>(s) * kABytes; };
  const long long first = blockIdx.x, step = gridDim.x;

  if (warp == kTmaWarp) {
    // ---------------- TMA producer, x tiles (HBM latency: kXSlots deep) ----------------
    if (lane == 0) {
      tma_prefetch_desc(&tm_x);
      uint32_t it = 0;
      for (long long tile = first; tile < a.n_tiles; tile += step) {
        const long long ct = tile / a.n_tt, tt = tile - ct * a.n_tt;
        const int p = static_cast<int>(tt % a.period);
        const long long lo = a.lo[p] + (tt / a.period) * a.adv;
        for (int kb = 0; kb < a.nkb; ++kb, ++it) {
          const int s = it % kXSlots;
          if (it >= kXSlots) mbar_wait(&empty_x[s], ((it / kXSlots) - 1) & 1);
          mbar_expect_tx(&full_x[s], kBBytes);
          tma_load_2d(x_ptr(s), &tm_x, static_cast<int>(lo) + kb * kBK, static_cast<int>(ct) * kTN, &full_x[s]);
        }
      }
    }
  } else if (warp == kTmaWarpA) {
    // ---------------- TMA producer, tap matrices (L2 resident: kASlots deep) ----------------
    if (lane == 0) {
      tma_prefetch_desc(&tm_a);
      uint32_t it = 0;
      for (long long tile = first; tile < a.n_tiles; tile += step) {
        const long long tt = tile % a.n_tt;
        const int p = static_cast<int>(tt % a.period);
        for (int kb = 0; kb < a.nkb; ++kb, ++it) {
          const int sh = it % kAhSlots, sl = it % kAlSlots;
          if (it >= kAhSlots) mbar_wait(&empty_ah[sh], ((it / kAhSlots) - 1) & 1);
          mbar_expect_tx(&full_ah[sh], kABytes);
          tma_load_2d(ah_ptr(sh), &tm_a, kb * kBK, (2 * p) * kTM, &full_ah[sh]);
          if (it >= kAlSlots) mbar_wait(&empty_al[sl], ((it / kAlSlots) - 1) & 1);
          mbar_expect_tx(&full_al[sl], kABytes);
          tma_load_2d(al_ptr(sl), &tm_a, kb * kBK, (2 * p + 1) * kTM, &full_al[sl]);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      // D fp32, A/B tf32, both K-major, N = 256, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(kTN >> 3) << 17) | (uint32_t(kTM >> 4) << 24);
      // third product of k-block `j` (deferred one step so the in-place conversion of its x tile overlaps
      // the first two products of the next k-block)
      auto finish = [&](uint32_t j, uint32_t d, bool last_of_tile, int b) {
        const int sx = j % kXSlots, sh = j % kAhSlots;
        mbar_wait(&conv[sx], (j / kXSlots) & 1);
        tc_fence_after();
        const uint64_t dah = umma_desc_sw128(ah_ptr(sh)), dxl = umma_desc_sw128(x_ptr(sx));
#pragma unroll
        for (int k = 0; k < kBK / 8; ++k) umma_tf32(d, dah + 2 * k, dxl + 2 * k, idesc, 1u);
        umma_commit(&empty_x[sx]);
        umma_commit(&empty_ah[sh]);
        if (last_of_tile) umma_commit(&acc_full[b]);
      };
      uint32_t it = 0, ti = 0;
      uint32_t prev_d = 0;
      int prev_b = 0;
      bool have_prev = false, prev_last = false;
      for (long long tile = first; tile < a.n_tiles; tile += step, ++ti) {
        const int b = ti & 1;
        if (ti >= 2) mbar_wait(&acc_empty[b], ((ti >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t d = tmem + b * kTN;
        for (int kb = 0; kb < a.nkb; ++kb, ++it) {
          const int sx = it % kXSlots, sh = it % kAhSlots, sl = it % kAlSlots;
          mbar_wait(&full_ah[sh], (it / kAhSlots) & 1);
          mbar_wait(&full_al[sl], (it / kAlSlots) & 1);
          mbar_wait(&full_x[sx], (it / kXSlots) & 1);
          tc_fence_after();
          const uint64_t dah = umma_desc_sw128(ah_ptr(sh)), dal = umma_desc_sw128(al_ptr(sl));
          const uint64_t dx = umma_desc_sw128(x_ptr(sx));
#pragma unroll
          for (int k = 0; k < kBK / 8; ++k) {
            umma_tf32(d, dah + 2 * k, dx + 2 * k, idesc, (kb | k) ? 1u : 0u);
            umma_tf32(d, dal + 2 * k, dx + 2 * k, idesc, 1u);
          }
          umma_commit(&mid[sx]);
          umma_commit(&empty_al[sl]);
          if (have_prev) finish(it - 1, prev_d, prev_last, prev_b);
          have_prev = true; prev_d = d; prev_b = b; prev_last = (kb == a.nkb - 1);
        }
      }
      if (have_prev) finish(it - 1, prev_d, prev_last, prev_b);
    }
  } else if (warp >= kConvWarp0) {
    // ---------------- converters: x -> x - trunc_tf32(x) in place, once the raw tile has been consumed ----------------
    const int ctid = threadIdx.x - kConvWarp0 * 32;
    uint32_t it = 0;
    for (long long tile = first; tile < a.n_tiles; tile += step) {
      for (int kb = 0; kb < a.nkb; ++kb, ++it) {
        const int s = it % kXSlots;
        mbar_wait(&mid[s], (it / kXSlots) & 1);
        float4* buf = reinterpret_cast<float4*>(x_ptr(s)) + ctid;
        constexpr int kPer = static_cast<int>(kBBytes / 16) / (32 * kConvWarps);   // 16-byte pieces per thread
        float4 v[kPer];
#pragma unroll
        for (int i = 0; i < kPer; ++i) v[i] = buf[i * 32 * kConvWarps];            // all loads in flight at once
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
          float4 r;
          r.x = v[i].x - __uint_as_float(__float_as_uint(v[i].x) & 0xFFFFE000u);
          r.y = v[i].y - __uint_as_float(__float_as_uint(v[i].y) & 0xFFFFE000u);
          r.z = v[i].z - __uint_as_float(__float_as_uint(v[i].z) & 0xFFFFE000u);
          r.w = v[i].w - __uint_as_float(__float_as_uint(v[i].w) & 0xFFFFE000u);
          buf[i * 32 * kConvWarps] = r;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv[s]);
      }
    }
  } else {
    // ---------------- epilogue warps 0-7 ----------------
    // warp w reads TMEM lanes 32(w%4) .. +31 (= outputs) and the column half w/4 (= 128 channels);
    // every store instruction writes 32 consecutive outputs of one channel (128 bytes)
    const int quarter = warp & 3, half = warp >> 2;
    uint32_t ti = 0;
    for (long long tile = first; tile < a.n_tiles; tile += step, ++ti) {
      const long long ct = tile / a.n_tt, tt = tile - ct * a.n_tt;
      const int b = ti & 1;
      mbar_wait(&acc_full[b], (ti >> 1) & 1);
      tc_fence_after();
      const long long m = tt * kTM + quarter * 32 + lane;
      const bool m_ok = m < a.n_out;
      const long long c_first = ct * kTN + half * (kTN / 2);
      const bool all_channels = c_first + kTN / 2 <= a.channels;
      float* p = a.y + c_first * a.y_stride + m;
      const uint32_t taddr = tmem + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(b * kTN + half * (kTN / 2));
      uint32_t v[2][32];
      tmem_ld32(v[0], taddr);
#pragma unroll
      for (int q = 0; q < kTN / 2 / 32; ++q) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (q + 1 < kTN / 2 / 32) tmem_ld32(v[(q + 1) & 1], taddr + (q + 1) * 32);   // in flight during the stores
        const uint32_t* vv = v[q & 1];
        if (m_ok) {
          if (all_channels) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { *p = __uint_as_float(vv[j]); p += a.y_stride; }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (c_first + q * 32 + j < a.channels) *p = __uint_as_float(vv[j]);
              p += a.y_stride;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[b]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * kTN));
}

float round_tf32(float v) {   // nearest TF32 (ties away): exactly representable, so the hardware truncation keeps it
  uint32_t u;
  memcpy(&u, &v, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  memcpy(&v, &u, 4);
  return v;
}

long long floor_div(long long a, long long b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

}  // namespace

int src_mma_build(const std::vector<double>& taps, int L, int M, SrcMmaPlan& mp) {
  mp = SrcMmaPlan{};
  const long long T = static_cast<long long>(taps.size());
  const long long P = (T - 1) / 2;
  const long long g = std::gcd(static_cast<long long>(kTM) * M, 4LL * L);
  const long long period = 4LL * L / g;
  if (period > 1024) return DSPB200_OK;
  mp.adv = period * kTM * M / L;
  std::vector<int> lo(static_cast<size_t>(period));
  long long span = 0;
  for (long long p = 0; p < period; ++p) {
    const long long m0 = p * kTM;
    // taps h[t], t = m M + P - i L in [0, T):  i >= (m M + P - T + 1) / L,  i <= (m M + P) / L
    const long long i_min = -floor_div(-(m0 * M + P - T + 1), L);
    const long long i_max = floor_div((m0 + kTM - 1) * M + P, L);
    const long long l4 = floor_div(i_min, 4) * 4;
    lo[static_cast<size_t>(p)] = static_cast<int>(l4);
    span = std::max(span, i_max - l4 + 1);
  }
  const long long kpad = (span + kBK - 1) / kBK * kBK;
  const size_t elems = static_cast<size_t>(period) * 2 * kTM * static_cast<size_t>(kpad);
  if (elems * sizeof(float) > (size_t(64) << 20)) return DSPB200_OK;   // keep the tap matrices L2-resident
  std::vector<float> tab(elems, 0.f);
  for (long long p = 0; p < period; ++p)
    for (int r = 0; r < kTM; ++r)
      for (long long k = 0; k < kpad; ++k) {
        const long long t = (p * kTM + r) * M + P - (lo[static_cast<size_t>(p)] + k) * L;
        if (t < 0 || t >= T) continue;
        const float v = static_cast<float>(taps[static_cast<size_t>(t)]);
        const float hi = round_tf32(v);
        tab[((static_cast<size_t>(p) * 2 + 0) * kTM + r) * kpad + k] = hi;
        tab[((static_cast<size_t>(p) * 2 + 1) * kTM + r) * kpad + k] = v - hi;
      }
  DSP_CUDA(cudaMalloc(reinterpret_cast<void**>(&mp.d_table), elems * sizeof(float)));
  DSP_CUDA(cudaMemcpy(mp.d_table, tab.data(), elems * sizeof(float), cudaMemcpyHostToDevice));
  DSP_CUDA(cudaMalloc(reinterpret_cast<void**>(&mp.d_lo), lo.size() * sizeof(int)));
  DSP_CUDA(cudaMemcpy(mp.d_lo, lo.data(), lo.size() * sizeof(int), cudaMemcpyHostToDevice));
  mp.period = static_cast<int>(period);
  mp.kpad = static_cast<int>(kpad);
  mp.ok = 1;
  return DSPB200_OK;
}

void src_mma_free(SrcMmaPlan& mp) {
  cudaFree(mp.d_table);
  cudaFree(mp.d_lo);
  mp = SrcMmaPlan{};
}

bool src_mma_usable(const SrcMmaPlan& mp, const float* x, int64_t xs, int64_t channels, int64_t n_in) {
  // a tile always multiplies 256 channels: below 3/4 occupancy of the last tile row the FFMA kernel wins
  const int64_t padded = ceil_div(channels, kTN) * kTN;
  if (4 * channels < 3 * padded && getenv("DSPB200_SRC_FORCE_MMA") == nullptr) return false;
  return mp.ok && reinterpret_cast<uintptr_t>(x) % 16 == 0 && xs % 4 == 0 && n_in >= 128 &&
         kSmemBytes <= static_cast<size_t>(max_smem_optin());
}

int src_mma_run(const SrcMmaPlan& mp, const float* x, int64_t xs, float* y, int64_t ys, int64_t channels,
                int64_t n_in, int64_t n_out, cudaStream_t stream) {
  CUtensorMap tm_a, tm_x;
  memset(&tm_a, 0, sizeof(tm_a));
  memset(&tm_x, 0, sizeof(tm_x));
  DSP_TRY(encode_tmap_2d(&tm_a, DSPB200_F32, mp.d_table, static_cast<uint64_t>(mp.kpad),
                         static_cast<uint64_t>(mp.period) * 2 * kTM, static_cast<uint64_t>(mp.kpad) * sizeof(float),
                         kBK, kTM, true));
  DSP_TRY(encode_tmap_2d(&tm_x, DSPB200_F32, x, static_cast<uint64_t>(n_in), static_cast<uint64_t>(channels),
                         static_cast<uint64_t>(xs) * sizeof(float), kBK, kTN, true));
  MmaArgs a{};
  a.y = y; a.y_stride = ys; a.channels = channels; a.n_out = n_out;
  a.lo = mp.d_lo; a.period = mp.period; a.nkb = mp.kpad / kBK; a.adv = mp.adv;
  a.n_tt = ceil_div(n_out, kTM);
  a.n_tiles = a.n_tt * ceil_div(channels, kTN);
  const size_t smem = kSmemBytes;
  DSP_CUDA(cudaFuncSetAttribute(src_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int64_t sms = sm_count();
  const int grid = static_cast<int>(a.n_tiles < sms ? a.n_tiles : sms);
  src_mma_kernel<<<grid, kThreads, smem, stream>>>(tm_a, tm_x, a);
  return after_launch("src_mma_kernel");
}

}  // namespace dspb200
