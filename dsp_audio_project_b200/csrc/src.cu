// K1 -- L/M polyphase sample-rate converter.
//
// Replaces conversion_tasa_muestreo (dsp_core.py:133-173): zero-stuff by L,
// Blackman-sinc FIR of 40*max(L,M)+1 taps, 'same' convolution, keep every M-th.
// Closed form, never materialising the stuffed signal:
//     y[c, m] = sum_j h[p + j*L] * x[c, i0 - j],   q = m*M + P, i0 = q / L, p = q % L
// with P = (min(N*L, T) - 1) / 2 the 'same'-mode centre offset.
//
// Two kernels:
//  * src_tiled_kernel -- the fast path.  Outputs are grouped 8 at a time
//    ("groups"); the taps each group needs, for every input position of the
//    group's window, are laid out on the host as rows of 8 ([group][pos][8],
//    zero where a tap does not exist) so the inner loop is a dense rank-1
//    update  acc[8 outputs][RC channels] += taps[pos][8] (x) x[channels][pos].
//    Lanes index channels, so tap loads are warp-wide broadcasts from shared
//    memory and every lane reuses each tap for RC channels and each input for
//    8 outputs.  fp32 uses the packed FFMA2 (fma.rn.f32x2) with the input sample
//    as the broadcast operand.  Input windows [CH channels][PITCH samples] are
//    staged by TMA (cp.async.bulk.tensor.2d, zero fill outside the signal = the
//    'same' zero padding) into a double-buffered ring, one elected thread
//    issuing, an mbarrier per stage; PITCH*sizeof(T) == 16 (mod 128) makes the
//    16-byte per-lane reads bank-conflict free.  Outputs leave as 16-byte
//    vectors (8 consecutive outputs per lane per channel).
//  * src_generic_kernel -- one thread per output, any (L, M, N); used for short
//    inputs (N*L < T, where numpy swaps operands), for ratios whose tap table or
//    window does not fit shared memory, and as the in-library cross-check.
//
// Roofline (C2: 1024 x 441000 -> 480000, L=160, M=147): algorithmic bytes
// sizeof(T)*(n_in + n_out) per channel; 40 FMA per output -> 10.4 flop/B fp32,
// i.e. HBM and the FP32 pipe are co-critical.
#include <algorithm>
#include <cstdlib>
#include <new>
#include <vector>

#include "design.cuh"
#include "internal.cuh"

namespace dspb200 {

constexpr int kRM = 8;            // outputs per group
constexpr int kSrcMaxWarps = 16;  // groups per tile = warps per CTA

template <typename T> struct SrcCfg;
template <> struct SrcCfg<float> { static constexpr int RC = 4; };
template <> struct SrcCfg<double> { static constexpr int RC = 2; };

struct SrcTiledGeom {
  int ok = 0;
  int GP = 0, PO = 0, PI = 0;  // groups / outputs / inputs per period
  int W = 0, WROWS = 0;
  int GT = 0;                  // groups per tile
  int PITCH = 0;
  int CH = 0;
  size_t smem_bytes = 0;
  size_t stage_bytes = 0;
  size_t table_bytes = 0;
};

template <typename T> struct SrcTiledArgs {
  const T* x;      // only used by the non-TMA loader
  long long x_stride;
  T* y;
  long long y_stride;
  long long channels, n_in, n_out;
  const T* table;        // [GP][WROWS][8]
  const int* group_lo;   // [GP]
  const int* tile_lo;    // [GP]
  int GP, PI, W, WROWS, GT, PITCH, CH;
  int n_tt;              // time tiles per channel tile
  long long n_tiles;
  int y_vec_ok;          // y rows are 16-byte aligned
  int tma_store;         // outputs leave through TMA (swizzled staging in the consumed stage)
};

template <typename T, int RC> struct Acc;
template <int RC> struct Acc<float, RC> {
  float2 a[kRM / 2][RC];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int r = 0; r < kRM / 2; ++r)
#pragma unroll
      for (int c = 0; c < RC; ++c) a[r][c] = make_float2(0.f, 0.f);
  }
  // taps t[0..7] for one input position, x for channel slot c
  __device__ __forceinline__ void step(const float4 t0, const float4 t1, const float xv, const int c) {
    ffma2_bcast(a[0][c], make_float2(t0.x, t0.y), xv);
    ffma2_bcast(a[1][c], make_float2(t0.z, t0.w), xv);
    ffma2_bcast(a[2][c], make_float2(t1.x, t1.y), xv);
    ffma2_bcast(a[3][c], make_float2(t1.z, t1.w), xv);
  }
  __device__ __forceinline__ float get(int r, int c) const { return (r & 1) ? a[r >> 1][c].y : a[r >> 1][c].x; }
};
template <int RC> struct Acc<double, RC> {
  double a[kRM][RC];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int r = 0; r < kRM; ++r)
#pragma unroll
      for (int c = 0; c < RC; ++c) a[r][c] = 0.0;
  }
  __device__ __forceinline__ double get(int r, int c) const { return a[r][c]; }
};

// Rank-1 update over one group's window: acc[8][RC] += taps[pos][8] (x) x[RC][pos].
template <typename T, int RC>
__device__ __forceinline__ void src_group_mac(Acc<T, RC>& acc, const T* __restrict__ tab,
                                              const T* __restrict__ xrow, size_t slot_stride, int nquads) {
  acc.zero();
  if constexpr (sizeof(T) == 4) {
#pragma unroll 1
    for (int qd = 0; qd < nquads; ++qd) {
      float4 xv[RC];
#pragma unroll
      for (int c = 0; c < RC; ++c)
        xv[c] = *reinterpret_cast<const float4*>(xrow + c * slot_stride + 4 * qd);
      const float4* tp = reinterpret_cast<const float4*>(tab + static_cast<size_t>(qd) * 4 * kRM);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float4 t0 = tp[2 * e], t1 = tp[2 * e + 1];
#pragma unroll
        for (int c = 0; c < RC; ++c) {
          const float xe = e == 0 ? xv[c].x : (e == 1 ? xv[c].y : (e == 2 ? xv[c].z : xv[c].w));
          acc.step(t0, t1, xe, c);
        }
      }
    }
  } else {
#pragma unroll 1
    for (int qd = 0; qd < nquads; ++qd) {
      double xv[RC][4];
#pragma unroll
      for (int c = 0; c < RC; ++c) {
        const double2 lo2 = *reinterpret_cast<const double2*>(xrow + c * slot_stride + 4 * qd);
        const double2 hi2 = *reinterpret_cast<const double2*>(xrow + c * slot_stride + 4 * qd + 2);
        xv[c][0] = lo2.x; xv[c][1] = lo2.y; xv[c][2] = hi2.x; xv[c][3] = hi2.y;
      }
      const double2* tp = reinterpret_cast<const double2*>(tab + static_cast<size_t>(qd) * 4 * kRM);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        double t[kRM];
#pragma unroll
        for (int r = 0; r < kRM / 2; ++r) {
          const double2 tt2 = tp[4 * e + r];
          t[2 * r] = tt2.x;
          t[2 * r + 1] = tt2.y;
        }
#pragma unroll
        for (int c = 0; c < RC; ++c)
#pragma unroll
          for (int r = 0; r < kRM; ++r) acc.a[r][c] = fma(t[r], xv[c][e], acc.a[r][c]);
      }
    }
  }
}

// Direct register -> global stores: 8 consecutive outputs per channel slot.
template <typename T, int RC>
__device__ __forceinline__ void src_store_direct(const Acc<T, RC>& acc, const SrcTiledArgs<T>& a, int ct,
                                                 int lane, long long m0) {
#pragma unroll
  for (int c = 0; c < RC; ++c) {
    const long long ch = static_cast<long long>(ct) * a.CH + c * 32 + lane;
    if (ch < a.channels) {
      T* yp = a.y + ch * a.y_stride + m0;
      if (a.y_vec_ok && m0 + kRM <= a.n_out) {
        if constexpr (sizeof(T) == 4) {
          *reinterpret_cast<float4*>(yp) = make_float4(acc.get(0, c), acc.get(1, c), acc.get(2, c), acc.get(3, c));
          *reinterpret_cast<float4*>(yp + 4) = make_float4(acc.get(4, c), acc.get(5, c), acc.get(6, c), acc.get(7, c));
        } else {
#pragma unroll
          for (int r = 0; r < kRM; r += 2)
            *reinterpret_cast<double2*>(yp + r) = make_double2(acc.get(r, c), acc.get(r + 1, c));
        }
      } else {
#pragma unroll
        for (int r = 0; r < kRM; ++r)
          if (m0 + r < a.n_out) yp[r] = acc.get(r, c);
      }
    }
  }
}

// kTma = true : warp-specialised.  Warps 0..GT-1 compute one output group each
//   per tile; warp GT is the TMA producer: it streams the input windows into a
//   two-stage ring (mbarrier full[]) and, when a.tma_store is set, writes each
//   finished output tile back with TMA from the stage the tile was computed
//   from (128-byte swizzled boxes, so the per-lane 16-byte shared-memory writes
//   are conflict free and the global writes are full lines, clipped at the
//   signal end by the tensor map).
// kTma = false: plain cooperative loader for inputs TMA cannot describe.
template <typename T, int RC, bool kTma>
__global__ void __launch_bounds__((kSrcMaxWarps + (kTma ? 1 : 0)) * 32, 1)
src_tiled_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_y,
                 const SrcTiledArgs<T> a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const size_t stage_elems = static_cast<size_t>(a.CH) * a.PITCH;
  T* xs0 = reinterpret_cast<T*>(smem);
  T* xs1 = xs0 + stage_elems;
  T* s_table = xs1 + stage_elems;
  int* s_group_lo = reinterpret_cast<int*>(s_table + static_cast<size_t>(a.GP) * a.WROWS * kRM);
  int* s_tile_lo = s_group_lo + a.GP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(
      (reinterpret_cast<uintptr_t>(s_tile_lo + a.GP) + 7) & ~static_cast<uintptr_t>(7));
  uint64_t* full = bars;        // [2] input window landed (TMA complete_tx)
  uint64_t* done = bars + 2;    // [2] consumers finished with the stage (outputs staged / window consumed)

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int nthreads = blockDim.x;
  constexpr int VEC = 16 / static_cast<int>(sizeof(T));

  for (int i = tid; i < a.GP * a.WROWS * kRM; i += nthreads) s_table[i] = a.table[i];
  for (int i = tid; i < a.GP; i += nthreads) {
    s_group_lo[i] = a.group_lo[i];
    s_tile_lo[i] = a.tile_lo[i];
  }
  if (kTma && tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(&done[0], a.GT);
    mbar_init(&done[1], a.GT);
    fence_mbar_init();
  }
  __syncthreads();

  auto tile_coords = [&](long long tile, int& ct, int& tt, int& lo) {
    ct = static_cast<int>(tile / a.n_tt);
    tt = static_cast<int>(tile - static_cast<long long>(ct) * a.n_tt);
    const int g0 = tt * a.GT;
    const int k0 = g0 / a.GP;
    const int gi0 = g0 - k0 * a.GP;
    lo = k0 * a.PI + s_tile_lo[gi0];
    lo -= (lo & (VEC - 1));   // TMA boxes must start on a 16-byte boundary of the row
  };

  const long long first = blockIdx.x;
  const long long step = gridDim.x;

  if constexpr (kTma) {
    if (warp == a.GT) {
      // ---------------- producer warp ----------------
      if (lane == 0) {
        const uint32_t stage_bytes = static_cast<uint32_t>(stage_elems * sizeof(T));
        tma_prefetch_desc(&tmap);
        if (a.tma_store) tma_prefetch_desc(&tmap_y);
        auto issue = [&](long long tile, int stage) {
          int ct, tt, lo;
          tile_coords(tile, ct, tt, lo);
          mbar_expect_tx(&full[stage], stage_bytes);
          tma_load_2d(stage ? xs1 : xs0, &tmap, lo, ct * a.CH, &full[stage]);
        };
        if (first < a.n_tiles) issue(first, 0);
        if (first + step < a.n_tiles) issue(first + step, 1);
        constexpr int OUTS_PER_BOX = 128 / static_cast<int>(sizeof(T));
        const int boxes = a.GT * kRM / OUTS_PER_BOX;
        const size_t box_elems = static_cast<size_t>(a.CH) * OUTS_PER_BOX;
        long long it = 0;
        for (long long tile = first; tile < a.n_tiles; tile += step, ++it) {
          const int stage = static_cast<int>(it & 1);
          const uint32_t phase = static_cast<uint32_t>((it >> 1) & 1);
          mbar_wait(&done[stage], phase);
          if (a.tma_store) {
            int ct, tt, lo;
            tile_coords(tile, ct, tt, lo);
            const T* src = stage ? xs1 : xs0;
            for (int b = 0; b < boxes; ++b)
              tma_store_2d(&tmap_y, src + b * box_elems, tt * a.GT * kRM + b * OUTS_PER_BOX, ct * a.CH);
            tma_store_commit();
            tma_store_wait_read0();   // the stage may be overwritten once TMA has read it
          }
          const long long nxt = tile + 2 * step;
          if (nxt < a.n_tiles) issue(nxt, stage);
        }
        if (a.tma_store) tma_store_wait_all0();
      }
      return;
    }
  }

  // ---------------- consumer warps ----------------
  const size_t slot_stride = static_cast<size_t>(32) * a.PITCH;
  long long it = 0;
  for (long long tile = first; tile < a.n_tiles; tile += step, ++it) {
    const int stage = static_cast<int>(it & 1);
    const uint32_t phase = static_cast<uint32_t>((it >> 1) & 1);
    T* xs = stage ? xs1 : xs0;
    int ct, tt, tile_lo;
    tile_coords(tile, ct, tt, tile_lo);
    if constexpr (kTma) {
      mbar_wait(&full[stage], phase);
    } else {
      const int total = a.CH * a.PITCH;
      for (int i = tid; i < total; i += nthreads) {
        const int row = i / a.PITCH;
        const int col = i - row * a.PITCH;
        const long long ch = static_cast<long long>(ct) * a.CH + row;
        const long long gi = static_cast<long long>(tile_lo) + col;
        T v = T(0);
        if (ch < a.channels && gi >= 0 && gi < a.n_in) v = a.x[ch * a.x_stride + gi];
        xs[i] = v;
      }
      __syncthreads();
    }

    const int G = tt * a.GT + warp;   // global output group
    const long long m0 = static_cast<long long>(G) * kRM;
    const bool active = m0 < a.n_out;
    Acc<T, RC> acc;
    if (active) {
      const int k = G / a.GP;
      const int gi = G - k * a.GP;
      const int off = k * a.PI + s_group_lo[gi] - tile_lo;
      const int phi = off & 3;
      const int nquads = (phi + a.W + 3) >> 2;
      const T* tab = s_table + (static_cast<size_t>(gi) * a.WROWS + (3 - phi)) * kRM;
      const T* xrow = xs + static_cast<size_t>(lane) * a.PITCH + (off - phi);
      src_group_mac<T, RC>(acc, tab, xrow, slot_stride, nquads);
    }

    if constexpr (kTma) {
      if (a.tma_store) {
        // every consumer is done reading this stage -> reuse it as the output tile
        asm volatile("bar.sync 1, %0;" ::"r"(a.GT * 32) : "memory");
        if (active) {
          constexpr int OUTS_PER_BOX = 128 / static_cast<int>(sizeof(T));
          constexpr int GROUPS_PER_BOX = OUTS_PER_BOX / kRM;
          constexpr int CHUNKS = kRM * static_cast<int>(sizeof(T)) / 16;   // 16-byte chunks per group
          unsigned char* box = reinterpret_cast<unsigned char*>(xs) +
                               static_cast<size_t>(warp / GROUPS_PER_BOX) * a.CH * 128;
          const int chunk0 = (warp % GROUPS_PER_BOX) * CHUNKS;
#pragma unroll
          for (int c = 0; c < RC; ++c) {
            const int row = c * 32 + lane;
            unsigned char* rp = box + static_cast<size_t>(row) * 128;
#pragma unroll
            for (int q = 0; q < CHUNKS; ++q) {
              void* dst = rp + (((chunk0 + q) ^ (row & 7)) << 4);
              if constexpr (sizeof(T) == 4)
                *reinterpret_cast<float4*>(dst) = make_float4(acc.get(4 * q, c), acc.get(4 * q + 1, c),
                                                               acc.get(4 * q + 2, c), acc.get(4 * q + 3, c));
              else
                *reinterpret_cast<double2*>(dst) = make_double2(acc.get(2 * q, c), acc.get(2 * q + 1, c));
            }
          }
          fence_proxy_async();
        }
      } else if (active) {
        src_store_direct<T, RC>(acc, a, ct, lane, m0);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&done[stage]);
    } else {
      if (active) src_store_direct<T, RC>(acc, a, ct, lane, m0);
      __syncthreads();  // every warp is done with this stage
    }
  }
}

// One thread per output sample; any geometry.
template <typename T>
__global__ void __launch_bounds__(256)
src_generic_kernel(const T* __restrict__ x, long long x_stride, T* __restrict__ y, long long y_stride,
                   long long channels, long long n_in, long long n_out, int L, int M, long long P,
                   int n_taps, const T* __restrict__ h) {
  const long long m = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (m >= n_out) return;
  const long long q = m * M + P;
  const long long i0 = q / L;
  const int p = static_cast<int>(q - i0 * L);
  long long jlo = i0 - (n_in - 1);
  if (jlo < 0) jlo = 0;
  long long jhi = (n_taps - 1 - p) / L;
  if (p > n_taps - 1) jhi = -1;
  if (jhi > i0) jhi = i0;
  for (long long ch = blockIdx.y; ch < channels; ch += gridDim.y) {
    const T* xc = x + ch * x_stride;
    T acc = T(0);
    for (long long j = jlo; j <= jhi; ++j) acc += h[p + j * L] * xc[i0 - j];
    y[ch * y_stride + m] = acc;
  }
}

}  // namespace dspb200

struct dspb200_src_plan {
  uint32_t magic = dspb200::kMagicSrc;   // first member: checked by every entry point
  int L, M, dtype, n_taps;
  int device;
  std::vector<double> taps;   // h * L, float64
  void* d_taps = nullptr;     // device copy in the plan dtype
  dspb200::SrcTiledGeom geom;
  void* d_table = nullptr;
  int* d_group_lo = nullptr;
  int* d_tile_lo = nullptr;
  dspb200::SrcMmaPlan mma;   // fp32: tensor-core form (src_mma.cu)
};

namespace dspb200 {

static void src_geometry(int L, int M, int64_t n_in, int& T, int64_t& P, int64_t& n_out) {
  const int big = L > M ? L : M;
  T = 40 * big + 1;
  const int64_t n_exp = n_in * L;
  P = ((n_exp < T ? n_exp : T) - 1) / 2;
  const int64_t n_filt = n_exp > T ? n_exp : T;
  n_out = (n_filt + M - 1) / M;
}

// Build the tap-row table and tile geometry of the tiled kernel for the
// long-signal case (P = (T-1)/2).  Returns geom.ok = 0 when it does not fit.
template <typename T>
static void build_tiled(const std::vector<double>& h, int L, int M, int smem_limit, SrcTiledGeom& g,
                        std::vector<T>& table, std::vector<int>& group_lo, std::vector<int>& tile_lo) {
  const int Tt = static_cast<int>(h.size());
  const int64_t P = (Tt - 1) / 2;
  auto gcd = [](int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; };
  const int gg = gcd(L, kRM);
  g.GP = L / gg;
  g.PO = kRM * g.GP;
  g.PI = kRM * M / gg;
  g.CH = 32 * SrcCfg<T>::RC;
  struct GInfo { int64_t i0[kRM]; int p[kRM]; int nt[kRM]; int64_t lo, hi; };
  std::vector<GInfo> info(static_cast<size_t>(g.GP));
  int W = 0;
  for (int gi = 0; gi < g.GP; ++gi) {
    GInfo& I = info[static_cast<size_t>(gi)];
    I.lo = INT64_MAX; I.hi = INT64_MIN;
    for (int r = 0; r < kRM; ++r) {
      const int64_t q = static_cast<int64_t>(kRM * gi + r) * M + P;
      I.i0[r] = q / L;
      I.p[r] = static_cast<int>(q % L);
      I.nt[r] = I.p[r] <= Tt - 1 ? (Tt - 1 - I.p[r]) / L + 1 : 0;
      if (I.nt[r] > 0) {
        I.lo = std::min(I.lo, I.i0[r] - I.nt[r] + 1);
        I.hi = std::max(I.hi, I.i0[r]);
      }
    }
    if (I.lo == INT64_MAX) { I.lo = I.i0[0]; I.hi = I.i0[0]; }
    W = std::max(W, static_cast<int>(I.hi - I.lo + 1));
  }
  g.W = W;
  g.WROWS = 4 * ((W + 6) / 4) + 8;
  g.table_bytes = static_cast<size_t>(g.GP) * g.WROWS * kRM * sizeof(T);
  table.assign(static_cast<size_t>(g.GP) * g.WROWS * kRM, T(0));
  group_lo.resize(static_cast<size_t>(g.GP));
  for (int gi = 0; gi < g.GP; ++gi) {
    const GInfo& I = info[static_cast<size_t>(gi)];
    group_lo[static_cast<size_t>(gi)] = static_cast<int>(I.lo);
    for (int s = 0; s < W; ++s) {
      const int64_t i = I.lo + s;
      for (int r = 0; r < kRM; ++r) {
        const int64_t j = I.i0[r] - i;
        if (j >= 0 && j < I.nt[r])
          table[(static_cast<size_t>(gi) * g.WROWS + static_cast<size_t>(s + 3)) * kRM + r] =
              static_cast<T>(h[static_cast<size_t>(I.p[r] + j * L)]);
      }
    }
  }
  // choose the largest tile (groups per tile) whose window fits a TMA box and shared memory
  const int vec = 16 / static_cast<int>(sizeof(T));
  for (int GT = kSrcMaxWarps; GT >= 2; GT >>= 1) {
    tile_lo.assign(static_cast<size_t>(g.GP), 0);
    int span = 0;
    for (int gi0 = 0; gi0 < g.GP; ++gi0) {
      int64_t lo = INT64_MAX;
      for (int t = 0; t < GT; ++t) {
        const int gl = gi0 + t;
        const int64_t in_lo = static_cast<int64_t>(gl / g.GP) * g.PI + info[static_cast<size_t>(gl % g.GP)].lo;
        lo = std::min(lo, in_lo);
      }
      tile_lo[static_cast<size_t>(gi0)] = static_cast<int>(lo);
      for (int t = 0; t < GT; ++t) {
        const int gl = gi0 + t;
        const int64_t in_lo = static_cast<int64_t>(gl / g.GP) * g.PI + info[static_cast<size_t>(gl % g.GP)].lo;
        span = std::max(span, static_cast<int>(in_lo - lo) + W + 3 + (vec - 1));
      }
    }
    int pitch = static_cast<int>(round_up(span, vec));
    while ((static_cast<size_t>(pitch) * sizeof(T)) % 128 != 16) pitch += vec;
    if (pitch > 256) continue;
    const size_t stage = static_cast<size_t>(g.CH) * pitch * sizeof(T);
    const size_t total = 2 * stage + g.table_bytes + 2 * static_cast<size_t>(g.GP) * sizeof(int) + 64;  // + 4 mbarriers
    if (total > static_cast<size_t>(smem_limit)) continue;
    g.GT = GT;
    g.PITCH = pitch;
    g.stage_bytes = stage;
    g.smem_bytes = total;
    g.ok = 1;
    return;
  }
  g.ok = 0;
}

template <typename T>
int src_run(const dspb200_src_plan* plan, const T* x, int64_t xs, T* y, int64_t ys,
            int64_t channels, int64_t n_in, cudaStream_t stream, int force_kind) {
  DSP_PLAN(plan, kMagicSrc, "src");
  DSP_CHECK(plan->dtype == DType<T>::id, "plan dtype %d does not match the entry point", plan->dtype);
  DSP_CHECK(channels >= 0, "negative channel count");
  DSP_CHECK(n_in >= 1, "input must hold at least one sample (numpy.convolve rejects empty input)");
  DSP_CHECK(n_in < (1LL << 31) / (plan->L > plan->M ? plan->L : plan->M), "signal too long for 32-bit tile coordinates");
  if (channels == 0) return DSPB200_OK;
  DSP_CHECK(x != nullptr && y != nullptr, "NULL buffer");
  DSP_TRY(ensure_device());
  DSP_TRY(check_plan_device(plan->device, "src"));
  int Tt; int64_t P, n_out;
  src_geometry(plan->L, plan->M, n_in, Tt, P, n_out);
  DSP_CHECK(xs >= n_in && ys >= n_out, "channel stride smaller than the row length");
  const SrcTiledGeom& g = plan->geom;
  const bool long_signal = n_in * plan->L >= Tt;
  bool tiled = g.ok && long_signal;
  if (force_kind == 0) tiled = false;
  if constexpr (sizeof(T) == 4) {
    if (long_signal && force_kind != 0 && force_kind != 1 && getenv("DSPB200_SRC_NO_MMA") == nullptr &&
        src_mma_usable(plan->mma, x, xs, channels, n_in))
      return src_mma_run(plan->mma, x, xs, y, ys, channels, n_in, n_out, stream);
  }
  if (!tiled) {
    const int threads = 256;
    dim3 grid(static_cast<unsigned>(ceil_div(n_out, threads)), static_cast<unsigned>(channels < 65535 ? channels : 65535));
    src_generic_kernel<T><<<grid, threads, 0, stream>>>(x, xs, y, ys, channels, n_in, n_out, plan->L, plan->M, P,
                                                        Tt, static_cast<const T*>(plan->d_taps));
    return after_launch("src_generic_kernel");
  }
  SrcTiledArgs<T> a{};
  a.x = x; a.x_stride = xs; a.y = y; a.y_stride = ys;
  a.channels = channels; a.n_in = n_in; a.n_out = n_out;
  a.table = static_cast<const T*>(plan->d_table);
  a.group_lo = plan->d_group_lo; a.tile_lo = plan->d_tile_lo;
  a.GP = g.GP; a.PI = g.PI; a.W = g.W; a.WROWS = g.WROWS; a.GT = g.GT; a.PITCH = g.PITCH; a.CH = g.CH;
  const int64_t n_groups = ceil_div(n_out, kRM);
  a.n_tt = static_cast<int>(ceil_div(n_groups, g.GT));
  const int64_t n_ct = ceil_div(channels, g.CH);
  a.n_tiles = n_ct * a.n_tt;
  const int vec = 16 / static_cast<int>(sizeof(T));
  a.y_vec_ok = (reinterpret_cast<uintptr_t>(y) % 16 == 0 && ys % vec == 0) ? 1 : 0;
  const bool tma_ok = reinterpret_cast<uintptr_t>(x) % 16 == 0 && (xs * sizeof(T)) % 16 == 0;
  CUtensorMap tmap, tmap_y;
  memset(&tmap, 0, sizeof(tmap));
  memset(&tmap_y, 0, sizeof(tmap_y));
  if (tma_ok)
    DSP_TRY(encode_tmap_2d(&tmap, plan->dtype, x, static_cast<uint64_t>(n_in), static_cast<uint64_t>(channels),
                           static_cast<uint64_t>(xs) * sizeof(T), static_cast<uint32_t>(g.PITCH),
                           static_cast<uint32_t>(g.CH)));
  // TMA store path: y rows 16-byte aligned, the output tile fits the consumed
  // input stage, and the tile is a whole number of 128-byte swizzled boxes.
  const int outs_per_box = 128 / static_cast<int>(sizeof(T));
  a.tma_store = 0;
  if (tma_ok && a.y_vec_ok && g.GT * kRM <= g.PITCH && (g.GT * kRM) % outs_per_box == 0 &&
      getenv("DSPB200_SRC_NO_TMA_STORE") == nullptr) {
    DSP_TRY(encode_tmap_2d(&tmap_y, plan->dtype, y, static_cast<uint64_t>(n_out), static_cast<uint64_t>(channels),
                           static_cast<uint64_t>(ys) * sizeof(T), static_cast<uint32_t>(outs_per_box),
                           static_cast<uint32_t>(g.CH), true));
    a.tma_store = 1;
  }
  const int64_t sms = sm_count();
  const int grid = static_cast<int>(a.n_tiles < sms ? a.n_tiles : sms);
  constexpr int RC = SrcCfg<T>::RC;
  if (tma_ok) {
    auto kern = src_tiled_kernel<T, RC, true>;
    DSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(g.smem_bytes)));
    kern<<<grid, (g.GT + 1) * 32, g.smem_bytes, stream>>>(tmap, tmap_y, a);
  } else {
    auto kern = src_tiled_kernel<T, RC, false>;
    DSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(g.smem_bytes)));
    kern<<<grid, g.GT * 32, g.smem_bytes, stream>>>(tmap, tmap_y, a);
  }
  return after_launch("src_tiled_kernel");
}

template int src_run<float>(const dspb200_src_plan*, const float*, int64_t, float*, int64_t, int64_t, int64_t, cudaStream_t, int);
template int src_run<double>(const dspb200_src_plan*, const double*, int64_t, double*, int64_t, int64_t, int64_t, cudaStream_t, int);

int src_plan_ratio(const dspb200_src_plan* plan, int* L, int* M, int* dtype) {
  DSP_PLAN(plan, kMagicSrc, "src");
  *L = plan->L; *M = plan->M; *dtype = plan->dtype;
  return DSPB200_OK;
}

const std::vector<double>* src_plan_taps(const dspb200_src_plan* plan) {
  return (plan && plan->magic == kMagicSrc) ? &plan->taps : nullptr;
}

template <typename T>
static int upload(const std::vector<T>& v, void** dptr) {
  *dptr = nullptr;
  if (v.empty()) return DSPB200_OK;
  DSP_CUDA(cudaMalloc(dptr, v.size() * sizeof(T)));
  DSP_CUDA(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return DSPB200_OK;
}

template <typename T>
static int plan_tables(dspb200_src_plan* p) {
  std::vector<T> taps_t(p->taps.size());
  for (size_t i = 0; i < taps_t.size(); ++i) taps_t[i] = static_cast<T>(p->taps[i]);
  DSP_TRY(upload(taps_t, &p->d_taps));
  std::vector<T> table;
  std::vector<int> group_lo, tile_lo;
  build_tiled<T>(p->taps, p->L, p->M, max_smem_optin(), p->geom, table, group_lo, tile_lo);
  if (p->geom.ok) {
    DSP_TRY(upload(table, &p->d_table));
    void* q = nullptr;
    DSP_TRY(upload(group_lo, &q)); p->d_group_lo = static_cast<int*>(q);
    DSP_TRY(upload(tile_lo, &q)); p->d_tile_lo = static_cast<int*>(q);
  }
  if constexpr (sizeof(T) == 4) DSP_TRY(src_mma_build(p->taps, p->L, p->M, p->mma));
  return DSPB200_OK;
}

// Host buffers through a plan the caller keeps (the drop-in module caches one per ratio and device) or, with
// plan == NULL, through a plan built for the call.
template <typename T>
static int src_host(const dspb200_src_plan* given, int L, int M, const T* x, int64_t channels, int64_t n_in, T* y,
                    int64_t y_cap, int64_t* n_out_p) {
  DSP_CHECK(n_out_p != nullptr, "n_out is NULL");
  DSP_CHECK(channels >= 0 && n_in >= 1, "bad shape");
  if (given) {
    int dt = 0;
    DSP_TRY(src_plan_ratio(given, &L, &M, &dt));
    DSP_CHECK(dt == DType<T>::id, "plan dtype %d does not match the entry point", dt);
  }
  int Tt; int64_t P, n_out;
  src_geometry(L, M, n_in, Tt, P, n_out);
  *n_out_p = n_out;
  DSP_CHECK(y_cap >= n_out, "y capacity %lld < %lld outputs", (long long)y_cap, (long long)n_out);
  if (channels == 0) return DSPB200_OK;
  DSP_CHECK(x != nullptr && y != nullptr, "NULL buffer");
  dspb200_src_plan* own = nullptr;
  if (!given) DSP_TRY(dspb200_src_plan_create(L, M, DType<T>::id, &own));
  const dspb200_src_plan* plan = given ? given : own;
  const int vec = 16 / static_cast<int>(sizeof(T));
  const int64_t xp = round_up(n_in, vec), yp = round_up(n_out, vec);
  T *dx = nullptr, *dy = nullptr;
  cudaError_t e = cudaMalloc(&dx, static_cast<size_t>(channels) * xp * sizeof(T));
  if (e == cudaSuccess) e = cudaMalloc(&dy, static_cast<size_t>(channels) * yp * sizeof(T));
  int rc = DSPB200_OK;
  if (e == cudaSuccess)
    e = cudaMemcpy2DAsync(dx, xp * sizeof(T), x, n_in * sizeof(T), n_in * sizeof(T), channels, cudaMemcpyHostToDevice, 0);
  if (e == cudaSuccess) rc = src_run<T>(plan, dx, xp, dy, yp, channels, n_in, nullptr, -1);
  if (e == cudaSuccess && rc == DSPB200_OK)
    e = cudaMemcpy2DAsync(y, y_cap * sizeof(T), dy, yp * sizeof(T), n_out * sizeof(T), channels, cudaMemcpyDeviceToHost, 0);
  if (e == cudaSuccess) e = cudaStreamSynchronize(0);
  cudaFree(dx);
  cudaFree(dy);
  if (own) dspb200_src_plan_destroy(own);
  if (e != cudaSuccess) return fail(DSPB200_ERR_CUDA, "src host path: %s", cudaGetErrorString(e));
  return rc;
}

}  // namespace dspb200

using namespace dspb200;

extern "C" {

int dspb200_src_geometry(int L, int M, int64_t n_in, int* n_taps, int64_t* centre, int64_t* n_out) {
  DSP_CHECK(L >= 1 && M >= 1, "L and M must be >= 1 (got L=%d M=%d)", L, M);
  DSP_CHECK(n_in >= 0, "n_in must be >= 0");
  int T; int64_t P, no;
  src_geometry(L, M, n_in, T, P, no);
  if (n_taps) *n_taps = T;
  if (centre) *centre = P;
  if (n_out) *n_out = no;
  return DSPB200_OK;
}

int dspb200_src_plan_create(int L, int M, int dtype, dspb200_src_plan** plan) {
  DSP_CHECK(plan != nullptr, "plan output pointer is NULL");
  DSP_CHECK(L >= 1 && M >= 1, "L and M must be >= 1 (got L=%d M=%d)", L, M);
  DSP_CHECK(!(L == 1 && M == 1), "L == M == 1 is the reference's bypass: return the input object instead");
  DSP_CHECK(L <= 4096 && M <= 4096, "L and M above 4096 are not supported");
  DSP_CHECK(dtype == DSPB200_F32 || dtype == DSPB200_F64, "dtype must be 0 (f32) or 1 (f64)");
  DSP_TRY(ensure_device());
  dspb200_src_plan* p = new (std::nothrow) dspb200_src_plan();
  if (!p) return fail(DSPB200_ERR_ALLOC, "out of host memory");
  p->L = L; p->M = M; p->dtype = dtype;
  p->taps = src_filter(L, M);
  p->n_taps = static_cast<int>(p->taps.size());
  cudaGetDevice(&p->device);
  int rc = dtype == DSPB200_F32 ? plan_tables<float>(p) : plan_tables<double>(p);
  if (rc != DSPB200_OK) {
    dspb200_src_plan_destroy(p);
    return rc;
  }
  *plan = p;
  return DSPB200_OK;
}

int dspb200_src_plan_destroy(dspb200_src_plan* p) {
  if (!p) return DSPB200_OK;
  DSP_PLAN(p, kMagicSrc, "src");
  p->magic = 0;
  cudaFree(p->d_taps);
  cudaFree(p->d_table);
  cudaFree(p->d_group_lo);
  cudaFree(p->d_tile_lo);
  src_mma_free(p->mma);
  delete p;
  return DSPB200_OK;
}

int dspb200_src_plan_kernel_kind(const dspb200_src_plan* plan, int64_t channels, int64_t n_in,
                                 int64_t x_stride, int* kind) {
  DSP_PLAN(plan, kMagicSrc, "src");
  DSP_CHECK(kind != nullptr, "NULL argument");
  (void)channels; (void)x_stride;
  const int T = plan->n_taps;
  *kind = (plan->geom.ok && n_in * plan->L >= T) ? 1 : 0;
  if (plan->dtype == DSPB200_F32 && n_in * plan->L >= T && getenv("DSPB200_SRC_NO_MMA") == nullptr &&
      src_mma_usable(plan->mma, nullptr, x_stride, channels, n_in))
    *kind = 2;
  return DSPB200_OK;
}

int dspb200_src_run_f32(const dspb200_src_plan* plan, const float* x, int64_t xs, float* y, int64_t ys,
                        int64_t channels, int64_t n_in, void* stream) {
  return src_run<float>(plan, x, xs, y, ys, channels, n_in, static_cast<cudaStream_t>(stream), -1);
}
int dspb200_src_run_f64(const dspb200_src_plan* plan, const double* x, int64_t xs, double* y, int64_t ys,
                        int64_t channels, int64_t n_in, void* stream) {
  return src_run<double>(plan, x, xs, y, ys, channels, n_in, static_cast<cudaStream_t>(stream), -1);
}
/* test hooks: force the generic kernel (kind 0) / the tiled FFMA kernel (kind 1) regardless of geometry */
int dspb200_src_run_tiled_f32(const dspb200_src_plan* plan, const float* x, int64_t xs, float* y, int64_t ys,
                              int64_t channels, int64_t n_in, void* stream) {
  return src_run<float>(plan, x, xs, y, ys, channels, n_in, static_cast<cudaStream_t>(stream), 1);
}
int dspb200_src_run_generic_f32(const dspb200_src_plan* plan, const float* x, int64_t xs, float* y, int64_t ys,
                                int64_t channels, int64_t n_in, void* stream) {
  return src_run<float>(plan, x, xs, y, ys, channels, n_in, static_cast<cudaStream_t>(stream), 0);
}
int dspb200_src_run_generic_f64(const dspb200_src_plan* plan, const double* x, int64_t xs, double* y, int64_t ys,
                                int64_t channels, int64_t n_in, void* stream) {
  return src_run<double>(plan, x, xs, y, ys, channels, n_in, static_cast<cudaStream_t>(stream), 0);
}
int dspb200_src_host_f32(int L, int M, const float* x, int64_t channels, int64_t n_in, float* y,
                         int64_t y_cap, int64_t* n_out) {
  return src_host<float>(nullptr, L, M, x, channels, n_in, y, y_cap, n_out);
}
int dspb200_src_host_f64(int L, int M, const double* x, int64_t channels, int64_t n_in, double* y,
                         int64_t y_cap, int64_t* n_out) {
  return src_host<double>(nullptr, L, M, x, channels, n_in, y, y_cap, n_out);
}
int dspb200_src_plan_host_f32(const dspb200_src_plan* plan, const float* x, int64_t channels, int64_t n_in, float* y,
                              int64_t y_cap, int64_t* n_out) {
  DSP_CHECK(plan != nullptr, "NULL plan");
  return src_host<float>(plan, 1, 1, x, channels, n_in, y, y_cap, n_out);
}
int dspb200_src_plan_host_f64(const dspb200_src_plan* plan, const double* x, int64_t channels, int64_t n_in, double* y,
                              int64_t y_cap, int64_t* n_out) {
  DSP_CHECK(plan != nullptr, "NULL plan");
  return src_host<double>(plan, 1, 1, x, channels, n_in, y, y_cap, n_out);
}

}  // extern "C"
