// K2 on the tensor cores: the biquad cascade as a chunked linear system.
//
// The whole cascade of sistema_ecualizador (dsp_core.py:216-254; up to 8 second-order sections)
// is ONE linear time-invariant system with S = 2 * sections states.  Over a chunk of 112 samples
//     z_k = T x_k + O s_k ,      s_(k+1) = Phi s_k + K x_k
// with T [112 x 112] the lower-triangular Toeplitz matrix of the cascade's impulse response,
// K [S x 112] the state reached from a zero start, O [112 x S] the free response and Phi = A^112.
// [T; K] x_k is a GEMM: D[128 x 256 channels] = [T; K][128 x 112] . X[256 x 112]^T on tcgen05.mma
// (kind::tf32 with the three-product split of src_mma.cu, fp32 accumulators in TMEM).  Rows
// 0..111 of D are the zero-state outputs, rows 112..112+S-1 the zero-state end states u_k.
// The sequential part that remains is s_(k+1) = Phi s_k + u_k, S^2 FMA per channel and chunk.
//
// Tiles are (channel group of 256, chunk) pairs, ordered chunk-major and handed out by an atomic
// counter, so every tile a tile waits for is owned by a CTA that is already running.  The state
// crosses CTAs by decoupled look-back: the epilogue threads (one per channel) publish the tile's
// aggregate u_k, walk back over the predecessors' aggregates until they meet a published start
// state, compose forward with Phi, publish their own end state and hand the start state to the
// drain through shared memory.  The drain adds O s_k in fp32 FMAs (row r of O lives in the
// registers of the thread that owns TMEM lane r), clips once (dsp_core.py:254) and stores 32
// consecutive samples of one channel per instruction.
//
// The tables are general ([period] tile phases, window starts, input advance), so the same kernel
// also runs SRC and EQ fused: [T; K] . A_p with A_p the resampler's banded tap matrix.
//
// Non-finite inputs poison their whole 112-sample chunk (0 * NaN inside the GEMM), not only the
// samples after them as the sequential reference does.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <numeric>
#include <vector>

#include "design.cuh"
#include "internal.cuh"

namespace dspb200 {

namespace {

constexpr int kTM = 128;        // MMA M: kRows output rows + state rows
constexpr int kRows = 112;      // samples per chunk
constexpr int kTN = 256;        // channels per tile (MMA N)
constexpr int kBK = 32;         // k-values per stage (one 128-byte swizzle row)
constexpr int kXSlots = 4;
constexpr int kAlSlots = 2;
constexpr int kEpiWarps = 8;
constexpr int kConvWarps = 4;
// warp roles: 0-7 drain, 8 tile dispenser + TMA producer of x, 9 MMA issuer, 10 TMA producer of the coefficient
// tiles, 11 and 15 state warps (TMEM lane quarter 3, one per channel half), 12-14 and 16 converters
constexpr int kTmaWarp = 8, kMmaWarp = 9, kTmaWarpA = 10, kStateWarp0 = 11, kStateWarp1 = 15;
constexpr int kWarps = 17;
constexpr int kThreads = kWarps * 32;
constexpr int kTileQ = 8;       // tile queue depth; no role runs more than 4 tiles ahead of the drain
constexpr int kStagePitch = 33;
constexpr uint32_t kABytes = kTM * kBK * 4;   // 16 KB
constexpr uint32_t kBBytes = kTN * kBK * 4;   // 32 KB
template <int kS> struct LtiSmem {
  static constexpr int kAhSlots = kS > 12 ? 2 : 3;
  static constexpr size_t kSBytes = static_cast<size_t>(kTN) * kS * 4;              // start states, by channel pair
  static constexpr size_t kStageBytes = static_cast<size_t>(2) * kS * kStagePitch * 4;   // u transposition, one per state warp
  static constexpr size_t kBytes = kXSlots * kBBytes + (kAhSlots + kAlSlots) * kABytes + kSBytes + kStageBytes + 1024;
};

struct LtiArgs {
  float* z; long long z_stride;
  long long channels, n_out;
  const int* lo;                // [period] window start of tile phase p
  int period, nkb, kvalid;      // kvalid: GEMM depth actually multiplied (multiple of 8)
  long long adv;                // input samples per `period` chunks
  long long n_tt, n_groups, n_tiles;
  const float* o_tab;           // [128][16] free-response rows
  float* ring;                  // [n_tiles][2][kS][256]: aggregate u, end state; written once per tile
  unsigned* flags;              // [n_tiles][half]: 0 nothing, 1 aggregate published, 2 end state published
  unsigned* counter;            // tile dispenser
  unsigned long long* prof;     // development: cycles per epilogue phase (NULL = off)
  int clip;
  float phi[kLtiMaxStates * kLtiMaxStates];
};

__device__ __forceinline__ uint64_t umma_desc_sw128(const void* p) {
  const uint64_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  return ((addr >> 4) & 0x3FFF) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar))) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t* v, uint32_t taddr) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ float4 lds128(uint32_t addr) {   // explicit shared-space load: ptxas may move it across global stores
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float clip_unit(float y) {   // NaN passes through like np.clip
  float r;
  asm("max.NaN.f32 %0, %1, 0fBF800000;\n\tmin.NaN.f32 %0, %0, 0f3F800000;" : "=f"(r) : "f"(y));
  return r;
}

// s <- Phi s + u.  Phi sits in the kernel parameters: every FFMA takes its coefficient from the constant bank.
template <int kS>
__device__ __forceinline__ void advance_state(float (&s)[kS], const float* __restrict__ phi, const float (&u)[kS]) {
  float t[kS];
#pragma unroll
  for (int i = 0; i < kS; ++i) {
    float acc = u[i];
#pragma unroll
    for (int j = 0; j < kS; ++j) acc = fmaf(phi[i * kLtiMaxStates + j], s[j], acc);
    t[i] = acc;
  }
#pragma unroll
  for (int i = 0; i < kS; ++i) s[i] = t[i];
}

template <int kS>
__global__ void __launch_bounds__(kThreads, 1)
lti_mma_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_x,
               const __grid_constant__ LtiArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  constexpr int kAhSlots = LtiSmem<kS>::kAhSlots;
  __shared__ __align__(8) uint64_t bars[4 * kXSlots + 2 * kAhSlots + 2 * kAlSlots + 4 + kTileQ + 4];
  __shared__ long long tile_q[kTileQ];
  __shared__ uint32_t tmem_base_s;
  uint64_t* full_x = bars;
  uint64_t* mid = full_x + kXSlots;
  uint64_t* conv = mid + kXSlots;
  uint64_t* empty_x = conv + kXSlots;
  uint64_t* full_ah = empty_x + kXSlots;
  uint64_t* empty_ah = full_ah + kAhSlots;
  uint64_t* full_al = empty_ah + kAhSlots;
  uint64_t* empty_al = full_al + kAlSlots;
  uint64_t* acc_full = empty_al + kAlSlots;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* tile_full = acc_empty + 2;
  uint64_t* s_full = tile_full + kTileQ;      // [half] start states of the tile are in shared memory
  uint64_t* s_free = s_full + 2;              // [half] the drain warps of that half are done with them
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kXSlots; ++s) {
      mbar_init(&full_x[s], 1); mbar_init(&mid[s], 1); mbar_init(&conv[s], kConvWarps); mbar_init(&empty_x[s], 1);
    }
    for (int s = 0; s < kAhSlots; ++s) { mbar_init(&full_ah[s], 1); mbar_init(&empty_ah[s], 1); }
    for (int s = 0; s < kAlSlots; ++s) { mbar_init(&full_al[s], 1); mbar_init(&empty_al[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], kEpiWarps); }
    for (int s = 0; s < kTileQ; ++s) mbar_init(&tile_full[s], 1);
    for (int h = 0; h < 2; ++h) { mbar_init(&s_full[h], 32); mbar_init(&s_free[h], kEpiWarps / 2); }
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
  float* s_s = reinterpret_cast<float*>(al_ptr(kAlSlots));                       // [128 pairs][kS][2]
  float* stage_s = s_s + static_cast<size_t>(kTN) * kS;                          // [2][kS][33]
  // the i-th tile of this CTA, as dispensed by the x producer (-1: no more work)
  auto next_tile = [&](uint32_t i) -> long long {
    mbar_wait(&tile_full[i % kTileQ], (i / kTileQ) & 1);
    return *reinterpret_cast<volatile long long*>(&tile_q[i % kTileQ]);
  };

  if (warp == kTmaWarp) {
    // ---------------- tile dispenser + TMA producer of x tiles ----------------
    if (lane == 0) {
      tma_prefetch_desc(&tm_x);
      uint32_t it = 0;
      for (uint32_t ti = 0;; ++ti) {
        long long tile = static_cast<long long>(atomicAdd(a.counter, 1u));
        if (tile >= a.n_tiles) tile = -1;
        *reinterpret_cast<volatile long long*>(&tile_q[ti % kTileQ]) = tile;
        mbar_arrive(&tile_full[ti % kTileQ]);
        if (tile < 0) break;
        const long long tt = tile / a.n_groups, g = tile - tt * a.n_groups;
        const int p = static_cast<int>(tt % a.period);
        const long long lo = a.lo[p] + (tt / a.period) * a.adv;
        for (int kb = 0; kb < a.nkb; ++kb, ++it) {
          const int s = it % kXSlots;
          if (it >= kXSlots) mbar_wait(&empty_x[s], ((it / kXSlots) - 1) & 1);
          mbar_expect_tx(&full_x[s], kBBytes);
          tma_load_2d(x_ptr(s), &tm_x, static_cast<int>(lo) + kb * kBK, static_cast<int>(g) * kTN, &full_x[s]);
        }
      }
    }
  } else if (warp == kTmaWarpA) {
    // ---------------- TMA producer, coefficient tiles (L2 resident) ----------------
    if (lane == 0) {
      tma_prefetch_desc(&tm_a);
      uint32_t it = 0;
      for (uint32_t ti = 0;; ++ti) {
        const long long tile = next_tile(ti);
        if (tile < 0) break;
        const long long tt = tile / a.n_groups;
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
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(kTN >> 3) << 17) | (uint32_t(kTM >> 4) << 24);
      auto ksteps = [&](int kb) -> int {
        const int left = a.kvalid - kb * kBK;
        return left >= kBK ? kBK / 8 : left / 8;
      };
      auto finish = [&](uint32_t j, int kb, uint32_t d, bool last_of_tile, int b) {
        const int sx = j % kXSlots, sh = j % kAhSlots;
        mbar_wait(&conv[sx], (j / kXSlots) & 1);
        tc_fence_after();
        const uint64_t dah = umma_desc_sw128(ah_ptr(sh)), dxl = umma_desc_sw128(x_ptr(sx));
        const int ks = ksteps(kb);
        for (int k = 0; k < ks; ++k) umma_tf32(d, dah + 2 * k, dxl + 2 * k, idesc, 1u);
        umma_commit(&empty_x[sx]);
        umma_commit(&empty_ah[sh]);
        if (last_of_tile) umma_commit(&acc_full[b]);
      };
      uint32_t it = 0;
      uint32_t prev_d = 0;
      int prev_b = 0, prev_kb = 0;
      bool have_prev = false, prev_last = false;
      for (uint32_t ti = 0;; ++ti) {
        const long long tile = next_tile(ti);
        if (tile < 0) break;
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
          const int ks = ksteps(kb);
          for (int k = 0; k < ks; ++k) {
            umma_tf32(d, dah + 2 * k, dx + 2 * k, idesc, (kb | k) ? 1u : 0u);
            umma_tf32(d, dal + 2 * k, dx + 2 * k, idesc, 1u);
          }
          umma_commit(&mid[sx]);
          umma_commit(&empty_al[sl]);
          if (have_prev) finish(it - 1, prev_kb, prev_d, prev_last, prev_b);
          have_prev = true; prev_d = d; prev_b = b; prev_kb = kb; prev_last = (kb == a.nkb - 1);
        }
      }
      if (have_prev) finish(it - 1, prev_kb, prev_d, prev_last, prev_b);
    }
  } else if (warp == kStateWarp0 || warp == kStateWarp1) {
    // ---------------- state warps: one per channel half, lane j carries channels h*128 + q*32 + j ----------------
    // u_k (TMEM rows 112..) -> aggregate record; look back to the nearest published end state; start state to
    // the drain warps of the same half through shared memory; end state to the successors.  These warps have
    // no other global stores in flight, so their fences are cheap, and they run a tile ahead of the drain.
    const int h = warp == kStateWarp0 ? 0 : 1;
    float* stg = stage_s + h * (kS * kStagePitch);
    const size_t rec = static_cast<size_t>(2) * kS * kTN;   // floats per tile record: aggregate, end state
    const int st_row = lane - (kRows - 96);
    for (uint32_t ti = 0;; ++ti) {
      const long long tile = next_tile(ti);
      if (tile < 0) break;
      const long long tt = tile / a.n_groups;
      const int b = ti & 1;
      long long t0 = 0;
      mbar_wait(&acc_full[b], (ti >> 1) & 1);
      tc_fence_after();
      if (a.prof && h == 0 && lane == 0) t0 = clock64();
      const uint32_t taddr = tmem + (static_cast<uint32_t>(96) << 16) + static_cast<uint32_t>(b * kTN + h * (kTN / 2));
      float* my_rec = a.ring + static_cast<size_t>(tile) * rec + h * (kTN / 2);
      unsigned* my_flag = a.flags + 2 * tile + h;
      // u of block q (32 channels): TMEM -> staging rows -> one channel per lane
      auto load_u = [&](int q, float (&u)[kS]) {
        uint32_t v[32];
        tmem_ld32(v, taddr + q * 32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        __syncwarp();
        if (st_row >= 0 && st_row < kS) {
#pragma unroll
          for (int j = 0; j < 32; ++j) stg[st_row * kStagePitch + j] = __uint_as_float(v[j]);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < kS; ++i) u[i] = stg[i * kStagePitch + lane];
      };
      // is the predecessor's end state already there?  (lane 0 decides for the warp)
      long long pt = tile - a.n_groups;
      unsigned f = 2u;
      if (tt > 0 && lane == 0) f = ld_acquire(a.flags + 2 * pt + h);
      f = __shfl_sync(0xffffffffu, f, 0);
      const bool direct = f == 2u;
      if (!direct) {
        // aggregate first, so successors can hop over this tile while it is still looking back
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
          float u[kS];
          load_u(q, u);
#pragma unroll
          for (int i = 0; i < kS; ++i) __stcg(my_rec + i * kTN + q * 32 + lane, u[i]);
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) {
          st_release(my_flag, 1u);
          for (;;) {   // chunk 0 only ever publishes an end state, so the walk ends there at the latest
            do { f = ld_acquire(a.flags + 2 * pt + h); } while (f == 0u);
            if (f == 2u) break;
            pt -= a.n_groups;
          }
        }
        pt = __shfl_sync(0xffffffffu, pt, 0);
      }
      __syncwarp();
      float s0[4][kS];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float s[kS], u[kS];
        if (tt == 0) {
#pragma unroll
          for (int i = 0; i < kS; ++i) s[i] = 0.f;
        } else {
          const float* r = a.ring + static_cast<size_t>(pt) * rec + kS * kTN + h * (kTN / 2) + q * 32 + lane;
#pragma unroll
          for (int i = 0; i < kS; ++i) s[i] = __ldcg(r + i * kTN);
          for (long long w = pt + a.n_groups; w < tile; w += a.n_groups) {
            const float* ra = a.ring + static_cast<size_t>(w) * rec + h * (kTN / 2) + q * 32 + lane;
            float g[kS];
#pragma unroll
            for (int i = 0; i < kS; ++i) g[i] = __ldcg(ra + i * kTN);
            advance_state<kS>(s, a.phi, g);
          }
        }
#pragma unroll
        for (int i = 0; i < kS; ++i) s0[q][i] = s[i];
        load_u(q, u);
        advance_state<kS>(s, a.phi, u);
#pragma unroll
        for (int i = 0; i < kS; ++i) __stcg(my_rec + (kS + i) * kTN + q * 32 + lane, s[i]);
      }
      __threadfence();
      __syncwarp();
      if (lane == 0) st_release(my_flag, 2u);
      // hand the start states to the drain warps of this half once they are done with the previous tile's
      if (ti > 0) mbar_wait(&s_free[h], (ti - 1) & 1);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = h * (kTN / 2) + q * 32 + lane;
        float* d = s_s + (static_cast<size_t>(c >> 1) * kS) * 2 + (c & 1);
#pragma unroll
        for (int i = 0; i < kS; ++i) d[2 * i] = s0[q][i];
      }
      mbar_arrive(&s_full[h]);
      if (a.prof && h == 0 && lane == 0) {
        atomicAdd(a.prof + 3, static_cast<unsigned long long>(clock64() - t0));
        atomicAdd(a.prof + 4, direct ? 1ull : 0ull);
      }
    }
  } else if (warp >= 12) {
    // ---------------- converters: x -> x - trunc_tf32(x) in place ----------------
    const int ctid = (warp == 16 ? 3 : warp - 12) * 32 + lane;
    uint32_t it = 0;
    for (uint32_t ti = 0;; ++ti) {
      const long long tile = next_tile(ti);
      if (tile < 0) break;
      for (int kb = 0; kb < a.nkb; ++kb, ++it) {
        const int s = it % kXSlots;
        mbar_wait(&mid[s], (it / kXSlots) & 1);
        float4* buf = reinterpret_cast<float4*>(x_ptr(s)) + ctid;
        constexpr int kPer = static_cast<int>(kBBytes / 16) / (32 * kConvWarps);
        float4 v[kPer];
#pragma unroll
        for (int i = 0; i < kPer; ++i) v[i] = buf[i * 32 * kConvWarps];
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
    // ---------------- drain warps 0-7: z = clip(D + O s) ----------------
    // warp w reads TMEM lanes 32(w%4).. (= chunk rows) and the column half w/4 (= 128 channels); every store
    // instruction writes 32 consecutive samples of one channel.  Two channels per packed FMA.
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;
    float o_r[kS];
#pragma unroll
    for (int i = 0; i < kS; ++i) o_r[i] = a.o_tab[row * kLtiMaxStates + i];
    const uint32_t s_addr = static_cast<uint32_t>(__cvta_generic_to_shared(s_s)) + static_cast<uint32_t>(half * (kTN / 4) * kS * 8);
    for (uint32_t ti = 0;; ++ti) {
      long long t0 = 0, t1 = 0, t2 = 0;
      const bool prof = a.prof && threadIdx.x == 0;
      if (prof) t0 = clock64();
      const long long tile = next_tile(ti);
      if (tile < 0) break;
      const long long tt = tile / a.n_groups, g = tile - tt * a.n_groups;
      const int b = ti & 1;
      mbar_wait(&acc_full[b], (ti >> 1) & 1);
      if (prof) t1 = clock64();
      mbar_wait(&s_full[half], ti & 1);
      tc_fence_after();
      if (prof) t2 = clock64();
      const long long m = tt * kRows + row;
      const bool m_ok = row < kRows && m < a.n_out;
      const long long c_first = g * kTN + half * (kTN / 2);
      const long long c_rest = a.channels - c_first;
      const int c_left = c_rest > kTN ? kTN : static_cast<int>(c_rest);   // channels of this half that exist
      float* p = a.z + c_first * a.z_stride + m;
      const uint32_t taddr = tmem + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(b * kTN + half * (kTN / 2));
      uint32_t v[2][32];
      tmem_ld32(v[0], taddr);
#pragma unroll
      for (int q = 0; q < kTN / 2 / 32; ++q) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (q + 1 < kTN / 2 / 32) tmem_ld32(v[(q + 1) & 1], taddr + (q + 1) * 32);
        const uint32_t* vv = v[q & 1];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float2 acc = make_float2(__uint_as_float(vv[j]), __uint_as_float(vv[j + 1]));
          const uint32_t sp = s_addr + static_cast<uint32_t>((q * 16 + (j >> 1)) * kS * 8);
#pragma unroll
          for (int i = 0; i < kS / 2; ++i) {
            const float4 s4 = lds128(sp + i * 16);     // states 2i, 2i+1 of both channels; same address for the warp
            acc = ffma2s(make_float2(s4.x, s4.y), o_r[2 * i], acc);
            acc = ffma2s(make_float2(s4.z, s4.w), o_r[2 * i + 1], acc);
          }
          if (a.clip) { acc.x = clip_unit(acc.x); acc.y = clip_unit(acc.y); }
          if (m_ok && q * 32 + j < c_left) *p = acc.x;
          if (m_ok && q * 32 + j + 1 < c_left) p[a.z_stride] = acc.y;
          p += 2 * a.z_stride;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&acc_empty[b]); mbar_arrive(&s_free[half]); }
      if (prof) {
        const long long t3 = clock64();
        atomicAdd(a.prof + 0, static_cast<unsigned long long>(t1 - t0));   // waiting for the accumulator
        atomicAdd(a.prof + 1, static_cast<unsigned long long>(t2 - t1));   // waiting for the start states
        atomicAdd(a.prof + 2, static_cast<unsigned long long>(t3 - t2));   // drain
        atomicAdd(a.prof + 5, 1ull);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * kTN));
}

float round_tf32(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  memcpy(&v, &u, 4);
  return v;
}

using Mat = std::vector<double>;   // row-major n x n

Mat mat_mul(const Mat& x, const Mat& y, int n) {
  Mat r(static_cast<size_t>(n) * n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < n; ++k) {
      const double v = x[static_cast<size_t>(i) * n + k];
      if (v == 0.0) continue;
      for (int j = 0; j < n; ++j) r[static_cast<size_t>(i) * n + j] += v * y[static_cast<size_t>(k) * n + j];
    }
  return r;
}

std::once_flag g_pool_once;

template <int kS>
int launch(const CUtensorMap& tm_a, const CUtensorMap& tm_x, const LtiArgs& a, int grid, cudaStream_t stream) {
  constexpr size_t smem = LtiSmem<kS>::kBytes;
  DSP_CUDA(cudaFuncSetAttribute(lti_mma_kernel<kS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  lti_mma_kernel<kS><<<grid, kThreads, smem, stream>>>(tm_a, tm_x, a);
  return after_launch("lti_mma_kernel");
}

}  // namespace

// The cascade as one state-space system (A, B, C, D), states ordered by section.
// Section form (design.cuh):  q' = A_i q + [b0, b1] v ,  w = c . q + d v .
void cascade_state_space(const Section* sec, int ns, std::vector<double>& A, std::vector<double>& B,
                         std::vector<double>& C, double& D) {
  const int n = 2 * ns;
  A.assign(static_cast<size_t>(n) * n, 0.0);
  B.assign(static_cast<size_t>(n), 0.0);
  C.assign(static_cast<size_t>(n), 0.0);
  D = 1.0;
  for (int i = 0; i < ns; ++i) {
    const Section& s = sec[i];
    const double bi[2] = {s.b0, s.b1};
    for (int r = 0; r < 2; ++r) {
      double* row = &A[static_cast<size_t>(2 * i + r) * n];
      for (int j = 0; j < 2 * i; ++j) row[j] = bi[r] * C[static_cast<size_t>(j)];   // driven by the output so far
      row[2 * i] = s.a[2 * r];
      row[2 * i + 1] = s.a[2 * r + 1];
      B[static_cast<size_t>(2 * i + r)] = bi[r] * D;
    }
    for (int j = 0; j < 2 * i; ++j) C[static_cast<size_t>(j)] *= s.d;
    C[static_cast<size_t>(2 * i)] = s.c[0];
    C[static_cast<size_t>(2 * i + 1)] = s.c[1];
    D *= s.d;
  }
}

int lti_mma_build_eq(const Section* sec, int ns, LtiMmaPlan& mp) {
  mp = LtiMmaPlan{};
  if (ns < 1 || 2 * ns > kLtiMaxStates) return DSPB200_OK;
  const int n = 2 * ns;
  std::vector<double> A, B, C;
  double D;
  cascade_state_space(sec, ns, A, B, C, D);
  // A^k B and C A^k for k = 0..kRows
  std::vector<std::vector<double>> akb(kRows + 1, std::vector<double>(static_cast<size_t>(n)));
  std::vector<std::vector<double>> cak(kRows + 1, std::vector<double>(static_cast<size_t>(n)));
  akb[0] = B;
  cak[0] = C;
  for (int k = 1; k <= kRows; ++k)
    for (int i = 0; i < n; ++i) {
      long double s1 = 0.0L, s2 = 0.0L;
      for (int j = 0; j < n; ++j) {
        s1 += static_cast<long double>(A[static_cast<size_t>(i) * n + j]) * akb[k - 1][static_cast<size_t>(j)];
        s2 += static_cast<long double>(cak[k - 1][static_cast<size_t>(j)]) * A[static_cast<size_t>(j) * n + i];
      }
      akb[k][static_cast<size_t>(i)] = static_cast<double>(s1);
      cak[k][static_cast<size_t>(i)] = static_cast<double>(s2);
    }
  Mat phi(static_cast<size_t>(n) * n, 0.0);
  for (int i = 0; i < n; ++i) phi[static_cast<size_t>(i) * n + i] = 1.0;
  {
    Mat base = A;
    for (int k = kRows; k > 0; k >>= 1) {
      if (k & 1) phi = mat_mul(phi, base, n);
      base = mat_mul(base, base, n);
    }
  }
  // impulse response h[0] = D, h[m] = C A^(m-1) B
  std::vector<double> h(kRows);
  h[0] = D;
  for (int m = 1; m < kRows; ++m) {
    long double s = 0.0L;
    for (int j = 0; j < n; ++j) s += static_cast<long double>(cak[m - 1][static_cast<size_t>(j)]) * B[static_cast<size_t>(j)];
    h[m] = static_cast<double>(s);
  }
  // balance the states: unit row norms of K (the GEMM rows), so every state is computed at full relative precision
  std::vector<double> scale(static_cast<size_t>(n), 1.0);
  for (int i = 0; i < n; ++i) {
    long double ss = 0.0L;
    for (int k = 0; k < kRows; ++k) ss += static_cast<long double>(akb[k][static_cast<size_t>(i)]) * akb[k][static_cast<size_t>(i)];
    const double nrm = std::sqrt(static_cast<double>(ss));
    if (nrm > 0.0 && std::isfinite(nrm)) scale[static_cast<size_t>(i)] = 1.0 / nrm;
  }
  const int kpad = 128;
  std::vector<float> tab(static_cast<size_t>(2) * kTM * kpad, 0.f);
  auto put = [&](int r, int k, double val) {
    const float v = static_cast<float>(val);
    const float hi = round_tf32(v);
    tab[(static_cast<size_t>(0) * kTM + r) * kpad + k] = hi;
    tab[(static_cast<size_t>(1) * kTM + r) * kpad + k] = v - hi;
  };
  for (int r = 0; r < kRows; ++r)
    for (int k = 0; k <= r; ++k) put(r, k, h[r - k]);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < kRows; ++k) put(kRows + i, k, scale[static_cast<size_t>(i)] * akb[kRows - 1 - k][static_cast<size_t>(i)]);
  std::vector<float> otab(static_cast<size_t>(kTM) * kLtiMaxStates, 0.f);
  for (int r = 0; r < kRows; ++r)
    for (int i = 0; i < n; ++i) otab[static_cast<size_t>(r) * kLtiMaxStates + i] = static_cast<float>(cak[r][static_cast<size_t>(i)] / scale[static_cast<size_t>(i)]);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      mp.phi[i * kLtiMaxStates + j] = static_cast<float>(phi[static_cast<size_t>(i) * n + j] * scale[static_cast<size_t>(i)] / scale[static_cast<size_t>(j)]);
  for (float v : tab) if (!std::isfinite(v)) return DSPB200_OK;
  for (float v : otab) if (!std::isfinite(v)) return DSPB200_OK;
  const int lo0 = 0;
  DSP_CUDA(cudaMalloc(reinterpret_cast<void**>(&mp.d_table), tab.size() * sizeof(float)));
  DSP_CUDA(cudaMemcpy(mp.d_table, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice));
  DSP_CUDA(cudaMalloc(reinterpret_cast<void**>(&mp.d_otab), otab.size() * sizeof(float)));
  DSP_CUDA(cudaMemcpy(mp.d_otab, otab.data(), otab.size() * sizeof(float), cudaMemcpyHostToDevice));
  DSP_CUDA(cudaMalloc(reinterpret_cast<void**>(&mp.d_lo), sizeof(int)));
  DSP_CUDA(cudaMemcpy(mp.d_lo, &lo0, sizeof(int), cudaMemcpyHostToDevice));
  mp.period = 1;
  mp.kpad = kpad;
  mp.kvalid = kRows;
  mp.adv = kRows;
  mp.states = (n + 3) / 4 * 4;
  mp.ok = 1;
  return DSPB200_OK;
}

void lti_mma_free(LtiMmaPlan& mp) {
  cudaFree(mp.d_table);
  cudaFree(mp.d_otab);
  cudaFree(mp.d_lo);
  mp = LtiMmaPlan{};
}

bool lti_mma_usable(const LtiMmaPlan& mp, const float* x, int64_t xs, int64_t channels, int64_t n_in) {
  if (!mp.ok || reinterpret_cast<uintptr_t>(x) % 16 != 0 || xs % 4 != 0 || n_in < kRows) return false;
  if (LtiSmem<12>::kBytes + 2048 > static_cast<size_t>(max_smem_optin())) return false;   // the largest layout
  if (getenv("DSPB200_EQ_FORCE_MMA") != nullptr) return true;
  // a tile multiplies 256 channels, and the look-back walks ceil(SMs / groups) - 1 predecessors
  // at states^2 FMA each: below ~4.7k channels the FFMA scan kernel wins
  const int64_t groups = ceil_div(channels, kTN);
  if (4 * channels < 3 * groups * kTN) return false;
  return ceil_div(sm_count(), groups) <= 8;
}

int lti_mma_run(const LtiMmaPlan& mp, const float* x, int64_t xs, float* z, int64_t zs, int64_t channels,
                int64_t n_in, int64_t n_out, bool clip, cudaStream_t stream) {
  CUtensorMap tm_a, tm_x;
  memset(&tm_a, 0, sizeof(tm_a));
  memset(&tm_x, 0, sizeof(tm_x));
  DSP_TRY(encode_tmap_2d(&tm_a, DSPB200_F32, mp.d_table, static_cast<uint64_t>(mp.kpad),
                         static_cast<uint64_t>(mp.period) * 2 * kTM, static_cast<uint64_t>(mp.kpad) * sizeof(float),
                         kBK, kTM, true));
  DSP_TRY(encode_tmap_2d(&tm_x, DSPB200_F32, x, static_cast<uint64_t>(n_in), static_cast<uint64_t>(channels),
                         static_cast<uint64_t>(xs) * sizeof(float), kBK, kTN, true));
  LtiArgs a{};
  a.z = z; a.z_stride = zs; a.channels = channels; a.n_out = n_out;
  a.lo = mp.d_lo; a.period = mp.period; a.nkb = mp.kpad / kBK; a.kvalid = mp.kvalid; a.adv = mp.adv;
  a.n_tt = ceil_div(n_out, kRows);
  a.n_groups = ceil_div(channels, kTN);
  a.n_tiles = a.n_tt * a.n_groups;
  DSP_CHECK(a.n_tiles < (1ll << 31) && a.n_tt < (1ll << 29), "shape too large for the tensor-core EQ kernel");
  a.o_tab = mp.d_otab;
  a.clip = clip ? 1 : 0;
  memcpy(a.phi, mp.phi, sizeof(a.phi));
  const int64_t sms = sm_count();
  const int grid = static_cast<int>(a.n_tiles < sms ? a.n_tiles : sms);
  // stream-ordered scratch: one state record and flag per tile (never reused within a launch, so a
  // slow reader can not be overtaken by a writer), and the tile counter
  std::call_once(g_pool_once, [] {
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  });
  const size_t slots = static_cast<size_t>(a.n_tiles);
  const size_t flag_bytes = static_cast<size_t>(round_up(static_cast<int64_t>((2 * slots + 1) * sizeof(unsigned)), 256));
  const size_t ring_bytes = slots * 2 * mp.states * kTN * sizeof(float);
  unsigned char* scratch = nullptr;
  if (cudaMallocAsync(reinterpret_cast<void**>(&scratch), flag_bytes + ring_bytes, stream) != cudaSuccess) {
    cudaGetLastError();
    return kLtiNoScratch;   // the caller runs the scan kernel instead
  }
  DSP_CUDA(cudaMemsetAsync(scratch, 0, flag_bytes, stream));
  a.flags = reinterpret_cast<unsigned*>(scratch);
  a.counter = a.flags + 2 * slots;
  a.ring = reinterpret_cast<float*>(scratch + flag_bytes);
  unsigned long long* prof = nullptr;
  if (getenv("DSPB200_LTI_PROF") != nullptr) {
    cudaMalloc(reinterpret_cast<void**>(&prof), 8 * sizeof(unsigned long long));
    cudaMemset(prof, 0, 8 * sizeof(unsigned long long));
  }
  a.prof = prof;
  int rc;
  switch (mp.states) {
    case 4: rc = launch<4>(tm_a, tm_x, a, grid, stream); break;
    case 8: rc = launch<8>(tm_a, tm_x, a, grid, stream); break;
    case 12: rc = launch<12>(tm_a, tm_x, a, grid, stream); break;
    case 16: rc = launch<16>(tm_a, tm_x, a, grid, stream); break;
    default: rc = fail(DSPB200_ERR_INVALID, "internal: bad state count %d", mp.states);
  }
  cudaFreeAsync(scratch, stream);
  if (prof) {
    unsigned long long h[8];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(prof);
    const double nt = h[5] ? static_cast<double>(h[5]) : 1.0;
    fprintf(stderr, "lti_mma: %llu tiles, %d CTAs; cycles per tile: drain warps wait acc %.0f, wait state %.0f, drain %.0f; state warp %.0f; direct %.1f%%\n",
            h[5], grid, h[0] / nt, h[1] / nt, h[2] / nt, h[3] / nt, 100.0 * h[4] / nt);
  }
  return rc;
}

}  // namespace dspb200
