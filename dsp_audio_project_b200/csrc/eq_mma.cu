// K2 on the tensor cores: the biquad cascade as a chunked linear system.
//
// The whole cascade of sistema_ecualizador (dsp_core.py:216-254; up to 8 second-order sections)
// is ONE linear time-invariant system with S = 2 * sections states.  Over a chunk of 96 samples
//     z_k = T x_k + O s_k ,      s_(k+1) = Phi s_k + K x_k
// with T [96 x 96] the lower-triangular Toeplitz matrix of the cascade's impulse response,
// K [S x 96] the state reached from a zero start, O [96 x S] the free response and Phi = A^96.
// Both products are GEMMs on tcgen05.mma (kind::tf32, fp32 accumulators in TMEM).  One tile is
//     D[128 channels x 112] = X[128 x 96] . [T; K]^T  +  S[128 x 4S] . O'^T
// * first term: the three-product TF32 split of src_mma.cu (X_hi [T;K]_hi + X_hi [T;K]_lo +
//   X_lo [T;K]_hi; X_lo is formed in shared memory, in place, by the converter warps);
// * second term: the start state of the chunk, split into three TF32 pieces by the thread that
//   owns the channel and stored to tensor memory, is the A operand read from TMEM;
//   [s1 | s2 | s3 | s1] . [O_hi | O_hi | O_hi | O_lo]^T is exact to fp32 rounding.
// TMEM lane c then holds, for channel c, the chunk's 96 finished outputs in columns 0..95 and the
// zero-state end state u_k in columns 96..96+S-1.  The thread that owns the lane owns the
// channel: it keeps s_k in registers, advances s_(k+1) = Phi s_k + u_k (S^2 FMA per channel and
// chunk is all that is left of the recurrence on the FMA pipe), hands its split to the MMA warp,
// clips once (dsp_core.py:254) and lays 128-byte runs of its channel into a swizzled staging tile that
// leaves as TMA stores of [128 channels x 32 samples] (direct 32-byte vector stores from 32 channels per
// instruction ran at a third of the speed; a shuffle transpose at a fifth).  A CTA walks a group of 128
// channels through time, so the state stays in registers; the form is used when there are enough
// channel groups to fill the GPU (about 15k channels), narrower batches stay on the scan kernel.
// With more groups than SMs the time axis is cut into slices, handed out by an atomic counter, so that
// the last round is full; a later slice picks up the end state its predecessor left in global memory.
//
// Warp roles (one persistent CTA per SM): warps 0-3 epilogue (one TMEM lane quarter each), warp 4
// TMA producer (the coefficient tiles once -- they stay resident in shared memory -- then x tiles
// through a 7-deep ring), warp 5 MMA issuer, warps 6-7 converters.  Four accumulators let the MMAs run up to
// three chunks ahead of the epilogue; the only serial link per chunk is
// accumulator -> state update -> tcgen05.st -> S/2 free-response MMAs of the next chunk.
//
// Non-finite inputs poison their whole 96-sample chunk (0 * NaN inside the GEMM), not only the
// samples after them as the sequential reference does.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "design.cuh"
#include "internal.cuh"
#include "umma.cuh"

namespace dspb200 {

namespace {

constexpr int kTM = 128;        // channels per tile (MMA M, TMEM lanes)
constexpr int kRows = 96;       // samples per chunk (three 32-sample k-blocks, three 128-byte lines per channel)
constexpr int kTNn = kRows + kLtiMaxStates;   // 112 coefficient rows (MMA N): outputs + end states
constexpr int kBK = 32;         // k-values per stage (one 128-byte swizzle row)
constexpr int kNkb = kRows / kBK;
constexpr int kXSlots = 7;
constexpr int kStages = 1;       // staging tiles of the TMA stores
constexpr int kJobQ = 16;        // job queue depth; no role runs more than ~7 items ahead of the epilogue
constexpr int kAccs = 4;
constexpr int kSCol = kAccs * kTNn;           // TMEM columns 448..511: the split start states (A operand of the correction)
constexpr int kEpiWarps = 4;
constexpr int kConvWarps = 2;      // 8 warps in all: registers are allotted in groups of four warps, 10 warps would cap the kernel at 168
constexpr int kTmaWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1, kConvWarp0 = kEpiWarps + 2;
constexpr int kThreads = (kConvWarp0 + kConvWarps) * 32;
constexpr uint32_t kXBytes = kTM * kBK * 4;      // 16 KB: one k-block of x
// resident coefficient tiles, 128 bytes (32 k-values) per row.  T is lower triangular: the tile of k-block kb only
// holds rows 32 kb .. 111 (112, 80, 48 rows), hi and lo; then two k-blocks of the free-response operand, 96 rows each
constexpr uint32_t tab_rows(int kb) { return kTNn - kb * kBK; }
constexpr uint32_t tab_off(int kb, int hl) {
  uint32_t off = 0;
  for (int j = 0; j < kb; ++j) off += 2 * tab_rows(j) * 128;
  return off + hl * tab_rows(kb) * 128;
}
constexpr uint32_t kOOff = tab_off(kNkb, 0);
constexpr uint32_t kOBytes = kRows * 128;
constexpr uint32_t kTabTotal = kOOff + 2 * kOBytes;   // 84 KB
constexpr size_t kSmemBytes = static_cast<size_t>(kTabTotal) + (kXSlots + kStages) * kXBytes + 1024;

struct LtiArgs {
  float* z; long long z_stride;
  long long channels, n_out;
  long long n_tt, n_groups;
  // work items: (channel group, time slice) pairs, slice-major, handed out by an atomic counter so that a slice only
  // ever waits for a lower-numbered item, which a running CTA already owns
  long long n_jobs, chunks_per_job;
  unsigned* counter;            // job dispenser
  unsigned* done;               // [n_jobs] 1 once the item's end state is in `xfer`
  float* xfer;                  // [n_jobs][kS][128] end states handed to the group's next slice
  float* state;                 // streaming form: [channels][16] state after the block (NULL: not kept)
  int state_in;                 // streaming form: start from `state` instead of zero
  int warm;                     // > 0: slices are independent; a later slice starts `warm` chunks early from a zero
                                // state and only stores from its own first chunk on (no hand-over between slices)
  int clip;
  unsigned long long* prof;     // development: cycles per epilogue phase (NULL = off)
  float phi[kLtiMaxStates * kLtiMaxStates];
};

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory"); }

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
               const __grid_constant__ CUtensorMap tm_z, const __grid_constant__ LtiArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));   // swizzle atoms: 1024-byte aligned
  __shared__ __align__(8) uint64_t bars[4 * kXSlots + 2 * kAccs + 2 + kJobQ];
  __shared__ long long job_q[kJobQ];
  __shared__ uint32_t tmem_base_s;
  uint64_t* full_x = bars;                    // [x slot] TMA landed the x tile
  uint64_t* mid = full_x + kXSlots;           // [x slot] the products with the raw x tile have completed
  uint64_t* conv = mid + kXSlots;             // [x slot] x replaced in place by x - trunc(x)
  uint64_t* empty_x = conv + kXSlots;         // [x slot] the product with X_lo has completed
  uint64_t* acc_full = empty_x + kXSlots;     // [acc] accumulator complete (free response included)
  uint64_t* acc_empty = acc_full + kAccs;     // [acc] accumulator drained by the epilogue warps
  uint64_t* tab_full = acc_empty + kAccs;     // coefficient tiles resident
  uint64_t* s_ready = tab_full + 1;           // the next chunk's start states are in tensor memory
  uint64_t* job_full = s_ready + 1;           // [queue slot] the dispenser has filled it
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kXSlots; ++s) {
      mbar_init(&full_x[s], 1); mbar_init(&mid[s], 1); mbar_init(&conv[s], kConvWarps); mbar_init(&empty_x[s], 1);
    }
    for (int b = 0; b < kAccs; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], kEpiWarps); }
    mbar_init(tab_full, 1);
    mbar_init(s_ready, kEpiWarps);
    for (int q = 0; q < kJobQ; ++q) mbar_init(&job_full[q], 1);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(&tmem_base_s))), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  unsigned char* xring = smem + kTabTotal;
  auto x_ptr = [&](int s) -> unsigned char* { return xring + static_cast<size_t>(s) * kXBytes; };
  unsigned char* stage0 = x_ptr(kXSlots);                                         // kStages x [128 channels][32 samples], swizzled
  auto tab_ptr = [&](int kb, int hl) -> unsigned char* { return smem + tab_off(kb, hl); };   // first row held: 32 kb
  auto o_ptr = [&](int kb) -> unsigned char* { return smem + kOOff + kb * kOBytes; };
  // the i-th work item of this CTA (-1: none left); group, slice and chunk range of an item
  auto next_job = [&](uint32_t i) -> long long {
    mbar_wait(&job_full[i % kJobQ], (i / kJobQ) & 1);
    return *reinterpret_cast<volatile long long*>(&job_q[i % kJobQ]);
  };
  struct Job { int g, h, tw, t0, t1; };   // chunks [tw, t0) are the warm-up of an overlapping slice (tw == t0 otherwise)
  auto job_of = [&](long long j) -> Job {
    Job r;
    const int jj = static_cast<int>(j), ng = static_cast<int>(a.n_groups), cpj = static_cast<int>(a.chunks_per_job);
    r.h = jj / ng;
    r.g = jj - r.h * ng;
    r.t0 = r.h * cpj;
    r.t1 = r.t0 + cpj < static_cast<int>(a.n_tt) ? r.t0 + cpj : static_cast<int>(a.n_tt);
    r.tw = a.warm > 0 ? (r.t0 > a.warm ? r.t0 - a.warm : 0) : r.t0;
    return r;
  };

  if (warp == kTmaWarp) {
    // ---------------- TMA producer: coefficient tiles once, then x tiles ----------------
    if (lane == 0) {
      tma_prefetch_desc(&tm_a);
      tma_prefetch_desc(&tm_x);
      mbar_expect_tx(tab_full, kTabTotal);
      for (int kb = 0; kb < kNkb; ++kb)
        for (int hl = 0; hl < 2; ++hl)
          for (int r = kb * kBK; r < kTNn; r += 16)         // boxes of 16 rows
            tma_load_2d(tab_ptr(kb, hl) + (r - kb * kBK) * 128, &tm_a, kb * kBK, hl * kTNn + r, tab_full);
      for (int kb = 0; kb < 2; ++kb)
        for (int r = 0; r < kRows; r += 16) tma_load_2d(o_ptr(kb) + r * 128, &tm_a, kb * kBK, 2 * kTNn + r, tab_full);
      uint32_t it = 0;
      for (uint32_t ji = 0;; ++ji) {
        long long j = static_cast<long long>(atomicAdd(a.counter, 1u));
        if (j >= a.n_jobs) j = -1;
        *reinterpret_cast<volatile long long*>(&job_q[ji % kJobQ]) = j;
        mbar_arrive(&job_full[ji % kJobQ]);
        if (j < 0) break;
        const Job jb = job_of(j);
        for (int tt = jb.tw; tt < jb.t1; ++tt)
          for (int kb = 0; kb < kNkb; ++kb, ++it) {
            const int s = it % kXSlots;
            if (it >= kXSlots) mbar_wait(&empty_x[s], ((it / kXSlots) - 1) & 1);
            mbar_expect_tx(&full_x[s], kXBytes);
            tma_load_2d(x_ptr(s), &tm_x, tt * kRows + kb * kBK, jb.g * kTM, &full_x[s]);
          }
      }
    }
  } else if (warp == kMmaWarp) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      // D fp32, A/B tf32, both K-major, M = 128 channels, N = 112 coefficient rows
      auto idesc_n = [](int n) -> uint32_t {
        return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(kTM >> 4) << 24);
      };
      // T is lower triangular: k-block kb (inputs 32 kb ..) only reaches outputs 32 kb .. and the state rows, so its
      // MMAs skip the first 32 kb coefficient rows (N = 112, 80, 48; accumulator columns and B rows shifted alike)
      uint32_t n_corr = 0;
      // third product of k-block `j` (deferred one step so the in-place conversion of its x tile overlaps the
      // first two products of the next k-block); after a chunk's last one, the free response of its start state
      auto finish = [&](uint32_t j, int kb, uint32_t d, bool last_of_tile, int b, bool corr) {
        const int sx = j % kXSlots;
        mbar_wait(&conv[sx], (j / kXSlots) & 1);
        tc_fence_after();
        const uint64_t dxl = umma_desc_sw128(x_ptr(sx)), dth = umma_desc_sw128(tab_ptr(kb, 0));
        const uint32_t idk = idesc_n(kTNn - kb * kBK);
#pragma unroll
        for (int k = 0; k < kBK / 8; ++k) umma_tf32(d + kb * kBK, dxl + 2 * k, dth + 2 * k, idk, 1u);
        umma_commit(&empty_x[sx]);
        if (last_of_tile) {
          if (corr) {
            // z += [s1 | s2 | s3 | s1] . [O_hi | O_hi | O_hi | O_lo]^T with the start state split in three TF32 pieces
            mbar_wait(s_ready, n_corr & 1);
            ++n_corr;
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < kS / 2; ++k)
              umma_tf32_ts(d, tmem + kSCol + 8 * k, umma_desc_sw128(o_ptr((8 * k) / kBK)) + 2 * (((8 * k) % kBK) / 8), idesc_n(kRows), 1u);   // the end-state columns take no free response
          }
          umma_commit(&acc_full[b]);
        }
      };
      mbar_wait(tab_full, 0);
      uint32_t it = 0, ti = 0;
      uint32_t prev_d = 0;
      int prev_b = 0, prev_kb = 0;
      bool have_prev = false, prev_last = false, prev_corr = false;
      for (uint32_t ji = 0;; ++ji) {
        const long long j = next_job(ji);
        if (j < 0) break;
        const Job jb = job_of(j);
        for (int tt = jb.tw; tt < jb.t1; ++tt, ++ti) {
          const int b = ti % kAccs;
          if (ti >= kAccs) mbar_wait(&acc_empty[b], ((ti / kAccs) - 1) & 1);
          tc_fence_after();
          const uint32_t d = tmem + b * kTNn;
          for (int kb = 0; kb < kNkb; ++kb, ++it) {
            const int sx = it % kXSlots;
            mbar_wait(&full_x[sx], (it / kXSlots) & 1);
            tc_fence_after();
            const uint64_t dx = umma_desc_sw128(x_ptr(sx));
            const uint64_t dth = umma_desc_sw128(tab_ptr(kb, 0));
            const uint64_t dtl = umma_desc_sw128(tab_ptr(kb, 1));
            const uint32_t idk = idesc_n(kTNn - kb * kBK);
#pragma unroll
            for (int k = 0; k < kBK / 8; ++k) {
              umma_tf32(d + kb * kBK, dx + 2 * k, dth + 2 * k, idk, (kb | k) ? 1u : 0u);
              umma_tf32(d + kb * kBK, dx + 2 * k, dtl + 2 * k, idk, 1u);
            }
            umma_commit(&mid[sx]);
            if (have_prev) finish(it - 1, prev_kb, prev_d, prev_last, prev_b, prev_corr);
            have_prev = true; prev_d = d; prev_b = b; prev_kb = kb; prev_last = (kb == kNkb - 1); prev_corr = tt > 0 || a.state_in != 0;   // chunk 0 starts from a zero state unless one was passed in
          }
        }
      }
      if (have_prev) finish(it - 1, prev_kb, prev_d, prev_last, prev_b, prev_corr);
    }
  } else if (warp >= kConvWarp0) {
    // ---------------- converters: x -> x - trunc_tf32(x) in place, once the raw tile has been consumed ----------------
    const int ctid = threadIdx.x - kConvWarp0 * 32;
    uint32_t it = 0;
    for (uint32_t ji = 0;; ++ji) {
      const long long j = next_job(ji);
      if (j < 0) break;
      const Job jb = job_of(j);
      for (int n_kb = (jb.t1 - jb.tw) * kNkb; n_kb > 0; --n_kb, ++it) {
          const int s = it % kXSlots;
          mbar_wait(&mid[s], (it / kXSlots) & 1);
          float4* buf = reinterpret_cast<float4*>(x_ptr(s)) + ctid;
          constexpr int kPer = static_cast<int>(kXBytes / 16) / (32 * kConvWarps);   // 16-byte pieces per thread
          float4 v[kPer];
#pragma unroll
          for (int i = 0; i < kPer; ++i) v[i] = buf[i * 32 * kConvWarps];
#pragma unroll
          for (int i = 0; i < kPer; ++i) {
            float4 r;
            r.x = v[i].x - trunc_tf32(v[i].x);
            r.y = v[i].y - trunc_tf32(v[i].y);
            r.z = v[i].z - trunc_tf32(v[i].z);
            r.w = v[i].w - trunc_tf32(v[i].w);
            buf[i * 32 * kConvWarps] = r;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&conv[s]);
      }
    }
  } else {
    // ---------------- epilogue warps 0-3: thread = TMEM lane = channel ----------------
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    constexpr int kBlocks = kRows / 16;
    uint32_t ti = 0, n_st = 0;
    // the start state of the coming chunk, split in three TF32 pieces, to tensor memory: [s1 | s2 | s3 | s1]
    auto hand_over = [&](const float (&s)[kS]) {
#pragma unroll
      for (int q = 0; q < kS / 4; ++q) {
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int col = 16 * q + j, i = col % kS, piece = col / kS;
          const float s1 = trunc_tf32(s[i]);
          const float r1 = s[i] - s1;
          const float s2 = trunc_tf32(r1);
          w[j] = __float_as_uint(piece == 1 ? s2 : (piece == 2 ? r1 - s2 : s1));
        }
        tmem_st16(lane_base + kSCol + 16 * q, w);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_ready);
    };
    for (uint32_t ji = 0;; ++ji) {
      const long long job = next_job(ji);
      if (job < 0) break;
      const Job jb = job_of(job);
      const int g = jb.g;
      float s[kS];
      const long long chan = static_cast<long long>(g) * kTM + warp * 32 + lane;
      if (jb.tw == 0 && !a.state_in) {
#pragma unroll
        for (int i = 0; i < kS; ++i) s[i] = 0.f;     // zero initial state per channel (lfilter, dsp_core.py:214)
      } else if (a.warm > 0 && jb.tw > 0) {
        // overlapping slice: zero state `warm` chunks before its first own chunk; what the true state would add has
        // decayed below float32 resolution by then (LtiMmaPlan::warm_chunks)
#pragma unroll
        for (int i = 0; i < kS; ++i) s[i] = 0.f;
        hand_over(s);
      } else if (jb.tw == 0) {
        // streaming form: the state the previous block of these channels ended in
#pragma unroll
        for (int i = 0; i < kS; ++i) s[i] = chan < a.channels ? __ldcg(a.state + chan * kLtiMaxStates + i) : 0.f;
        hand_over(s);
      } else {
        // later slice of the group: its start state is the end state of the slice before, finished rounds ago
        const long long pj = job - a.n_groups;
        while (ld_acquire(a.done + pj) == 0u) {
        }
        const float* r = a.xfer + static_cast<size_t>(pj) * (kS * kTM) + warp * 32 + lane;
#pragma unroll
        for (int i = 0; i < kS; ++i) s[i] = __ldcg(r + i * kTM);
        hand_over(s);
      }
      for (int tt = jb.tw; tt < jb.t1; ++tt, ++ti) {
        const int b = ti % kAccs;
        long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
        const bool prof = a.prof && threadIdx.x == 0;
        if (prof) t0 = clock64();
        mbar_wait(&acc_full[b], (ti / kAccs) & 1);
        tc_fence_after();
        if (prof) t1 = clock64();
        // the whole lane (96 outputs + end state) in one round trip to tensor memory; the accumulator is free again
        const uint32_t taddr = lane_base + static_cast<uint32_t>(b * kTNn);
        uint32_t uu[16], v[kBlocks][16];
        tmem_ld16(uu, taddr + kRows);
#pragma unroll
        for (int blk = 0; blk < kBlocks; ++blk) tmem_ld16(v[blk], taddr + blk * 16);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[b]);
        if (prof) t2 = clock64();
        // the next chunk's start state first: the MMA warp is waiting for it
        {
          float u[kS];
#pragma unroll
          for (int i = 0; i < kS; ++i) u[i] = __uint_as_float(uu[i]);
          advance_state<kS>(s, a.phi, u);
        }
        if (tt + 1 < jb.t1) hand_over(s);
        if (prof) t3 = clock64();
        // clip; each thread lays its channel's 32-sample runs into a 128-byte-swizzled staging tile, which goes out
        // as one TMA store of [128 channels x 32 samples] (full lines, clipped at the tensor's edges by the hardware)
        const int r = warp * 32 + lane;
        if (tt >= jb.t0)                                      // warm-up chunks of an overlapping slice store nothing
#pragma unroll
        for (int blk = 0; blk < kRows / 32; ++blk, ++n_st) {
          unsigned char* stage = stage0 + static_cast<size_t>(n_st % kStages) * kXBytes;
          const uint32_t row_addr = static_cast<uint32_t>(__cvta_generic_to_shared(stage)) + static_cast<uint32_t>(r) * 128u;
          if (threadIdx.x == 0)                               // the store that last used this staging tile has read it
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kStages - 1) : "memory");
          epi_bar();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              o[q] = __uint_as_float(v[2 * blk + (j >> 2)][(j & 3) * 4 + q]);
              if (a.clip) o[q] = clip_unit(o[q]);
            }
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};"
                         ::"r"(row_addr + static_cast<uint32_t>((j ^ (r & 7)) << 4)), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]) : "memory");
          }
          fence_proxy_async();
          epi_bar();
          if (threadIdx.x == 0) {
            tma_store_2d(&tm_z, stage, tt * kRows + blk * 32, g * kTM);
            tma_store_commit();
          }
        }
        if (prof) {
          atomicAdd(a.prof + 0, static_cast<unsigned long long>(t1 - t0));           // waiting for the accumulator
          atomicAdd(a.prof + 1, static_cast<unsigned long long>(t2 - t1));           // tensor memory -> registers
          atomicAdd(a.prof + 3, static_cast<unsigned long long>(t3 - t2));           // state hand-over
          atomicAdd(a.prof + 4, static_cast<unsigned long long>(clock64() - t3));    // clip + stores
          atomicAdd(a.prof + 2, 1ull);
        }
      }
      if (jb.t1 == static_cast<int>(a.n_tt) && a.state != nullptr && chan < a.channels) {
#pragma unroll
        for (int i = 0; i < kS; ++i) a.state[chan * kLtiMaxStates + i] = s[i];
      }
      if (jb.t1 < static_cast<int>(a.n_tt) && a.warm == 0) {
        // the group goes on in a later slice: leave the end state where that slice will pick it up
        float* r = a.xfer + static_cast<size_t>(job) * (kS * kTM) + warp * 32 + lane;
#pragma unroll
        for (int i = 0; i < kS; ++i) __stcg(r + i * kTM, s[i]);
        __threadfence();
        epi_bar();
        if (threadIdx.x == 0) st_release(a.done + job, 1u);
      }
    }
  }
  if (threadIdx.x == 0) tma_store_wait_all0();
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
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
int launch(const CUtensorMap& tm_a, const CUtensorMap& tm_x, const CUtensorMap& tm_z, const LtiArgs& a, int grid, cudaStream_t stream) {
  constexpr size_t smem = kSmemBytes;
  DSP_CUDA(cudaFuncSetAttribute(lti_mma_kernel<kS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  lti_mma_kernel<kS><<<grid, kThreads, smem, stream>>>(tm_a, tm_x, tm_z, a);
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

// The chunk system of a cascade in float64 (host only): z = T x + O s, s' = Phi s + K x over `rows` samples (0: the
// tensor-core EQ's 96), states
// rescaled to unit row norms of K so that every state row of the GEMM is computed at full relative precision.
int lti_chunk_system(const Section* sec, int ns, LtiChunkSystem& cs, int rows) {
  cs = LtiChunkSystem{};
  const int kRows = rows > 0 ? rows : dspb200::kRows;   // shadows the tensor-core EQ's chunk length
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
  std::vector<double> scale(static_cast<size_t>(n), 1.0);
  for (int i = 0; i < n; ++i) {
    long double ss = 0.0L;
    for (int k = 0; k < kRows; ++k) ss += static_cast<long double>(akb[k][static_cast<size_t>(i)]) * akb[k][static_cast<size_t>(i)];
    const double nrm = std::sqrt(static_cast<double>(ss));
    if (nrm > 0.0 && std::isfinite(nrm)) scale[static_cast<size_t>(i)] = 1.0 / nrm;
  }
  cs.rows = kRows;
  cs.states = n;
  cs.tk.assign(static_cast<size_t>(kRows + kLtiMaxStates) * kRows, 0.0);
  cs.o.assign(static_cast<size_t>(kRows) * kLtiMaxStates, 0.0);
  cs.phi.assign(static_cast<size_t>(kLtiMaxStates) * kLtiMaxStates, 0.0);
  for (int r = 0; r < kRows; ++r)
    for (int k = 0; k <= r; ++k) cs.tk[static_cast<size_t>(r) * kRows + k] = h[r - k];
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < kRows; ++k)
      cs.tk[static_cast<size_t>(kRows + i) * kRows + k] = scale[static_cast<size_t>(i)] * akb[kRows - 1 - k][static_cast<size_t>(i)];
  for (int r = 0; r < kRows; ++r)
    for (int i = 0; i < n; ++i) cs.o[static_cast<size_t>(r) * kLtiMaxStates + i] = cak[r][static_cast<size_t>(i)] / scale[static_cast<size_t>(i)];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      cs.phi[static_cast<size_t>(i) * kLtiMaxStates + j] = phi[static_cast<size_t>(i) * n + j] * scale[static_cast<size_t>(i)] / scale[static_cast<size_t>(j)];
  return DSPB200_OK;
}

// How many chunks after a start from a ZERO state the output is within 2^-24 of max|x| of the true one (float64,
// host).  With |x| <= 1 the true state is bounded elementwise by b = sum_k |Phi^k K| 1; what a wrong start state e
// still contributes w chunks later is O Phi^w e, so the first w with max_r (|O| |Phi^w| b)_r <= 2^-24 is enough
// (elementwise bounds: conservative by a small factor, i.e. a few hundred samples).  0: no such w within 4096 chunks.
int lti_warm_chunks(const LtiChunkSystem& cs) {
  const int n = cs.states, rows = cs.rows, ld = kLtiMaxStates;
  if (n < 1 || rows < 1) return 0;
  auto phi_at = [&](int i, int j) { return cs.phi[static_cast<size_t>(i) * ld + j]; };
  std::vector<double> m(static_cast<size_t>(n) * rows), m2(static_cast<size_t>(n) * rows), b(static_cast<size_t>(n), 0.0);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < rows; ++k) m[static_cast<size_t>(i) * rows + k] = cs.tk[static_cast<size_t>(rows + i) * rows + k];
  bool converged = false;
  for (int it = 0; it < 8192 && !converged; ++it) {
    double rmax = 0.0, bmax = 0.0;
    for (int i = 0; i < n; ++i) {
      double r = 0.0;
      for (int k = 0; k < rows; ++k) r += std::fabs(m[static_cast<size_t>(i) * rows + k]);
      b[static_cast<size_t>(i)] += r;
      rmax = r > rmax ? r : rmax;
      bmax = b[static_cast<size_t>(i)] > bmax ? b[static_cast<size_t>(i)] : bmax;
    }
    if (!(rmax == rmax) || !std::isfinite(bmax)) return 0;
    if (rmax <= 1e-13 * bmax) { converged = true; break; }
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < rows; ++k) {
        double acc = 0.0;
        for (int j = 0; j < n; ++j) acc += phi_at(i, j) * m[static_cast<size_t>(j) * rows + k];
        m2[static_cast<size_t>(i) * rows + k] = acc;
      }
    m.swap(m2);
  }
  if (!converged) return 0;
  Mat p(static_cast<size_t>(n) * n, 0.0), phi(static_cast<size_t>(n) * n);
  for (int i = 0; i < n; ++i) {
    p[static_cast<size_t>(i) * n + i] = 1.0;
    for (int j = 0; j < n; ++j) phi[static_cast<size_t>(i) * n + j] = phi_at(i, j);
  }
  const double tol = 1.0 / 16777216.0;
  for (int w = 0; w <= 4096; ++w) {
    double err = 0.0;
    for (int r = 0; r < rows; ++r) {
      double acc = 0.0;
      for (int i = 0; i < n; ++i) {
        double e = 0.0;
        for (int j = 0; j < n; ++j) e += std::fabs(p[static_cast<size_t>(i) * n + j]) * b[static_cast<size_t>(j)];
        acc += std::fabs(cs.o[static_cast<size_t>(r) * ld + i]) * e;
      }
      err = acc > err ? acc : err;
    }
    if (err <= tol) return w < 1 ? 1 : w;
    p = mat_mul(phi, p, n);
  }
  return 0;
}

int lti_mma_build_eq(const Section* sec, int ns, LtiMmaPlan& mp) {
  mp = LtiMmaPlan{};
  LtiChunkSystem cs;
  DSP_TRY(lti_chunk_system(sec, ns, cs));
  if (cs.rows != kRows) return DSPB200_OK;
  const int n = cs.states;
  const int kpad = kRows;
  const int ks = (n + 3) / 4 * 4;                     // states padded to a multiple of 4
  // three sections of kTNn rows: [T; K] rounded to TF32, its remainder, and the free-response operand
  std::vector<float> tab(static_cast<size_t>(3) * kTNn * kpad, 0.f);
  for (int r = 0; r < kRows + n; ++r)
    for (int k = 0; k < kRows; ++k) {
      const float v = static_cast<float>(cs.tk[static_cast<size_t>(r) * kRows + k]);
      const float hi = round_tf32(v);
      tab[(static_cast<size_t>(0) * kTNn + r) * kpad + k] = hi;
      tab[(static_cast<size_t>(1) * kTNn + r) * kpad + k] = v - hi;
    }
  // z += [s1 | s2 | s3 | s1] . [O_hi | O_hi | O_hi | O_lo]^T  (s = s1 + s2 + s3 in TF32 pieces, O = C A^r in the scaled basis);
  // rows kRows.. stay zero: the end-state columns of the accumulator take no free response
  for (int r = 0; r < kRows; ++r)
    for (int i = 0; i < n; ++i) {
      const float v = static_cast<float>(cs.o[static_cast<size_t>(r) * kLtiMaxStates + i]);
      const float hi = round_tf32(v);
      float* row = &tab[(static_cast<size_t>(2) * kTNn + r) * kpad];
      row[i] = hi; row[ks + i] = hi; row[2 * ks + i] = hi; row[3 * ks + i] = v - hi;
    }
  for (int i = 0; i < kLtiMaxStates * kLtiMaxStates; ++i) mp.phi[i] = static_cast<float>(cs.phi[static_cast<size_t>(i)]);
  for (float v : tab) if (!std::isfinite(v)) return DSPB200_OK;
  for (float v : mp.phi) if (!std::isfinite(v)) return DSPB200_OK;
  DSP_CUDA(cudaMalloc(reinterpret_cast<void**>(&mp.d_table), tab.size() * sizeof(float)));
  DSP_CUDA(cudaMemcpy(mp.d_table, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice));
  mp.kpad = kpad;
  mp.states = ks;
  mp.ok = 1;
  mp.warm_chunks = lti_warm_chunks(cs);
  return DSPB200_OK;
}

void lti_mma_free(LtiMmaPlan& mp) {
  cudaFree(mp.d_table);
  mp = LtiMmaPlan{};
}

bool lti_mma_possible(const LtiMmaPlan& mp, const float* x, int64_t xs, const float* z, int64_t zs) {
  return mp.ok && reinterpret_cast<uintptr_t>(x) % 16 == 0 && xs % 4 == 0 && reinterpret_cast<uintptr_t>(z) % 16 == 0 &&
         zs % 4 == 0 && mp.kpad == kNkb * kBK && kSmemBytes + 2048 <= static_cast<size_t>(max_smem_optin());
}

int lti_mma_chunk() { return kRows; }

// How a batch is cut into work items (channel group x time slice) and what share of the SMs' time does useful work.
//  * more groups than SMs: exact slices, a later slice picks up the end state its predecessor left (see the kernel);
//  * fewer groups than SMs and `overlap` allowed: independent slices that start plan.warm_chunks early from a zero
//    state, so a narrow batch (C2's 1024 channels: 8 groups) still fills the GPU, at the price of the warm-up.
struct LtiSlicing { int64_t slices = 1; int warm = 0; double eff = 0.0; };
static LtiSlicing lti_slicing(const LtiMmaPlan& mp, int64_t channels, int64_t n_out, bool overlap) {
  const int64_t sms = sm_count(), n_tt = ceil_div(n_out, kRows), groups = ceil_div(channels, kTM);
  LtiSlicing r;
  if (groups > sms) {
    // with at least one group per SM a slice's predecessor ran a whole round earlier, so its end state is waiting in
    // memory; the smallest count that fills >= 95 % of the last round (or the best one up to 16)
    for (int64_t h = 1; h <= 16 && ceil_div(n_tt, h) >= 8; ++h) {
      const int64_t jobs = groups * ceil_div(n_tt, ceil_div(n_tt, h));
      const double eff = static_cast<double>(jobs) / static_cast<double>(ceil_div(jobs, sms) * sms);
      if (eff > r.eff + 1e-9) { r.eff = eff; r.slices = h; }
      if (eff >= 0.95) break;
    }
    return r;
  }
  r.eff = static_cast<double>(groups) / static_cast<double>(sms);
  if (!overlap || mp.warm_chunks <= 0 || getenv("DSPB200_EQ_NO_OVERLAP") != nullptr) return r;
  const int64_t w = mp.warm_chunks;
  for (int64_t h = 2; h <= 256 && ceil_div(n_tt, h) >= 8; ++h) {
    const int64_t cpj = ceil_div(n_tt, h), hh = ceil_div(n_tt, cpj), jobs = groups * hh;
    const int64_t rounds = ceil_div(jobs, sms);
    // a round lasts as long as its longest item: cpj + w chunks; ideal: groups * n_tt chunks spread over all SMs
    const double eff = static_cast<double>(groups * n_tt) / (static_cast<double>(sms * rounds) * static_cast<double>(cpj + w));
    if (eff > r.eff + 1e-9) { r.eff = eff; r.slices = hh; r.warm = static_cast<int>(w); }
  }
  return r;
}

static bool buffers_overlap(const float* x, int64_t xs, const float* z, int64_t zs, int64_t channels) {
  if (x == nullptr || z == nullptr) return false;
  const uintptr_t x0 = reinterpret_cast<uintptr_t>(x), x1 = x0 + static_cast<uintptr_t>(channels * xs) * 4;
  const uintptr_t z0 = reinterpret_cast<uintptr_t>(z), z1 = z0 + static_cast<uintptr_t>(channels * zs) * 4;
  return x0 < z1 && z0 < x1;
}

bool lti_mma_usable(const LtiMmaPlan& mp, const float* x, int64_t xs, const float* z, int64_t zs, int64_t channels,
                    int64_t n_in) {
  if (!mp.ok || reinterpret_cast<uintptr_t>(x) % 16 != 0 || xs % 4 != 0 || n_in < kRows) return false;
  if (reinterpret_cast<uintptr_t>(z) % 16 != 0 || zs % 4 != 0) return false;
  if (mp.kpad != kNkb * kBK) return false;
  if (kSmemBytes + 2048 > static_cast<size_t>(max_smem_optin())) return false;
  if (getenv("DSPB200_EQ_FORCE_MMA") != nullptr) return true;
  // a CTA walks a group of 128 channels through time: the form pays once the work items keep the SMs about 80 % busy
  // -- at least 0.8 groups per SM (about 15k channels on 148 SMs), or a narrower batch whose time axis is long enough
  // to be cut into overlapping slices (out of place only: a slice re-reads the inputs before its own range); the
  // FFMA scan kernel (47-50 % of the HBM roofline) serves the rest.
  const int64_t groups = ceil_div(channels, kTM), sms = sm_count();
  if (5 * groups >= 4 * sms) return true;
  const LtiSlicing sl = lti_slicing(mp, channels, n_in, !buffers_overlap(x, xs, z, zs, channels));
  return sl.warm > 0 && sl.eff >= 0.62;
}

int lti_mma_run(const LtiMmaPlan& mp, const float* x, int64_t xs, float* z, int64_t zs, int64_t channels,
                int64_t n_in, int64_t n_out, bool clip, float* state, bool state_in, cudaStream_t stream) {
  CUtensorMap tm_a, tm_x, tm_z;
  memset(&tm_a, 0, sizeof(tm_a));
  memset(&tm_x, 0, sizeof(tm_x));
  memset(&tm_z, 0, sizeof(tm_z));
  DSP_TRY(encode_tmap_2d(&tm_a, DSPB200_F32, mp.d_table, static_cast<uint64_t>(mp.kpad),
                         static_cast<uint64_t>(3) * kTNn, static_cast<uint64_t>(mp.kpad) * sizeof(float),
                         kBK, 16, true));
  DSP_TRY(encode_tmap_2d(&tm_x, DSPB200_F32, x, static_cast<uint64_t>(n_in), static_cast<uint64_t>(channels),
                         static_cast<uint64_t>(xs) * sizeof(float), kBK, kTM, true));
  DSP_TRY(encode_tmap_2d(&tm_z, DSPB200_F32, z, static_cast<uint64_t>(n_out), static_cast<uint64_t>(channels),
                         static_cast<uint64_t>(zs) * sizeof(float), kBK, kTM, true));
  LtiArgs a{};
  a.z = z; a.z_stride = zs; a.channels = channels; a.n_out = n_out;
  a.n_tt = ceil_div(n_out, kRows);
  a.n_groups = ceil_div(channels, kTM);
  DSP_CHECK(n_in < (1ll << 31) - 256 && channels < (1ll << 31) - 256 && n_out < (1ll << 31) - 256,
            "shape too large for the tensor-core EQ kernel");
  a.clip = clip ? 1 : 0;
  a.state = state;
  a.state_in = (state != nullptr && state_in) ? 1 : 0;
  memcpy(a.phi, mp.phi, sizeof(a.phi));
  const int64_t sms = sm_count();
  // exact hand-over slices for wide batches, overlapping warm-up slices for narrow ones (never in place, never in the
  // streaming form, whose blocks reproduce one pass bit for bit)
  const LtiSlicing sl = lti_slicing(mp, channels, n_out, state == nullptr && !buffers_overlap(x, xs, z, zs, channels));
  int64_t slices = sl.slices;
  a.warm = sl.warm;
  if (const char* e = getenv("DSPB200_EQ_SLICES")) {   // development: force the slice count
    const long v = atol(e);
    if (v >= 1 && v <= a.n_tt) slices = v;
  }
  a.chunks_per_job = ceil_div(a.n_tt, slices);
  slices = ceil_div(a.n_tt, a.chunks_per_job);
  a.n_jobs = a.n_groups * slices;
  DSP_CHECK(a.n_jobs < (1ll << 31), "shape too large for the tensor-core EQ kernel");
  const int grid = static_cast<int>(a.n_jobs < sms ? a.n_jobs : sms);
  // stream-ordered scratch: the job counter, one flag per work item and, when sliced, the handed-over end states
  std::call_once(g_pool_once, [] {
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  });
  const size_t flag_bytes = static_cast<size_t>(round_up((a.n_jobs + 1) * static_cast<int64_t>(sizeof(unsigned)), 256));
  const size_t xfer_bytes = (slices > 1 && a.warm == 0) ? static_cast<size_t>(a.n_jobs) * mp.states * kTM * sizeof(float) : 0;
  unsigned char* scratch = nullptr;
  DSP_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&scratch), flag_bytes + xfer_bytes, stream));
  DSP_CUDA(cudaMemsetAsync(scratch, 0, flag_bytes, stream));
  a.counter = reinterpret_cast<unsigned*>(scratch);
  a.done = a.counter + 1;
  a.xfer = reinterpret_cast<float*>(scratch + flag_bytes);
  unsigned long long* prof = nullptr;
  if (getenv("DSPB200_LTI_PROF") != nullptr) {
    cudaMalloc(reinterpret_cast<void**>(&prof), 8 * sizeof(unsigned long long));
    cudaMemset(prof, 0, 8 * sizeof(unsigned long long));
  }
  a.prof = prof;
  int rc;
  switch (mp.states) {
    case 4: rc = launch<4>(tm_a, tm_x, tm_z, a, grid, stream); break;
    case 8: rc = launch<8>(tm_a, tm_x, tm_z, a, grid, stream); break;
    case 12: rc = launch<12>(tm_a, tm_x, tm_z, a, grid, stream); break;
    case 16: rc = launch<16>(tm_a, tm_x, tm_z, a, grid, stream); break;
    default: rc = fail(DSPB200_ERR_INVALID, "internal: bad state count %d", mp.states);
  }
  cudaFreeAsync(scratch, stream);
  if (prof) {
    unsigned long long h[8];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(prof);
    const double nt = h[2] ? static_cast<double>(h[2]) : 1.0;
    fprintf(stderr, "lti_mma: %d CTAs, %lld groups x %lld slices (warm-up %d chunks); epilogue cycles per tile: waiting for the accumulator %.0f, tensor memory -> registers %.0f, state hand-over %.0f, clip + stores %.0f\n",
            grid, static_cast<long long>(a.n_groups), static_cast<long long>(slices), a.warm, h[0] / nt, h[1] / nt, h[3] / nt, h[4] / nt);
  }
  return rc;
}

}  // namespace dspb200
