// K1 + K2 in ONE pass over HBM: x -> z without materialising y (app.py:164-167: the resampler's output is only ever
// the equaliser's input).
//
// Both stages are linear, so over a chunk of C = 80 outputs of the 160/147 resampler
//     [z; u] = G_ph x_win + O s ,      s' = Phi s + u ,      G_ph = [T; K] A_ph
// where A_ph [80 x 113] is the banded tap matrix of conversion_tasa_muestreo (dsp_core.py:148-170; phase ph = chunk mod 2,
// because 80 outputs advance the input by 73.5 samples), x_win the 113 input samples the chunk's FIR windows reach,
// and (T, K, O, Phi) the 80-sample chunk system of the biquad cascade of sistema_ecualizador (dsp_core.py:230-254,
// lti_chunk_system in eq_mma.cu).  G_ph is formed once per plan in float64.  One tile is
//     D[128 channels x 96] = X[128 x 128] . G_ph^T  +  S[128 x 64] . O'^T            (tcgen05.mma, fp32 accumulators in TMEM)
// evaluated in fp16 pieces: x 2^6 = x_hi + x_lo and G 2^g = G_hi + G_lo (round to nearest), three products
// x_hi G_hi + x_hi G_lo + x_lo G_hi (each exact in fp32; the dropped x_lo G_lo is 2^-22 relative), the state in three
// pieces [s1 | s2 | s3 | s1] . [O_hi | O_hi | O_hi | O_lo].  fp16 rather than the TF32 split of the two separate kernels:
// twice the tensor throughput and half the shared-memory footprint, which is what lets both G_ph (hi and lo) stay
// RESIDENT in shared memory (98 KB) -- the coefficient tiles are never re-streamed from L2.
//
// Orientation: M = channels.  Both operands' "A" side lives in TENSOR MEMORY: a converter thread owns a channel
// (= TMEM lane), reads its new input samples from the TMA-staged x boxes, splits them and writes packed fp16 pairs
// with tcgen05.st; the 40-sample overlap between consecutive windows stays in its registers, so every x sample
// crosses shared memory once.  The MMAs therefore read only the B operand (coefficients) from shared memory.  G_ph is
// causal (an input only reaches outputs at or after its own time): k-step j's MMAs skip the first r0[j] coefficient rows.
// The epilogue thread that owns the lane keeps s in registers (s' = Phi s + u: 144 FMA per chunk), hands the split state
// to the MMA warp through tensor memory, clips once (dsp_core.py:254) and lays 128-byte runs into a swizzled staging
// tile that leaves as TMA stores of [128 channels x 32 samples] (two chunks = five full boxes).
//
// Warp roles (one persistent CTA per SM): warps 0-3 epilogue, 4-7 converters (one TMEM lane quarter each), 8 TMA
// producer (coefficient tiles once, then 32-sample x boxes through a 7-deep ring), 9 MMA issuer.  Two window buffers and
// two accumulators in tensor memory: conversion of chunk k+1 and the epilogue of chunk k-1 overlap the MMAs of chunk k.
//
// Domain: |x| < 1023 and |state| < 1023 (fp16 pieces; audio is |x| <= 1) -- larger values overflow to Inf/NaN, as do
// non-finite inputs for their whole chunk.  The three-kernel cascade (DSPB200_CHAIN_NO_FUSED=1) has no such limit.
#include <cuda_fp16.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "design.cuh"
#include "internal.cuh"
#include "umma.cuh"

namespace dspb200 {

namespace {

constexpr int kTM = 128;            // channels per group (MMA M, TMEM lanes)
constexpr int kC = 80;              // outputs per chunk
constexpr int kND = 96;             // accumulator columns = coefficient rows: 80 outputs + 16 state slots
constexpr int kNks = 8;             // k-steps of 16 window samples
constexpr int kWin = 114;           // window samples per chunk: the 113 a chunk's FIR windows reach, plus one so that both phases
                                    // have the same shape (40 carried + 74 new); the extra column of G_ph is zero
constexpr int kNew = 74;            // samples read per chunk; consecutive windows start 73 (after an even chunk) or 74 apart
constexpr int kAdv0 = 73, kAdv1 = 74;
constexpr int kHalo = 40;
static_assert(kNew == kWin - kHalo && kAdv1 == kNew && kAdv0 == kNew - 1, "window = carried halo + new samples");
constexpr int kXSlots = 6;             // x boxes in shared memory (the converter holds up to four); further ahead the boxes are
constexpr int kPrefetch = 0;            // prefetched into L2, so a landing only has to cover the L2 latency
constexpr int kStages = 2;            // staging tiles of the TMA stores
constexpr uint32_t kBoxBytes = kTM * 32 * 4;   // [128 channels x 32 samples] fp32, 128-byte rows, swizzled
constexpr int kEpiWarps = 4, kConvWarps = 4;
constexpr int kConvWarp0 = 4, kTmaWarp = 8, kMmaWarp = 9;
constexpr int kThreads = 12 * 32;   // three warpgroups: epilogue, converters, {TMA, MMA, two idle warps}
// registers are re-divided between the warpgroups at kernel start (setmaxnreg); the pool is what the launch allotted,
// 384 threads x 168: the three counts must not add up to more than 3 x 168 = 504 or an increase waits for ever
constexpr int kRegsEpi = 200, kRegsConv = 208, kRegsAux = 96;
static_assert(kRegsEpi + kRegsConv + kRegsAux <= 3 * 168, "setmaxnreg budget");
constexpr uint32_t kColX = 0;       // window buffers: [2][hi 64 | lo 64] columns
constexpr uint32_t kColD = 256;     // accumulators: [2][96]
constexpr uint32_t kColS = 448;     // split state pieces, 32 columns: [s1 | s2 | s3 | s1 | s2] kS fp16 each (kS <= 12), else [s1 | s2 | s3] x 16
constexpr int kXScaleExp = 6, kSScaleExp = 6;
// Advance the state inside the free-response MMAs (rows 80.. of that operand hold Phi, five fp16 pieces of the state)?
// Correct and 350 cycles per chunk shorter in the epilogue, but the 16 extra coefficient rows (2 KB) do not fit next to six
// x boxes and two staging tiles, and giving up either costs more than it gains (5 boxes: 16.9 ms, one staging tile:
// 16.7 ms against 12.7 ms per 18944-clip wave).  Kept for a cascade of at most 12 states should the budget change.
constexpr bool kUseStateMma = false;

struct XzArgs {
  long long channels, n_in, n_out;
  int n_chunks, n_groups;
  int n_boxes, box0;            // aligned 32-sample x boxes per group; index of the first (negative: left zero padding)
  int first_new;                // first sample that is not part of chunk 0's carried halo
  int clip;
  float unscale, x_scale, s_scale;
  uint32_t tab_bytes, o_off;
  uint32_t blk_off[2][2][2];    // [phase][hi, lo][64-sample block]: byte offset of the tile in the table
  int blk_row0[2][2];           // first coefficient row a tile holds
  int r0[2][kNks];              // first coefficient row the k-step's MMAs touch (multiple of 16)
  unsigned long long* prof;     // development (DSPB200_XZ_PROF=1): cycles per phase of each role, CTA 0
  float phi[kLtiMaxStates * kLtiMaxStates];
};

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {   // two fp16 (round to nearest even), `lo` in the low half
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) {
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}

// (v0, v1) x scale -> packed fp16 hi and lo (round to nearest even): six instructions per pair with the packed fp32 forms
__device__ __forceinline__ void split_pair(float v0, float v1, float scale, uint32_t& hi, uint32_t& lo) {
  const float2 as = fmul2s(make_float2(v0, v1), scale);
  hi = pack_h2(as.x, as.y);
  const float2 d = ffma2s(unpack_h2(hi), -1.f, as);
  lo = pack_h2(d.x, d.y);
}

// s <- Phi s + u.  phi_t is Phi transposed (phi_t[j][i] = Phi[i][j]) in the kernel parameters: a packed FFMA takes the
// coefficient pair (Phi[2i][j], Phi[2i+1][j]) from the constant bank and s[j] as the broadcast operand.
template <int kS>
__device__ __forceinline__ void advance_state(float (&s)[kS], const float* __restrict__ phi_t, const float (&u)[kS]) {
  float2 t[kS / 2];
#pragma unroll
  for (int i = 0; i < kS / 2; ++i) t[i] = make_float2(u[2 * i], u[2 * i + 1]);
#pragma unroll
  for (int j = 0; j < kS; ++j)
#pragma unroll
    for (int i = 0; i < kS / 2; ++i)
      ffma2_bcast(t[i], *reinterpret_cast<const float2*>(phi_t + j * kLtiMaxStates + 2 * i), s[j]);
#pragma unroll
  for (int i = 0; i < kS / 2; ++i) { s[2 * i] = t[i].x; s[2 * i + 1] = t[i].y; }
}

// The converter's view of the x ring: boxes are numbered from the group's first one; `base` is the running box count
// of this CTA at the start of the group (ring slot and mbarrier parity follow from base + index).
struct XRing {
  uint32_t xring;          // shared-memory address of slot 0
  uint64_t* full;
  uint64_t* empty;
  uint32_t base;
  int box0;
  int waited, released;    // boxes of this group waited for / handed back so far
};

// r[0 .. 4 N4 - 4] <- samples pos, pos + 1, ... of this thread's channel row (N4 aligned 16-byte reads starting at the
// multiple of 4 at or below pos, then a shift by pos mod 4)
template <int N4>
__device__ __forceinline__ void read_run(float (&r)[4 * N4], int pos, XRing& q, int row) {
  const int pos4 = pos & ~3;
  const int bi0 = (pos4 >> 5) - q.box0;
  const int c0 = (pos4 >> 2) & 7;
  const int bi_last = bi0 + ((c0 + N4 - 1) >> 3);
  for (int b = q.waited; b <= bi_last; ++b) {
    const uint32_t it = q.base + static_cast<uint32_t>(b);
    mbar_wait(&q.full[it % kXSlots], (it / kXSlots) & 1);
  }
  if (bi_last + 1 > q.waited) q.waited = bi_last + 1;
  uint32_t slot = (q.base + static_cast<uint32_t>(bi0)) % kXSlots;
  uint32_t ch = static_cast<uint32_t>(c0);
  const uint32_t rx = static_cast<uint32_t>(row & 7);
  uint32_t rowaddr = q.xring + slot * kBoxBytes + static_cast<uint32_t>(row) * 128u;
#pragma unroll
  for (int i = 0; i < N4; ++i) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r[4 * i]), "=f"(r[4 * i + 1]), "=f"(r[4 * i + 2]), "=f"(r[4 * i + 3])
                 : "r"(rowaddr + ((ch ^ rx) << 4)));
    if (++ch == 8) {
      ch = 0;
      slot = slot + 1 == kXSlots ? 0 : slot + 1;
      rowaddr = q.xring + slot * kBoxBytes + static_cast<uint32_t>(row) * 128u;
    }
  }
  const int sh = pos - pos4;
  if (sh & 1) {
#pragma unroll
    for (int i = 0; i + 1 < 4 * N4; ++i) r[i] = r[i + 1];
  }
  if (sh & 2) {
#pragma unroll
    for (int i = 0; i + 2 < 4 * N4; ++i) r[i] = r[i + 2];
  }
}

// boxes that end at or before `next_pos` (rounded down to a multiple of 4) are not read again
__device__ __forceinline__ void release_upto(XRing& q, int next_pos, int lane) {
  const int upto = ((next_pos & ~3) >> 5) - q.box0;
  __syncwarp();
  for (int b = q.released; b < upto; ++b) {
    const uint32_t it = q.base + static_cast<uint32_t>(b);
    if (lane == 0) mbar_arrive(&q.empty[it % kXSlots]);
  }
  if (upto > q.released) q.released = upto;
}

// One chunk of the converter: window = [40 carried samples | 74 new samples]; split every sample x 2^6 into fp16 hi + lo and
// store packed pairs to the window buffer in tensor memory (column c = samples 2c, 2c + 1).  The next window starts
// `adv` = 73 or 74 samples on: its first 40 samples are kept in `carry`.  One code path for both phases: the kernel is
// bound by instruction fetch before anything else.
__device__ __forceinline__ void convert_chunk(float (&carry)[kHalo], int pos, int adv, XRing& q, int row, int lane, float x_scale,
                                              uint32_t tbuf, unsigned long long* prof) {
  float r[80];
  long long t0 = 0;
  if (prof) t0 = clock64();
  read_run<20>(r, pos, q, row);
  if (prof) atomicAdd(prof + 2, static_cast<unsigned long long>(clock64() - t0));   // box waits + shared-memory reads
  release_upto(q, pos + adv, lane);     // the samples are in registers: hand the boxes behind the next chunk's start back now
  auto win = [&](int w) -> float {   // compile-time index after unrolling
    return w < kHalo ? carry[w] : (w < kWin ? r[w - kHalo] : 0.f);
  };
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const int w = 32 * g + 2 * c;
      if (w < kWin) {
        split_pair(win(w), win(w + 1), x_scale, hi[c], lo[c]);
      } else {
        hi[c] = 0u;
        lo[c] = 0u;
      }
    }
    tmem_st16(tbuf + 16 * g, hi);
    tmem_st16(tbuf + 64 + 16 * g, lo);
  }
  if (adv == kAdv1) {
#pragma unroll
    for (int j = 0; j < kHalo; ++j) carry[j] = r[kAdv1 - kHalo + j];
  } else {
#pragma unroll
    for (int j = 0; j < kHalo; ++j) carry[j] = r[kAdv0 - kHalo + j];
  }
}

__device__ __forceinline__ bool elect_one() {   // one lane of a converged warp
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
  return pred != 0;
}

// kProf: the development instantiation (DSPB200_XZ_PROF=1) counts cycles per phase of each role; in the production one
// `prof_p` is a compile-time NULL and every profiling branch leaves the binary (the kernel is bound by instruction fetch).
template <int kS, bool kProf>
__global__ void __launch_bounds__(kThreads, 1)
xz_mma_kernel(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_x,
              const __grid_constant__ CUtensorMap tm_z, const __grid_constant__ XzArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];   // no static shared memory in this kernel: the window starts 1024-aligned
  // layout: coefficient tiles | x ring | staging tiles | barriers | MMA parameter table
  unsigned char* xring = smem + a.tab_bytes;
  unsigned char* stage = xring + kXSlots * kBoxBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage + kStages * kBoxBytes);
  uint64_t* full_x = bars;                    // [x slot] TMA landed the box
  uint64_t* empty_x = full_x + kXSlots;       // [x slot] all four converter warps are done with it
  uint64_t* x_ready = empty_x + kXSlots;      // [2] window buffer written to tensor memory
  uint64_t* x_free = x_ready + 2;             // [2] the MMAs that read it have completed
  uint64_t* acc_full = x_free + 2;            // [2] accumulator complete (free response included)
  uint64_t* acc_empty = acc_full + 2;         // [2] accumulator drained by the epilogue warps
  uint64_t* tab_full = acc_empty + 2;         // coefficient tiles resident
  uint64_t* s_ready = tab_full + 1;           // the next chunk's start states are in tensor memory
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(s_ready + 1);
  uint4* mma_tab = reinterpret_cast<uint4*>(bars + 32);     // [2 phases][8 k-steps]
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // warp-uniform for the compiler too
  const int lane = threadIdx.x & 31;
  unsigned long long* const prof_p = kProf ? a.prof : nullptr;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0u) __trap();   // the swizzle atoms need the 1024-byte alignment
    for (int s = 0; s < kXSlots; ++s) { mbar_init(&full_x[s], 1); mbar_init(&empty_x[s], kConvWarps); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&x_ready[b], kConvWarps); mbar_init(&x_free[b], 1);
      mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], kEpiWarps);
    }
    mbar_init(tab_full, 1);
    mbar_init(s_ready, kEpiWarps);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_base_s;
  long long prof_c0 = 0;
  unsigned long long prof_g0 = 0;
  if (prof_p != nullptr && threadIdx.x == 0) {
    prof_c0 = clock64();
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(prof_g0));
  }
  const int n_local = (a.n_groups - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp >= 2 * 4) {
   // third warpgroup (TMA producer, MMA issuer, two idle warps): gives most of its registers to the other two.  Both roles
   // run as whole converged warps and elect one lane per instruction: tcgen05 / TMA instructions issued from a divergent
   // `lane == 0` region cost a seven-instruction election loop each.
   asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsAux));
   if (warp == kTmaWarp) {
    // ---------------- TMA producer: coefficient tiles once, then the x boxes of each group in time order ----------------
    if (elect_one()) {
      tma_prefetch_desc(&tm_g);
      tma_prefetch_desc(&tm_x);
      mbar_expect_tx(tab_full, a.tab_bytes);
      for (uint32_t off = 0; off < a.tab_bytes; off += 16 * 128)      // boxes of 16 rows x 64 fp16
        tma_load_2d(smem + off, &tm_g, 0, static_cast<int>(off / 128), tab_full);
    }
    __syncwarp();
    uint32_t it = 0;
    for (int gi = 0; gi < n_local; ++gi) {
      const int g = static_cast<int>(blockIdx.x) + gi * static_cast<int>(gridDim.x);
      if (elect_one())
        for (int bi = 0; bi < kPrefetch && bi < a.n_boxes; ++bi) tma_prefetch_l2_2d(&tm_x, (a.box0 + bi) * 32, g * kTM);
      __syncwarp();
      for (int bi = 0; bi < a.n_boxes; ++bi, ++it) {
        const uint32_t s = it % kXSlots;
        if (it >= kXSlots) mbar_wait(&empty_x[s], ((it / kXSlots) - 1) & 1);
        if (elect_one()) {
          mbar_expect_tx(&full_x[s], kBoxBytes);
          tma_load_2d(xring + s * kBoxBytes, &tm_x, (a.box0 + bi) * 32, g * kTM, &full_x[s]);
          if (kPrefetch > 0 && bi + kPrefetch < a.n_boxes) tma_prefetch_l2_2d(&tm_x, (a.box0 + bi + kPrefetch) * 32, g * kTM);
        }
        __syncwarp();
      }
    }
   } else if (warp == kMmaWarp) {
    // ---------------- MMA issuer ----------------
    // loop invariants of the two phases -- accumulator column, instruction descriptor and the two coefficient descriptors
    // of every k-step -- in shared memory, so that the issue loop stays a loop
    if (lane < 2 * kNks) {
      const int ph = lane / kNks, j = lane % kNks;
      const int r0 = a.r0[ph][j], blk = j >> 2;
      const uint32_t rel = static_cast<uint32_t>(r0 - a.blk_row0[ph][blk]) * 128u;
      mma_tab[lane] = make_uint4(static_cast<uint32_t>(r0), umma_idesc_f16(kND - r0),
                                 static_cast<uint32_t>(umma_desc_sw128(smem + a.blk_off[ph][0][blk] + rel)) + 2 * (j & 3),
                                 static_cast<uint32_t>(umma_desc_sw128(smem + a.blk_off[ph][1][blk] + rel)) + 2 * (j & 3));
    }
    __syncwarp();
    const uint64_t desc_hi = umma_desc_sw128(smem) & 0xFFFFFFFF00000000ull;
    const uint64_t od = umma_desc_sw128(smem + a.o_off);
    // the free response of the start state; with 5 kS <= 64 the same four MMAs also advance the state (rows 80.. of the
    // operand hold Phi): the accumulator's state columns then hold s' = Phi s + K x and nothing of the recurrence is left
    // on the FMA pipe.  Wider cascades keep the state update in the epilogue and the state columns take no free response.
    constexpr bool kStateMma = kUseStateMma && 5 * kS <= 64;
    const uint32_t ido = umma_idesc_f16(kStateMma ? kND : kC);
    mbar_wait(tab_full, 0);
    uint32_t kk = 0, n_corr = 0;
    for (int gi = 0; gi < n_local; ++gi) {
      for (int k = 0; k < a.n_chunks; ++k, ++kk) {
        const uint32_t buf = kk & 1;
        const bool prof = prof_p != nullptr && blockIdx.x == 0 && lane == 0;
        long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
        if (prof) t0 = clock64();
        mbar_wait(&x_ready[buf], (kk >> 1) & 1);
        if (prof) t1 = clock64();
        if (kk >= 2) mbar_wait(&acc_empty[buf], ((kk >> 1) - 1) & 1);
        tc_fence_after();
        if (prof) t2 = clock64();
        const uint32_t d = tmem + kColD + buf * kND;
        const uint32_t xh = tmem + kColX + buf * 128, xl = xh + 64;
        const uint4* tab = mma_tab + (k & 1) * kNks;
#pragma unroll 1
        for (int j = 0; j < kNks; ++j) {
          const uint4 t = tab[j];
          const uint64_t gh = desc_hi | t.z, gl = desc_hi | t.w;
          if (elect_one()) {
            umma_f16_ts(d + t.x, xh + 8 * j, gh, t.y, j ? 1u : 0u);   // k-step 0 reaches every row: it overwrites the accumulator
            umma_f16_ts(d + t.x, xh + 8 * j, gl, t.y, 1u);
            umma_f16_ts(d + t.x, xl + 8 * j, gh, t.y, 1u);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(&x_free[buf]);
        __syncwarp();
        if (prof) t3 = clock64();
        if (k > 0) {
          // [z; s'] += [s1 | s2 | s3 | s1 | s2] . [O'_hi | O'_hi | O'_hi | O'_lo | O'_lo]^T (kS <= 12; else four pieces, O only);
          // chunk 0 starts from a zero state (lfilter, dsp_core.py:214)
          mbar_wait(s_ready, n_corr & 1);
          ++n_corr;
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int p = 0; p < 4; ++p)
              umma_f16_ts(d, tmem + kColS + 8 * ((!kStateMma && p == 3) ? 0 : p), od + 2 * p, ido, 1u);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(&acc_full[buf]);
        __syncwarp();
        if (prof) {
          atomicAdd(prof_p + 9, static_cast<unsigned long long>(t1 - t0));    // waiting for the window
          atomicAdd(prof_p + 10, static_cast<unsigned long long>(t2 - t1));   // waiting for the accumulator
          atomicAdd(prof_p + 11, static_cast<unsigned long long>(t3 - t2));   // issuing the main products
          atomicAdd(prof_p + 12, static_cast<unsigned long long>(clock64() - t3));   // state + free response
          atomicAdd(prof_p + 13, 1ull);
        }
      }
    }
   }
  } else if (warp >= kConvWarp0) {
    // ---------------- converters: thread = channel = TMEM lane ----------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsConv));
    const int row = (warp - kConvWarp0) * 32 + lane;
    const uint32_t lane_base = tmem + (static_cast<uint32_t>((warp - kConvWarp0) * 32) << 16);
    XRing q;
    q.xring = smem_u32(xring);
    q.full = full_x;
    q.empty = empty_x;
    q.box0 = a.box0;
    uint32_t kk = 0;
    for (int gi = 0; gi < n_local; ++gi) {
      q.base = static_cast<uint32_t>(gi) * static_cast<uint32_t>(a.n_boxes);
      q.waited = 0;
      q.released = 0;
      float carry[kHalo];
      {
        float r[44];
        read_run<11>(r, a.first_new - kHalo, q, row);
#pragma unroll
        for (int j = 0; j < kHalo; ++j) carry[j] = r[j];
      }
      int pos = a.first_new;
      for (int k = 0; k < a.n_chunks; ++k, ++kk) {
        const uint32_t buf = kk & 1;
        const bool prof = prof_p != nullptr && blockIdx.x == 0 && threadIdx.x == kConvWarp0 * 32;
        long long t0 = 0, t1 = 0;
        if (prof) t0 = clock64();
        if (kk >= 2) {
          mbar_wait(&x_free[buf], ((kk >> 1) - 1) & 1);
          tc_fence_after();
        }
        if (prof) t1 = clock64();
        const uint32_t tbuf = lane_base + kColX + buf * 128;
        const int adv = (k & 1) ? kAdv1 : kAdv0;
        convert_chunk(carry, pos, adv, q, row, lane, a.x_scale, tbuf, prof ? prof_p : nullptr);
        pos += adv;
        long long t2 = 0;
        if (prof) t2 = clock64();
        tmem_wait_st();
        if (prof) atomicAdd(prof_p + 14, static_cast<unsigned long long>(clock64() - t2));   // tcgen05.wait::st
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&x_ready[buf]);
        if (prof) {
          atomicAdd(prof_p + 0, static_cast<unsigned long long>(t1 - t0));            // waiting for the window buffer
          atomicAdd(prof_p + 1, static_cast<unsigned long long>(clock64() - t1));     // boxes, conversion, tensor-memory stores
          atomicAdd(prof_p + 3, 1ull);
        }
      }
      // hand the group's remaining boxes back so the producer can go on with the next group
      __syncwarp();
      for (int b = q.released; b < a.n_boxes; ++b) {
        if (b >= q.waited) {
          const uint32_t it = q.base + static_cast<uint32_t>(b);
          mbar_wait(&full_x[it % kXSlots], (it / kXSlots) & 1);
        }
        if (lane == 0) mbar_arrive(&empty_x[(q.base + static_cast<uint32_t>(b)) % kXSlots]);
      }
    }
  } else {
    // ---------------- epilogue warps 0-3: thread = TMEM lane = channel ----------------
    // Each warp stages and stores its own 32 channels (TMA boxes of [32 channels x 32 samples] out of its quarter of the
    // staging tiles): no barrier between the four warps, only the mbarriers towards the MMA warp.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsEpi));
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const int r = warp * 32 + lane;
    const uint32_t row_addr0 = smem_u32(stage) + static_cast<uint32_t>(r) * 128u;
    const uint32_t rx = static_cast<uint32_t>(r & 7);
    unsigned char* const warp_stage = stage + warp * (32 * 128);
    uint32_t kk = 0, n_box = 0;      // boxes stored so far: box n goes through staging tile n mod kStages
    // the start state of the coming chunk, x 2^6 and split in three fp16 pieces, to tensor memory
    constexpr bool kStateMma = kUseStateMma && 5 * kS <= 64;
    auto hand_over = [&](const float (&s)[kS]) {
      uint32_t w[3][kS / 2];
#pragma unroll
      for (int c = 0; c < kS / 2; ++c) {
        const float v0 = s[2 * c] * a.s_scale, v1 = s[2 * c + 1] * a.s_scale;
        w[0][c] = pack_h2(v0, v1);
        const float2 f1 = unpack_h2(w[0][c]);
        const float r0 = v0 - f1.x, r1 = v1 - f1.y;
        w[1][c] = pack_h2(r0, r1);
        const float2 f2 = unpack_h2(w[1][c]);
        w[2][c] = pack_h2(r0 - f2.x, r1 - f2.y);
      }
      uint32_t cols[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        if constexpr (kStateMma) {       // [s1 | s2 | s3 | s1 | s2 | 0], kS fp16 each
          const int q = c / (kS / 2);
          cols[c] = q < 5 ? w[q % 3][c % (kS / 2)] : 0u;
        } else {                         // [s1 | s2 | s3 | unused], 16 fp16 each
          cols[c] = (c < 24 && (c % 8) < kS / 2) ? w[c / 8][c % 8] : 0u;
        }
      }
      tmem_st16(lane_base + kColS, cols);
      tmem_st16(lane_base + kColS + 16, cols + 16);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_ready);
    };
    // the warp's store that last used the coming staging tile has read it
    auto wait_stage = [&]() {
      long long tw = 0;
      const bool pw = prof_p != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
      if (pw) tw = clock64();
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kStages - 1) : "memory");
      __syncwarp();
      if (pw) atomicAdd(prof_p + 16, static_cast<unsigned long long>(clock64() - tw));
    };
    for (int gi = 0; gi < n_local; ++gi) {
      const int g = static_cast<int>(blockIdx.x) + gi * static_cast<int>(gridDim.x);
      float s[kS];
#pragma unroll
      for (int i = 0; i < kS; ++i) s[i] = 0.f;
      for (int k = 0; k < a.n_chunks; ++k, ++kk) {
        const uint32_t buf = kk & 1;
        const bool prof = prof_p != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
        long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
        if (prof) t0 = clock64();
        mbar_wait(&acc_full[buf], (kk >> 1) & 1);
        tc_fence_after();
        if (prof) t1 = clock64();
        const uint32_t taddr = lane_base + kColD + buf * kND;
        uint32_t v[6][16];
        tmem_ld16(v[5], taddr + kC);
#pragma unroll
        for (int blk = 0; blk < 5; ++blk) tmem_ld16(v[blk], taddr + blk * 16);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
        if (prof) t2 = clock64();
        // the next chunk's start state first: the MMA warp is waiting for it
        if constexpr (kStateMma) {      // the accumulator's state columns already hold Phi s + K x
#pragma unroll
          for (int i = 0; i < kS; ++i) s[i] = __uint_as_float(v[5][i]) * a.unscale;
        } else {
          float u[kS];
#pragma unroll
          for (int i = 0; i < kS; ++i) u[i] = __uint_as_float(v[5][i]) * a.unscale;
          advance_state<kS>(s, a.phi, u);
        }
        long long t2b = 0;
        if (prof) t2b = clock64();
        if (k + 1 < a.n_chunks) hand_over(s);
        if (prof) t3 = clock64();
        if (prof) atomicAdd(prof_p + 15, static_cast<unsigned long long>(t3 - t2b));   // of "state": the hand-over
        // unscale, clip, stage, store.  Two chunks = 160 outputs = five boxes of 32: an even chunk fills boxes 0, 1 and the
        // first half of box 2 of its pair, the odd chunk the rest.
        auto put4 = [&](int q, uint32_t pp0) {   // pieces 4q .. 4q+3 (outputs 16q .. 16q+15) to piece positions pp0 .. pp0+3 of the staging row
          float o[16];
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const float2 m = fmul2s(make_float2(__uint_as_float(v[q][e]), __uint_as_float(v[q][e + 1])), a.unscale);
            o[e] = m.x;
            o[e + 1] = m.y;
          }
          if (a.clip) {
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = clip_unit(o[e]);
          }
          const uint32_t base = row_addr0 + (n_box % kStages) * kBoxBytes;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};"
                         ::"r"(base + (((pp0 + j) ^ rx) << 4)), "f"(o[4 * j]), "f"(o[4 * j + 1]), "f"(o[4 * j + 2]), "f"(o[4 * j + 3]) : "memory");
        };
        auto send = [&](int box) {
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tm_z, warp_stage + (n_box % kStages) * kBoxBytes, box * 32, g * kTM + warp * 32);
            tma_store_commit();
          }
          ++n_box;
        };
        // staging position of the chunk's first piece and the pieces at which a box starts / is complete
        const uint32_t off = (k & 1) ? 4u : 0u;
        const uint32_t start_mask = (k & 1) ? 0x01010u : 0x10101u;           // odd: pieces 4, 12; even: 0, 8, 16
        const uint32_t end_mask = (k & 1) ? 0x80808u : (k + 1 == a.n_chunks ? 0x88080u : 0x08080u);   // even: 7, 15 (+ 19 when the signal ends here: the hardware clips the rest)
        int box = 5 * (k >> 1) + ((k & 1) ? 2 : 0);
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          if ((start_mask >> (4 * q)) & 1u) wait_stage();
          put4(q, (static_cast<uint32_t>(4 * q) + off) & 7u);
          if ((end_mask >> (4 * q + 3)) & 1u) send(box++);
        }
        if (prof) {
          atomicAdd(prof_p + 4, static_cast<unsigned long long>(t1 - t0));           // waiting for the accumulator
          atomicAdd(prof_p + 5, static_cast<unsigned long long>(t2 - t1));           // tensor memory -> registers
          atomicAdd(prof_p + 6, static_cast<unsigned long long>(t3 - t2));           // state update + hand-over
          atomicAdd(prof_p + 7, static_cast<unsigned long long>(clock64() - t3));    // clip + staging + stores
          atomicAdd(prof_p + 8, 1ull);
        }
      }
    }
    if (lane == 0) tma_store_wait_all0();
    if (prof_p != nullptr && threadIdx.x == 0) {
      unsigned long long g1;
      unsigned smid;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
      asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
      if (blockIdx.x == 0) {
        prof_p[20] = static_cast<unsigned long long>(clock64() - prof_c0);
        prof_p[21] = g1 - prof_g0;
      }
      if (blockIdx.x < 160) {      // per CTA: start and end time (ns), SM
        prof_p[32 + 3 * blockIdx.x] = prof_g0;
        prof_p[33 + 3 * blockIdx.x] = g1;
        prof_p[34 + 3 * blockIdx.x] = smid;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, 512);
}

long long floor_div(long long x, long long y) { return x >= 0 ? x / y : -((-x + y - 1) / y); }

size_t xz_smem_bytes(const XzPlan& xp) {
  return static_cast<size_t>(xp.tab_bytes) + (kXSlots + kStages) * static_cast<size_t>(kBoxBytes) + 32 * sizeof(uint64_t) + 2 * kNks * sizeof(uint4);
}

template <int kS>
int launch(const CUtensorMap& tm_g, const CUtensorMap& tm_x, const CUtensorMap& tm_z, const XzArgs& a, size_t smem, int grid,
           cudaStream_t stream) {
  if (a.prof != nullptr) {
    DSP_CUDA(cudaFuncSetAttribute(xz_mma_kernel<kS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    xz_mma_kernel<kS, true><<<grid, kThreads, smem, stream>>>(tm_g, tm_x, tm_z, a);
  } else {
    DSP_CUDA(cudaFuncSetAttribute(xz_mma_kernel<kS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    xz_mma_kernel<kS, false><<<grid, kThreads, smem, stream>>>(tm_g, tm_x, tm_z, a);
  }
  return after_launch("xz_mma_kernel");
}

}  // namespace

// Coefficient tables of the fused form for (taps of L/M, cascade).  xp.ok stays 0 when the shape is not the one the kernel
// is laid out for (chunks of 80 outputs in two phases with 113-sample windows: 160/147) or a value does not fit fp16.
int xz_build(const std::vector<double>& taps, int L, int M, const Section* sec, int ns, XzPlan& xp) {
  xp = XzPlan{};
  if (ns < 1 || 2 * ns > kLtiMaxStates) return DSPB200_OK;
  const long long T = static_cast<long long>(taps.size()), P = (T - 1) / 2;
  if ((static_cast<long long>(2 * kC) * M) % L != 0 || (static_cast<long long>(kC) * M) % L == 0) return DSPB200_OK;   // two phases
  // windows of the two phases: the 113 samples a chunk's FIR windows reach, widened to 114 so that both phases are
  // [40 carried | 74 new]: an even chunk gets one sample more at the end, an odd chunk one at the start (zero columns of G)
  long long start[2];
  for (int ph = 0; ph < 2; ++ph) {
    const long long m0 = static_cast<long long>(ph) * kC;
    const long long first = -floor_div(-(m0 * M + P - T + 1), L), last = floor_div((m0 + kC - 1) * M + P, L);
    if (last - first + 1 != kWin - 1) return DSPB200_OK;
    start[ph] = first - ph;
  }
  const long long adv = static_cast<long long>(2 * kC) * M / L;
  if (start[1] - start[0] != kAdv0 || start[0] + adv - start[1] != kAdv1) return DSPB200_OK;
  LtiChunkSystem cs;
  DSP_TRY(lti_chunk_system(sec, ns, cs, kC));
  if (cs.rows != kC) return DSPB200_OK;
  const int n = cs.states;
  // G_ph = [T; K] A_ph, float64 (long double accumulation)
  std::vector<double> G(static_cast<size_t>(2) * kND * 128, 0.0);
  double gmax = 0.0;
  for (int ph = 0; ph < 2; ++ph) {
    std::vector<double> A(static_cast<size_t>(kC) * kWin, 0.0);
    for (int r = 0; r < kC; ++r)
      for (int w = 0; w < kWin; ++w) {
        const long long t = (static_cast<long long>(ph) * kC + r) * M + P - (start[ph] + w) * L;
        if (t >= 0 && t < T) A[static_cast<size_t>(r) * kWin + w] = taps[static_cast<size_t>(t)];
      }
    for (int row = 0; row < kC + n; ++row)
      for (int w = 0; w < kWin; ++w) {
        long double acc = 0.0L;
        for (int j = 0; j < kC; ++j)
          acc += static_cast<long double>(cs.tk[static_cast<size_t>(row) * kC + j]) * A[static_cast<size_t>(j) * kWin + w];
        const double v = static_cast<double>(acc);
        G[(static_cast<size_t>(ph) * kND + row) * 128 + w] = v;
        gmax = std::fmax(gmax, std::fabs(v));
      }
  }
  const int ks = (n + 3) / 4 * 4;                       // states padded to a multiple of 4
  const bool state_mma = kUseStateMma && 5 * ks <= 64;  // see the kernel: the state advances inside the free-response MMAs
  double omax = 0.0;
  for (double v : cs.o) omax = std::fmax(omax, std::fabs(v));
  if (state_mma)
    for (double v : cs.phi) omax = std::fmax(omax, std::fabs(v));
  if (!(gmax > 0.0) || !std::isfinite(gmax) || !std::isfinite(omax)) return DSPB200_OK;
  const int ge = static_cast<int>(std::floor(std::log2(8192.0 / gmax)));       // |G 2^ge| in [4096, 8192)
  const int oe = kXScaleExp + ge - kSScaleExp;                                  // O' = O 2^oe, state pieces x 2^6
  if (omax * std::ldexp(1.0, oe) >= 32768.0 || ge < -8 || ge > 24) return DSPB200_OK;
  // which rows does a k-step reach?  r0 = the largest multiple of 16 at or below its first non-zero row
  for (int ph = 0; ph < 2; ++ph)
    for (int j = 0; j < kNks; ++j) {
      int first = kND;
      for (int row = 0; row < kND && first == kND; ++row)
        for (int w = 16 * j; w < 16 * j + 16; ++w)
          if (G[(static_cast<size_t>(ph) * kND + row) * 128 + w] != 0.0) { first = row; break; }
      int r0 = first / 16 * 16;
      if (r0 > kND - 16) r0 = kND - 16;
      if (j == 0) r0 = 0;     // the first k-step overwrites the whole accumulator
      xp.r0[ph][j] = r0;
    }
  // table layout (= shared-memory layout): rows of 64 fp16 (128 bytes); per phase and per hi/lo two 64-sample blocks, each
  // holding the coefficient rows from the smallest r0 of its four k-steps on; then the free-response operand (80 rows)
  uint32_t rows_total = 0;
  for (int ph = 0; ph < 2; ++ph)
    for (int b = 0; b < 2; ++b) {
      int row0 = kND;
      for (int j = 4 * b; j < 4 * b + 4; ++j) row0 = xp.r0[ph][j] < row0 ? xp.r0[ph][j] : row0;
      xp.blk_row0[ph][b] = row0;
    }
  for (int ph = 0; ph < 2; ++ph)
    for (int hl = 0; hl < 2; ++hl)
      for (int b = 0; b < 2; ++b) {
        xp.blk_off[ph][hl][b] = rows_total * 128u;
        rows_total += static_cast<uint32_t>(kND - xp.blk_row0[ph][b]);
      }
  xp.o_off = rows_total * 128u;
  rows_total += state_mma ? kND : kC;
  xp.tab_rows = static_cast<int>(rows_total);
  xp.tab_bytes = rows_total * 128u;
  std::vector<__half> tab(static_cast<size_t>(rows_total) * 64, __float2half_rn(0.f));
  bool finite = true;
  auto split = [&](double v, __half& hi, __half& lo) {
    const float f = static_cast<float>(v);
    hi = __float2half_rn(f);
    const float hf = __half2float(hi);
    lo = __float2half_rn(f - hf);
    if (!std::isfinite(hf)) finite = false;
  };
  const double gs = std::ldexp(1.0, ge), os = std::ldexp(1.0, oe);
  for (int ph = 0; ph < 2; ++ph)
    for (int b = 0; b < 2; ++b)
      for (int row = xp.blk_row0[ph][b]; row < kND; ++row)
        for (int c = 0; c < 64; ++c) {
          __half hi, lo;
          split(G[(static_cast<size_t>(ph) * kND + row) * 128 + 64 * b + c] * gs, hi, lo);
          const size_t rr = static_cast<size_t>(row - xp.blk_row0[ph][b]);
          tab[(xp.blk_off[ph][0][b] / 128 + rr) * 64 + static_cast<size_t>(c)] = hi;
          tab[(xp.blk_off[ph][1][b] / 128 + rr) * 64 + static_cast<size_t>(c)] = lo;
        }
  // free-response operand, 96 rows: outputs 0..79 take O, and (state_mma) rows 80.. take Phi, so that the same MMAs advance the
  // state.  K layout: [s1 | s2 | s3 | s1 | s2] . [hi | hi | hi | lo | lo] with ks fp16 per piece, or, for more than 12
  // states, [s1 | s2 | s3 | s1] . [hi | hi | hi | lo] with 16 per piece (the s2 . lo term, 2^-22 relative, is dropped there)
  for (int r = 0; r < (state_mma ? kC + n : kC); ++r)
    for (int i = 0; i < n; ++i) {
      const double v = r < kC ? cs.o[static_cast<size_t>(r) * kLtiMaxStates + i]
                              : cs.phi[static_cast<size_t>(r - kC) * kLtiMaxStates + i];
      __half hi, lo;
      split(v * os, hi, lo);
      __half* row = &tab[(xp.o_off / 128 + static_cast<size_t>(r)) * 64];
      if (state_mma) { row[i] = hi; row[ks + i] = hi; row[2 * ks + i] = hi; row[3 * ks + i] = lo; row[4 * ks + i] = lo; }
      else { row[i] = hi; row[16 + i] = hi; row[32 + i] = hi; row[48 + i] = lo; }
    }
  for (int i = 0; i < kLtiMaxStates * kLtiMaxStates; ++i) {
    xp.phi[i] = static_cast<float>(cs.phi[static_cast<size_t>(i)]);
    if (!std::isfinite(xp.phi[i])) finite = false;
  }
  if (!finite) return DSPB200_OK;
  DSP_CUDA(cudaMalloc(reinterpret_cast<void**>(&xp.d_table), tab.size() * sizeof(__half)));
  DSP_CUDA(cudaMemcpy(xp.d_table, tab.data(), tab.size() * sizeof(__half), cudaMemcpyHostToDevice));
  xp.states = ks;
  xp.unscale = static_cast<float>(std::ldexp(1.0, -(kXScaleExp + ge)));
  xp.first_new = static_cast<int>(start[0]) + kHalo;
  xp.start0 = static_cast<int>(start[0]);
  xp.L = L;
  xp.M = M;
  cudaGetDevice(&xp.device);
  xp.ok = 1;
  return DSPB200_OK;
}

void xz_free(XzPlan& xp) {
  cudaFree(xp.d_table);
  xp = XzPlan{};
}

int xz_chunk() { return kC; }

bool xz_possible(const XzPlan& xp, const float* x, int64_t xs, const float* z, int64_t zs, int64_t n_in, int n_taps) {
  return xp.ok && reinterpret_cast<uintptr_t>(x) % 16 == 0 && xs % 4 == 0 && reinterpret_cast<uintptr_t>(z) % 16 == 0 &&
         zs % 4 == 0 && n_in * xp.L >= n_taps && n_in < (1ll << 30) && xz_smem_bytes(xp) <= static_cast<size_t>(max_smem_optin());
}

bool xz_usable(const XzPlan& xp, const float* x, int64_t xs, const float* z, int64_t zs, int64_t channels, int64_t n_in,
               int n_taps) {
  if (!xz_possible(xp, x, xs, z, zs, n_in, n_taps)) return false;
  if (getenv("DSPB200_CHAIN_FORCE_FUSED") != nullptr) return true;
  // a CTA walks a group of 128 channels through time: the form pays once the groups fill at least 80 % of the SMs
  const int64_t groups = ceil_div(channels, kTM), sms = sm_count();
  return 5 * groups >= 4 * sms;
}

int xz_run(const XzPlan& xp, const float* x, int64_t xs, float* z, int64_t zs, int64_t channels, int64_t n_in, int64_t n_out,
           bool clip, cudaStream_t stream) {
  DSP_TRY(check_plan_device(xp.device, "fused chain"));
  CUtensorMap tm_g, tm_x, tm_z;
  memset(&tm_g, 0, sizeof(tm_g));
  memset(&tm_x, 0, sizeof(tm_x));
  memset(&tm_z, 0, sizeof(tm_z));
  DSP_TRY(encode_tmap_2d_f16(&tm_g, xp.d_table, 64, static_cast<uint64_t>(xp.tab_rows), 128, 64, 16));
  DSP_TRY(encode_tmap_2d(&tm_x, DSPB200_F32, x, static_cast<uint64_t>(n_in), static_cast<uint64_t>(channels),
                         static_cast<uint64_t>(xs) * sizeof(float), 32, kTM, true));
  DSP_TRY(encode_tmap_2d(&tm_z, DSPB200_F32, z, static_cast<uint64_t>(n_out), static_cast<uint64_t>(channels),
                         static_cast<uint64_t>(zs) * sizeof(float), 32, 32, true));     // one store per epilogue warp
  XzArgs a{};
  a.channels = channels; a.n_in = n_in; a.n_out = n_out;
  a.n_chunks = static_cast<int>(ceil_div(n_out, kC));
  a.n_groups = static_cast<int>(ceil_div(channels, kTM));
  DSP_CHECK(channels < (1ll << 31) - 256 && n_out < (1ll << 31) - 256, "shape too large for the fused SRC->EQ kernel");
  a.first_new = xp.first_new;
  const long long first_read = static_cast<long long>(xp.first_new) - kHalo;         // the prologue fills the halo from here
  a.box0 = static_cast<int>(floor_div(floor_div(first_read, 4) * 4, 32));
  // last sample any chunk's aligned reads touch: start of its new range rounded down to 4, plus 80
  long long pos_last = xp.first_new;
  for (int k = 0; k + 1 < a.n_chunks; ++k) pos_last += (k & 1) ? kAdv1 : kAdv0;
  const long long last_read = floor_div(pos_last, 4) * 4 + 79;
  a.n_boxes = static_cast<int>(floor_div(last_read, 32) - a.box0 + 1);
  a.clip = clip ? 1 : 0;
  a.unscale = xp.unscale;
  a.x_scale = static_cast<float>(std::ldexp(1.0, kXScaleExp));
  a.s_scale = static_cast<float>(std::ldexp(1.0, kSScaleExp));
  a.tab_bytes = xp.tab_bytes;
  a.o_off = xp.o_off;
  memcpy(a.blk_off, xp.blk_off, sizeof(a.blk_off));
  memcpy(a.blk_row0, xp.blk_row0, sizeof(a.blk_row0));
  memcpy(a.r0, xp.r0, sizeof(a.r0));
  for (int i = 0; i < kLtiMaxStates; ++i)
    for (int j = 0; j < kLtiMaxStates; ++j) a.phi[j * kLtiMaxStates + i] = xp.phi[i * kLtiMaxStates + j];   // transposed: see advance_state
  const int64_t sms = sm_count();
  const int grid = static_cast<int>(a.n_groups < sms ? a.n_groups : sms);
  const size_t smem = xz_smem_bytes(xp);
  unsigned long long* prof = nullptr;
  if (getenv("DSPB200_XZ_PROF") != nullptr) {
    cudaMalloc(reinterpret_cast<void**>(&prof), 512 * sizeof(unsigned long long));
    cudaMemset(prof, 0, 512 * sizeof(unsigned long long));
  }
  a.prof = prof;
  int rc;
  switch (xp.states) {
    case 4: rc = launch<4>(tm_g, tm_x, tm_z, a, smem, grid, stream); break;
    case 8: rc = launch<8>(tm_g, tm_x, tm_z, a, smem, grid, stream); break;
    case 12: rc = launch<12>(tm_g, tm_x, tm_z, a, smem, grid, stream); break;
    case 16: rc = launch<16>(tm_g, tm_x, tm_z, a, smem, grid, stream); break;
    default: rc = fail(DSPB200_ERR_INVALID, "internal: bad state count %d", xp.states);
  }
  if (prof) {
    unsigned long long h[512];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(prof);
    const double nc = h[3] ? static_cast<double>(h[3]) : 1.0, ne = h[8] ? static_cast<double>(h[8]) : 1.0,
                 nm = h[13] ? static_cast<double>(h[13]) : 1.0;
    fprintf(stderr,
            "xz_mma: %d CTAs x %d chunks; cycles per chunk (CTA 0) -- converter: window buffer wait %.0f, boxes + conversion %.0f "
            "(box waits + reads %.0f, wait::st %.0f); "
            "epilogue: accumulator wait %.0f, tensor memory -> registers %.0f, state %.0f (hand-over %.0f), stores %.0f (tile waits %.0f); "
            "mma: window wait %.0f, accumulator wait %.0f, main issue %.0f, state + free response %.0f; "
            "CTA 0 ran %.3f ms at %.0f MHz\n",
            grid, a.n_chunks, h[0] / nc, h[1] / nc, h[2] / nc, h[14] / nc, h[4] / ne, h[5] / ne, h[6] / ne, h[15] / ne, h[7] / ne,
            h[16] / ne, h[9] / nm, h[10] / nm, h[11] / nm,
            h[12] / nm, h[21] * 1e-6, h[21] ? h[20] * 1e3 / static_cast<double>(h[21]) : 0.0);
    unsigned long long t_first = ~0ull;
    for (int c = 0; c < grid && c < 160; ++c) t_first = h[32 + 3 * c] < t_first ? h[32 + 3 * c] : t_first;
    fprintf(stderr, "xz_mma per CTA [cta: sm start end] ms:");
    for (int c = 0; c < grid && c < 160; ++c)
      fprintf(stderr, " [%d: %llu %.2f %.2f]", c, h[34 + 3 * c], (h[32 + 3 * c] - t_first) * 1e-6, (h[33 + 3 * c] - t_first) * 1e-6);
    fprintf(stderr, "\n");
  }
  return rc;
}

}  // namespace dspb200
