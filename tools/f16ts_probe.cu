// Bring-up probe for the fused SRC->EQ kernel (csrc/xz_mma.cu).  One CTA, no pipelining.  Checks
//  (1) tcgen05.mma kind::f16 with the A operand in tensor memory: rows = lanes, two fp16 per 32-bit column
//      (k even in the low half), written with tcgen05.st by the thread that owns the lane;
//  (2) the accuracy of the three-product fp16 split  A_hi B_hi + A_hi B_lo + A_lo B_hi  (fp32 accumulation);
//  (3) N-trimmed MMAs: accumulator column offset + B row offset (8-row groups), N a multiple of 16;
//  (4) TMA loads of fp32 boxes whose start coordinate is not a multiple of 4 samples (and negative), with the
//      128-byte swizzle [32 samples x 128 rows] and the 64-byte swizzle [16 samples x 128 rows].
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/f16ts_probe.cu -o tools/bin/f16ts_probe
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

constexpr int kM = 128, kN = 96, kK = 64;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n));
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t phase) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}"
                 : "=r"(ok) : "r"(s32(b)), "r"(phase) : "memory");
}
__device__ __forceinline__ uint64_t smem_desc_sw128(const void* p) {
  const uint64_t addr = s32(p);
  return ((addr >> 4) & 0x3FFF) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t* v, uint32_t taddr) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {   // low half = first argument
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// D[128 x 96] (+)= A[128 x 64] . B[96 x 64]^T ;  mode 0: full N; mode 1: k-step j only touches rows 16 j .. 95
__global__ void __launch_bounds__(128, 1)
probe_mma(const __grid_constant__ CUtensorMap tmB, const float* A, float* D, int mode) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sBh = smem;                    // [96 rows][64 fp16] swizzled, 12 KB
  unsigned char* sBl = smem + kN * 128;
  __shared__ __align__(8) uint64_t full_bar, mma_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&full_bar, 1);
    mbar_init(&mma_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  const uint32_t colAh = 0, colAl = kK / 2, colD = 128;
  // thread = row of A: split into fp16 hi/lo pairs and store to tensor memory
  {
    const int m = warp * 32 + lane;
    const uint32_t lane_base = tmem + (uint32_t(warp * 32) << 16);
    for (int c0 = 0; c0 < kK / 2; c0 += 8) {
      uint32_t h[8], l[8];
      for (int j = 0; j < 8; ++j) {
        const float a0 = A[m * kK + 2 * (c0 + j)], a1 = A[m * kK + 2 * (c0 + j) + 1];
        h[j] = pack_h2(a0, a1);
        const __half2 hh = *reinterpret_cast<__half2*>(&h[j]);
        const float2 hf = __half22float2(hh);
        l[j] = pack_h2(a0 - hf.x, a1 - hf.y);
      }
      tmem_st8(lane_base + colAh + c0, h);
      tmem_st8(lane_base + colAl + c0, l);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full_bar)), "r"(2 * kN * 128) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(s32(sBh)), "l"((uint64_t)&tmB), "r"(0), "r"(0), "r"(s32(&full_bar)) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(s32(sBl)), "l"((uint64_t)&tmB), "r"(0), "r"(kN), "r"(s32(&full_bar)) : "memory");
    mbar_wait(&full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    auto idesc_n = [](int n) -> uint32_t {   // D fp32, A/B fp16, K-major, M = 128
      return (1u << 4) | (0u << 7) | (0u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(kM >> 4) << 24);
    };
    auto mma = [&](uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                   "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
                   ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    };
    for (int j = 0; j < kK / 16; ++j) {
      const int r0 = mode ? 16 * j : 0;           // first B row / accumulator column this k-step touches
      const uint32_t id = idesc_n(kN - r0);
      const uint64_t bh = smem_desc_sw128(sBh + r0 * 128) + 2 * j, bl = smem_desc_sw128(sBl + r0 * 128) + 2 * j;
      mma(tmem + colD + r0, tmem + colAh + 8 * j, bh, id, j > 0);
      mma(tmem + colD + r0, tmem + colAh + 8 * j, bl, id, 1);
      mma(tmem + colD + r0, tmem + colAl + 8 * j, bh, id, 1);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&mma_bar)) : "memory");
    mbar_wait(&mma_bar, 0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const int m = warp * 32 + lane;
  for (int c0 = 0; c0 < kN; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(v, tmem + (uint32_t(warp * 32) << 16) + colD + c0);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) D[(size_t)m * kN + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// dump of a TMA-loaded box: raw shared memory image, the host undoes the swizzle
__global__ void probe_tma(const __grid_constant__ CUtensorMap tm, int c0, int c1, int bytes, float* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(s32(smem)), "l"((uint64_t)&tm), "r"(c0), "r"(c1), "r"(s32(&bar)) : "memory");
  }
  mbar_wait(&bar, 0);
  const float* s = reinterpret_cast<const float*>(smem);
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = s[i];
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int test_mma(EncodeFn enc, int mode, float amp_a) {
  std::mt19937 rng(7 + mode);
  std::uniform_real_distribution<float> u(-1.f, 1.f);
  std::vector<float> A((size_t)kM * kK), B((size_t)kN * kK);
  for (auto& v : A) v = amp_a * u(rng);
  for (auto& v : B) v = 4096.f * u(rng) * std::pow(10.f, -3.f * std::fabs(u(rng)));
  if (mode)   // trimmed: k-step j may only feed rows >= 16 j, so the others must be zero for the result to be A.B^T
    for (int n = 0; n < kN; ++n)
      for (int k = 0; k < kK; ++k)
        if (n < 16 * (k / 16)) B[(size_t)n * kK + k] = 0.f;
  std::vector<__half> Bs((size_t)2 * kN * kK);
  for (size_t i = 0; i < B.size(); ++i) {
    const __half h = __float2half_rn(B[i]);
    Bs[i] = h;
    Bs[B.size() + i] = __float2half_rn(B[i] - __half2float(h));
  }
  float *dA, *dD; __half* dB;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, Bs.size() * 2); cudaMalloc(&dD, (size_t)kM * kN * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bs.data(), Bs.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, (size_t)kM * kN * 4);
  CUtensorMap tmB;
  cuuint64_t dims[2] = {(cuuint64_t)kK, (cuuint64_t)2 * kN};
  cuuint64_t strides[1] = {(cuuint64_t)kK * 2};
  cuuint32_t box[2] = {kK, kN}, es[2] = {1, 1};
  CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r) { printf("encode failed %d\n", (int)r); return 1; }
  const int smem = 2 * kN * 128 + 1024;
  cudaFuncSetAttribute(probe_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_mma<<<1, 128, smem>>>(tmB, dA, dD, mode);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("probe_mma mode %d failed: %s\n", mode, cudaGetErrorString(e)); return 1; }
  std::vector<float> D((size_t)kM * kN);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double emax = 0, rmax = 0, e1 = 0;
  for (int m = 0; m < kM; ++m)
    for (int n = 0; n < kN; ++n) {
      double s = 0, s1 = 0;
      for (int k = 0; k < kK; ++k) {
        s += (double)A[(size_t)m * kK + k] * B[(size_t)n * kK + k];
        s1 += (double)__half2float(__float2half_rn(A[(size_t)m * kK + k])) * __half2float(__float2half_rn(B[(size_t)n * kK + k]));
      }
      emax = fmax(emax, fabs(D[(size_t)m * kN + n] - s));
      e1 = fmax(e1, fabs(s1 - s));
      rmax = fmax(rmax, fabs(s));
    }
  printf("f16 TS mma, mode %d, |A| <= %g: max|D - exact| = %.3e (relative to max|ref| %.3e: %.3e); one-product fp16 would give %.3e\n",
         mode, amp_a, emax, rmax, emax / rmax, e1 / rmax);
  printf("   D[0][0..3] = %g %g %g %g ; D[127][95] = %g\n", D[0], D[1], D[2], D[3], D[(size_t)127 * kN + 95]);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return 0;
}

static int test_tma(EncodeFn enc, int box0, int sw_bytes, int c0, int c1) {
  const int N = 1000, C = 300, rows = 128;
  std::vector<float> h((size_t)N * C);
  for (int c = 0; c < C; ++c)
    for (int i = 0; i < N; ++i) h[(size_t)c * N + i] = (float)(i + 1) + 4096.f * (float)c;
  float *x, *out;
  cudaMalloc(&x, h.size() * 4);
  cudaMemcpy(x, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  const int bytes = box0 * rows * 4;
  cudaMalloc(&out, bytes);
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)C};
  cuuint64_t strides[1] = {(cuuint64_t)N * 4};
  cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)rows}, es[2] = {1, 1};
  const CUtensorMapSwizzle sw = sw_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (sw_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE);
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r) { printf("tma encode failed %d\n", (int)r); return 1; }
  cudaFuncSetAttribute(probe_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes + 1024);
  probe_tma<<<1, 128, bytes + 1024>>>(tm, c0, c1, bytes, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("probe_tma box %d sw %d c0 %d failed: %s\n", box0, sw_bytes, c0, cudaGetErrorString(e)); return 1; }
  std::vector<float> o((size_t)box0 * rows);
  cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost);
  // hypotheses: 16-byte chunk index XORed with (row % 8) [128B], ((row / 2) % 4) [64B], none
  const int row_bytes = box0 * 4, chunks = row_bytes / 16;
  long bad = 0;
  for (int rr = 0; rr < rows; ++rr)
    for (int i = 0; i < box0; ++i) {
      const int ch = i / 4;
      int phys = ch;
      if (sw_bytes == 128) phys = ch ^ (rr % 8);
      else if (sw_bytes == 64) phys = ch ^ ((rr / 2) % 4);
      const float got = o[(size_t)rr * box0 + phys * 4 + (i % 4)];
      const int gi = c0 + i, gc = c1 + rr;
      const float want = (gi < 0 || gi >= N || gc < 0 || gc >= C) ? 0.f : (float)(gi + 1) + 4096.f * (float)gc;
      if (got != want) {
        if (bad < 3) printf("   mismatch row %d i %d: got %g want %g\n", rr, i, got, want);
        ++bad;
      }
    }
  printf("TMA box [%d x %d] swizzle %d at (%d, %d): %s (%ld mismatches, %d chunks/row)\n", box0, rows, sw_bytes, c0, c1,
         bad ? "MISMATCH" : "ok", bad, chunks);
  cudaFree(x); cudaFree(out);
  return 0;
}

int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fp;
  int rc = 0;
  rc |= test_mma(enc, 0, 64.f);
  rc |= test_mma(enc, 1, 64.f);
  rc |= test_mma(enc, 1, 0.01f);
  rc |= test_tma(enc, 32, 128, 0, 0);
  rc |= test_tma(enc, 32, 128, 5, 3);
  rc |= test_tma(enc, 32, 128, -3, 200);
  rc |= test_tma(enc, 32, 128, 990, 0);
  rc |= test_tma(enc, 16, 64, 0, 0);
  rc |= test_tma(enc, 16, 64, 7, 1);
  rc |= test_tma(enc, 16, 64, -5, 250);
  printf(rc ? "PROBE FAILED\n" : "PROBE DONE\n");
  return rc;
}
