// Bring-up probe for the tcgen05 path: D[128 x N] = A[128 x K] * B[N x K]^T in TF32 with fp32
// accumulation in TMEM.  K-major operands, 128-byte swizzle (TMA boxes of 32 floats), one CTA,
// no pipelining.  Checks (1) descriptor encodings, (2) how the tensor core narrows fp32 to TF32
// (truncation vs rounding), (3) the accuracy of the 3-term split  A_hi*B + A_lo*B + A_hi*B_lo
// expressed as one GEMM with the terms concatenated along K.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/umma_probe.cu -o /tmp/umma_probe -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

constexpr int kM = 128, kN = 256, kBK = 32;

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
  // K-major, 128-byte swizzle: 8-row groups are 1024 bytes apart (SBO), LBO unused, version 1 (sm100)
  const uint64_t addr = s32(p);
  return ((addr >> 4) & 0x3FFF) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
}

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* D, int K) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* sA = reinterpret_cast<float*>(smem);                       // [128][32] swizzled
  float* sB = reinterpret_cast<float*>(smem + kM * kBK * 4);        // [256][32] swizzled
  __shared__ __align__(8) uint64_t full_bar, mma_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&full_bar, 1);
    mbar_init(&mma_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base_s)), "r"(kN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;

  // instruction descriptor: D fp32, A/B tf32, both K-major, N, M
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(kN >> 3) << 17) | (uint32_t(kM >> 4) << 24);
  if (threadIdx.x == 0) {
    const int nkb = K / kBK;
    for (int kb = 0; kb < nkb; ++kb) {
      const uint32_t bytes = (kM + kN) * kBK * 4;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full_bar)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(s32(sA)), "l"((uint64_t)&tmA), "r"(kb * kBK), "r"(0), "r"(s32(&full_bar)) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(s32(sB)), "l"((uint64_t)&tmB), "r"(kb * kBK), "r"(0), "r"(s32(&full_bar)) : "memory");
      mbar_wait(&full_bar, kb & 1);
      asm volatile("tcgen05.fence::after_thread_sync;");
      const uint64_t da = smem_desc_sw128(sA), db = smem_desc_sw128(sB);
      for (int k = 0; k < kBK / 8; ++k) {
        const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                     ::"r"(tmem), "l"(da + uint64_t(k * 2)), "l"(db + uint64_t(k * 2)), "r"(idesc), "r"(acc) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&mma_bar)) : "memory");
      mbar_wait(&mma_bar, kb & 1);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  // epilogue: warp w reads TMEM lanes 32w..32w+31 (rows of D), 32 columns at a time
  const int m = warp * 32 + lane;
  for (int c0 = 0; c0 < kN; c0 += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem + (uint32_t(warp * 32) << 16) + uint32_t(c0);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) D[(size_t)m * kN + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kN));
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float round_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

static int run(EncodeFn enc, const std::vector<float>& A, const std::vector<float>& B, int K, std::vector<float>& D) {
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, (size_t)kM * kN * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, (size_t)kM * kN * 4);
  CUtensorMap tmA, tmB;
  cuuint64_t dimsA[2] = {(cuuint64_t)K, kM}, dimsB[2] = {(cuuint64_t)K, kN};
  cuuint64_t strides[1] = {(cuuint64_t)K * 4};
  cuuint32_t boxA[2] = {kBK, kM}, boxB[2] = {kBK, kN}, es[2] = {1, 1};
  CUresult r1 = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, dimsA, strides, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, dimsB, strides, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r1 || r2) { printf("encode failed %d %d\n", (int)r1, (int)r2); return 1; }
  const int smem = (kM + kN) * kBK * 4 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(tmA, tmB, dD, K);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
  D.resize((size_t)kM * kN);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return 0;
}

int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fp;
  const int K = 64;
  std::mt19937 rng(1);
  std::uniform_real_distribution<float> u(-1.f, 1.f);
  std::vector<float> A((size_t)kM * K), B((size_t)kN * K), D;
  for (auto& v : A) v = u(rng);
  for (auto& v : B) v = u(rng);
  if (run(enc, A, B, K, D)) return 1;
  double e_trunc = 0, e_round = 0, e_full = 0, ref_max = 0;
  for (int m = 0; m < kM; ++m)
    for (int n = 0; n < kN; ++n) {
      double st = 0, sr = 0, sf = 0;
      for (int k = 0; k < K; ++k) {
        const float a = A[(size_t)m * K + k], b = B[(size_t)n * K + k];
        st += (double)trunc_tf32(a) * trunc_tf32(b);
        sr += (double)round_tf32(a) * round_tf32(b);
        sf += (double)a * b;
      }
      const double d = D[(size_t)m * kN + n];
      e_trunc = fmax(e_trunc, fabs(d - st)); e_round = fmax(e_round, fabs(d - sr)); e_full = fmax(e_full, fabs(d - sf));
      ref_max = fmax(ref_max, fabs(sf));
    }
  printf("single TF32 GEMM K=%d: max|D-ref| vs truncated-input ref %.3e, rounded-input ref %.3e, exact ref %.3e (max|ref| %.3f)\n",
         K, e_trunc, e_round, e_full, ref_max);
  printf("D[0][0..3] = %g %g %g %g ; D[127][255] = %g\n", D[0], D[1], D[2], D[3], D[(size_t)127 * kN + 255]);

  // 3-term split as one GEMM of depth 3K:  [A_hi | A_lo | A_hi] . [B | B | B_lo]
  std::vector<float> A3((size_t)kM * 3 * K), B3((size_t)kN * 3 * K);
  for (int m = 0; m < kM; ++m)
    for (int k = 0; k < K; ++k) {
      const float a = A[(size_t)m * K + k], hi = trunc_tf32(a), lo = a - hi;
      A3[(size_t)m * 3 * K + k] = hi; A3[(size_t)m * 3 * K + K + k] = lo; A3[(size_t)m * 3 * K + 2 * K + k] = hi;
    }
  for (int n = 0; n < kN; ++n)
    for (int k = 0; k < K; ++k) {
      const float b = B[(size_t)n * K + k], hi = trunc_tf32(b), lo = b - hi;
      B3[(size_t)n * 3 * K + k] = b; B3[(size_t)n * 3 * K + K + k] = b; B3[(size_t)n * 3 * K + 2 * K + k] = lo;
    }
  if (run(enc, A3, B3, 3 * K, D)) return 1;
  double e3 = 0;
  for (int m = 0; m < kM; ++m)
    for (int n = 0; n < kN; ++n) {
      double sf = 0;
      for (int k = 0; k < K; ++k) sf += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
      e3 = fmax(e3, fabs(D[(size_t)m * kN + n] - sf));
    }
  printf("3-term split (raw B as the hi operand): max|D-exact| %.3e  (relative to max|ref|: %.3e)\n", e3, e3 / ref_max);
  return 0;
}
