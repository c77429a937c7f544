// Shared helpers for libdspb200: error plumbing, launch accounting, small
// device utilities (cp.async, mbarrier, TMA, packed-fp32 FMA).  sm_100a only.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/dspb200.h"

namespace dspb200 {

// ---- error handling ------------------------------------------------------
std::string& last_error();
int fail(int code, const char* fmt, ...);
extern std::atomic<long long> g_launches;

#define DSP_CUDA(call)                                                                  \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess)                                                             \
      return ::dspb200::fail(DSPB200_ERR_CUDA, "%s failed: %s (%s:%d)", #call,          \
                             cudaGetErrorString(e__), __FILE__, __LINE__);              \
  } while (0)

#define DSP_CHECK(cond, ...)                                                            \
  do {                                                                                  \
    if (!(cond)) return ::dspb200::fail(DSPB200_ERR_INVALID, __VA_ARGS__);              \
  } while (0)

#define DSP_TRY(expr)                                                                   \
  do {                                                                                  \
    int rc__ = (expr);                                                                  \
    if (rc__ != DSPB200_OK) return rc__;                                                \
  } while (0)

inline int after_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return fail(DSPB200_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  return DSPB200_OK;
}

// Every plan struct starts with a type tag: an entry point handed the wrong (or a destroyed) handle answers
// DSPB200_ERR_INVALID instead of reading another plan's layout.  Plans that own device tables also record the
// device they were built on; running them with another device current is an error, not a stray pointer.
constexpr uint32_t kMagicSrc = 0x31435253u;   // "SRC1"
constexpr uint32_t kMagicEq = 0x31205145u;    // "EQ 1"
constexpr uint32_t kMagicFft = 0x31544646u;   // "FFT1"
#define DSP_PLAN(p, MAGIC, what)                                                                    \
  do {                                                                                              \
    DSP_CHECK((p) != nullptr, what " plan is NULL");                                                \
    DSP_CHECK((p)->magic == (MAGIC), "handle %p is not a live " what " plan", (const void*)(p));    \
  } while (0)
int check_plan_device(int plan_device, const char* what);   // plan_device < 0: the plan owns no device memory

int ensure_device();            // checks a usable sm_100 device is current
int sm_count();                 // SMs of the current device (cached per device)
int max_smem_optin();           // opt-in shared memory per block

template <typename T> struct DType;
template <> struct DType<float> { static constexpr int id = DSPB200_F32; };
template <> struct DType<double> { static constexpr int id = DSPB200_F64; };

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// TMA descriptor encode (driver entry point fetched through the runtime, so
// the library does not link libcuda).
int encode_tmap_2d(CUtensorMap* map, int dtype, const void* base, uint64_t dim0, uint64_t dim1,
                   uint64_t stride1_bytes, uint32_t box0, uint32_t box1, bool swizzle128 = false);
int encode_tmap_2d_sw(CUtensorMap* map, int dtype, const void* base, uint64_t dim0, uint64_t dim1,
                      uint64_t stride1_bytes, uint32_t box0, uint32_t box1, int swizzle_bytes);   // 0, 64 or 128
constexpr int kTmapF16 = 16;   // internal dtype code of encode_tmap_2d_sw: fp16 tables of the fused chain kernel
int encode_tmap_2d_f16(CUtensorMap* map, const void* base, uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes,
                       uint32_t box0, uint32_t box1);                                              // 128-byte swizzle

#ifdef __CUDACC__
// ---- device utilities ----------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 16-byte async copy global->shared; bytes beyond src_bytes are zero-filled.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// mbarrier + TMA (cp.async.bulk.tensor) -- the Blackwell/Hopper async-proxy path.
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Non-suspending poll (test_wait): for hand-offs between warps of one CTA that
// are expected to be ready within tens of cycles.
__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3}], [%4];\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
// TMA prefetch of a box into L2 only (no shared memory, no completion to wait for)
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];\n" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(smem_u32(smem_src))
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// Packed fp32 FMA (Blackwell FFMA2): d.xy += a.xy * s  with a scalar s that
// ptxas folds into the instruction's broadcast operand (no extra MOV).
__device__ __forceinline__ void ffma2_bcast(float2& d, const float2 a, const float s) {
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  const float2 sv = make_float2(s, s);
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(dd)
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)),
        "l"(*reinterpret_cast<const unsigned long long*>(&sv)));
  d = *reinterpret_cast<float2*>(&dd);
}
// r.xy = a.xy * s + c.xy   and   r.xy = a.xy * s   (scalar s broadcast to both halves)
__device__ __forceinline__ float2 ffma2s(const float2 a, const float s, const float2 c) {
  unsigned long long r;
  const float2 sv = make_float2(s, s);
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(r)
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)),
        "l"(*reinterpret_cast<const unsigned long long*>(&sv)),
        "l"(*reinterpret_cast<const unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 fmul2s(const float2 a, const float s) {
  unsigned long long r;
  const float2 sv = make_float2(s, s);
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(r)
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)),
        "l"(*reinterpret_cast<const unsigned long long*>(&sv)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(dd)
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)),
        "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  d = *reinterpret_cast<float2*>(&dd);
}
#endif  // __CUDACC__

}  // namespace dspb200
