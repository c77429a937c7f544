// Library core: error state, device checks, TMA descriptor encoding.
#include "common.cuh"

#include <mutex>

namespace dspb200 {

std::atomic<long long> g_launches{0};

std::string& last_error() {
  static thread_local std::string s;
  return s;
}

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}

namespace {
struct DevCache {
  int valid = 0;
  int sms = 0;
  int smem_optin = 0;
  int cc_major = 0;
};
DevCache g_dev[64];
std::mutex g_dev_mu;

int fill_cache(int dev) {
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (g_dev[dev].valid) return DSPB200_OK;
  cudaDeviceProp prop;
  DSP_CUDA(cudaGetDeviceProperties(&prop, dev));
  g_dev[dev].sms = prop.multiProcessorCount;
  g_dev[dev].smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
  g_dev[dev].cc_major = prop.major;
  g_dev[dev].valid = 1;
  return DSPB200_OK;
}
}  // namespace

int ensure_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(DSPB200_ERR_NO_DEVICE,
                "no CUDA device is available (%s); libdspb200 has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  int dev = 0;
  DSP_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(DSPB200_ERR_NO_DEVICE, "device index %d out of range", dev);
  DSP_TRY(fill_cache(dev));
  if (g_dev[dev].cc_major != 10)
    return fail(DSPB200_ERR_NO_DEVICE,
                "device %d has compute capability %d.x; this library is built for sm_100a only",
                dev, g_dev[dev].cc_major);
  return DSPB200_OK;
}

int check_plan_device(int plan_device, const char* what) {
  if (plan_device < 0) return DSPB200_OK;
  int dev = -1;
  DSP_CUDA(cudaGetDevice(&dev));
  if (dev != plan_device)
    return fail(DSPB200_ERR_INVALID, "%s plan was built on device %d but device %d is current: its tables live on device %d",
                what, plan_device, dev, plan_device);
  return DSPB200_OK;
}

int sm_count() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!g_dev[dev].valid && fill_cache(dev) != DSPB200_OK) return 148;
  return g_dev[dev].sms;
}

int max_smem_optin() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 227 * 1024;
  if (!g_dev[dev].valid && fill_cache(dev) != DSPB200_OK) return 227 * 1024;
  return g_dev[dev].smem_optin;
}

int encode_tmap_2d(CUtensorMap* map, int dtype, const void* base, uint64_t dim0, uint64_t dim1,
                   uint64_t stride1_bytes, uint32_t box0, uint32_t box1, bool swizzle128) {
  return encode_tmap_2d_sw(map, dtype, base, dim0, dim1, stride1_bytes, box0, box1, swizzle128 ? 128 : 0);
}

int encode_tmap_2d_sw(CUtensorMap* map, int dtype, const void* base, uint64_t dim0, uint64_t dim1,
                      uint64_t stride1_bytes, uint32_t box0, uint32_t box1, int swizzle_bytes) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                               const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
    else
      cudaGetLastError();
  });
  if (!fn) return fail(DSPB200_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {dim0, dim1};
  cuuint64_t strides[1] = {stride1_bytes};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = dtype == DSPB200_F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64
                                 : (dtype == kTmapF16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
  CUresult r = fn(map, dt,
                  2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(DSPB200_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return DSPB200_OK;
}

int encode_tmap_2d_f16(CUtensorMap* map, const void* base, uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes,
                       uint32_t box0, uint32_t box1) {
  return encode_tmap_2d_sw(map, kTmapF16, base, dim0, dim1, stride1_bytes, box0, box1, 128);
}

}  // namespace dspb200

using namespace dspb200;

extern "C" {

int dspb200_version(void) { return DSPB200_VERSION; }

const char* dspb200_last_error_string(void) { return last_error().c_str(); }

int64_t dspb200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int dspb200_device_count(int* count) {
  DSP_CHECK(count != nullptr, "count is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    *count = 0;
    return fail(DSPB200_ERR_NO_DEVICE, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
  }
  *count = n;
  return DSPB200_OK;
}

int dspb200_current_device(int* device) {
  DSP_CHECK(device != nullptr, "device is NULL");
  *device = -1;
  DSP_TRY(ensure_device());
  int d = -1;
  DSP_CUDA(cudaGetDevice(&d));
  *device = d;
  return DSPB200_OK;
}

int dspb200_device_info(int device, char* name, int name_len, int* sms, int* cc_major, int* cc_minor,
                        size_t* total_mem) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(DSPB200_ERR_NO_DEVICE, "cudaGetDeviceProperties(%d) failed: %s", device,
                cudaGetErrorString(e));
  }
  if (name && name_len > 0) {
    strncpy(name, prop.name, static_cast<size_t>(name_len) - 1);
    name[name_len - 1] = 0;
  }
  if (sms) *sms = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (total_mem) *total_mem = prop.totalGlobalMem;
  return DSPB200_OK;
}

}  // extern "C"
