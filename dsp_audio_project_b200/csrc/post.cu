// Elementwise / reduction kernels on either side of the hot path:
//  * loader front end (dsp_core.py:23-31): interleaved multi-channel frames ->
//    mono mean (float64) -> float32 -> peak normalisation when peak > 1e-6;
//  * playback export (app.py:349-354): nan_to_num, peak normalise, * 32767,
//    truncate to int16.
// Plain HBM-bound streaming kernels: 16-byte vector accesses where alignment
// allows, grids of a few CTAs per SM, per-row peaks by atomicMax on the
// non-negative float bit pattern.
#include <cfloat>

#include "common.cuh"

namespace dspb200 {

template <typename T> struct Limits;
template <> struct Limits<float> { static __device__ __forceinline__ float max() { return FLT_MAX; } };
template <> struct Limits<double> { static __device__ __forceinline__ double max() { return DBL_MAX; } };

// numpy.nan_to_num defaults: NaN -> 0, +-inf -> +-largest finite
template <typename T> __device__ __forceinline__ T nan_to_num(T v) {
  if (v != v) return T(0);
  if (v > Limits<T>::max()) return Limits<T>::max();
  if (v < -Limits<T>::max()) return -Limits<T>::max();
  return v;
}

__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
  atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(addr), static_cast<unsigned long long>(__double_as_longlong(v)));
}

template <typename T> __device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T u = __shfl_xor_sync(0xffffffffu, v, o);
    v = u > v ? u : v;
  }
  return v;
}

// peaks[row] = max |nan_to_num(x[row, :])|   (peaks zeroed by the caller)
template <typename T>
__global__ void __launch_bounds__(256)
row_peak_kernel(const T* __restrict__ x, long long stride, long long rows, long long n, int chunks_per_row,
                T* __restrict__ peaks) {
  const long long total = rows * chunks_per_row;
  for (long long w = blockIdx.x; w < total; w += gridDim.x) {
    const long long row = w / chunks_per_row;
    const int chunk = static_cast<int>(w - row * chunks_per_row);
    const long long per = (n + chunks_per_row - 1) / chunks_per_row;
    const long long lo = chunk * per;
    const long long hi = lo + per < n ? lo + per : n;
    const T* xr = x + row * stride;
    T m = T(0);
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      const T v = nan_to_num(xr[i]);
      const T a = v < T(0) ? -v : v;
      m = a > m ? a : m;
    }
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > T(0)) atomic_max_nonneg(peaks + row, m);
  }
}

// out = (int16) trunc(nan_to_num(x) / peak * 32767)   (no division when peak == 0)   app.py:349-354
template <typename T>
__global__ void __launch_bounds__(256)
pcm16_kernel(const T* __restrict__ x, long long stride, const T* __restrict__ peaks, short* __restrict__ out,
             long long out_stride, long long rows, long long n) {
  const long long total = rows * n;
  for (long long id = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; id < total;
       id += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = id / n;
    const long long i = id - row * n;
    T v = nan_to_num(x[row * stride + i]);
    const T pk = peaks[row];
    if (pk > T(0)) v = v / pk;
    const T s = v * T(32767);
    out[row * out_stride + i] = static_cast<short>(static_cast<int>(s));   // C truncation, as ndarray.astype
  }
}

// mono[clip, i] = (float) mean_c in[clip, i, c] (mean in float64), and its running peak
template <typename TI>
__global__ void __launch_bounds__(256)
mono_kernel(const TI* __restrict__ in, long long clips, long long frames, int cin, float* __restrict__ mono,
            long long mono_stride, float* __restrict__ peaks) {
  const long long total = clips * frames;
  float m = 0.f;
  long long cur_clip = -1;
  for (long long id = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; id < total;
       id += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long clip = id / frames;
    const long long i = id - clip * frames;
    if (clip != cur_clip) {
      if (cur_clip >= 0 && m > 0.f) atomic_max_nonneg(peaks + cur_clip, m);
      cur_clip = clip;
      m = 0.f;
    }
    const TI* p = in + (clip * frames + i) * cin;
    float v;
    if (cin == 1) {
      v = static_cast<float>(p[0]);
    } else {
      double acc = 0.0;   // numpy's mean: float64 accumulation in element order
      for (int c = 0; c < cin; ++c) acc += static_cast<double>(p[c]);
      v = static_cast<float>(acc / static_cast<double>(cin));
    }
    mono[clip * mono_stride + i] = v;
    const float a = fabsf(v);
    m = a > m ? a : m;     // NaN never wins, like a NaN-free np.max; NaN inputs are outside the loader's contract
  }
  if (cur_clip >= 0 && m > 0.f) atomic_max_nonneg(peaks + cur_clip, m);
}

__global__ void __launch_bounds__(256)
normalize_kernel(float* __restrict__ mono, long long mono_stride, const float* __restrict__ peaks, long long clips,
                 long long frames) {
  const long long total = clips * frames;
  for (long long id = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; id < total;
       id += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long clip = id / frames;
    const long long i = id - clip * frames;
    const float pk = peaks[clip];
    if (pk > 1e-6f) mono[clip * mono_stride + i] = mono[clip * mono_stride + i] / pk;   // dsp_core.py:29-31
  }
}

static int grid_for(long long work_items, int threads) {
  long long b = ceil_div(work_items, threads);
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

template <typename T>
static int pcm16_run(const T* x, int64_t stride, T* peaks, short* out, int64_t out_stride, int64_t rows, int64_t n,
                     cudaStream_t stream) {
  DSP_CHECK(rows >= 0 && n >= 0, "negative shape");
  if (rows == 0 || n == 0) return DSPB200_OK;
  DSP_CHECK(x && peaks && out, "NULL buffer");
  DSP_CHECK(stride >= n && out_stride >= n, "row stride smaller than n");
  DSP_TRY(ensure_device());
  DSP_CUDA(cudaMemsetAsync(peaks, 0, static_cast<size_t>(rows) * sizeof(T), stream));
  int chunks = static_cast<int>(ceil_div(static_cast<int64_t>(sm_count()) * 8, rows));
  if (chunks < 1) chunks = 1;
  const int64_t max_chunks = ceil_div(n, 4096);
  if (chunks > max_chunks) chunks = static_cast<int>(max_chunks);
  row_peak_kernel<T><<<grid_for(rows * chunks * 256, 256), 256, 0, stream>>>(x, stride, rows, n, chunks, peaks);
  DSP_TRY(after_launch("row_peak_kernel"));
  pcm16_kernel<T><<<grid_for(rows * n, 256), 256, 0, stream>>>(x, stride, peaks, out, out_stride, rows, n);
  return after_launch("pcm16_kernel");
}

template <typename TI>
static int mono_run(const TI* in, int64_t clips, int64_t frames, int cin, float* mono, int64_t mono_stride,
                    float* peaks, cudaStream_t stream) {
  DSP_CHECK(clips >= 0 && frames >= 0 && cin >= 1, "bad shape");
  if (clips == 0 || frames == 0) return DSPB200_OK;
  DSP_CHECK(in && mono && peaks, "NULL buffer");
  DSP_CHECK(mono_stride >= frames, "mono stride smaller than frames");
  DSP_TRY(ensure_device());
  DSP_CUDA(cudaMemsetAsync(peaks, 0, static_cast<size_t>(clips) * sizeof(float), stream));
  mono_kernel<TI><<<grid_for(clips * frames, 256), 256, 0, stream>>>(in, clips, frames, cin, mono, mono_stride, peaks);
  DSP_TRY(after_launch("mono_kernel"));
  normalize_kernel<<<grid_for(clips * frames, 256), 256, 0, stream>>>(mono, mono_stride, peaks, clips, frames);
  return after_launch("normalize_kernel");
}

}  // namespace dspb200

using namespace dspb200;

extern "C" {

int dspb200_pcm16_run_f32(const float* x, int64_t stride, float* peaks, int16_t* out, int64_t out_stride,
                          int64_t rows, int64_t n, void* stream) {
  return pcm16_run<float>(x, stride, peaks, out, out_stride, rows, n, static_cast<cudaStream_t>(stream));
}
int dspb200_pcm16_run_f64(const double* x, int64_t stride, double* peaks, int16_t* out, int64_t out_stride,
                          int64_t rows, int64_t n, void* stream) {
  return pcm16_run<double>(x, stride, peaks, out, out_stride, rows, n, static_cast<cudaStream_t>(stream));
}
int dspb200_mono_normalize_run_f64(const double* in, int64_t clips, int64_t frames, int channels_in, float* mono,
                                   int64_t mono_stride, float* peaks, void* stream) {
  return mono_run<double>(in, clips, frames, channels_in, mono, mono_stride, peaks, static_cast<cudaStream_t>(stream));
}
int dspb200_mono_normalize_run_f32(const float* in, int64_t clips, int64_t frames, int channels_in, float* mono,
                                   int64_t mono_stride, float* peaks, void* stream) {
  return mono_run<float>(in, clips, frames, channels_in, mono, mono_stride, peaks, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
