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

// Four consecutive elements per thread and iteration: one 16-byte (fp32) or two
// 16-byte (fp64) loads when the row is 16-byte aligned, scalar loads otherwise.
template <typename T, bool kVec> struct Quad;
template <bool kVec> struct Quad<float, kVec> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    if (kVec) {
      const float4 q = *reinterpret_cast<const float4*>(p);
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
      v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; v[3] = p[3];
    }
  }
};
template <bool kVec> struct Quad<double, kVec> {
  static __device__ __forceinline__ void load(const double* p, double (&v)[4]) {
    if (kVec) {
      const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
      v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; v[3] = p[3];
    }
  }
};

// peaks[row] = max |nan_to_num(x[row, :])|   (peaks zeroed by the caller)
// grid: (chunks along the row, rows), 256 threads, each chunk a whole number of quads
template <typename T, bool kVec>
__global__ void __launch_bounds__(256)
row_peak_kernel(const T* __restrict__ x, long long stride, long long rows, long long n, long long chunk,
                T* __restrict__ peaks) {
  for (long long row = blockIdx.y; row < rows; row += gridDim.y) {
    const T* xr = x + row * stride;
    const long long lo = static_cast<long long>(blockIdx.x) * chunk;
    const long long hi = lo + chunk < n ? lo + chunk : n;
    T m = T(0);
    long long i = lo + 4LL * threadIdx.x;
#pragma unroll 4
    for (; i + 3 < hi; i += 4LL * blockDim.x) {
      T v[4];
      Quad<T, kVec>::load(xr + i, v);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const T w = nan_to_num(v[k]);
        const T a = w < T(0) ? -w : w;
        m = a > m ? a : m;
      }
    }
    for (; i < hi; ++i) {          // at most one partial quad per chunk (the row tail)
      const T w = nan_to_num(xr[i]);
      const T a = w < T(0) ? -w : w;
      m = a > m ? a : m;
    }
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > T(0)) atomic_max_nonneg(peaks + row, m);
  }
}

// out = (int16) trunc(nan_to_num(x) / peak * 32767)   (no division when peak == 0)   app.py:349-354
template <typename T, bool kVec>
__global__ void __launch_bounds__(256)
pcm16_kernel(const T* __restrict__ x, long long stride, const T* __restrict__ peaks, short* __restrict__ out,
             long long out_stride, long long rows, long long n, long long chunk) {
  for (long long row = blockIdx.y; row < rows; row += gridDim.y) {
    const T* xr = x + row * stride;
    short* orow = out + row * out_stride;
    const T pk = peaks[row];
    const bool scale = pk > T(0);
    const long long lo = static_cast<long long>(blockIdx.x) * chunk;
    const long long hi = lo + chunk < n ? lo + chunk : n;
    auto quant = [&](T v) -> short {
      v = nan_to_num(v);
      if (scale) v = v / pk;
      return static_cast<short>(static_cast<int>(v * T(32767)));   // C truncation, as ndarray.astype
    };
    long long i = lo + 4LL * threadIdx.x;
#pragma unroll 4
    for (; i + 3 < hi; i += 4LL * blockDim.x) {
      T v[4];
      Quad<T, kVec>::load(xr + i, v);
      const short s0 = quant(v[0]), s1 = quant(v[1]), s2 = quant(v[2]), s3 = quant(v[3]);
      if (kVec) {
        uint2 pkd;
        pkd.x = (static_cast<unsigned>(static_cast<unsigned short>(s1)) << 16) | static_cast<unsigned short>(s0);
        pkd.y = (static_cast<unsigned>(static_cast<unsigned short>(s3)) << 16) | static_cast<unsigned short>(s2);
        *reinterpret_cast<uint2*>(orow + i) = pkd;
      } else {
        orow[i] = s0; orow[i + 1] = s1; orow[i + 2] = s2; orow[i + 3] = s3;
      }
    }
    for (; i < hi; ++i) orow[i] = quant(xr[i]);
  }
}

// mono[clip, i] = (float) mean_c in[clip, i, c] (mean in float64) and the clip's running peak.
// grid: (chunks along the clip, clips)
template <typename TI, int CIN>
__global__ void __launch_bounds__(256)
mono_kernel(const TI* __restrict__ in, long long clips, long long frames, int cin_rt, float* __restrict__ mono,
            long long mono_stride, float* __restrict__ peaks, long long chunk) {
  const int cin = CIN > 0 ? CIN : cin_rt;
  for (long long clip = blockIdx.y; clip < clips; clip += gridDim.y) {
    const TI* src = in + clip * frames * cin;
    float* dst = mono + clip * mono_stride;
    const long long lo = static_cast<long long>(blockIdx.x) * chunk;
    const long long hi = lo + chunk < frames ? lo + chunk : frames;
    float m = 0.f;
#pragma unroll 4
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      float v;
      if (CIN == 1) {
        v = static_cast<float>(src[i]);
      } else if (CIN == 2 && sizeof(TI) == 8) {
        const double2 p = *reinterpret_cast<const double2*>(src + 2 * i);     // one 16-byte load per frame
        v = static_cast<float>((p.x + p.y) / 2.0);
      } else if (CIN == 2) {
        const float2 p = *reinterpret_cast<const float2*>(src + 2 * i);
        v = static_cast<float>((static_cast<double>(p.x) + static_cast<double>(p.y)) / 2.0);
      } else {
        double acc = 0.0;      // numpy's mean: float64 accumulation in element order
        for (int c = 0; c < cin; ++c) acc += static_cast<double>(src[i * cin + c]);
        v = static_cast<float>(acc / static_cast<double>(cin));
      }
      dst[i] = v;
      const float a = fabsf(v);
      m = a > m ? a : m;
    }
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomic_max_nonneg(peaks + clip, m);
  }
}

template <bool kVec>
__global__ void __launch_bounds__(256)
normalize_kernel(float* __restrict__ mono, long long mono_stride, const float* __restrict__ peaks, long long clips,
                 long long frames, long long chunk) {
  for (long long clip = blockIdx.y; clip < clips; clip += gridDim.y) {
    const float pk = peaks[clip];
    if (!(pk > 1e-6f)) continue;                                   // dsp_core.py:29-31
    float* row = mono + clip * mono_stride;
    const long long lo = static_cast<long long>(blockIdx.x) * chunk;
    const long long hi = lo + chunk < frames ? lo + chunk : frames;
    long long i = lo + 4LL * threadIdx.x;
#pragma unroll 4
    for (; i + 3 < hi; i += 4LL * blockDim.x) {
      if (kVec) {
        float4 q = *reinterpret_cast<float4*>(row + i);
        q.x = q.x / pk; q.y = q.y / pk; q.z = q.z / pk; q.w = q.w / pk;
        *reinterpret_cast<float4*>(row + i) = q;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) row[i + k] = row[i + k] / pk;
      }
    }
    for (; i < hi; ++i) row[i] = row[i] / pk;
  }
}

// chunk length (multiple of 1024 elements) and 2-D grid giving ~16 CTAs per SM
static void tile_rows(long long rows, long long n, long long& chunk, dim3& grid) {
  const long long want = static_cast<long long>(sm_count()) * 16;
  long long per_row = ceil_div(want, rows > 0 ? rows : 1);
  const long long max_chunks = ceil_div(n, 4096);
  if (per_row > max_chunks) per_row = max_chunks;
  if (per_row < 1) per_row = 1;
  chunk = round_up(ceil_div(n, per_row), 1024);
  per_row = ceil_div(n, chunk);
  grid = dim3(static_cast<unsigned>(per_row), static_cast<unsigned>(rows < 65535 ? rows : 65535), 1);
}

template <typename T>
int pcm16_run(const T* x, int64_t stride, T* peaks, short* out, int64_t out_stride, int64_t rows, int64_t n,
              cudaStream_t stream) {
  DSP_CHECK(rows >= 0 && n >= 0, "negative shape");
  if (rows == 0 || n == 0) return DSPB200_OK;
  DSP_CHECK(x && peaks && out, "NULL buffer");
  DSP_CHECK(stride >= n && out_stride >= n, "row stride smaller than n");
  DSP_TRY(ensure_device());
  DSP_CUDA(cudaMemsetAsync(peaks, 0, static_cast<size_t>(rows) * sizeof(T), stream));
  long long chunk;
  dim3 grid;
  tile_rows(rows, n, chunk, grid);
  const int vec = 16 / static_cast<int>(sizeof(T));
  const bool vin = reinterpret_cast<uintptr_t>(x) % 16 == 0 && stride % vec == 0;
  const bool vout = reinterpret_cast<uintptr_t>(out) % 8 == 0 && out_stride % 4 == 0;
  if (vin) row_peak_kernel<T, true><<<grid, 256, 0, stream>>>(x, stride, rows, n, chunk, peaks);
  else row_peak_kernel<T, false><<<grid, 256, 0, stream>>>(x, stride, rows, n, chunk, peaks);
  DSP_TRY(after_launch("row_peak_kernel"));
  if (vin && vout) pcm16_kernel<T, true><<<grid, 256, 0, stream>>>(x, stride, peaks, out, out_stride, rows, n, chunk);
  else pcm16_kernel<T, false><<<grid, 256, 0, stream>>>(x, stride, peaks, out, out_stride, rows, n, chunk);
  return after_launch("pcm16_kernel");
}
template int pcm16_run<float>(const float*, int64_t, float*, short*, int64_t, int64_t, int64_t, cudaStream_t);
template int pcm16_run<double>(const double*, int64_t, double*, short*, int64_t, int64_t, int64_t, cudaStream_t);


template <typename TI>
static int mono_run(const TI* in, int64_t clips, int64_t frames, int cin, float* mono, int64_t mono_stride,
                    float* peaks, cudaStream_t stream) {
  DSP_CHECK(clips >= 0 && frames >= 0 && cin >= 1, "bad shape");
  if (clips == 0 || frames == 0) return DSPB200_OK;
  DSP_CHECK(in && mono && peaks, "NULL buffer");
  DSP_CHECK(mono_stride >= frames, "mono stride smaller than frames");
  DSP_TRY(ensure_device());
  DSP_CUDA(cudaMemsetAsync(peaks, 0, static_cast<size_t>(clips) * sizeof(float), stream));
  long long chunk;
  dim3 grid;
  tile_rows(clips, frames, chunk, grid);
  const bool pair_ok = reinterpret_cast<uintptr_t>(in) % (2 * sizeof(TI)) == 0;
  if (cin == 1) mono_kernel<TI, 1><<<grid, 256, 0, stream>>>(in, clips, frames, cin, mono, mono_stride, peaks, chunk);
  else if (cin == 2 && pair_ok) mono_kernel<TI, 2><<<grid, 256, 0, stream>>>(in, clips, frames, cin, mono, mono_stride, peaks, chunk);
  else mono_kernel<TI, 0><<<grid, 256, 0, stream>>>(in, clips, frames, cin, mono, mono_stride, peaks, chunk);
  DSP_TRY(after_launch("mono_kernel"));
  const bool vec = reinterpret_cast<uintptr_t>(mono) % 16 == 0 && mono_stride % 4 == 0;
  if (vec) normalize_kernel<true><<<grid, 256, 0, stream>>>(mono, mono_stride, peaks, clips, frames, chunk);
  else normalize_kernel<false><<<grid, 256, 0, stream>>>(mono, mono_stride, peaks, clips, frames, chunk);
  return after_launch("normalize_kernel");
}

}  // namespace dspb200

using namespace dspb200;

// ---- synthetic clips for the throughput configurations ("generated on device per wave") -------------
// x[c, i] = lo + (hi - lo) * u,  u = top 24 bits of splitmix64(seed + golden * (c * n + i + 1)) / 2^24: counter based, so a
// wave is reproducible on the host (tests) whatever the launch shape.
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// lo + span * u with two roundings (no contraction into an FMA), so that the numpy twin reproduces it bit for bit
__device__ __forceinline__ float affine_rn(float lo, float span, float u) { return __fadd_rn(lo, __fmul_rn(span, u)); }
__device__ __forceinline__ double affine_rn(double lo, double span, double u) { return __dadd_rn(lo, __dmul_rn(span, u)); }
template <typename T>
__global__ void __launch_bounds__(128)
generate_uniform_kernel(T* __restrict__ x, long long stride, long long channels, long long n, long long first_channel,
                        unsigned long long seed, T lo, T span) {
  const long long quads = (n + 3) / 4;
  const unsigned long long pairs_per_row = static_cast<unsigned long long>((n + 1) / 2);
  for (long long c = blockIdx.y; c < channels; c += gridDim.y) {
    T* row = x + c * stride;
    // one 64-bit hash per PAIR of samples (two 64-bit multiplies each: per sample they were as long as the store)
    const unsigned long long base = static_cast<unsigned long long>(first_channel + c) * pairs_per_row;
    for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < quads; q += static_cast<long long>(gridDim.x) * blockDim.x) {
      T v[4];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const unsigned long long h = splitmix64(seed + 0x9E3779B97F4A7C15ull * (base + static_cast<unsigned long long>(2 * q + e) + 1ull));
        v[2 * e] = affine_rn(lo, span, static_cast<T>(static_cast<unsigned>(h >> 40)) * static_cast<T>(1.0 / 16777216.0));
        v[2 * e + 1] = affine_rn(lo, span, static_cast<T>(static_cast<unsigned>(h & 0xffffffffull) >> 8) * static_cast<T>(1.0 / 16777216.0));
      }
      if (4 * q + 3 < n && (reinterpret_cast<uintptr_t>(row + 4 * q) % (4 * sizeof(T))) == 0) {
        if (sizeof(T) == 4) {
          *reinterpret_cast<float4*>(row + 4 * q) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
          *reinterpret_cast<double2*>(row + 4 * q) = make_double2(v[0], v[1]);
          *reinterpret_cast<double2*>(row + 4 * q + 2) = make_double2(v[2], v[3]);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (4 * q + e < n) row[4 * q + e] = v[e];
      }
    }
  }
}

template <typename T>
static int generate_run(T* x, int64_t stride, int64_t channels, int64_t n, int64_t first_channel, uint64_t seed, double lo,
                        double hi, cudaStream_t stream) {
  DSP_CHECK(channels >= 0 && n >= 0 && first_channel >= 0, "negative shape");
  if (channels == 0 || n == 0) return DSPB200_OK;
  DSP_CHECK(x != nullptr, "NULL buffer");
  DSP_CHECK(stride >= n, "channel stride smaller than n");
  DSP_TRY(ensure_device());
  const long long quads = (n + 3) / 4;
  const long long per_row = ceil_div(quads, 128);
  // a resident grid (16 CTAs of 128 threads per SM) that strides over rows and row chunks: a million eight-iteration
  // CTAs spent as long being scheduled as writing (9.1 ms per 18944 x 441000 wave against 5.1 ms of HBM time).
  // 128-thread CTAs of 32 registers: one of them still fits on an SM next to the five resident CTAs of the 4096-point
  // FFT kernel (57.6k of the 64k registers), so a wave generated on a side stream overlaps the previous wave's spectra
  const long long want_x = per_row < 8 ? per_row : 8;
  const long long slots = static_cast<long long>(sm_count()) * 16 / (want_x > 0 ? want_x : 1);
  const long long want_y = channels < slots ? channels : slots;
  dim3 grid(static_cast<unsigned>(want_x > 0 ? want_x : 1), static_cast<unsigned>(want_y > 0 ? want_y : 1));
  generate_uniform_kernel<T><<<grid, 128, 0, stream>>>(x, stride, channels, n, first_channel, seed, static_cast<T>(lo),
                                                        static_cast<T>(hi - lo));
  return after_launch("generate_uniform_kernel");
}

extern "C" {

int dspb200_generate_uniform_f32(float* x, int64_t stride, int64_t channels, int64_t n, int64_t first_channel, uint64_t seed,
                                 double lo, double hi, void* stream) {
  return generate_run<float>(x, stride, channels, n, first_channel, seed, lo, hi, static_cast<cudaStream_t>(stream));
}
int dspb200_generate_uniform_f64(double* x, int64_t stride, int64_t channels, int64_t n, int64_t first_channel, uint64_t seed,
                                 double lo, double hi, void* stream) {
  return generate_run<double>(x, stride, channels, n, first_channel, seed, lo, hi, static_cast<cudaStream_t>(stream));
}
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
