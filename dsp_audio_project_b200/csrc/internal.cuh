// Cross-translation-unit entry points used by the chain driver.
#pragma once
#include "common.cuh"

struct dspb200_src_plan;
struct dspb200_eq_plan;
struct dspb200_fft_plan;

namespace dspb200 {
template <typename T>
int src_run(const dspb200_src_plan* plan, const T* x, int64_t xs, T* y, int64_t ys, int64_t channels,
            int64_t n_in, cudaStream_t stream, int force_kind);
template <typename T>
int eq_run(const dspb200_eq_plan* plan, const T* x, int64_t xs, T* z, int64_t zs, int64_t channels,
           int64_t n, cudaStream_t stream);
template <typename T>
int fftmag_run(const dspb200_fft_plan* p, const T* x, int64_t xs, int64_t n_valid, int64_t offset,
               int64_t hop, int64_t n_frames, T* mag, int64_t mfs, int64_t mcs, int64_t channels,
               void* ws, size_t ws_bytes, cudaStream_t stream);
int src_plan_ratio(const dspb200_src_plan* plan, int* L, int* M, int* dtype);
int fft_plan_info(const dspb200_fft_plan* plan, int* n_fft, int* dtype);
int eq_plan_dtype(const dspb200_eq_plan* plan);
}  // namespace dspb200
