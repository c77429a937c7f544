// The app's cascade (app.py:161-167, :202-205) on a batch of clips:
//   y = SRC(x);  z = EQ(y);  mag = |FFT(hann * frames(z))|
// Device form: three kernels on one stream (y may live in scratch).  Host form:
// channel slabs pipelined over three streams so the H2D copy of slab k+1, the
// kernels of slab k and the D2H copies of slab k-1 overlap.
#include <cstdlib>

#include "internal.cuh"

namespace dspb200 {

static int64_t src_out_len(int L, int M, int64_t n_in) {
  int T = 0;
  int64_t P = 0, n_out = 0;
  dspb200_src_geometry(L, M, n_in, &T, &P, &n_out);
  return n_out;
}

static int64_t src_out_len_of(const dspb200_src_plan* src, int64_t n_in) {
  int L = 1, M = 1, dt = 0;
  if (src_plan_ratio(src, &L, &M, &dt) != DSPB200_OK) return 0;
  return src_out_len(L, M, n_in);
}

struct ChainShape {
  int L = 1, M = 1, n_fft = 0;
  int64_t n_out = 0, n_frames = 0, bins = 0;
  size_t y_bytes = 0, fft_ws = 0;
};

template <typename T>
static int chain_shape(const dspb200_src_plan* src, const dspb200_fft_plan* fft, int64_t channels, int64_t n_in,
                       ChainShape& s) {
  int dt = DType<T>::id;
  if (src) {
    int d2;
    DSP_TRY(src_plan_ratio(src, &s.L, &s.M, &d2));
    DSP_CHECK(d2 == dt, "src plan dtype does not match");
    s.n_out = src_out_len(s.L, s.M, n_in);
  } else {
    s.n_out = n_in;
  }
  s.y_bytes = static_cast<size_t>(round_up(static_cast<int64_t>(channels) * s.n_out * sizeof(T), 256));
  if (fft) {
    int d2;
    DSP_TRY(fft_plan_info(fft, &s.n_fft, &d2));
    DSP_CHECK(d2 == dt, "fft plan dtype does not match");
    s.n_frames = s.n_out / s.n_fft;
    s.bins = s.n_fft / 2 + 1;
    DSP_TRY(dspb200_fft_workspace_bytes(fft, channels * s.n_frames, &s.fft_ws));
  }
  return DSPB200_OK;
}

// Would a cascade of this shape run SRC and EQ as the one fused kernel (xz_mma.cu)?  *xp is NULL when not.
template <typename T>
static int chain_fused_plan(const dspb200_src_plan* src, const dspb200_eq_plan* eq, const T* x, int64_t xs, const T* z,
                            int64_t zs, int64_t channels, int64_t n_in, bool force, const XzPlan** xp) {
  *xp = nullptr;
  if (sizeof(T) != 4 || !src || !eq || getenv("DSPB200_CHAIN_NO_FUSED") != nullptr) return DSPB200_OK;
  const XzPlan* p = nullptr;
  DSP_TRY(eq_plan_xz(eq, src, &p));
  if (!p) return DSPB200_OK;
  const std::vector<double>* taps = src_plan_taps(src);
  const int n_taps = taps ? static_cast<int>(taps->size()) : 0;
  const float* xf = reinterpret_cast<const float*>(x);
  const float* zf = reinterpret_cast<const float*>(z);
  const bool ok = force ? xz_possible(*p, xf, xs, zf, zs, n_in, n_taps) : xz_usable(*p, xf, xs, zf, zs, channels, n_in, n_taps);
  if (ok) *xp = p;
  return DSPB200_OK;
}

template <typename T>
static int chain_run(const dspb200_src_plan* src, const dspb200_eq_plan* eq, const dspb200_fft_plan* fft,
                     const T* x, int64_t xs, int64_t channels, int64_t n_in, T* y, T* z, T* mag, void* ws,
                     size_t ws_bytes, cudaStream_t stream) {
  DSP_CHECK(channels >= 0 && n_in >= 1, "bad shape");
  if (channels == 0) return DSPB200_OK;
  DSP_CHECK(x != nullptr && z != nullptr, "NULL buffer");
  ChainShape s;
  DSP_TRY(chain_shape<T>(src, fft, channels, n_in, s));
  if (eq) DSP_CHECK(eq_plan_dtype(eq) == DType<T>::id, "eq plan dtype does not match");
  unsigned char* wsp = static_cast<unsigned char*>(ws);
  const XzPlan* xp = nullptr;
  if (!y) DSP_TRY(chain_fused_plan<T>(src, eq, x, xs, z, s.n_out, channels, n_in, false, &xp));
  if (!xp && !y && src && eq && sizeof(T) == 4 && ws != nullptr && ws_bytes >= static_cast<size_t>(round_up(static_cast<int64_t>(s.fft_ws), 256)) + s.y_bytes) {
    // a narrow float32 batch whose equaliser runs the tensor-core form only out of place (overlapping time slices):
    // the resampler writes into the scratch the caller provided behind the FFT's workspace
    bool oop = false;
    DSP_TRY(eq_prefers_out_of_place(eq, channels, s.n_out, s.n_out, &oop));
    if (oop) y = reinterpret_cast<T*>(wsp + round_up(static_cast<int64_t>(s.fft_ws), 256));
  }
  if (xp) {
    // SRC and EQ as one kernel: x is read once, z written once, y never exists (app.py:164-167)
    DSP_TRY(xz_run(*xp, reinterpret_cast<const float*>(x), xs, reinterpret_cast<float*>(z), s.n_out, channels, n_in, s.n_out,
                   eq_plan_clip(eq) != 0, stream));
  } else {
    const T* yp = x;
    int64_t y_stride = xs;
    if (src) {
      // without a caller-provided y the resampler writes into z and the equaliser runs in place: no scratch
      T* ybuf = y ? y : z;
      DSP_TRY(src_run<T>(src, x, xs, ybuf, s.n_out, channels, n_in, stream, -1));
      yp = ybuf;
      y_stride = s.n_out;
    }
    if (eq) {
      DSP_TRY(eq_run<T>(eq, yp, y_stride, z, s.n_out, channels, s.n_out, stream));
    } else if (yp != z) {
      DSP_CUDA(cudaMemcpy2DAsync(z, s.n_out * sizeof(T), yp, y_stride * sizeof(T), s.n_out * sizeof(T), channels,
                                 cudaMemcpyDeviceToDevice, stream));
    }
  }
  if (fft && mag && s.n_frames > 0) {
    DSP_CHECK(s.fft_ws == 0 || (ws != nullptr && ws_bytes >= s.fft_ws), "workspace too small for the FFT");
    DSP_TRY(fftmag_run<T>(fft, z, s.n_out, s.n_out, 0, s.n_fft, s.n_frames, mag, s.bins, s.n_frames * s.bins,
                          channels, wsp, s.fft_ws, stream));
  }
  return DSPB200_OK;
}

// Streams and device buffers of the host-form pipeline, kept per calling thread between
// calls (cudaMalloc/cudaFree and stream creation are milliseconds; a chain call is tens).
struct HostPipe {
  static constexpr int kStreams = 3;
  int device = -1;
  cudaStream_t st[kStreams] = {nullptr, nullptr, nullptr};
  static constexpr int kBufs = 6;
  void* buf[kStreams][kBufs] = {};      // x, z, mag, workspace, int16 z, row peaks
  size_t cap[kStreams][kBufs] = {};
  void release() {
    for (int i = 0; i < kStreams; ++i) {
      for (int j = 0; j < kBufs; ++j) { if (buf[i][j]) cudaFree(buf[i][j]); buf[i][j] = nullptr; cap[i][j] = 0; }
      if (st[i]) cudaStreamDestroy(st[i]);
      st[i] = nullptr;
    }
    device = -1;
  }
  cudaError_t ensure(int i, int j, size_t bytes) {
    if (bytes <= cap[i][j]) return cudaSuccess;
    if (buf[i][j]) cudaFree(buf[i][j]);
    buf[i][j] = nullptr; cap[i][j] = 0;
    cudaError_t e = cudaMalloc(&buf[i][j], bytes);
    if (e == cudaSuccess) cap[i][j] = bytes;
    return e;
  }
  ~HostPipe() {}   // buffers die with the context; freeing at thread exit could race its teardown
};
static thread_local HostPipe g_pipe;

template <typename T>
static int chain_host(const dspb200_src_plan* src, const dspb200_eq_plan* eq, const dspb200_fft_plan* fft,
                      const T* x, int64_t channels, int64_t n_in, T* z, T* mag, int16_t* z_pcm = nullptr,
                      T* peaks = nullptr) {
  // z_pcm != NULL: the export form.  z stays on the device; what crosses PCIe is the int16 signal app.py:349-354
  // writes to the WAV (half the bytes of float32 z) and, optionally, the per-clip peaks it was divided by.
  DSP_CHECK(channels >= 0 && n_in >= 1, "bad shape");
  if (channels == 0) return DSPB200_OK;
  DSP_CHECK(x != nullptr && (z != nullptr || z_pcm != nullptr), "NULL buffer");
  DSP_TRY(ensure_device());
  int64_t slab = 32;   // measured: 32-channel slabs pipeline best against PCIe (63 ms vs 65 ms at 64, 68 ms at 128 per 1024-clip wave)
  if (const char* e = getenv("DSPB200_CHAIN_SLAB")) {
    const long v = atol(e);
    if (v > 0) slab = v;
  }
  if (slab > channels) slab = channels;
  ChainShape s;
  DSP_TRY(chain_shape<T>(src, fft, slab, n_in, s));
  const int vec = 16 / static_cast<int>(sizeof(T));
  const int64_t xp = round_up(n_in, vec);
  constexpr int kStreams = HostPipe::kStreams;
  const int n_streams = static_cast<int>(ceil_div(channels, slab) < kStreams ? ceil_div(channels, slab) : kStreams);
  HostPipe& hp = g_pipe;
  int dev_now = 0;
  DSP_CUDA(cudaGetDevice(&dev_now));
  if (hp.device != dev_now) {
    hp.release();
    hp.device = dev_now;
  }
  cudaStream_t* st = hp.st;
  T* dx[kStreams] = {nullptr, nullptr, nullptr};
  T* dz[kStreams] = {nullptr, nullptr, nullptr};
  T* dm[kStreams] = {nullptr, nullptr, nullptr};
  void* dw[kStreams] = {nullptr, nullptr, nullptr};
  int16_t* dq[kStreams] = {nullptr, nullptr, nullptr};
  T* dp[kStreams] = {nullptr, nullptr, nullptr};
  const size_t ws_bytes = s.fft_ws;
  const size_t mag_elems = static_cast<size_t>(slab) * s.n_frames * s.bins;
  cudaError_t e = cudaSuccess;
  int rc = DSPB200_OK;
  for (int i = 0; i < n_streams && e == cudaSuccess; ++i) {
    if (!st[i]) e = cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = hp.ensure(i, 0, static_cast<size_t>(slab) * xp * sizeof(T));
    if (e == cudaSuccess) e = hp.ensure(i, 1, static_cast<size_t>(slab) * s.n_out * sizeof(T));
    if (e == cudaSuccess && mag_elems && mag) e = hp.ensure(i, 2, mag_elems * sizeof(T));
    if (e == cudaSuccess && ws_bytes) e = hp.ensure(i, 3, ws_bytes);
    if (e == cudaSuccess && z_pcm) e = hp.ensure(i, 4, static_cast<size_t>(slab) * round_up(s.n_out, 8) * sizeof(int16_t));
    if (e == cudaSuccess && z_pcm) e = hp.ensure(i, 5, static_cast<size_t>(slab) * sizeof(T));
    dq[i] = z_pcm ? static_cast<int16_t*>(hp.buf[i][4]) : nullptr;
    dp[i] = z_pcm ? static_cast<T*>(hp.buf[i][5]) : nullptr;
    dx[i] = static_cast<T*>(hp.buf[i][0]);
    dz[i] = static_cast<T*>(hp.buf[i][1]);
    dm[i] = (mag_elems && mag) ? static_cast<T*>(hp.buf[i][2]) : nullptr;
    dw[i] = ws_bytes ? hp.buf[i][3] : nullptr;
  }
  int k = 0;
  for (int64_t c0 = 0; c0 < channels && e == cudaSuccess && rc == DSPB200_OK; c0 += slab, ++k) {
    const int i = k % n_streams;
    const int64_t nc = (channels - c0) < slab ? (channels - c0) : slab;
    e = cudaMemcpy2DAsync(dx[i], xp * sizeof(T), x + c0 * n_in, n_in * sizeof(T), n_in * sizeof(T), nc,
                          cudaMemcpyHostToDevice, st[i]);
    if (e != cudaSuccess) break;
    rc = chain_run<T>(src, eq, fft, dx[i], xp, nc, n_in, nullptr, dz[i], dm[i], dw[i], ws_bytes, st[i]);
    if (rc != DSPB200_OK) break;
    if (z_pcm) {
      const int64_t qp = round_up(s.n_out, 8);     // 16-byte aligned int16 rows on the device
      rc = pcm16_run<T>(dz[i], s.n_out, dp[i], dq[i], qp, nc, s.n_out, st[i]);
      if (rc != DSPB200_OK) break;
      if (qp == s.n_out)
        e = cudaMemcpyAsync(z_pcm + c0 * s.n_out, dq[i], static_cast<size_t>(nc) * s.n_out * sizeof(int16_t),
                            cudaMemcpyDeviceToHost, st[i]);
      else
        e = cudaMemcpy2DAsync(z_pcm + c0 * s.n_out, s.n_out * sizeof(int16_t), dq[i], qp * sizeof(int16_t),
                              s.n_out * sizeof(int16_t), nc, cudaMemcpyDeviceToHost, st[i]);
      if (e == cudaSuccess && peaks)
        e = cudaMemcpyAsync(peaks + c0, dp[i], static_cast<size_t>(nc) * sizeof(T), cudaMemcpyDeviceToHost, st[i]);
    }
    if (e == cudaSuccess && z)
      e = cudaMemcpyAsync(z + c0 * s.n_out, dz[i], static_cast<size_t>(nc) * s.n_out * sizeof(T),
                          cudaMemcpyDeviceToHost, st[i]);
    if (e == cudaSuccess && dm[i])
      e = cudaMemcpyAsync(mag + c0 * s.n_frames * s.bins, dm[i],
                          static_cast<size_t>(nc) * s.n_frames * s.bins * sizeof(T), cudaMemcpyDeviceToHost, st[i]);
  }
  for (int i = 0; i < n_streams; ++i) {
    if (st[i]) {
      cudaError_t e2 = cudaStreamSynchronize(st[i]);
      if (e == cudaSuccess) e = e2;
    }
  }
  if (e != cudaSuccess) {
    hp.release();
    return fail(DSPB200_ERR_CUDA, "chain host path: %s", cudaGetErrorString(e));
  }
  return rc;
}

}  // namespace dspb200

using namespace dspb200;

extern "C" {

int dspb200_chain_workspace_bytes(const dspb200_src_plan* src, const dspb200_eq_plan* eq, const dspb200_fft_plan* fft,
                                  int64_t channels, int64_t n_in, int keep_y, size_t* bytes) {
  DSP_CHECK(bytes != nullptr, "bytes is NULL");
  DSP_CHECK(channels >= 0 && n_in >= 1, "bad shape");
  int dtype = DSPB200_F32, L, M, nf;
  if (src) DSP_TRY(src_plan_ratio(src, &L, &M, &dtype));
  else if (fft) DSP_TRY(fft_plan_info(fft, &nf, &dtype));
  ChainShape s;
  if (dtype == DSPB200_F32) DSP_TRY(chain_shape<float>(src, fft, channels, n_in, s));
  else DSP_TRY(chain_shape<double>(src, fft, channels, n_in, s));
  // without a y buffer the resampler writes into z and the equaliser runs in place -- unless the batch is narrow enough
  // that the equaliser's tensor-core form only pays out of place: then a y scratch rides behind the FFT's workspace
  *bytes = s.fft_ws;
  if (!keep_y && dtype == DSPB200_F32 && src && eq) {
    bool oop = false;
    DSP_TRY(eq_prefers_out_of_place(eq, channels, s.n_out, s.n_out, &oop));
    if (oop) *bytes = round_up(static_cast<int64_t>(s.fft_ws), 256) + s.y_bytes;
  }
  return DSPB200_OK;
}

int dspb200_chain_run_f32(const dspb200_src_plan* src, const dspb200_eq_plan* eq, const dspb200_fft_plan* fft,
                          const float* x, int64_t xs, int64_t channels, int64_t n_in, float* y, float* z,
                          float* mag, void* ws, size_t ws_bytes, void* stream) {
  return chain_run<float>(src, eq, fft, x, xs, channels, n_in, y, z, mag, ws, ws_bytes,
                          static_cast<cudaStream_t>(stream));
}
int dspb200_chain_run_f64(const dspb200_src_plan* src, const dspb200_eq_plan* eq, const dspb200_fft_plan* fft,
                          const double* x, int64_t xs, int64_t channels, int64_t n_in, double* y, double* z,
                          double* mag, void* ws, size_t ws_bytes, void* stream) {
  return chain_run<double>(src, eq, fft, x, xs, channels, n_in, y, z, mag, ws, ws_bytes,
                           static_cast<cudaStream_t>(stream));
}
int dspb200_chain_kernel_kind(const dspb200_src_plan* src, const dspb200_eq_plan* eq, int64_t channels, int64_t n_in,
                              int64_t x_stride, int* kind) {
  DSP_CHECK(kind != nullptr, "kind is NULL");
  *kind = 0;
  if (!src || !eq) return DSPB200_OK;
  DSP_TRY(ensure_device());
  const XzPlan* xp = nullptr;
  const int64_t n_out = src_out_len_of(src, n_in);
  DSP_TRY(chain_fused_plan<float>(src, eq, nullptr, x_stride, nullptr, n_out, channels, n_in, false, &xp));
  *kind = xp ? 1 : 0;
  return DSPB200_OK;
}
/* test hook: the fused SRC->EQ kernel whatever the batch width; DSPB200_ERR_UNSUPPORTED when the pair has no fused form */
int dspb200_chain_fused_f32(const dspb200_src_plan* src, const dspb200_eq_plan* eq, const float* x, int64_t xs,
                            int64_t channels, int64_t n_in, float* z, int64_t zs, void* stream) {
  DSP_CHECK(src != nullptr && eq != nullptr, "NULL plan");
  DSP_CHECK(channels >= 0 && n_in >= 1, "bad shape");
  if (channels == 0) return DSPB200_OK;
  DSP_CHECK(x != nullptr && z != nullptr, "NULL buffer");
  DSP_TRY(ensure_device());
  const XzPlan* xp = nullptr;
  DSP_TRY(chain_fused_plan<float>(src, eq, x, xs, z, zs, channels, n_in, true, &xp));
  if (!xp) return fail(DSPB200_ERR_UNSUPPORTED, "no fused SRC->EQ form for this plan pair / alignment");
  const int64_t n_out = src_out_len_of(src, n_in);
  DSP_CHECK(xs >= n_in && zs >= n_out, "channel stride smaller than the row length");
  return xz_run(*xp, x, xs, z, zs, channels, n_in, n_out, eq_plan_clip(eq) != 0, static_cast<cudaStream_t>(stream));
}
int dspb200_chain_host_f32(const dspb200_src_plan* src, const dspb200_eq_plan* eq, const dspb200_fft_plan* fft,
                           const float* x, int64_t channels, int64_t n_in, float* z, float* mag) {
  return chain_host<float>(src, eq, fft, x, channels, n_in, z, mag);
}
int dspb200_chain_host_pcm16_f32(const dspb200_src_plan* src, const dspb200_eq_plan* eq, const dspb200_fft_plan* fft,
                                 const float* x, int64_t channels, int64_t n_in, int16_t* z_pcm, float* peaks,
                                 float* mag) {
  DSP_CHECK(z_pcm != nullptr, "z_pcm is NULL");
  return chain_host<float>(src, eq, fft, x, channels, n_in, nullptr, mag, z_pcm, peaks);
}
int dspb200_chain_host_pcm16_f64(const dspb200_src_plan* src, const dspb200_eq_plan* eq, const dspb200_fft_plan* fft,
                                 const double* x, int64_t channels, int64_t n_in, int16_t* z_pcm, double* peaks,
                                 double* mag) {
  DSP_CHECK(z_pcm != nullptr, "z_pcm is NULL");
  return chain_host<double>(src, eq, fft, x, channels, n_in, nullptr, mag, z_pcm, peaks);
}
int dspb200_host_release(void) {
  g_pipe.release();
  return DSPB200_OK;
}

int dspb200_chain_host_f64(const dspb200_src_plan* src, const dspb200_eq_plan* eq, const dspb200_fft_plan* fft,
                           const double* x, int64_t channels, int64_t n_in, double* z, double* mag) {
  return chain_host<double>(src, eq, fft, x, channels, n_in, z, mag);
}

}  // extern "C"
