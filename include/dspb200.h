/*
 * dspb200.h -- C ABI of libdspb200.so: the B200 (sm_100a) implementation of the
 * numeric hot path of Renatovela-ctrl/dsp-audio-project, modules/dsp_core.py.
 *
 * Boundary.  The reference exposes eight module-level Python
 * functions; the three that do arithmetic are replaced here, batched over
 * channels.  Citations are into /root/reference/modules/dsp_core.py:
 *
 *   conversion_tasa_muestreo   :133-173  -> dspb200_src_*        (kernel K1)
 *   generar_respuesta_impulso_sinc :104-131 -> dspb200_design_sinc_taps (host)
 *   sistema_ecualizador        :216-254  -> dspb200_eq_*         (kernel K2)
 *   disenar_coeficientes_diferencias :179-203 -> dspb200_design_peaking_biquad (host)
 *   aplicar_ecuacion_diferencias :205-214 -> dspb200_eq_* with one raw section
 *   calcular_espectro_magnitud :68-98    -> dspb200_fftmag_*     (kernel K3)
 *   fft_diezmado_en_tiempo     :41-66    -> dspb200_fft_c2c_*    (kernel K3, complex)
 *   app.py:161-167, :202-205 cascade     -> dspb200_chain_*
 *
 * Conventions
 *   - plain C types only; every function returns an int status (0 = OK) and
 *     never throws; dspb200_last_error_string() describes the last failure on
 *     the calling thread.
 *   - "device" entry points take DEVICE pointers owned by the caller, element
 *     strides between channels, and a cudaStream_t passed as void* (NULL =
 *     default stream).  They only enqueue work.
 *   - "host" entry points take HOST pointers, do the H2D/D2H copies and the
 *     synchronisation themselves (what a ctypes/cgo caller without a CUDA
 *     allocator uses; this is the path bench.py's e2e number times).
 *   - layout: [channels, time], time fastest; dtype 0 = float32, 1 = float64.
 *   - there is NO CPU fallback: without a CUDA device the calls fail with
 *     DSPB200_ERR_NO_DEVICE / DSPB200_ERR_CUDA.
 *   - plans are immutable after creation and may be shared between threads;
 *     a plan belongs to the device that was current when it was created.
 */
#ifndef DSPB200_H_
#define DSPB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSPB200_VERSION 100 /* 0.1.0 */

enum {
  DSPB200_OK = 0,
  DSPB200_ERR_INVALID = 1,     /* bad argument (maps to ValueError in the Python shim) */
  DSPB200_ERR_CUDA = 2,        /* a CUDA runtime/driver call failed */
  DSPB200_ERR_UNSUPPORTED = 3, /* valid request this build cannot serve */
  DSPB200_ERR_NO_DEVICE = 4,   /* no CUDA device / not an sm_100 part */
  DSPB200_ERR_ALLOC = 5
};

enum { DSPB200_F32 = 0, DSPB200_F64 = 1 };

#define DSPB200_MAX_SECTIONS 16 /* biquad sections per EQ plan */
#define DSPB200_EQ_BANDS 6      /* the reference's fixed band table, dsp_core.py:225-228 */

typedef struct dspb200_src_plan dspb200_src_plan;
typedef struct dspb200_eq_plan dspb200_eq_plan;
typedef struct dspb200_fft_plan dspb200_fft_plan;

/* ---- library / device ------------------------------------------------- */
int dspb200_version(void);
const char* dspb200_last_error_string(void);
int dspb200_device_count(int* count);
/* The CUDA device current on the calling thread: the device a plan created now belongs to (plans hold device
 * tables; run entry points return DSPB200_ERR_INVALID when another device is current). */
int dspb200_current_device(int* device);
/* name_len bytes are written to name (NUL terminated). */
int dspb200_device_info(int device, char* name, int name_len, int* sm_count,
                        int* cc_major, int* cc_minor, size_t* total_mem_bytes);

/* ---- host-side design, float64 (dsp_core.py:104-131, :179-203) -------- */
/* Blackman-windowed sinc; even n_taps is bumped to n_taps+1 (:114).  h must
 * hold h_capacity doubles; *n_out receives the length actually written. */
int dspb200_design_sinc_taps(double w_c_norm, int n_taps, double* h, int h_capacity, int* n_out);
/* The resampler's filter for (L, M): cutoff 1/max(L,M), 40*max(L,M)+1 taps,
 * scaled by L (:155-162). */
int dspb200_design_src_filter(int L, int M, double* h, int h_capacity, int* n_out);
/* RBJ peaking biquad, alpha = sin(w0)/2, normalised by a0 (:187-201). */
int dspb200_design_peaking_biquad(double fc, double fs, double gain_db, double b[3], double a[3]);
/* The cascade's band rules (:233-251) for n_bands (centre, gain) pairs given in
 * the caller's dict order: |g| > 0.1, clamp to 0.9*fs/2, skip fc_eff <= 10 Hz.
 * Writes the surviving (fc_eff, gain) pairs; *bypass = 1 iff every |g| < 0.1
 * (:222-223, the caller must then return its input object untouched). */
int dspb200_eq_select_sections(double fs, const double* fc_nominal, const double* gains_db,
                               int n_bands, double* fc_eff, double* gain_eff, int* n_active,
                               int* bypass);

/* ---- K1: L/M polyphase sample-rate converter (dsp_core.py:133-173) ---- */
/* Geometry of the reference's zero-stuff / 'same' convolution / stride-M pick
 * for n_in input samples: taps T = 40*max(L,M)+1, centre offset
 * P = (min(n_in*L, T)-1)/2, output length ceil(max(n_in*L, T)/M), and
 * fs_out = (int)(fs*L/M) is left to the caller. */
int dspb200_src_geometry(int L, int M, int64_t n_in, int* n_taps, int64_t* centre, int64_t* n_out);
int dspb200_src_plan_create(int L, int M, int dtype, dspb200_src_plan** plan);
int dspb200_src_plan_destroy(dspb200_src_plan* plan);
/* y[c, m] = sum_i h[m*M + P - i*L] * x[c, i]; never materialises the
 * zero-stuffed signal.  x: [channels, n_in] with x_stride elements between
 * channels; y: [channels, n_out] with y_stride.  L == M == 1 is rejected with
 * DSPB200_ERR_INVALID: the reference returns the input object itself there
 * (:144-145) and the caller must do the same. */
int dspb200_src_run_f32(const dspb200_src_plan* plan, const float* x, int64_t x_stride,
                        float* y, int64_t y_stride, int64_t channels, int64_t n_in, void* stream);
int dspb200_src_run_f64(const dspb200_src_plan* plan, const double* x, int64_t x_stride,
                        double* y, int64_t y_stride, int64_t channels, int64_t n_in, void* stream);
/* Which kernel a run of this shape would use: 0 = generic, 1 = tiled (TMA + FFMA2),
 * 2 = tensor (fp32 only: banded-Toeplitz GEMM on tcgen05, 3-term TF32 split). */
int dspb200_src_plan_kernel_kind(const dspb200_src_plan* plan, int64_t channels, int64_t n_in,
                                 int64_t x_stride, int* kind);
/* Host buffers, dense [channels, n_in] -> [channels, n_out]. */
int dspb200_src_host_f32(int L, int M, const float* x, int64_t channels, int64_t n_in, float* y,
                         int64_t y_capacity_per_channel, int64_t* n_out);
int dspb200_src_host_f64(int L, int M, const double* x, int64_t channels, int64_t n_in, double* y,
                         int64_t y_capacity_per_channel, int64_t* n_out);

/* The same through a plan the caller keeps (its tap tables are built once, not per call). */
int dspb200_src_plan_host_f32(const dspb200_src_plan* plan, const float* x, int64_t channels, int64_t n_in,
                              float* y, int64_t y_capacity_per_channel, int64_t* n_out);
int dspb200_src_plan_host_f64(const dspb200_src_plan* plan, const double* x, int64_t channels, int64_t n_in,
                              double* y, int64_t y_capacity_per_channel, int64_t* n_out);

/* ---- K2: biquad equaliser cascade (dsp_core.py:205-254) --------------- */
/* Sections given as peaking (fc_eff, gain_db) pairs at rate fs, already
 * selected by dspb200_eq_select_sections (or any caller-side rule).  The final
 * clip to [-1, 1] (:254) is applied iff clip != 0.  n_sections may be 0
 * (clip-only copy: the reference's |g| == 0.1 corner). */
int dspb200_eq_plan_create(double fs, const double* fc_eff, const double* gains_db, int n_sections,
                           int clip, int dtype, dspb200_eq_plan** plan);
/* Sections given as raw difference-equation coefficients b[3], a[3] per
 * section (ba = n_sections x 6 doubles: b0 b1 b2 a0 a1 a2), the engine behind
 * aplicar_ecuacion_diferencias (:205-214). */
int dspb200_eq_plan_create_raw(const double* ba, int n_sections, int clip, int dtype,
                               dspb200_eq_plan** plan);
/* The reference's six-band rule applied to gains given in band order
 * Sub-Bass..Brilliance; *bypass as in dspb200_eq_select_sections (the plan is
 * still created, as a clip-free identity, when *bypass = 1). */
int dspb200_eq_plan_create_bands(double fs, const double gains_db[DSPB200_EQ_BANDS], int dtype,
                                 dspb200_eq_plan** plan, int* bypass);
int dspb200_eq_plan_destroy(dspb200_eq_plan* plan);
/* z[c, :] = clip(cascade(x[c, :])); zero initial state per channel; one pass
 * over HBM for all sections.  In place (z == x) is allowed.  Two forms: the
 * chunked linear-recurrence scan on the FMA pipe (any dtype, any shape), and,
 * for fp32 batches wide enough to fill the GPU (about 15k channels; at most 8
 * sections; 16-byte aligned rows), the cascade as one linear system advanced 96
 * samples per tcgen05 GEMM tile with the state carried between tiles.  Narrower
 * batches take the same form when the signal is long enough to be cut into
 * independent time slices that fill the GPU (out of place only): a slice then
 * starts dspb200_eq_plan_warm_chunks() chunks early from a zero state, by which
 * time what the true state would add is below 2^-24 of max|x| -- inside float32
 * rounding, not bit-identical to one sequential pass. */
int dspb200_eq_run_f32(const dspb200_eq_plan* plan, const float* x, int64_t x_stride, float* z,
                       int64_t z_stride, int64_t channels, int64_t n, void* stream);
int dspb200_eq_run_f64(const dspb200_eq_plan* plan, const double* x, int64_t x_stride, double* z,
                       int64_t z_stride, int64_t channels, int64_t n, void* stream);
/* Streaming form (float32, tensor-core kernel): the cascade over consecutive
 * time blocks of the same channels, e.g. config C3's 65536 x 2.88 M samples in
 * blocks that fit in HBM.  state: [channels][16] floats, opaque (the plan's
 * internal basis), receives the state after the block; first != 0 starts from
 * zero (dsp_core.py:214), first == 0 from `state` as the previous block left
 * it.  Every block but the last must be a multiple of
 * dspb200_eq_stream_chunk() samples; the blocks then reproduce one pass over
 * the whole signal bit for bit.  The clip (:254) is applied per block, which
 * is the same thing.  DSPB200_ERR_UNSUPPORTED for plans without a tensor form
 * (no sections, more than 8) or rows that are not 16-byte aligned. */
int dspb200_eq_stream_chunk(void);
int dspb200_eq_run_stream_f32(const dspb200_eq_plan* plan, const float* x, int64_t x_stride, float* z,
                              int64_t z_stride, int64_t channels, int64_t n, float* state, int first,
                              void* stream);
int dspb200_eq_host_f32(const dspb200_eq_plan* plan, const float* x, float* z, int64_t channels,
                        int64_t n);
int dspb200_eq_host_f64(const dspb200_eq_plan* plan, const double* x, double* z, int64_t channels,
                        int64_t n);
/* Which kernel a run of this shape would use: 0 = scan (FMA pipe), 1 = tensor. */
int dspb200_eq_plan_kernel_kind(const dspb200_eq_plan* plan, int64_t channels, int64_t n,
                                int64_t x_stride, int* kind);
/* Introspection for tests (host only, no device needed): the chunk system the
 * tensor-core form multiplies, in float64:  z = T x + O s,  s' = Phi s + K x
 * over *rows = 96 samples with *states = 2 * sections states.  tk receives
 * [T; K] as (rows + 16) x rows, o as rows x 16, phi as 16 x 16 (row-major, unused
 * state rows / columns zero); any of them may be NULL.  *rows = 0 when the plan
 * has no tensor form (no sections, or more than 8). */
int dspb200_eq_plan_chunk_system(const dspb200_eq_plan* plan, int* rows, int* states, double* tk,
                                 double* o, double* phi);
/* Introspection for tests (host only): the number of 96-sample chunks after which the cascade, started from a zero
 * state, is within 2^-24 of max|x| of its true output -- the overlap with which the tensor-core form cuts the time
 * axis of batches too narrow to fill the GPU into independent slices.  0: not established (no tensor form). */
int dspb200_eq_plan_warm_chunks(const dspb200_eq_plan* plan, int* chunks);
/* Introspection for tests: number of sections and, per section, 9 doubles
 * (a00 a01 a10 a11 b0 b1 c0 c1 d) of the state-space form the kernel runs. */
int dspb200_eq_plan_describe(const dspb200_eq_plan* plan, int* n_sections, double* state_space,
                             int capacity_sections);

/* ---- K3: radix-2 FFT (dsp_core.py:41-98) ------------------------------ */
/* n_fft: power of two, 1 <= n_fft <= DSPB200_FFT_MAX.  flags: DSPB200_FFT_HANN
 * applies the reference's symmetric Hann 0.5-0.5cos(2 pi k/(n_fft-1)) (:85-87);
 * DSPB200_FFT_DB makes the magnitude entry points write 20*log10(mag + 1e-12)
 * (the app's dB conversion, app.py:207-210) instead of the magnitude. */
#define DSPB200_FFT_MAX (1 << 17)
#define DSPB200_FFT_HANN 1
#define DSPB200_FFT_DB 2
int dspb200_fft_plan_create(int n_fft, int flags, int dtype, dspb200_fft_plan** plan);
int dspb200_fft_plan_destroy(dspb200_fft_plan* plan);
/* Bytes of device workspace a run with this many transforms needs (0 for
 * sizes that fit one CTA's shared memory). */
int dspb200_fft_workspace_bytes(const dspb200_fft_plan* plan, int64_t n_transforms, size_t* bytes);
/* Magnitude spectra of real frames.  Frame f of channel c covers samples
 * [offset + f*hop, offset + f*hop + n_fft) of x[c, :]; samples at or beyond
 * n_valid read as zero (the reference's zero padding, :79-82).  Output
 * mag[c, f, k], k = 0..n_fft/2, with mag_frame_stride / mag_channel_stride in
 * elements. */
int dspb200_fftmag_run_f32(const dspb200_fft_plan* plan, const float* x, int64_t x_stride,
                           int64_t n_valid, int64_t offset, int64_t hop, int64_t n_frames,
                           float* mag, int64_t mag_frame_stride, int64_t mag_channel_stride,
                           int64_t channels, void* workspace, size_t workspace_bytes, void* stream);
int dspb200_fftmag_run_f64(const dspb200_fft_plan* plan, const double* x, int64_t x_stride,
                           int64_t n_valid, int64_t offset, int64_t hop, int64_t n_frames,
                           double* mag, int64_t mag_frame_stride, int64_t mag_channel_stride,
                           int64_t channels, void* workspace, size_t workspace_bytes, void* stream);
/* Complex transform, natural order in and out, interleaved (re, im); batch
 * transforms laid out back to back.  No window. */
int dspb200_fft_c2c_run_f32(const dspb200_fft_plan* plan, const float* in, float* out,
                            int64_t batch, void* workspace, size_t workspace_bytes, void* stream);
int dspb200_fft_c2c_run_f64(const dspb200_fft_plan* plan, const double* in, double* out,
                            int64_t batch, void* workspace, size_t workspace_bytes, void* stream);
/* Host-buffer forms (dense: x [channels, n_samples]; mag [channels, n_frames, n_fft/2+1]). */
int dspb200_fftmag_host_f32(const dspb200_fft_plan* plan, const float* x, int64_t channels,
                            int64_t n_samples, int64_t offset, int64_t hop, int64_t n_frames,
                            float* mag);
int dspb200_fftmag_host_f64(const dspb200_fft_plan* plan, const double* x, int64_t channels,
                            int64_t n_samples, int64_t offset, int64_t hop, int64_t n_frames,
                            double* mag);
int dspb200_fft_c2c_host_f64(const dspb200_fft_plan* plan, const double* in, double* out,
                             int64_t batch);

/* ---- the app's cascade SRC -> EQ -> framed spectra (app.py:161-167) ---- */
/* Device form: x [channels, n_in] -> y [channels, n_out] (SRC output, may be
 * NULL when the caller does not keep it), z [channels, n_out] (EQ output), mag
 * [channels, n_frames, n_fft/2+1] with non-overlapping frames (hop = n_fft,
 * tail dropped).  eq may be NULL (bypass: z = y).
 * With y == NULL (what app.py:164-167 does: y is only ever the EQ's input) wide
 * float32 batches of a 160/147-shaped ratio run SRC and EQ as ONE kernel that
 * reads x once and writes z once (xz_mma.cu; DSPB200_CHAIN_NO_FUSED=1 in the
 * environment disables it, DSPB200_CHAIN_FORCE_FUSED=1 uses it at any batch
 * width); otherwise the resampler writes into z and the equaliser runs in
 * place.  The fused form needs |x| < 1023 (fp16 operand pieces); its stores
 * are clipped at 16-byte granularity, which never leaves a row of the dense
 * z (n_out % 4 == 0 is a condition of the form).
 * `workspace` serves long FFTs (dspb200_fft_workspace_bytes) and, for float32
 * batches too narrow to give every SM a channel group, a scratch for the
 * resampler's output so that the equaliser can run its tensor-core form on
 * overlapping time slices (out of place only); dspb200_chain_workspace_bytes
 * says how much.  A smaller (or NULL) workspace is accepted when no long FFT
 * needs it: the equaliser then runs in place. */
int dspb200_chain_workspace_bytes(const dspb200_src_plan* src, const dspb200_eq_plan* eq,
                                  const dspb200_fft_plan* fft, int64_t channels, int64_t n_in,
                                  int keep_y, size_t* bytes);
int dspb200_chain_run_f32(const dspb200_src_plan* src, const dspb200_eq_plan* eq,
                          const dspb200_fft_plan* fft, const float* x, int64_t x_stride,
                          int64_t channels, int64_t n_in, float* y, float* z, float* mag,
                          void* workspace, size_t workspace_bytes, void* stream);
int dspb200_chain_run_f64(const dspb200_src_plan* src, const dspb200_eq_plan* eq,
                          const dspb200_fft_plan* fft, const double* x, int64_t x_stride,
                          int64_t channels, int64_t n_in, double* y, double* z, double* mag,
                          void* workspace, size_t workspace_bytes, void* stream);
/* 1 in *kind when dspb200_chain_run_f32 with y == NULL would use the fused
 * SRC->EQ kernel for this shape (16-byte aligned rows assumed), else 0. */
int dspb200_chain_kernel_kind(const dspb200_src_plan* src, const dspb200_eq_plan* eq, int64_t channels,
                              int64_t n_in, int64_t x_stride, int* kind);
/* Host form: pinned or pageable host buffers in, host buffers out; copies are
 * pipelined against the kernels in channel slabs.  z and mag are dense. */
int dspb200_chain_host_f32(const dspb200_src_plan* src, const dspb200_eq_plan* eq,
                           const dspb200_fft_plan* fft, const float* x, int64_t channels,
                           int64_t n_in, float* z, float* mag);
int dspb200_chain_host_f64(const dspb200_src_plan* src, const dspb200_eq_plan* eq,
                           const dspb200_fft_plan* fft, const double* x, int64_t channels,
                           int64_t n_in, double* z, double* mag);

/* Export form of the host cascade: what app.py does with z after the chain
 * (app.py:349-354: nan_to_num, divide by the clip's peak when it is > 0, times
 * 32767, truncate to int16) runs on the device, so the signal crosses PCIe as
 * int16 (half the bytes of float32 z).  z_pcm: dense [channels, n_out] int16;
 * peaks: [channels] values of the working type (the divisors; may be NULL);
 * mag as above -- build the FFT plan with DSPB200_FFT_DB for the dB spectra of
 * app.py:207-210 -- and may be NULL. */
int dspb200_chain_host_pcm16_f32(const dspb200_src_plan* src, const dspb200_eq_plan* eq,
                                 const dspb200_fft_plan* fft, const float* x, int64_t channels,
                                 int64_t n_in, int16_t* z_pcm, float* peaks, float* mag);
int dspb200_chain_host_pcm16_f64(const dspb200_src_plan* src, const dspb200_eq_plan* eq,
                                 const dspb200_fft_plan* fft, const double* x, int64_t channels,
                                 int64_t n_in, int16_t* z_pcm, double* peaks, double* mag);

/* The host form keeps its streams and device slabs per calling thread between
 * calls; this frees the calling thread's. */
int dspb200_host_release(void);

/* ---- either side of the path (device buffers) -------------- */
/* Playback export, app.py:349-354: per row nan_to_num, divide by the row peak
 * when it is > 0, times 32767, truncate to int16.  peaks: rows scratch values
 * of the input type (receives the peaks). */
int dspb200_pcm16_run_f32(const float* x, int64_t stride, float* peaks, int16_t* out, int64_t out_stride,
                          int64_t rows, int64_t n, void* stream);
int dspb200_pcm16_run_f64(const double* x, int64_t stride, double* peaks, int16_t* out, int64_t out_stride,
                          int64_t rows, int64_t n, void* stream);
/* Loader front end, dsp_core.py:23-31: in [clips, frames, channels_in]
 * interleaved -> mono mean (float64) -> float32 -> divided by the clip's peak
 * when that is > 1e-6.  mono [clips, frames] float32, peaks [clips] float32. */
int dspb200_mono_normalize_run_f64(const double* in, int64_t clips, int64_t frames, int channels_in, float* mono,
                                   int64_t mono_stride, float* peaks, void* stream);
int dspb200_mono_normalize_run_f32(const float* in, int64_t clips, int64_t frames, int channels_in, float* mono,
                                   int64_t mono_stride, float* peaks, void* stream);

/* Synthetic clips for the throughput configurations (inputs are
 * generated on the device wave by wave).  x[c, i] = lo + (hi - lo) * u / 2^24 with
 * h = splitmix64(seed + 0x9E3779B97F4A7C15 * ((first_channel + c) * ceil(n / 2) + i / 2 + 1)) and
 * u = h >> 40 for even i, (h & 0xffffffff) >> 8 for odd i (one hash per pair of samples):
 * counter based, reproducible on the host. */
int dspb200_generate_uniform_f32(float* x, int64_t stride, int64_t channels, int64_t n, int64_t first_channel,
                                 uint64_t seed, double lo, double hi, void* stream);
int dspb200_generate_uniform_f64(double* x, int64_t stride, int64_t channels, int64_t n, int64_t first_channel,
                                 uint64_t seed, double lo, double hi, void* stream);

/* Number of kernels this library has launched on the calling process since
 * load (bench.py's gpu_launches). */
int64_t dspb200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DSPB200_H_ */
