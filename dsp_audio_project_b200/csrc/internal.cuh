// Cross-translation-unit entry points used by the chain driver.
#pragma once
#include <vector>

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
// K3, 4096-point magnitude frames with 32 points per thread (fft_r32.cu).  fp32 only.
struct FftR32Plan {
  int ok = 0;
  void* d_tables = nullptr;   // pass-1 twiddles, W_4096^t, per-thread Hann (A, B) pairs
  float hann_cos[32] = {}, hann_sin[32] = {};
};
int fft_r32_build(int n_fft, FftR32Plan& rp);
void fft_r32_free(FftR32Plan& rp);
int fft_r32_run(const FftR32Plan& rp, const float* x, int64_t xs, int64_t n_valid, int64_t offset, int64_t hop,
                int64_t n_frames, float* mag, int64_t mfs, int64_t mcs, int64_t channels, int hann, int db,
                cudaStream_t stream);
// K3, 2^16-point magnitude frames as three radix-32 passes (fft_long32.cu).  fp32 only.
struct FftLong32Plan {
  int ok = 0;
  void* d_tables = nullptr;   // W_1024 powers, split twiddles, column twiddles, per-column Hann (A, B) pairs
  float hann_cos[32] = {}, hann_sin[32] = {};
};
int fft_long32_build(int n_fft, FftLong32Plan& lp);
void fft_long32_free(FftLong32Plan& lp);
size_t fft_long32_workspace(int64_t n_transforms);
int fft_long32_run(const FftLong32Plan& lp, const float* x, int64_t xs, int64_t n_valid, int64_t offset, int64_t hop,
                   int64_t n_frames, float* mag, int64_t mfs, int64_t mcs, int64_t channels, int hann, int db, void* ws,
                   size_t ws_bytes, cudaStream_t stream);
int src_plan_ratio(const dspb200_src_plan* plan, int* L, int* M, int* dtype);
// playback export (post.cu; app.py:349-354): row peaks, then int16(trunc(nan_to_num(x) / peak * 32767))
template <typename T>
int pcm16_run(const T* x, int64_t stride, T* peaks, short* out, int64_t out_stride, int64_t rows, int64_t n,
              cudaStream_t stream);

// K1 on tcgen05 (src_mma.cu): per-plan tap matrices and the launch.  fp32 only.
struct SrcMmaPlan {
  int ok = 0;
  int period = 0;          // distinct tile phases
  int kpad = 0;            // GEMM depth per tile (multiple of 32)
  long long adv = 0;       // input samples per `period` tiles
  float* d_table = nullptr;   // [period][hi, lo][128][kpad]
  int* d_lo = nullptr;        // [period] window start of tile p
};
int src_mma_build(const std::vector<double>& taps, int L, int M, SrcMmaPlan& mp);
void src_mma_free(SrcMmaPlan& mp);
bool src_mma_usable(const SrcMmaPlan& mp, const float* x, int64_t xs, int64_t channels, int64_t n_in);
int src_mma_run(const SrcMmaPlan& mp, const float* x, int64_t xs, float* y, int64_t ys, int64_t channels,
                int64_t n_in, int64_t n_out, cudaStream_t stream);
// K2 on tcgen05 (eq_mma.cu): the cascade as a chunked linear system.  fp32 only.
constexpr int kLtiMaxStates = 16;
struct Section;
struct LtiMmaPlan {
  int ok = 0;
  int kpad = 0;            // GEMM depth per chunk in the table (= samples per chunk)
  int states = 0;          // 2 * sections, padded to a multiple of 4
  float* d_table = nullptr;   // [hi, lo, free response][112][kpad]: rows 0..95 outputs, 96.. end states
  float phi[kLtiMaxStates * kLtiMaxStates] = {};   // state transition over one chunk
  // chunks after which a slice started from a ZERO state is within 2^-24 of max|x| of the true output (the cascade
  // forgets its past at the rate of its slowest pole); 0: not established.  Lets narrow batches cut the time axis
  // into independent, overlapping slices (lti_mma_run).
  int warm_chunks = 0;
};
// z = T x + O s, s' = Phi s + K x over `rows` samples (float64, host): tk [(rows + 16) x rows] = [T; K],
// o [rows x 16], phi [16 x 16]; unused state rows/columns are zero.
struct LtiChunkSystem {
  int rows = 0, states = 0;
  std::vector<double> tk, o, phi;
};
int lti_chunk_system(const Section* sec, int ns, LtiChunkSystem& cs, int rows = 0);
void cascade_state_space(const Section* sec, int ns, std::vector<double>& A, std::vector<double>& B,
                         std::vector<double>& C, double& D);
int lti_mma_build_eq(const Section* sec, int ns, LtiMmaPlan& mp);
void lti_mma_free(LtiMmaPlan& mp);
int lti_warm_chunks(const LtiChunkSystem& cs);
bool lti_mma_usable(const LtiMmaPlan& mp, const float* x, int64_t xs, const float* z, int64_t zs, int64_t channels,
                    int64_t n_in);
// state (optional): [channels][16] floats in the plan's scaled basis, written with the state after the block;
// state_in: start from it instead of zero (streaming form)
int lti_mma_run(const LtiMmaPlan& mp, const float* x, int64_t xs, float* z, int64_t zs, int64_t channels,
                int64_t n_in, int64_t n_out, bool clip, float* state, bool state_in, cudaStream_t stream);
bool lti_mma_possible(const LtiMmaPlan& mp, const float* x, int64_t xs, const float* z, int64_t zs);
int lti_mma_chunk();
// K1 + K2 fused on tcgen05 (xz_mma.cu): x -> z in one pass, y never touches HBM.  fp32 I/O, 160/147-shaped ratios.
struct XzPlan {
  int ok = 0;
  int L = 0, M = 0, device = -1;
  int states = 0;          // 2 * sections, padded to a multiple of 4
  int tab_rows = 0;        // rows of 64 fp16 in d_table
  uint32_t tab_bytes = 0, o_off = 0;
  uint32_t blk_off[2][2][2] = {};
  int blk_row0[2][2] = {};
  int r0[2][8] = {};
  int first_new = 0, start0 = 0;
  float unscale = 0.f;
  void* d_table = nullptr;   // [phase][hi, lo][block] coefficient tiles of G_ph = [T; K] A_ph, then the free-response operand
  float phi[kLtiMaxStates * kLtiMaxStates] = {};
};
int xz_build(const std::vector<double>& taps, int L, int M, const Section* sec, int ns, XzPlan& xp);
void xz_free(XzPlan& xp);
int xz_chunk();
bool xz_possible(const XzPlan& xp, const float* x, int64_t xs, const float* z, int64_t zs, int64_t n_in, int n_taps);
bool xz_usable(const XzPlan& xp, const float* x, int64_t xs, const float* z, int64_t zs, int64_t channels, int64_t n_in,
               int n_taps);
int xz_run(const XzPlan& xp, const float* x, int64_t xs, float* z, int64_t zs, int64_t channels, int64_t n_in, int64_t n_out,
           bool clip, cudaStream_t stream);
// the fused tables of (src plan, eq plan), built on first use and cached in the eq plan; *xp is NULL when the pair has no
// fused form (other ratios, more than 8 sections, float64)
int eq_plan_xz(const dspb200_eq_plan* eq, const dspb200_src_plan* src, const XzPlan** xp);
int eq_plan_clip(const dspb200_eq_plan* plan);
int eq_prefers_out_of_place(const dspb200_eq_plan* plan, int64_t channels, int64_t n, int64_t stride, bool* prefers);
const std::vector<double>* src_plan_taps(const dspb200_src_plan* plan);
int fft_plan_info(const dspb200_fft_plan* plan, int* n_fft, int* dtype);
int eq_plan_dtype(const dspb200_eq_plan* plan);
}  // namespace dspb200
