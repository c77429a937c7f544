// Host-side float64 design: FIR taps, peaking biquads, state-space sections.
// These replace generar_respuesta_impulso_sinc (dsp_core.py:104-131) and
// disenar_coeficientes_diferencias (dsp_core.py:179-203); they run once per
// plan and their results are uploaded as device tables.
#include "design.cuh"

#include <cmath>

namespace dspb200 {

static const double kPi = 3.14159265358979323846264338327950288;

std::vector<double> sinc_taps(double w_c_norm, int n_taps) {
  if (n_taps % 2 == 0) n_taps += 1;  // odd length, linear phase (:114)
  const int half = n_taps / 2;
  std::vector<double> h(static_cast<size_t>(n_taps));
  long double total = 0.0L;
  for (int i = 0; i < n_taps; ++i) {
    const int n = i - half;
    const double arg = kPi * (w_c_norm * static_cast<double>(n));
    const double core = (arg == 0.0) ? 1.0 : std::sin(arg) / arg;  // numpy's normalised sinc (:120)
    double win = 1.0;
    if (n_taps > 1) {  // np.blackman (:123)
      const double ph = 2.0 * kPi * static_cast<double>(i) / static_cast<double>(n_taps - 1);
      win = 0.42 - 0.5 * std::cos(ph) + 0.08 * std::cos(2.0 * ph);
    }
    h[static_cast<size_t>(i)] = core * win;
    total += h[static_cast<size_t>(i)];
  }
  const double sum = static_cast<double>(total);
  if (sum != 0.0)  // unit DC gain (:127-129)
    for (auto& v : h) v /= sum;
  return h;
}

std::vector<double> src_filter(int L, int M) {
  const int big = L > M ? L : M;
  std::vector<double> h = sinc_taps(1.0 / static_cast<double>(big), 40 * big + 1);
  for (auto& v : h) v *= static_cast<double>(L);  // expansion gain (:162)
  return h;
}

void peaking_biquad(double fc, double fs, double gain_db, double b[3], double a[3]) {
  const double w0 = 2.0 * kPi * fc / fs;
  const double alpha = std::sin(w0) / 2.0;  // Q fixed (:188)
  const double A = std::pow(10.0, gain_db / 40.0);
  const double cw = std::cos(w0);
  const double a0 = 1.0 + alpha / A;
  b[0] = (1.0 + alpha * A) / a0;
  b[1] = (-2.0 * cw) / a0;
  b[2] = (1.0 - alpha * A) / a0;
  a[0] = a0 / a0;
  a[1] = (-2.0 * cw) / a0;
  a[2] = (1.0 - alpha / A) / a0;
}

void mat2_power(const double m[4], long long k, double out[4]) {
  long double r[4] = {1.0L, 0.0L, 0.0L, 1.0L};
  long double p[4] = {m[0], m[1], m[2], m[3]};
  while (k > 0) {
    if (k & 1) {
      long double t[4] = {r[0] * p[0] + r[1] * p[2], r[0] * p[1] + r[1] * p[3],
                          r[2] * p[0] + r[3] * p[2], r[2] * p[1] + r[3] * p[3]};
      for (int i = 0; i < 4; ++i) r[i] = t[i];
    }
    long double t[4] = {p[0] * p[0] + p[1] * p[2], p[0] * p[1] + p[1] * p[3],
                        p[2] * p[0] + p[3] * p[2], p[2] * p[1] + p[3] * p[3]};
    for (int i = 0; i < 4; ++i) p[i] = t[i];
    k >>= 1;
  }
  for (int i = 0; i < 4; ++i) out[i] = static_cast<double>(r[i]);
}

// lfilter's DF2T recurrence (dsp_core.py:214) as a state space:
//   s = [z0, z1];  y = z0 + b0 x;  s' = [[-a1, 1], [-a2, 0]] s + [b1 - a1 b0, b2 - a2 b0] x
// re-expressed in a well-conditioned basis so an all-fp32 run stays inside
// 1e-4 of full scale (the companion basis does not for the 40 Hz band).
Section section_from_ba(const double b_in[3], const double a_in[3]) {
  Section s{};
  const double a0 = a_in[0];
  const double b0 = b_in[0] / a0, b1 = b_in[1] / a0, b2 = b_in[2] / a0;
  const double a1 = a_in[1] / a0, a2 = a_in[2] / a0;
  const double B0 = b1 - a1 * b0, B1 = b2 - a2 * b0;
  s.d = b0;
  s.b0 = 1.0;
  const double disc = a1 * a1 - 4.0 * a2;
  if (disc < 0.0) {
    // complex pair sigma +- i omega: T = [[1,0],[-sigma,omega]] gives T^-1 A T =
    // [[sigma, omega], [-omega, sigma]]; a commuting rotation-scaling G then maps
    // the input vector onto e0, leaving c = [Bt0, -Bt1].
    const double sigma = -a1 / 2.0;
    const double omega = std::sqrt(-disc) / 2.0;
    const double bt0 = B0;
    const double bt1 = (sigma * B0 + B1) / omega;
    s.complex_poles = true;
    s.a[0] = sigma;  s.a[1] = omega;
    s.a[2] = -omega; s.a[3] = sigma;
    s.b1 = 0.0;
    s.c[0] = bt0;
    s.c[1] = -bt1;
    return s;
  }
  // real poles: orthogonal (Schur) triangularisation keeps the basis perfectly
  // conditioned even at critical damping (cuts near -12.04 dB), where
  // diagonalising would not.  The input vector stays general here.
  s.complex_poles = false;
  const double root = std::sqrt(disc);
  const double lam = (a1 <= 0.0) ? (-a1 + root) / 2.0 : (-a1 - root) / 2.0;  // larger |lambda| first
  double v0 = 1.0, v1 = lam + a1;  // eigenvector of [[-a1,1],[-a2,0]] for lam
  const double nv = std::sqrt(v0 * v0 + v1 * v1);
  v0 /= nv; v1 /= nv;
  // Q = [[v0,-v1],[v1,v0]];  U = Q^T A Q with A = [[-a1,1],[-a2,0]]
  const double A00 = -a1, A01 = 1.0, A10 = -a2, A11 = 0.0;
  const double AQ00 = A00 * v0 + A01 * v1, AQ01 = -A00 * v1 + A01 * v0;
  const double AQ10 = A10 * v0 + A11 * v1, AQ11 = -A10 * v1 + A11 * v0;
  s.a[0] = v0 * AQ00 + v1 * AQ10;
  s.a[1] = v0 * AQ01 + v1 * AQ11;
  s.a[2] = 0.0;  // -v1*AQ00 + v0*AQ10 vanishes analytically
  s.a[3] = -v1 * AQ01 + v0 * AQ11;
  s.b0 = v0 * B0 + v1 * B1;
  s.b1 = -v1 * B0 + v0 * B1;
  s.c[0] = v0;   // [1, 0] Q
  s.c[1] = -v1;
  return s;
}

}  // namespace dspb200

using namespace dspb200;

extern "C" {

int dspb200_design_sinc_taps(double w_c_norm, int n_taps, double* h, int h_capacity, int* n_out) {
  DSP_CHECK(n_taps >= 1, "n_taps must be >= 1 (got %d)", n_taps);
  DSP_CHECK(h != nullptr && n_out != nullptr, "NULL output pointer");
  const int n = n_taps + (n_taps % 2 == 0 ? 1 : 0);
  DSP_CHECK(h_capacity >= n, "h_capacity %d too small for %d taps", h_capacity, n);
  std::vector<double> t = sinc_taps(w_c_norm, n_taps);
  for (int i = 0; i < n; ++i) h[i] = t[static_cast<size_t>(i)];
  *n_out = n;
  return DSPB200_OK;
}

int dspb200_design_src_filter(int L, int M, double* h, int h_capacity, int* n_out) {
  DSP_CHECK(L >= 1 && M >= 1, "L and M must be >= 1 (got L=%d M=%d)", L, M);
  DSP_CHECK(h != nullptr && n_out != nullptr, "NULL output pointer");
  const int n = 40 * (L > M ? L : M) + 1;
  DSP_CHECK(h_capacity >= n, "h_capacity %d too small for %d taps", h_capacity, n);
  std::vector<double> t = src_filter(L, M);
  for (int i = 0; i < n; ++i) h[i] = t[static_cast<size_t>(i)];
  *n_out = n;
  return DSPB200_OK;
}

int dspb200_design_peaking_biquad(double fc, double fs, double gain_db, double b[3], double a[3]) {
  DSP_CHECK(b != nullptr && a != nullptr, "NULL output pointer");
  peaking_biquad(fc, fs, gain_db, b, a);
  return DSPB200_OK;
}

int dspb200_eq_select_sections(double fs, const double* fc_nominal, const double* gains_db,
                               int n_bands, double* fc_eff, double* gain_eff, int* n_active,
                               int* bypass) {
  DSP_CHECK(n_bands >= 0, "n_bands must be >= 0");
  DSP_CHECK(n_active != nullptr && bypass != nullptr, "NULL output pointer");
  DSP_CHECK(n_bands == 0 || (fc_nominal && gains_db && fc_eff && gain_eff), "NULL band arrays");
  bool all_small = true;  // :222
  for (int i = 0; i < n_bands; ++i)
    if (!(std::fabs(gains_db[i]) < 0.1)) all_small = false;
  *bypass = all_small ? 1 : 0;
  int k = 0;
  const double nyq = fs / 2.0;
  for (int i = 0; i < n_bands; ++i) {
    if (std::fabs(gains_db[i]) > 0.1) {  // :234
      const double ceiling = nyq * 0.90;  // :240
      const double fc = (fc_nominal[i] >= ceiling) ? ceiling : fc_nominal[i];
      if (fc > 10.0) {  // :249
        fc_eff[k] = fc;
        gain_eff[k] = gains_db[i];
        ++k;
      }
    }
  }
  *n_active = k;
  return DSPB200_OK;
}

}  // extern "C"
