// Host-side float64 design code shared by the plans (declarations).
#pragma once

#include <vector>

#include "common.cuh"

namespace dspb200 {

// dsp_core.py:104-131
std::vector<double> sinc_taps(double w_c_norm, int n_taps);
// dsp_core.py:155-162: cutoff 1/max(L,M), 40*max(L,M)+1 taps, times L
std::vector<double> src_filter(int L, int M);
// dsp_core.py:179-203
void peaking_biquad(double fc, double fs, double gain_db, double b[3], double a[3]);

// One second-order section in the state-space form the EQ kernel runs:
//   q' = A q + [1, b1]^T x ;  y = c . q + d x        (q is the state BEFORE x)
// A is a rotation-scaling (complex poles; b1 = 0) or upper-triangular (real
// poles; b1 in {0,1} after diagonal scaling) -- SURVEY.md section 7, "hard parts".
struct Section {
  double a[4];   // a00 a01 a10 a11
  double b0;     // 1 unless the section could not be normalised
  double b1;
  double c[2];
  double d;
  bool complex_poles;
};
Section section_from_ba(const double b[3], const double a[3]);
// M^k for a 2x2 matrix (long double accumulation), row-major.
void mat2_power(const double m[4], long long k, double out[4]);

}  // namespace dspb200
