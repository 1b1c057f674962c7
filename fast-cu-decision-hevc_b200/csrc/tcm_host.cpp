// tcm_host.cpp - host side of the outlier feature: the truncated-Laplacian / uniform mixture fit
// that turns one frequency's histogram of |coeff/8| into the outlier threshold Yc.
//
// Follows TCMprocessOneSequence and its helpers (TEncSlice.cpp:202-392).  It stays on the host on
// purpose: it is double-precision exp/log with data-dependent iteration counts,
// and the reference's results depend on the host libm; the inputs are 15 small histograms per picture
// that the GPU produces (feature_kernels.cu).  The expressions are evaluated in the same order as the
// reference so that the doubles, and therefore the integer thresholds, are identical.
#include "tcm_host.h"
#include <cmath>
#include <vector>

namespace cucd {

namespace {
constexpr double kLambdaDelta = 0.1;      // TEncSlice.cpp:179
constexpr double kMinLikelihood = -1.e30; // :180
constexpr double kStartPointProb = 0.1;   // :182
constexpr int kMaxAmp = 65536;            // :178

// fixed-point iteration for the Laplacian scale given the truncation point (TEncSlice.cpp:202-230)
double solve_lambda(double yc, double sumAbs, double count) {
  const double c = sumAbs / count;
  if (c / yc >= 0.95) return -1.0;
  double prev = c;
  double lam = c - yc * (1.0 - 1.0 / (1.0 - std::exp(-yc / prev)));
  for (int k = 0; k < 5; k++) {
    prev = lam;
    lam = c - yc * (1.0 - 1.0 / (1.0 - std::exp(-yc / prev)));
  }
  while (std::fabs(lam - prev) > kLambdaDelta) {
    prev = lam;
    lam = c - yc * (1.0 - 1.0 / (1.0 - std::exp(-yc / prev)));
  }
  return lam;
}
}  // namespace

double tcm_fit_one(const uint32_t* count, int nSamples) {
  int peak = 0;
  for (int k = 0; k < kTcmBins; k++) if (count[k]) peak = k;
  if (peak == 0 || peak >= kMaxAmp) return 0.0;                       // TEncSlice.cpp:304-313

  // cumulative sample count and cumulative |amplitude| up to each bin (:327-333); integers, exact in double
  std::vector<double> cumN(peak + 1), cumAmp(peak + 1);
  cumN[0] = count[0]; cumAmp[0] = 0.0;
  for (int k = 1; k <= peak; k++) {
    cumAmp[k] = cumAmp[k - 1] + (double)k * (double)count[k];
    cumN[k] = cumN[k - 1] + (double)count[k];
  }
  auto cnt = [&](int k) -> long long { return k <= peak ? (long long)count[k] : 0; };

  // first truncation point to try (:233-256)
  int start = peak;
  for (; start > 0; start--) {
    if (count[start] == 0) continue;
    if (cumN[start] < nSamples * (1.0 - kStartPointProb)) break;
  }
  const long long pct = nSamples / 100;
  if (cnt(0) > pct && cnt(1) > pct && cnt(2) > pct && cnt(3) > pct) { if (start < 3) start = 3; }
  else if (cnt(0) > pct && cnt(1) > pct && cnt(2) > pct) { if (start < 2) start = 2; }
  else { if (start < 1) start = 1; }
  if (start > peak) start = peak;

  auto likelihood = [&](int k) -> double {                            // :258-289
    const double n1 = cumN[k], n2 = nSamples - n1, yc = k, sumAbs = cumAmp[k];
    const double lam = solve_lambda(yc, sumAbs, n1);
    const double prob = n1 / (double)nSamples;
    if (lam > 0)
      return n2 * std::log(1 - prob) + n1 * std::log(prob) - n2 * std::log((peak - yc) * 2.0) - n1 * std::log(1 - std::exp(-yc / lam)) -
             n1 * std::log(2 * lam) - sumAbs / lam;
    return -kMinLikelihood;
  };

  double best = likelihood(start);                                    // :356-374
  int bestK = start;
  for (int k = start + 1; k <= peak; k++) {
    if (count[k] == 0) continue;
    const double l = likelihood(k);
    if (l > best) { best = l; bestK = k; }
  }
  return best > kMinLikelihood ? (double)bestK : 0.0;                 // :376-389
}

void tcm_fit_picture(const uint32_t* hist, int nBlocks, double* yc, int32_t* thr) {
  yc[0] = 0.0; thr[0] = 0;
  for (int f = 1; f < 16; f++) {
    yc[f] = tcm_fit_one(hist + (size_t)f * kTcmBins, nBlocks);
    thr[f] = (int32_t)(yc[f] * 8.0);                                  // Yc * DctScaling, :1010
  }
}

}  // namespace cucd
