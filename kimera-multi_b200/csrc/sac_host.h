// sac_host.h — host-side constants of the batched RANSAC (lcd.cu / api_core.cu): the pre-drawn
// sample stream, the table of the adaptive iteration bound k and the squared-distance threshold
// of the stereo problem.  Plain C++ so that the CPU suite can run them in front of the emulated
// kernels (tests/emu/).
#pragma once
#include <float.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <random>
#include <vector>

namespace kml {

// pre-drawn sample stream: mt19937(seed)() >> 1, what libstdc++'s
// uniform_int_distribution<int>(0, INT_MAX) returns; 8 values per draw for every
// draw Ransac::computeModel can consume: max_it + 1 counted trials and max_skip = 10 * max_it
// skipped samples, so the stream cannot run out before the loop's own limits end it
inline int sac_max_draws(int max_it) { return max_it + 1 + 10 * std::max(max_it, 0); }
inline void fill_raw_stream(uint32_t seed, int max_it, std::vector<uint32_t>* raw) {
  const size_t draws = (size_t)sac_max_draws(max_it);
  raw->resize(draws * 8);
  std::mt19937 mt(seed);
  for (auto& r : *raw) r = (uint32_t)mt() >> 1;
}
// draws a problem with sample size S can take from the stream
inline int sac_cap_draws(int raw_len, int S, int max_it) { return std::min(raw_len / S, sac_max_draws(max_it)); }

// k as a function of (N, best inlier count), [n1][n1]: the exact expressions of
// opengv::sac::Ransac::computeModel evaluated with the host libm, so that the device replay
// takes the same branches as a CPU run.
inline void fill_ktable(int n1, int sample_size, double prob, std::vector<double>* tab_out) {
  std::vector<double>& tab = *tab_out;
  tab.assign((size_t)n1 * n1, 1.0);
  const double lnum = std::log(1.0 - prob);
  for (int N = 1; N < n1; ++N)
    for (int n = 0; n <= N; ++n) {
      double w = (double)n / (double)N;
      double p_no = 1.0 - std::pow(w, (double)sample_size);
      p_no = std::max(DBL_EPSILON, p_no);
      p_no = std::min(1.0 - DBL_EPSILON, p_no);
      tab[(size_t)N * n1 + n] = lnum / std::log(p_no);
    }
}

// smallest double s with sqrt(s) >= thr  =>  (sqrt(d2) < thr) == (d2 < s)
inline double sq_crit_of(double thr) {
  if (!(thr > 0.0)) return 0.0;
  double s = thr * thr;
  while (std::sqrt(s) >= thr) s = std::nextafter(s, 0.0);
  while (std::sqrt(s) < thr) s = std::nextafter(s, INFINITY);
  return s;
}

}  // namespace kml
