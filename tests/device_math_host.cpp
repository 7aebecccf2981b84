// device_math_host.cpp — the DEVICE source of the mono / stereo minimal solvers
// (kimera-multi_b200/csrc/geom.cuh, fivept_thread.cuh, the very text nvcc compiles into
// libkml.so) compiled for the host with a shim for the CUDA keywords, one "thread" at a time
// (STRIDE = 1, no barriers).  Test infrastructure only (tests/test_device_math_host.py): it
// lets the CPU suite hold the device arithmetic to the oracle bit for bit without a GPU.
// Built with g++ -O2 -ffp-contract=off, the host equivalent of nvcc -fmad=false: every
// multiply and add rounds on its own; __fma_rn is the correctly rounded fma().
// The per-draw winner rule below restates mono_count_kernel's (ransac.cu).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>

#define KML_HOST_EMULATION 1
#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline void __syncthreads() {}  // never reached: SYNC = false
// the warp-cooperative Sturm fallback is not called here (a single host thread: the serial form is)
static inline void __syncwarp() {}
static inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
static inline int __all_sync(unsigned, bool p) { return p ? 1 : 0; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
using std::max;
using std::min;

#include "../kimera-multi_b200/csrc/fivept_thread.cuh"

using namespace kml::geom;

extern "C" {

void dmh_krsqrt(const double* x, int n, double* out) { for (int i = 0; i < n; ++i) out[i] = krsqrt(x[i]); }
void dmh_svd3(const double* A, double* U, double* S, double* V) { svd3(A, U, S, V); }
void dmh_svd3_r(const double* A, double* U, double* S, double* V) { svd3_r(A, U, S, V); }

// p1, p2: [3][3] rows = points; model 3x4 (p1 = R p2 + t)
void dmh_arun3(const double* p1, const double* p2, double* M) {
  arun3(p1, p1 + 3, p1 + 6, p2, p2 + 3, p2 + 6, M);
}
double dmh_arun_sqdist(const double* M, const double* p1, const double* p2) {
  return arun_sqdist(M, p1[0], p1[1], p1[2], p2[0], p2[1], p2[2]);
}
double dmh_mono_residual(const double* M, const double* f1, const double* f2) {
  double tinv[3];
  mono_tinv(M, tinv);
  const V3 a = {f1[0], f1[1], f1[2]}, b = {f2[0], f2[1], f2[2]};
  return mono_residual(M, tinv, a, b);
}
// 1 inlier / 0 outlier / -1 undecided
int dmh_mono_inlier_fast(const double* M, const double* f1, const double* f2, double thr) {
  double tinv[3];
  mono_tinv(M, tinv);
  const V3 a = {f1[0], f1[1], f1[2]}, b = {f2[0], f2[1], f2[2]};
  return mono_inlier_fast(M, tinv, a, b, inlier_margins(thr));
}

// One draw through stage 1 (front) -> stage 2 (isolate, deferred bisections included) ->
// stage 3 (one item per bracket) -> the count kernel's winner rule.  Returns 1 and the 3x4
// model if the draw yields one.  info[0] = roots chain 0, [1] = roots chain 1, [2] = deferred
// mask, [3] = refined items, [4] = winning item.
int dmh_mono_model(const double* f1, const double* f2, const uint16_t* sample8, int force_generic,
                   double* model, int* info) {
  double sm[kTphSlots], fo[kFrontOut], brk[2 * kMaxBrackets];
  mono_front_thread<1, false>(sm, f1, f2, sample8, true, fo);
  const int nr = mono_isolate_thread(fo, brk, force_generic != 0);
  const int R0 = nr & 255, R1 = (nr >> 8) & 255;
  for (int chain = 0; chain < 2; ++chain)
    if ((nr >> (16 + chain)) & 1) {
      const int R = chain ? R1 : R0;
      for (int j = 0; j < R; ++j) {
        double iso[kIsoSlots];
        mono_isolate_deferred_thread<1>(iso, fo, chain, j, R, brk + (chain ? 2 * R0 : 0));
      }
    }
  const int n = R0 + R1;
  double best = 1000000.0, bestM[12];
  int br = -1, refined = 0;
  for (int r = 0; r < n && r < kMaxBrackets; ++r) {
    double q = 0.0, M[12];
    const int status = mono_item(fo, r >= R0 ? 1 : 0, brk[2 * r], brk[2 * r + 1], f1, f2, sample8, &q, M);
    if (status >= 1 && refined < 10)
      if (status == 2 && q < best) { best = q; br = r; memcpy(bestM, M, sizeof M); }
    refined += status >= 1;
  }
  if (info) { info[0] = R0; info[1] = R1; info[2] = (nr >> 16) & 3; info[3] = refined; info[4] = br; }
  if (br < 0) return 0;
  memcpy(model, bestM, sizeof bestM);
  return 1;
}

}  // extern "C"
