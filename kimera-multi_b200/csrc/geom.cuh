// geom.cuh — fp64 building blocks shared by the RANSAC kernels (device side).
//
// Implements the ARITHMETIC CONTRACT of DESIGN.md §4 (operation order, sweep
// counts, tie rules) for:
//   - the 3x3 "proper" SVD used by both minimal solvers,
//   - opengv::point_cloud::threept_arun + 3-D residual (PointCloudSacProblem,
//     SURVEY.md A.8; /root/reference/images/kimera-multi.drawio:2595-2598, 2654),
//   - triangulate2 + bearing reprojection residual of
//     CentralRelativePoseSacProblem (drawio:2589-2592, 2646).
// Compiled with -fmad=false: the compiler contracts nothing; the fused multiply-adds of the
// contract are the explicit kfma() calls (the oracle writes fma() at the same places), every
// other multiply and add is rounded on its own.  fp64 div and sqrt are IEEE on sm_100a.
// Division, square root and the larger routines are deliberately NOT inlined:
// the mono kernel runs 16 warps per SM in different phases of a long program,
// so its instruction footprint must stay inside the SM's instruction cache
// (profiles/r01b: the fully inlined build stalled on instruction fetch).
#pragma once
#include <math.h>
#include <string.h>
#include <stdint.h>

#define KML_DI __device__ __forceinline__
#define KML_DN __device__ __noinline__

namespace kml {
namespace geom {

#ifdef KML_FILTER_STATS
__device__ unsigned long long g_fstats[12];
#endif
KML_DI double kfma(double a, double b, double c) { return __fma_rn(a, b, c); }
KML_DN double kdiv(double a, double b) { return a / b; }
// the contract's reciprocal square root: magic-constant start, four Newton steps — no sqrt, no division
KML_DI double krsqrt(double x) {
  unsigned long long i;
  memcpy(&i, &x, 8);
  i = 0x5FE6EB50C7B537A9ull - (i >> 1);
  double y;
  memcpy(&y, &i, 8);
  const double h = 0.5 * x;
  y = y * kfma(-(h * y), y, 1.5);
  y = y * kfma(-(h * y), y, 1.5);
  y = y * kfma(-(h * y), y, 1.5);
  y = y * kfma(-(h * y), y, 1.5);
  return y;
}
KML_DN double ksqrt(double a) { return sqrt(a); }

struct V3 {
  double x, y, z;
};
KML_DI double dot(const V3& a, const V3& b) { return kfma(a.z, b.z, kfma(a.y, b.y, a.x * b.x)); }
KML_DI V3 cross(const V3& a, const V3& b) {
  V3 c;
  c.x = kfma(a.y, b.z, -(a.z * b.y));
  c.y = kfma(a.z, b.x, -(a.x * b.z));
  c.z = kfma(a.x, b.y, -(a.y * b.x));
  return c;
}

// polynomial value, ascending coefficients
KML_DN double horner(const double* c, int deg, double x) {
  double r = c[deg];
  for (int i = deg - 1; i >= 0; --i) r = kfma(r, x, c[i]);
  return r;
}

// One Hestenes rotation on columns (p,q) of G (3x3 row-major), accumulated in W.
KML_DN bool jacobi_pair(double* G, double* W, int p, int q) {
  const double a = kfma(G[6 + p], G[6 + p], kfma(G[3 + p], G[3 + p], G[p] * G[p]));
  const double b = kfma(G[6 + q], G[6 + q], kfma(G[3 + q], G[3 + q], G[q] * G[q]));
  const double g = kfma(G[6 + p], G[6 + q], kfma(G[3 + p], G[3 + q], G[p] * G[q]));
  if (g * g <= 1e-30 * a * b) return false;
  const double h = b - a, tg = 2.0 * g;  // t = sgn(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = h / tg, times 2|g|: one division
  const double sgn = (h == 0.0 || (h > 0.0) == (g > 0.0)) ? 1.0 : -1.0;
  const double t = kdiv(sgn * fabs(tg), fabs(h) + ksqrt(kfma(h, h, tg * tg)));
  const double c = krsqrt(kfma(t, t, 1.0));
  const double s = c * t;
#pragma unroll 1
  for (int i = 0; i < 3; ++i) {
    const double gp = G[3 * i + p], gq = G[3 * i + q];
    G[3 * i + p] = kfma(c, gp, -(s * gq));
    G[3 * i + q] = kfma(s, gp, c * gq);
    const double wp = W[3 * i + p], wq = W[3 * i + q];
    W[3 * i + p] = kfma(c, wp, -(s * wq));
    W[3 * i + q] = kfma(s, wp, c * wq);
  }
  return true;
}

// Hestenes one-sided Jacobi, "proper" SVD: A = U diag(S) V^T, S descending,
// third columns completed by cross products (det U = det V = +1).
// U, V row-major 3x3 (columns = singular vectors).
KML_DN void svd3(const double* A, double* U, double* S, double* V) {
  double G[9], W[9];
#pragma unroll 1
  for (int i = 0; i < 9; ++i) {
    G[i] = A[i];
    W[i] = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
  }
#pragma unroll 1
  for (int sweep = 0; sweep < 12; ++sweep) {
    bool rotated = jacobi_pair(G, W, 0, 1);
    rotated = jacobi_pair(G, W, 0, 2) || rotated;
    rotated = jacobi_pair(G, W, 1, 2) || rotated;
    if (!rotated) break;
  }
  double n[3];
#pragma unroll 1
  for (int j = 0; j < 3; ++j) n[j] = ksqrt(kfma(G[6 + j], G[6 + j], kfma(G[3 + j], G[3 + j], G[j] * G[j])));
  // stable descending order of the three column norms
  int i0 = 0, i1 = 1, i2 = 2;
  if (n[i1] > n[i0]) { const int t = i0; i0 = i1; i1 = t; }
  if (n[i2] > n[i1]) { const int t = i1; i1 = i2; i2 = t; }
  if (n[i1] > n[i0]) { const int t = i0; i0 = i1; i1 = t; }
  S[0] = n[i0]; S[1] = n[i1]; S[2] = n[i2];
  V3 u0, u1;
  if (n[i0] > 0.0) {
    const double r0 = kdiv(1.0, n[i0]);
    u0.x = G[i0] * r0; u0.y = G[3 + i0] * r0; u0.z = G[6 + i0] * r0;
  } else {
    u0.x = 1.0; u0.y = 0.0; u0.z = 0.0;
  }
  if (n[i1] > 0.0) {
    const double r1 = kdiv(1.0, n[i1]);
    u1.x = G[i1] * r1; u1.y = G[3 + i1] * r1; u1.z = G[6 + i1] * r1;
  } else {
    int k = 0;
    double m = fabs(u0.x);
    if (fabs(u0.y) < m) { k = 1; m = fabs(u0.y); }
    if (fabs(u0.z) < m) { k = 2; }
    const V3 e = {k == 0 ? 1.0 : 0.0, k == 1 ? 1.0 : 0.0, k == 2 ? 1.0 : 0.0};
    const V3 c = cross(u0, e);
    const double nn = ksqrt(dot(c, c));
    u1.x = kdiv(c.x, nn); u1.y = kdiv(c.y, nn); u1.z = kdiv(c.z, nn);
  }
  const V3 u2 = cross(u0, u1);
  const V3 v0 = {W[i0], W[3 + i0], W[6 + i0]}, v1 = {W[i1], W[3 + i1], W[6 + i1]};
  const V3 v2 = cross(v0, v1);
  U[0] = u0.x; U[1] = u1.x; U[2] = u2.x;
  U[3] = u0.y; U[4] = u1.y; U[5] = u2.y;
  U[6] = u0.z; U[7] = u1.z; U[8] = u2.z;
  V[0] = v0.x; V[1] = v1.x; V[2] = v2.x;
  V[3] = v0.y; V[4] = v1.y; V[5] = v2.y;
  V[6] = v0.z; V[7] = v1.z; V[8] = v2.z;
}

// Register version of svd3 for kernels small enough to inline it (same operations in the same
// order as svd3 above; column pairs are template parameters so G and W never leave registers).
template <int P, int Q>
KML_DI bool jacobi_pair_r(double (&G)[9], double (&W)[9]) {
  const double a = kfma(G[6 + P], G[6 + P], kfma(G[3 + P], G[3 + P], G[P] * G[P]));
  const double b = kfma(G[6 + Q], G[6 + Q], kfma(G[3 + Q], G[3 + Q], G[Q] * G[Q]));
  const double g = kfma(G[6 + P], G[6 + Q], kfma(G[3 + P], G[3 + Q], G[P] * G[Q]));
  if (g * g <= 1e-30 * a * b) return false;
  const double h = b - a, tg = 2.0 * g;  // t = sgn(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = h / tg, times 2|g|: one division
  const double sgn = (h == 0.0 || (h > 0.0) == (g > 0.0)) ? 1.0 : -1.0;
  const double t = kdiv(sgn * fabs(tg), fabs(h) + ksqrt(kfma(h, h, tg * tg)));
  const double c = krsqrt(kfma(t, t, 1.0));
  const double s = c * t;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double gp = G[3 * i + P], gq = G[3 * i + Q];
    G[3 * i + P] = kfma(c, gp, -(s * gq));
    G[3 * i + Q] = kfma(s, gp, c * gq);
    const double wp = W[3 * i + P], wq = W[3 * i + Q];
    W[3 * i + P] = kfma(c, wp, -(s * wq));
    W[3 * i + Q] = kfma(s, wp, c * wq);
  }
  return true;
}
KML_DI V3 column_of(const double (&M)[9], int j) {  // column j of a row-major 3x3, j in registers
  V3 v;
  v.x = j == 0 ? M[0] : (j == 1 ? M[1] : M[2]);
  v.y = j == 0 ? M[3] : (j == 1 ? M[4] : M[5]);
  v.z = j == 0 ? M[6] : (j == 1 ? M[7] : M[8]);
  return v;
}
KML_DI void svd3_r(const double* A, double* U, double* S, double* V) {
  double G[9], W[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    G[i] = A[i];
    W[i] = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
  }
#pragma unroll 1
  for (int sweep = 0; sweep < 12; ++sweep) {
    bool rotated = jacobi_pair_r<0, 1>(G, W);
    rotated = jacobi_pair_r<0, 2>(G, W) || rotated;
    rotated = jacobi_pair_r<1, 2>(G, W) || rotated;
    if (!rotated) break;
  }
  double n[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) n[j] = ksqrt(kfma(G[6 + j], G[6 + j], kfma(G[3 + j], G[3 + j], G[j] * G[j])));
  // stable descending order of the three column norms
  int i0 = 0, i1 = 1, i2 = 2;
  double n0 = n[0], n1 = n[1], n2 = n[2];
  if (n1 > n0) { const int t = i0; i0 = i1; i1 = t; const double d = n0; n0 = n1; n1 = d; }
  if (n2 > n1) { const int t = i1; i1 = i2; i2 = t; const double d = n1; n1 = n2; n2 = d; }
  if (n1 > n0) { const int t = i0; i0 = i1; i1 = t; const double d = n0; n0 = n1; n1 = d; }
  S[0] = n0; S[1] = n1; S[2] = n2;
  V3 u0, u1;
  if (n0 > 0.0) {
    const V3 g0 = column_of(G, i0);
    const double r0 = kdiv(1.0, n0);
    u0.x = g0.x * r0; u0.y = g0.y * r0; u0.z = g0.z * r0;
  } else {
    u0.x = 1.0; u0.y = 0.0; u0.z = 0.0;
  }
  if (n1 > 0.0) {
    const V3 g1 = column_of(G, i1);
    const double r1 = kdiv(1.0, n1);
    u1.x = g1.x * r1; u1.y = g1.y * r1; u1.z = g1.z * r1;
  } else {
    int k = 0;
    double m = fabs(u0.x);
    if (fabs(u0.y) < m) { k = 1; m = fabs(u0.y); }
    if (fabs(u0.z) < m) { k = 2; }
    const V3 e = {k == 0 ? 1.0 : 0.0, k == 1 ? 1.0 : 0.0, k == 2 ? 1.0 : 0.0};
    const V3 c = cross(u0, e);
    const double nn = ksqrt(dot(c, c));
    u1.x = kdiv(c.x, nn); u1.y = kdiv(c.y, nn); u1.z = kdiv(c.z, nn);
  }
  const V3 u2 = cross(u0, u1);
  const V3 v0 = column_of(W, i0), v1 = column_of(W, i1);
  const V3 v2 = cross(v0, v1);
  U[0] = u0.x; U[1] = u1.x; U[2] = u2.x;
  U[3] = u0.y; U[4] = u1.y; U[5] = u2.y;
  U[6] = u0.z; U[7] = u1.z; U[8] = u2.z;
  V[0] = v0.x; V[1] = v1.x; V[2] = v2.x;
  V[3] = v0.y; V[4] = v1.y; V[5] = v2.y;
  V[6] = v0.z; V[7] = v1.z; V[8] = v2.z;
}

// ------------------------------------------------------------------ Arun
// p1 = R p2 + t from three correspondences; M = [R|t] row-major 3x4.
KML_DN void arun3(const double* a1, const double* b1, const double* c1, const double* a2,
                  const double* b2, const double* c2, double* M) {
  double m1[3], m2[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    m1[i] = ((a1[i] + b1[i]) + c1[i]) / 3.0;
    m2[i] = ((a2[i] + b2[i]) + c2[i]) / 3.0;
  }
  double H[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) H[i] = 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double* P1 = (k == 0) ? a1 : ((k == 1) ? b1 : c1);
    const double* P2 = (k == 0) ? a2 : ((k == 1) ? b2 : c2);
    double d1[3], d2[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      d1[i] = P1[i] - m1[i];
      d2[i] = P2[i] - m2[i];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) H[3 * r + c] = kfma(d2[r], d1[c], H[3 * r + c]);
  }
  double U[9], S[3], V[9];
  svd3_r(H, U, S, V);
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      M[4 * r + c] = kfma(V[3 * r + 2], U[3 * c + 2], kfma(V[3 * r + 1], U[3 * c + 1], V[3 * r + 0] * U[3 * c + 0]));
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double rc = kfma(M[4 * r + 2], m2[2], kfma(M[4 * r + 1], m2[1], M[4 * r + 0] * m2[0]));
    M[4 * r + 3] = m1[r] - rc;
  }
}

KML_DI double arun_sqdist(const double* M, double p1x, double p1y, double p1z, double p2x,
                          double p2y, double p2z) {
  const double x = kfma(M[2], p2z, kfma(M[1], p2y, kfma(M[0], p2x, M[3])));
  const double y = kfma(M[6], p2z, kfma(M[5], p2y, kfma(M[4], p2x, M[7])));
  const double z = kfma(M[10], p2z, kfma(M[9], p2y, kfma(M[8], p2x, M[11])));
  const double ex = p1x - x, ey = p1y - y, ez = p1z - z;
  return kfma(ez, ez, kfma(ey, ey, ex * ex));
}

// ------------------------------------------------------- mono residual
// (1 - f1.p^) + (1 - f2.q^), p = triangulate2(f1, f2 | R12, t12), q = R^T p - R^T t
KML_DI void mono_tinv(const double* M, double* tinv) {
  tinv[0] = -kfma(M[8], M[11], kfma(M[4], M[7], M[0] * M[3]));
  tinv[1] = -kfma(M[9], M[11], kfma(M[5], M[7], M[1] * M[3]));
  tinv[2] = -kfma(M[10], M[11], kfma(M[6], M[7], M[2] * M[3]));
}
KML_DI double mono_residual(const double* M, const double* tinv, const V3& f1, const V3& f2) {
  const V3 t = {M[3], M[7], M[11]};
  V3 f2u;
  f2u.x = kfma(M[2], f2.z, kfma(M[1], f2.y, M[0] * f2.x));
  f2u.y = kfma(M[6], f2.z, kfma(M[5], f2.y, M[4] * f2.x));
  f2u.z = kfma(M[10], f2.z, kfma(M[9], f2.y, M[8] * f2.x));
  const double b0 = dot(t, f1), b1 = dot(t, f2u);
  const double d12 = dot(f1, f2u);
  const double A00 = dot(f1, f1), A01 = -d12, A10 = d12, A11 = -dot(f2u, f2u);
  const double det = kfma(A00, A11, -(A01 * A10));
  const double rdet = kdiv(1.0, det);
  const double l0 = kfma(A11, b0, -(A01 * b1)) * rdet;
  const double l1 = kfma(A00, b1, -(A10 * b0)) * rdet;
  V3 p, q;
  p.x = 0.5 * kfma(l0, f1.x, kfma(l1, f2u.x, t.x));
  p.y = 0.5 * kfma(l0, f1.y, kfma(l1, f2u.y, t.y));
  p.z = 0.5 * kfma(l0, f1.z, kfma(l1, f2u.z, t.z));
  q.x = kfma(M[8], p.z, kfma(M[4], p.y, kfma(M[0], p.x, tinv[0])));
  q.y = kfma(M[9], p.z, kfma(M[5], p.y, kfma(M[1], p.x, tinv[1])));
  q.z = kfma(M[10], p.z, kfma(M[6], p.y, kfma(M[2], p.x, tinv[2])));
  const double e1 = 1.0 - dot(f1, p) * krsqrt(dot(p, p));
  const double e2 = 1.0 - dot(f2, q) * krsqrt(dot(q, q));
  return e1 + e2;
}

// ----------------------------------------------- inlier test, fast filter
// Decides `mono_residual(...) < thr` WITHOUT the six IEEE divisions / square roots
// for all but the correspondences whose residual lies within a rigorous error margin
// of the threshold (those, and every badly conditioned triangulation, return -1 and
// are re-evaluated with mono_residual, so the decision is always the exact one).
// Preconditions: unit bearings (SacState::unit_bearings).
// Error budget (e~ = filter value, e = value of mono_residual; eps = 2.2e-16, r = 1/det,
// |det| = sin^2(a) with a the angle between f1 and R f2, S = |l0| + |l1| + |t|_1).  The two
// evaluations differ by FMA contraction and <= 2 ulp reciprocal / rsqrt: |num~ - num| <= 8 eps |t|,
// |det~ - det| <= 4 eps, hence |dl0|, |dl1| <= 8 eps S |r|.  They move p along f1 and R f2, which
// are a apart: the component across the direction of p is <= eps S (4 sqrt|r| + 8 sin(x) |r|),
// x the TRUE angle between f1 and p.  With the guards |p|, |q| >= S / 1000 the direction of p (and
// of q, by the same argument with the cameras swapped) moves by dx <= 1000 eps (4 sqrt|r| + 8 sin(x) |r|),
// and e = (1 - cos x1) + (1 - cos x2) gives sin x_i <= u := sqrt(2 e), so
//     |e~ - e| <= g(e) := 2000 eps u (4 sqrt|r| + 8 u |r|).
// |r| is capped at min(1e6, 6000 / c), c := sqrt(2 (thr + 2e-6)).  Then g(e) <= 4e-8 c for every
// e <= thr + 2e-6, and dg/de < 1e-2 beyond, so with the margin m := 4e-7 c (10x):
//   true e <  thr  =>  e~ < thr + m            (never reported as outlier, which needs e~ > thr + m)
//   true e >= thr  =>  e~ >= e - g(e) > thr - m (never reported as inlier, which needs e~ < thr - m)
KML_DI double rcp_fast(double d, bool* ok) {
  double x;
#ifndef KML_HOST_EMULATION
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
#else  // tests/device_math_host.cpp: any start value within the Newton steps' basin gives the same decisions
  x = (double)(float)(1.0 / d);
#endif
  double e = __fma_rn(-d, x, 1.0);
  x = __fma_rn(x, e, x);
  e = __fma_rn(-d, x, 1.0);
  *ok = fabs(e) < 1e-8;  // remaining relative error e^2 < 1e-16
  return __fma_rn(x, e, x);
}
KML_DI double rsqrt_fast(double a, bool* ok) {
  double y;
#ifndef KML_HOST_EMULATION
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
#else
  y = (double)(float)(1.0 / sqrt(a));
#endif
  double e = __fma_rn(-(a * y), y, 1.0);
  y = __fma_rn(0.5 * y, e, y);
  e = __fma_rn(-(a * y), y, 1.0);
  *ok = fabs(e) < 1e-8;
  return __fma_rn(0.5 * y, e, y);
}
KML_DI double dotf(const V3& a, const V3& b) { return __fma_rn(a.z, b.z, __fma_rn(a.y, b.y, a.x * b.x)); }
struct InlierMargins {
  double lo, hi, det_min;  // inlier if e~ < lo, outlier if e~ > hi, else (or |det| < det_min) undecided
};
KML_DI InlierMargins inlier_margins(double thr) {
  InlierMargins m;
  const double c = ksqrt(2.0 * (thr + 2e-6));
  const double near = c * 4e-7 + 1e-12;
  m.lo = thr - near;
  m.hi = thr + near;
  m.det_min = fmax(1e-6, kdiv(c, 6000.0));  // |r| <= min(1e6, 6000 / c)
  return m;
}
// 1 = inlier, 0 = outlier, -1 = undecided (caller evaluates mono_residual)
KML_DI int mono_inlier_fast(const double* M, const double* tinv, const V3& f1, const V3& f2,
                            const InlierMargins& mg) {
  const V3 t = {M[3], M[7], M[11]};
  V3 f2u;
  f2u.x = __fma_rn(M[2], f2.z, __fma_rn(M[1], f2.y, M[0] * f2.x));
  f2u.y = __fma_rn(M[6], f2.z, __fma_rn(M[5], f2.y, M[4] * f2.x));
  f2u.z = __fma_rn(M[10], f2.z, __fma_rn(M[9], f2.y, M[8] * f2.x));
  const double b0 = dotf(t, f1), b1 = dotf(t, f2u), d12 = dotf(f1, f2u);
  const double A00 = dotf(f1, f1), n2 = dotf(f2u, f2u);
  const double det = __fma_rn(d12, d12, -(A00 * n2));
  bool ok0, ok1, ok2;
  const double r = rcp_fast(det, &ok0);
  const double l0 = __fma_rn(d12, b1, -(n2 * b0)) * r;
  const double l1 = __fma_rn(A00, b1, -(d12 * b0)) * r;
  V3 p, q;
  p.x = 0.5 * __fma_rn(l0, f1.x, __fma_rn(l1, f2u.x, t.x));
  p.y = 0.5 * __fma_rn(l0, f1.y, __fma_rn(l1, f2u.y, t.y));
  p.z = 0.5 * __fma_rn(l0, f1.z, __fma_rn(l1, f2u.z, t.z));
  q.x = __fma_rn(M[8], p.z, __fma_rn(M[4], p.y, __fma_rn(M[0], p.x, tinv[0])));
  q.y = __fma_rn(M[9], p.z, __fma_rn(M[5], p.y, __fma_rn(M[1], p.x, tinv[1])));
  q.z = __fma_rn(M[10], p.z, __fma_rn(M[6], p.y, __fma_rn(M[2], p.x, tinv[2])));
  const double pp = dotf(p, p), qq = dotf(q, q);
  const double s1 = (fabs(l0) + fabs(l1)) + ((fabs(t.x) + fabs(t.y)) + fabs(t.z));
  const double y1 = rsqrt_fast(pp, &ok1), y2 = rsqrt_fast(qq, &ok2);
  const double e = 2.0 - __fma_rn(dotf(f1, p), y1, dotf(f2, q) * y2);
  // conditioning guards (all false on NaN): |det| floor, |f2u| sane, |p|, |q| >= S / 1000
  const double floor2 = 1e-6 * (s1 * s1);
  const bool sane = ok0 & ok1 & ok2 & (fabs(det) >= mg.det_min) & (n2 > 0.5) & (n2 < 2.0) & (pp >= floor2) &
                    (qq >= floor2) & (pp < 1e280) & (qq < 1e280);
#ifdef KML_FILTER_STATS
  {
    atomicAdd(&g_fstats[0], 1ull);
    if (!ok0) atomicAdd(&g_fstats[1], 1ull);
    if (!(ok1 & ok2)) atomicAdd(&g_fstats[2], 1ull);
    if (!(fabs(det) >= mg.det_min)) atomicAdd(&g_fstats[3], 1ull);
    if (!((n2 > 0.5) & (n2 < 2.0))) atomicAdd(&g_fstats[4], 1ull);
    if (!(pp >= floor2)) atomicAdd(&g_fstats[5], 1ull);
    if (!(qq >= floor2)) atomicAdd(&g_fstats[6], 1ull);
    if (sane && !(e < mg.lo) && !(e > mg.hi)) atomicAdd(&g_fstats[7], 1ull);
    if (!sane) atomicAdd(&g_fstats[8], 1ull);
  }
#endif
  if (!sane) return -1;
  if (e < mg.lo) return 1;
  if (e > mg.hi) return 0;
  return -1;
}

}  // namespace geom
}  // namespace kml
