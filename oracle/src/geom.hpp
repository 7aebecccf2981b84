// oracle/src/geom.hpp — scalar fp64 geometry of the CPU oracle (TEST
// INFRASTRUCTURE, see oracle/kmo.h).  Restates, from SURVEY.md Appendix A
// (the upstream sources are not vendored in /root/reference):
//   A.6  opengv relative_pose::fivept_nister + essential decomposition +
//        triangulate2 + bearing reprojection residual
//        (opengv/src/relative_pose/methods.cpp, modules/fivept_nister/modules.cpp,
//         src/sac_problems/relative_pose/CentralRelativePoseSacProblem.cpp,
//         src/triangulation/methods.cpp, src/math/Sturm.cpp)
//   A.8  opengv point_cloud::threept_arun + 3-D residual
//        (opengv/src/point_cloud/methods.cpp, PointCloudSacProblem.cpp)
// The ARITHMETIC CONTRACT (operation order, fixed sweep counts, tie rules) is
// written out in DESIGN.md §4; the CUDA kernels implement the same contract
// independently.  Compile with -ffp-contract=off: the fused multiply-adds of the contract are the
// explicit kfma() calls, nothing else is contracted.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace kmo {

// The contract's fused multiply-add: a*b + c rounded once.  Written out wherever the contract fuses
// (never left to the compiler: -ffp-contract=off); the CUDA kernels write __fma_rn at the same places.
static inline double kfma(double a, double b, double c) { return __builtin_fma(a, b, c); }
// The contract's reciprocal square root: a fixed sequence of integer and IEEE operations (no sqrt, no
// division) — the magic-constant start value (3.4 % off at most) and four Newton steps y <- y (3/2 - (x/2) y^2),
// each squaring the error; the result is within a few ulp of 1/sqrt(x) and the same bits everywhere.
static inline double krsqrt(double x) {
  uint64_t i;
  std::memcpy(&i, &x, 8);
  i = 0x5FE6EB50C7B537A9ull - (i >> 1);
  double y;
  std::memcpy(&y, &i, 8);
  const double h = 0.5 * x;
  y = y * kfma(-(h * y), y, 1.5);
  y = y * kfma(-(h * y), y, 1.5);
  y = y * kfma(-(h * y), y, 1.5);
  y = y * kfma(-(h * y), y, 1.5);
  return y;
}

// ---------------------------------------------------------------- 3-vectors
static inline double dot3(const double* a, const double* b) {
  return kfma(a[2], b[2], kfma(a[1], b[1], a[0] * b[0]));
}
static inline void cross3(const double* a, const double* b, double* c) {
  c[0] = kfma(a[1], b[2], -(a[2] * b[1]));
  c[1] = kfma(a[2], b[0], -(a[0] * b[2]));
  c[2] = kfma(a[0], b[1], -(a[1] * b[0]));
}
// y = M x (M row-major 3x3)
static inline void matvec3(const double* M, const double* x, double* y) {
  y[0] = (M[0] * x[0] + M[1] * x[1]) + M[2] * x[2];
  y[1] = (M[3] * x[0] + M[4] * x[1]) + M[5] * x[2];
  y[2] = (M[6] * x[0] + M[7] * x[1]) + M[8] * x[2];
}
// y = M^T x
static inline void matTvec3(const double* M, const double* x, double* y) {
  y[0] = (M[0] * x[0] + M[3] * x[1]) + M[6] * x[2];
  y[1] = (M[1] * x[0] + M[4] * x[1]) + M[7] * x[2];
  y[2] = (M[2] * x[0] + M[5] * x[1]) + M[8] * x[2];
}

// ------------------------------------------------------------------- svd3
// One-sided (Hestenes) Jacobi SVD of a 3x3 matrix, "proper" form:
// A = U diag(S) V^T with S sorted descending, U = [u0 u1 u0xu1],
// V = [v0 v1 v0xv1] (both det +1).  The third singular value is reported as
// the norm of the third rotated column.  At most 12 sweeps over the pairs
// (0,1),(0,2),(1,2); a rotation is skipped when g*g <= 1e-30*a*b; t = sgn 2|g| / (|b-a| + sqrt((b-a)^2 + 4g^2)).
static const int kSvdSweeps = 12;
static inline void svd3(const double* A, double* U, double* S, double* V) {
  double G[9], W[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 9; ++i) G[i] = A[i];
  static const int P[3] = {0, 0, 1}, Q[3] = {1, 2, 2};
  for (int sweep = 0; sweep < kSvdSweeps; ++sweep) {
    bool rotated = false;
    for (int k = 0; k < 3; ++k) {
      const int p = P[k], q = Q[k];
      double a = kfma(G[6 + p], G[6 + p], kfma(G[3 + p], G[3 + p], G[p] * G[p]));
      double b = kfma(G[6 + q], G[6 + q], kfma(G[3 + q], G[3 + q], G[q] * G[q]));
      double g = kfma(G[6 + p], G[6 + q], kfma(G[3 + p], G[3 + q], G[p] * G[q]));
      if (g * g <= 1e-30 * a * b) continue;
      rotated = true;
      // t = sgn(zeta) / (|zeta| + sqrt(1 + zeta^2)) with zeta = (b - a) / (2 g), multiplied through by 2|g|:
      // one division instead of two
      const double h = b - a, tg = 2.0 * g;
      const double sgn = (h == 0.0 || (h > 0.0) == (g > 0.0)) ? 1.0 : -1.0;
      double t = (sgn * std::fabs(tg)) / (std::fabs(h) + std::sqrt(kfma(h, h, tg * tg)));
      double c = krsqrt(kfma(t, t, 1.0));
      double s = c * t;
      for (int i = 0; i < 3; ++i) {
        double gp = G[3 * i + p], gq = G[3 * i + q];
        G[3 * i + p] = kfma(c, gp, -(s * gq));
        G[3 * i + q] = kfma(s, gp, c * gq);
        double wp = W[3 * i + p], wq = W[3 * i + q];
        W[3 * i + p] = kfma(c, wp, -(s * wq));
        W[3 * i + q] = kfma(s, wp, c * wq);
      }
    }
    if (!rotated) break;
  }
  double n[3];
  for (int j = 0; j < 3; ++j)
    n[j] = std::sqrt(kfma(G[6 + j], G[6 + j], kfma(G[3 + j], G[3 + j], G[j] * G[j])));
  // stable descending order of the three column norms
  int i0 = 0, i1 = 1, i2 = 2;
  if (n[i1] > n[i0]) { int t = i0; i0 = i1; i1 = t; }
  if (n[i2] > n[i1]) { int t = i1; i1 = i2; i2 = t; }
  if (n[i1] > n[i0]) { int t = i0; i0 = i1; i1 = t; }
  S[0] = n[i0]; S[1] = n[i1]; S[2] = n[i2];
  double u0[3], u1[3], u2[3], v0[3], v1[3], v2[3];
  if (n[i0] > 0.0) {
    const double r0 = 1.0 / n[i0];
    for (int i = 0; i < 3; ++i) u0[i] = G[3 * i + i0] * r0;
  } else {
    u0[0] = 1.0; u0[1] = 0.0; u0[2] = 0.0;
  }
  if (n[i1] > 0.0) {
    const double r1 = 1.0 / n[i1];
    for (int i = 0; i < 3; ++i) u1[i] = G[3 * i + i1] * r1;
  } else {
    // rank <= 1: any unit vector orthogonal to u0 (deterministic choice)
    int k = 0;
    if (std::fabs(u0[1]) < std::fabs(u0[k])) k = 1;
    if (std::fabs(u0[2]) < std::fabs(u0[k])) k = 2;
    double e[3] = {0, 0, 0};
    e[k] = 1.0;
    cross3(u0, e, u1);
    double nn = std::sqrt(dot3(u1, u1));
    for (int i = 0; i < 3; ++i) u1[i] = u1[i] / nn;
  }
  cross3(u0, u1, u2);
  for (int i = 0; i < 3; ++i) { v0[i] = W[3 * i + i0]; v1[i] = W[3 * i + i1]; }
  cross3(v0, v1, v2);
  for (int i = 0; i < 3; ++i) {
    U[3 * i + 0] = u0[i]; U[3 * i + 1] = u1[i]; U[3 * i + 2] = u2[i];
    V[3 * i + 0] = v0[i]; V[3 * i + 1] = v1[i]; V[3 * i + 2] = v2[i];
  }
}

// ------------------------------------------------------------------- Arun
// A.8 arun_complete on 3 correspondences: p1 = R p2 + t.
// model = [R | t] row-major 3x4.
static inline void arun3(const double* a1, const double* b1, const double* c1,
                         const double* a2, const double* b2, const double* c2,
                         double* model) {
  double m1[3], m2[3];
  for (int i = 0; i < 3; ++i) {
    m1[i] = ((a1[i] + b1[i]) + c1[i]) / 3.0;
    m2[i] = ((a2[i] + b2[i]) + c2[i]) / 3.0;
  }
  const double* P1[3] = {a1, b1, c1};
  const double* P2[3] = {a2, b2, c2};
  double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = 0; k < 3; ++k) {
    double d1[3], d2[3];
    for (int i = 0; i < 3; ++i) { d1[i] = P1[k][i] - m1[i]; d2[i] = P2[k][i] - m2[i]; }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) H[3 * r + c] = kfma(d2[r], d1[c], H[3 * r + c]);
  }
  double U[9], S[3], V[9];
  svd3(H, U, S, V);
  // R = V U^T
  double* R = model;  // write into strided 3x4
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      R[4 * r + c] = kfma(V[3 * r + 2], U[3 * c + 2], kfma(V[3 * r + 1], U[3 * c + 1], V[3 * r + 0] * U[3 * c + 0]));
  for (int r = 0; r < 3; ++r) {
    double rc = kfma(R[4 * r + 2], m2[2], kfma(R[4 * r + 1], m2[1], R[4 * r + 0] * m2[0]));
    model[4 * r + 3] = m1[r] - rc;
  }
}

// squared 3-D residual |p1 - (R p2 + t)|^2 with the contract's op order
static inline double arun_sqdist(const double* M, const double* p1, const double* p2) {
  double x = kfma(M[2], p2[2], kfma(M[1], p2[1], kfma(M[0], p2[0], M[3])));
  double y = kfma(M[6], p2[2], kfma(M[5], p2[1], kfma(M[4], p2[0], M[7])));
  double z = kfma(M[10], p2[2], kfma(M[9], p2[1], kfma(M[8], p2[0], M[11])));
  double ex = p1[0] - x, ey = p1[1] - y, ez = p1[2] - z;
  return kfma(ez, ez, kfma(ey, ey, ex * ex));
}

// ---------------------------------------------------------- mono residual
// A.6 getSelectedDistancesToModel: (1 - f1.p^) + (1 - f2.q^),
// p = triangulate2(f1,f2 | R12,t12), q = R^T p - R^T t.
// M = [R12 | t12] row-major 3x4; tinv = -(R^T t) is passed precomputed.
static inline void mono_tinv(const double* M, double* tinv) {
  double t[3] = {M[3], M[7], M[11]};
  tinv[0] = -kfma(M[8], t[2], kfma(M[4], t[1], M[0] * t[0]));
  tinv[1] = -kfma(M[9], t[2], kfma(M[5], t[1], M[1] * t[0]));
  tinv[2] = -kfma(M[10], t[2], kfma(M[6], t[1], M[2] * t[0]));
}
static inline double mono_residual(const double* M, const double* tinv,
                                   const double* f1, const double* f2) {
  const double t[3] = {M[3], M[7], M[11]};
  double f2u[3];
  f2u[0] = kfma(M[2], f2[2], kfma(M[1], f2[1], M[0] * f2[0]));
  f2u[1] = kfma(M[6], f2[2], kfma(M[5], f2[1], M[4] * f2[0]));
  f2u[2] = kfma(M[10], f2[2], kfma(M[9], f2[1], M[8] * f2[0]));
  double b0 = dot3(t, f1), b1 = dot3(t, f2u);
  double d12 = dot3(f1, f2u);
  double A00 = dot3(f1, f1), A01 = -d12, A10 = d12, A11 = -dot3(f2u, f2u);
  double det = kfma(A00, A11, -(A01 * A10));
  const double rdet = 1.0 / det;
  double l0 = kfma(A11, b0, -(A01 * b1)) * rdet;
  double l1 = kfma(A00, b1, -(A10 * b0)) * rdet;
  double p[3], q[3];
  for (int i = 0; i < 3; ++i) p[i] = 0.5 * kfma(l0, f1[i], kfma(l1, f2u[i], t[i]));
  q[0] = kfma(M[8], p[2], kfma(M[4], p[1], kfma(M[0], p[0], tinv[0])));
  q[1] = kfma(M[9], p[2], kfma(M[5], p[1], kfma(M[1], p[0], tinv[1])));
  q[2] = kfma(M[10], p[2], kfma(M[6], p[1], kfma(M[2], p[0], tinv[2])));
  double e1 = 1.0 - dot3(f1, p) * krsqrt(dot3(p, p));
  double e2 = 1.0 - dot3(f2, q) * krsqrt(dot3(q, q));
  return e1 + e2;
}

// ------------------------------------------------------ 5-point (Nister)
// polynomial bookkeeping in (x,y,z):
//  deg1 [x y z 1]
//  deg2 [x2 xy xz y2 yz z2 x y z 1]
//  deg3 [x3 y3 x2y xy2 x2z x2 y2z y2 xyz xy | xz2 xz x yz2 yz y z3 z2 z 1]
// M12[i][j]: deg2 slot of (deg1 i)*(deg1 j); M23[i][j]: deg3 slot of (deg2 i)*(deg1 j)
static const int M12[4][4] = {{0, 1, 2, 6}, {1, 3, 4, 7}, {2, 4, 5, 8}, {6, 7, 8, 9}};
static const int M23[10][4] = {
    {0, 2, 4, 5},      // x2 * {x,y,z,1}
    {2, 3, 8, 9},      // xy
    {4, 8, 10, 11},    // xz
    {3, 1, 6, 7},      // y2
    {8, 6, 13, 14},    // yz
    {10, 13, 16, 17},  // z2
    {5, 9, 11, 12},    // x
    {9, 7, 14, 15},    // y
    {11, 14, 17, 18},  // z
    {12, 15, 18, 19}}; // 1

// out(deg2) = a(deg1)*b(deg1); products accumulated in (i,j) loop order
static inline void pmul11(const double* a, const double* b, double* out) {
  for (int k = 0; k < 10; ++k) out[k] = 0.0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) out[M12[i][j]] = kfma(a[i], b[j], out[M12[i][j]]);
}
// out(deg3) += a(deg2)*b(deg1)
static inline void pmul21_acc(const double* a, const double* b, double* out) {
  for (int i = 0; i < 10; ++i)
    for (int j = 0; j < 4; ++j) out[M23[i][j]] = kfma(a[i], b[j], out[M23[i][j]]);
}

// Null space of the 5x9 epipolar system by Householder QR of its transpose:
// basis[4][9] = last four columns of the 9x9 orthogonal factor.
static inline void nullspace5x9(const double Q[5][9], double basis[4][9]) {
  double A[9][5];
  for (int r = 0; r < 5; ++r)
    for (int c = 0; c < 9; ++c) A[c][r] = Q[r][c];
  double v[5][9];
  double vn2[5], beta[5];  // beta = 2 / |v|^2: one division per reflector
  for (int k = 0; k < 5; ++k) {
    double s2 = 0.0;
    for (int i = k; i < 9; ++i) s2 = kfma(A[i][k], A[i][k], s2);
    double nrm = std::sqrt(s2);
    double alpha = (A[k][k] >= 0.0) ? -nrm : nrm;
    for (int i = 0; i < 9; ++i) v[k][i] = (i < k) ? 0.0 : A[i][k];
    v[k][k] = v[k][k] - alpha;
    double n2 = 0.0;
    for (int i = k; i < 9; ++i) n2 = kfma(v[k][i], v[k][i], n2);
    vn2[k] = n2;
    beta[k] = 2.0 / n2;
    if (n2 > 0.0) {
      for (int j = k; j < 5; ++j) {
        double d = 0.0;
        for (int i = k; i < 9; ++i) d = kfma(v[k][i], A[i][j], d);
        double f = d * beta[k];
        for (int i = k; i < 9; ++i) A[i][j] = kfma(-f, v[k][i], A[i][j]);
      }
    }
  }
  for (int b = 0; b < 4; ++b) {
    double x[9];
    for (int i = 0; i < 9; ++i) x[i] = (i == 5 + b) ? 1.0 : 0.0;
    for (int k = 4; k >= 0; --k) {
      if (!(vn2[k] > 0.0)) continue;
      double d = 0.0;
      for (int i = k; i < 9; ++i) d = kfma(v[k][i], x[i], d);
      double f = d * beta[k];
      for (int i = k; i < 9; ++i) x[i] = kfma(-f, v[k][i], x[i]);
    }
    for (int i = 0; i < 9; ++i) basis[b][i] = x[i];
  }
}

// ---- univariate polynomials, coefficients ascending (c[0] + c[1] z + ...)
static inline double horner(const double* c, int deg, double x) {
  double r = c[deg];
  for (int i = deg - 1; i >= 0; --i) r = kfma(r, x, c[i]);
  return r;
}
// out = a*b  (degrees da, db), accumulation in (i,j) loop order
static inline void upmul(const double* a, int da, const double* b, int db, double* out) {
  for (int k = 0; k <= da + db; ++k) out[k] = 0.0;
  for (int i = 0; i <= da; ++i)
    for (int j = 0; j <= db; ++j) out[i + j] = kfma(a[i], b[j], out[i + j]);
}

// Sturm chain of a degree-10 polynomial.  chain[k] has degree deg[k];
// every remainder is negated and scaled by 1/|leading coefficient| (its own leading coefficient is -+1 exactly).
struct Sturm {
  double c[12][11];
  int deg[12];
  int len;
};
static inline void sturm_build(const double* p, int n, Sturm* st) {
  // strip exact-zero leading coefficients
  while (n > 0 && p[n] == 0.0) --n;
  for (int i = 0; i <= n; ++i) st->c[0][i] = p[i];
  st->deg[0] = n;
  st->len = 1;
  if (n < 1) return;
  for (int i = 0; i < n; ++i) st->c[1][i] = (double)(i + 1) * p[i + 1];
  st->deg[1] = n - 1;
  st->len = 2;
  while (st->deg[st->len - 1] > 0 && st->len < 12) {
    const double* a = st->c[st->len - 2];
    const double* b = st->c[st->len - 1];
    int da = st->deg[st->len - 2], db = st->deg[st->len - 1];
    double r[11];
    for (int i = 0; i <= da; ++i) r[i] = a[i];
    // a remainder's leading coefficient is exactly +-1 (below): dividing by it is multiplying by it
    const bool unit_lc = st->len - 1 >= 2;
    for (int d = da; d >= db; --d) {
      double f = unit_lc ? r[d] * b[db] : r[d] / b[db];
      for (int i = 0; i < db; ++i) r[d - db + i] = kfma(-f, b[i], r[d - db + i]);
      r[d] = 0.0;
    }
    int dr = db - 1;
    while (dr >= 0 && r[dr] == 0.0) --dr;
    if (dr < 0) break;  // exact gcd reached
    // negated and scaled by the reciprocal of |leading coefficient| (one division per remainder); the
    // leading coefficient itself is set to -+1 exactly
    const double inv = 1.0 / std::fabs(r[dr]);
    double* o = st->c[st->len];
    for (int i = 0; i < dr; ++i) o[i] = -(r[i] * inv);
    o[dr] = r[dr] > 0.0 ? -1.0 : 1.0;
    st->deg[st->len] = dr;
    st->len++;
  }
}
static inline int sturm_count(const Sturm* st, double x) {
  int changes = 0, last = 0;
  for (int k = 0; k < st->len; ++k) {
    double v = horner(st->c[k], st->deg[k], x);
    int s = (v > 0.0) - (v < 0.0);
    if (s != 0) {
      if (last != 0 && s != last) ++changes;
      last = s;
    }
  }
  return changes;
}

static const int kRootGrid = 32;    // sign-test cells on (-1,1]
static const int kRootGrid2 = 256;  // second, finer grid for the chains the first one does not separate
static int g_root_grid2 = 1;        // test knob (kmo_debug_root_grid2): 0 = skip it, every such chain is bisected
static const int kRootDepth = 48;   // max Sturm bisection depth per root
static const int kRootBisect = 6;   // sign bisection steps on the isolated bracket
static const int kRootNewton = 6;   // bracketed Newton steps that follow

// Real roots of p (degree n<=10) in (-1,1], ascending.  Root j (rank from the
// left) is isolated on its own by bisection on the Sturm count, then refined
// by 10 sign bisections and 8 bracket-safeguarded Newton steps on p.  Every
// root is an independent computation (one GPU lane per root).  Returns the
// number of roots found (<= 10); roots whose bracket shows no sign change
// (numerically multiple roots) are dropped.
static inline int roots_unit(const double* p, int n, double* roots) {
  Sturm st;
  sturm_build(p, n, &st);
  if (st.deg[0] < 1) return 0;
  const int vm1 = sturm_count(&st, -1.0), vp1 = sturm_count(&st, 1.0);
  int R = vm1 - vp1;
  if (R > 10) R = 10;
  int nr = 0;
  const double* c0 = st.c[0];
  const int d0 = st.deg[0];
  const double* c1 = st.c[1];  // p' (unscaled)
  const int d1 = st.deg[1];
  // Grid pass: signs of p on the 32 cells (x_{i-1}, x_i], x_i = -1 + i/16.  A
  // cell brackets a root if the sign changes across it or p(x_i) == 0.  If the
  // number of bracketing cells equals the Sturm count R, every one of them
  // holds exactly one distinct root and no other cell holds any, so the cells
  // ARE the isolating brackets; otherwise (two roots in one cell) each root is
  // isolated by bisection on the Sturm count.
  unsigned cells = 0u;
  int nb = 0;
  if (R > 0) {
    double fprev = horner(c0, d0, -1.0);
    for (int i = 1; i <= kRootGrid; ++i) {
      const double fi = horner(c0, d0, -1.0 + (double)i * (2.0 / kRootGrid));
      if ((fprev < 0.0 && fi > 0.0) || (fprev > 0.0 && fi < 0.0) || fi == 0.0) {
        cells |= 1u << (i - 1);
        ++nb;
      }
      fprev = fi;
    }
  }
  const bool grid_ok = (nb == R);
  // Second grid, 256 cells, for the ~8 % of the polynomials whose first grid shows fewer bracketing
  // cells than roots (two roots in one cell): the same rule on x_i = -1 + i/128.  Only if that fails
  // too (roots closer than 1/128, or a numerically multiple root) the Sturm count is bisected.
  bool cells2[kRootGrid2];
  int nb2 = 0;
  if (R > 0 && !grid_ok && g_root_grid2) {
    double fprev = horner(c0, d0, -1.0);
    for (int i = 1; i <= kRootGrid2; ++i) {
      const double fi = horner(c0, d0, -1.0 + (double)i * (2.0 / kRootGrid2));
      cells2[i - 1] = (fprev < 0.0 && fi > 0.0) || (fprev > 0.0 && fi < 0.0) || fi == 0.0;
      nb2 += cells2[i - 1] ? 1 : 0;
      fprev = fi;
    }
  }
  const bool grid2_ok = !grid_ok && g_root_grid2 && (nb2 == R);
  int cell = -1;
  for (int j = 0; j < R; ++j) {
    double lo = -1.0, hi = 1.0;
    if (grid_ok) {
      do { ++cell; } while (!((cells >> cell) & 1u));
      lo = -1.0 + (double)cell * (2.0 / kRootGrid);
      hi = -1.0 + (double)(cell + 1) * (2.0 / kRootGrid);
    } else if (grid2_ok) {
      do { ++cell; } while (!cells2[cell]);
      lo = -1.0 + (double)cell * (2.0 / kRootGrid2);
      hi = -1.0 + (double)(cell + 1) * (2.0 / kRootGrid2);
    } else {
      int vlo = vm1, vhi = vp1, jj = j;
      for (int depth = 0; depth < kRootDepth; ++depth) {
        if (vlo - vhi == 1) break;
        const double mid = 0.5 * (lo + hi);
        const int vm = sturm_count(&st, mid);
        const int left = vlo - vm;  // roots in (lo, mid]
        if (jj < left) { hi = mid; vhi = vm; } else { jj -= left; lo = mid; vlo = vm; }
      }
    }
    double flo = horner(c0, d0, lo);
    const double fhi = horner(c0, d0, hi);
    if (fhi == 0.0) { roots[nr++] = hi; continue; }
    if (!((flo < 0.0 && fhi > 0.0) || (flo > 0.0 && fhi < 0.0))) continue;
    for (int it = 0; it < kRootBisect; ++it) {
      const double mid = 0.5 * (lo + hi);
      const double fm = horner(c0, d0, mid);
      if ((fm < 0.0) == (flo < 0.0)) { lo = mid; flo = fm; } else { hi = mid; }
    }
    double x = 0.5 * (lo + hi);
    for (int it = 0; it < kRootNewton; ++it) {
      const double fx = horner(c0, d0, x);
      const double dfx = horner(c1, d1, x);
      if ((fx < 0.0) == (flo < 0.0)) { lo = x; flo = fx; } else { hi = x; }
      double xn = x - fx / dfx;
      if (!(xn >= lo && xn <= hi)) xn = 0.5 * (lo + hi);
      x = xn;
    }
    roots[nr++] = x;
  }
  return nr;
}

// null space basis B (X, Y, Z, W) and the 10x20 cubic constraint matrix A in Nister's column order
// [x3 y3 x2y xy2 x2z x2 y2z y2 xyz xy | xz2 xz x yz2 yz y z3 z2 z 1]
static inline void fivept_constraints(const double f1[5][3], const double f2[5][3], double B[4][9], double A[10][20]) {
  double Q[5][9];
  for (int k = 0; k < 5; ++k)
    for (int j = 0; j < 3; ++j)
      for (int i = 0; i < 3; ++i) Q[k][3 * j + i] = f2[k][i] * f1[k][j];
  nullspace5x9(Q, B);  // B[0]=X, B[1]=Y, B[2]=Z, B[3]=W (row-major 3x3 each)
  // entries of E(x,y,z) as deg1 polynomials
  double Ep[9][4];
  for (int e = 0; e < 9; ++e)
    for (int b = 0; b < 4; ++b) Ep[e][b] = B[b][e];
  for (int r = 0; r < 10; ++r)
    for (int c = 0; c < 20; ++c) A[r][c] = 0.0;
  // row 0: det(E)
  {
    double m[10], n[10], d[10];
    // E00 (E11 E22 - E12 E21)
    pmul11(Ep[4], Ep[8], m); pmul11(Ep[5], Ep[7], n);
    for (int k = 0; k < 10; ++k) d[k] = m[k] - n[k];
    pmul21_acc(d, Ep[0], A[0]);
    // - E01 (E10 E22 - E12 E20)  ==  + E01 (E12 E20 - E10 E22)
    pmul11(Ep[5], Ep[6], m); pmul11(Ep[3], Ep[8], n);
    for (int k = 0; k < 10; ++k) d[k] = m[k] - n[k];
    pmul21_acc(d, Ep[1], A[0]);
    // + E02 (E10 E21 - E11 E20)
    pmul11(Ep[3], Ep[7], m); pmul11(Ep[4], Ep[6], n);
    for (int k = 0; k < 10; ++k) d[k] = m[k] - n[k];
    pmul21_acc(d, Ep[2], A[0]);
  }
  // rows 1..9: (E E^T - 0.5 tr(E E^T) I) E
  {
    double EEt[3][3][10];
    for (int i = 0; i < 3; ++i)
      for (int j = i; j < 3; ++j) {
        double a[10], b[10], c[10];
        pmul11(Ep[3 * i + 0], Ep[3 * j + 0], a);
        pmul11(Ep[3 * i + 1], Ep[3 * j + 1], b);
        pmul11(Ep[3 * i + 2], Ep[3 * j + 2], c);
        for (int k = 0; k < 10; ++k) {
          EEt[i][j][k] = (a[k] + b[k]) + c[k];
          EEt[j][i][k] = EEt[i][j][k];
        }
      }
    double htr[10];
    for (int k = 0; k < 10; ++k) htr[k] = 0.5 * ((EEt[0][0][k] + EEt[1][1][k]) + EEt[2][2][k]);
    for (int i = 0; i < 3; ++i)
      for (int k = 0; k < 10; ++k) EEt[i][i][k] = EEt[i][i][k] - htr[k];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double* row = A[1 + 3 * i + j];
        for (int k = 0; k < 3; ++k) pmul21_acc(EEt[i][k], Ep[3 * k + j], row);
      }
  }
}

// Gauss-Jordan with partial pivoting (first maximum) on the first 10 columns; false = singular
static inline bool fivept_gauss_jordan(double A[10][20]) {
  for (int c = 0; c < 10; ++c) {
    int pr = c;
    double pv = std::fabs(A[c][c]);
    for (int r = c + 1; r < 10; ++r) {
      double v = std::fabs(A[r][c]);
      if (v > pv) { pv = v; pr = r; }
    }
    if (!(pv > 0.0)) return false;  // singular constraint system
    if (pr != c)
      for (int j = 0; j < 20; ++j) { double t = A[c][j]; A[c][j] = A[pr][j]; A[pr][j] = t; }
    const double inv = 1.0 / A[c][c];  // pivot row scaled by the reciprocal (one division per pivot)
    for (int j = 0; j < 20; ++j) A[c][j] = A[c][j] * inv;
    for (int r = 0; r < 10; ++r) {
      if (r == c) continue;
      double f = A[r][c];
      for (int j = 0; j < 20; ++j) A[r][j] = kfma(-f, A[c][j], A[r][j]);
    }
  }
  return true;
}

// A.6 fivept_nister: f1 = bearings in frame 1 (query), f2 = frame 2 (match);
// solves f1^T E f2 = 0 with E = [t12]x R12.  Returns #solutions, E row-major.
static inline int fivept_nister(const double f1[5][3], const double f2[5][3], double E[10][9]) {
  double B[4][9], A[10][20];
  fivept_constraints(f1, f2, B, A);
  if (!fivept_gauss_jordan(A)) return 0;
  // B(z): rows <k>=<e>-z<f>, <l>=<g>-z<h>, <m>=<i>-z<j>; columns [x y 1];
  // coefficients ascending in z.
  double Bz[3][3][5];
  static const int RE[3] = {4, 6, 8}, RF[3] = {5, 7, 9};
  for (int r = 0; r < 3; ++r) {
    const double* e = A[RE[r]];
    const double* f = A[RF[r]];
    for (int col = 0; col < 2; ++col) {
      int o = 10 + 3 * col;  // [.. z2, z, 1] blocks for x then y
      Bz[r][col][0] = e[o + 2];
      Bz[r][col][1] = e[o + 1] - f[o + 2];
      Bz[r][col][2] = e[o + 0] - f[o + 1];
      Bz[r][col][3] = -f[o + 0];
      Bz[r][col][4] = 0.0;
    }
    Bz[r][2][0] = e[19];
    Bz[r][2][1] = e[18] - f[19];
    Bz[r][2][2] = e[17] - f[18];
    Bz[r][2][3] = e[16] - f[17];
    Bz[r][2][4] = -f[16];
  }
  // cofactors of the third row -> p1,p2,p3; det via expansion along row 2
  double t1[9], t2[9], p1[8], p2[8], p3[7];
  upmul(Bz[0][1], 3, Bz[1][2], 4, t1); upmul(Bz[0][2], 4, Bz[1][1], 3, t2);
  for (int k = 0; k < 8; ++k) p1[k] = t1[k] - t2[k];
  upmul(Bz[0][2], 4, Bz[1][0], 3, t1); upmul(Bz[0][0], 3, Bz[1][2], 4, t2);
  for (int k = 0; k < 8; ++k) p2[k] = t1[k] - t2[k];
  upmul(Bz[0][0], 3, Bz[1][1], 3, t1); upmul(Bz[0][1], 3, Bz[1][0], 3, t2);
  for (int k = 0; k < 7; ++k) p3[k] = t1[k] - t2[k];
  double n1[11], n2[11], n3[11], nz[11];
  upmul(p1, 7, Bz[2][0], 3, n1);
  upmul(p2, 7, Bz[2][1], 3, n2);
  upmul(p3, 6, Bz[2][2], 4, n3);
  for (int k = 0; k < 11; ++k) nz[k] = (n1[k] + n2[k]) + n3[k];
  // reversed polynomial for |z| > 1
  double rz[11];
  for (int k = 0; k < 11; ++k) rz[k] = nz[10 - k];
  double zr[20];
  int nroots = 0;
  {
    double r[10];
    int n = roots_unit(nz, 10, r);
    for (int i = 0; i < n; ++i) zr[nroots++] = r[i];
    n = roots_unit(rz, 10, r);
    for (int i = 0; i < n; ++i) {
      if (r[i] == 1.0 || r[i] == 0.0) continue;  // z=1 already covered; u=0 is z=inf
      if (nroots < 10) zr[nroots++] = 1.0 / r[i];  // a degree-10 polynomial: keep the first 10
    }
  }
  int ns = 0;
  for (int k = 0; k < nroots && ns < 10; ++k) {
    double z = zr[k];
    double rd = 1.0 / horner(p3, 6, z);
    double x = horner(p1, 7, z) * rd;
    double y = horner(p2, 7, z) * rd;
    bool ok = true;
    for (int e = 0; e < 9; ++e) {
      double v = kfma(z, B[2][e], kfma(y, B[1][e], kfma(x, B[0][e], B[3][e])));
      if (!std::isfinite(v)) ok = false;
      E[ns][e] = v;
    }
    if (ok) ++ns;
  }
  return ns;
}

// Row f4: fivept_stewenius (opengv/src/relative_pose/modules/fivept_stewenius/modules.cpp), selected by
// ransac_2d2d_algorithm: 0 (/root/reference/params/D455/LcdParams.yaml:73).  Same null space and the same
// ten cubic constraints as Nister's solver, but the columns are ordered by degree
//   [x3 x2y x2z xy2 xyz xz2 y3 y2z yz2 z3 | x2 xy xz y2 yz z2 x y z 1],
// Gauss-Jordan on the ten cubic monomials expresses each of them in the quotient-ring basis
// b = [x2 xy xz y2 yz z2 x y z 1], and the ACTION MATRIX of multiplication by x in that basis,
//   rows 0..5 = -B[0..5] (x*x2 = x3, x*xy = x2y, x*xz = x2z, x*y2 = xy2, x*yz = xyz, x*z2 = xz2),
//   rows 6..9 = unit rows (x*x = x2, x*y = xy, x*z = xz, x*1 = x),
// has the solutions' x as eigenvalues and b(solution) as eigenvectors.  Upstream hands the 10x10
// matrix to Eigen's complex EigenSolver and keeps the real part of every (also complex) eigenvector;
// the restatement keeps the REAL solutions: eigenvalues = real roots of the characteristic polynomial
// (elimination to Hessenberg form with first-maximum pivoting, then the recurrence over leading
// minors), found by the same Sturm / bisection / Newton code as Nister's polynomial, and for every
// root the eigenvector from rows 0..5 of (M - x I) v = 0 with v = [x2, xy, xz, v3, v4, v5, x, y, z, 1]:
// six equations, five unknowns (v3, v4, v5, y, z), Gaussian elimination with first-maximum row pivoting.
static const int kStewCol[20] = {0, 6, 1, 3, 2, 10, 7, 13, 4, 11, 5, 12, 16, 8, 14, 17, 9, 15, 18, 19};  // new column of Nister column c

// characteristic polynomial det(lambda I - M) of a 10x10 matrix, ascending coefficients c[0..10] (c[10] = 1)
static inline void charpoly10(double H[10][10], double* c) {
  const int n = 10;
  for (int m = 1; m < n - 1; ++m) {
    double x = 0.0;
    int i = m;
    for (int j = m; j < n; ++j)
      if (std::fabs(H[j][m - 1]) > std::fabs(x)) { x = H[j][m - 1]; i = j; }
    if (i != m) {
      for (int j = m - 1; j < n; ++j) { const double t = H[i][j]; H[i][j] = H[m][j]; H[m][j] = t; }
      for (int j = 0; j < n; ++j) { const double t = H[j][i]; H[j][i] = H[j][m]; H[j][m] = t; }
    }
    if (x != 0.0) {
      for (int i2 = m + 1; i2 < n; ++i2) {
        double y = H[i2][m - 1];
        if (y != 0.0) {
          y = y / x;
          H[i2][m - 1] = y;
          for (int j = m; j < n; ++j) H[i2][j] = kfma(-y, H[m][j], H[i2][j]);
          for (int j = 0; j < n; ++j) H[j][m] = kfma(y, H[j][i2], H[j][m]);
        }
      }
    }
  }
  // P[k] = characteristic polynomial of the leading k x k block, monic: P[k][k] = 1
  double P[11][11];
  for (int k = 0; k <= n; ++k)
    for (int j = 0; j <= n; ++j) P[k][j] = 0.0;
  P[0][0] = 1.0;
  for (int k = 1; k <= n; ++k) {
    const int col = k - 1;
    const double h = H[col][col];
    for (int j = 0; j <= k; ++j) P[k][j] = j <= k - 1 ? kfma(-h, P[k - 1][j], j >= 1 ? P[k - 1][j - 1] : 0.0) : P[k - 1][j - 1];
    double prod = 1.0;
    for (int i = 1; i <= k - 1; ++i) {
      const int row = col - i;
      prod = prod * H[row + 1][row];
      const double sc = H[row][col] * prod;
      for (int j = 0; j <= k - 1 - i; ++j) P[k][j] = kfma(-sc, P[k - 1 - i][j], P[k][j]);
    }
  }
  for (int j = 0; j <= n; ++j) c[j] = P[n][j];
}

// y, z of the solution with eigenvalue x from the six non-trivial rows a[6][10] of the action matrix
static inline bool stewenius_yz(const double a[6][10], double x, double* y, double* z) {
  const double x2 = x * x;
  double C[6][5], d[6];
  for (int j = 0; j < 6; ++j) {
    C[j][0] = a[j][3]; C[j][1] = a[j][4]; C[j][2] = a[j][5];
    C[j][3] = kfma(a[j][1], x, a[j][7]);
    C[j][4] = kfma(a[j][2], x, a[j][8]);
    d[j] = -kfma(a[j][0], x2, kfma(a[j][6], x, a[j][9]));
  }
  d[0] = kfma(x, x2, d[0]);   // row 0: lambda * v0 = x * x2
  C[1][3] = C[1][3] - x2;  // row 1: lambda * v1 = x * (x y)
  C[2][4] = C[2][4] - x2;  // row 2: lambda * v2 = x * (x z)
  C[3][0] = C[3][0] - x;   // rows 3..5: lambda * v3, v4, v5
  C[4][1] = C[4][1] - x;
  C[5][2] = C[5][2] - x;
  for (int c = 0; c < 5; ++c) {
    int pr = c;
    double pv = std::fabs(C[c][c]);
    for (int r = c + 1; r < 6; ++r) {
      const double v = std::fabs(C[r][c]);
      if (v > pv) { pv = v; pr = r; }
    }
    if (!(pv > 0.0)) return false;
    if (pr != c) {
      for (int j = 0; j < 5; ++j) { const double t = C[c][j]; C[c][j] = C[pr][j]; C[pr][j] = t; }
      const double t = d[c]; d[c] = d[pr]; d[pr] = t;
    }
    for (int r = c + 1; r < 6; ++r) {
      const double f = C[r][c] / C[c][c];
      for (int j = c + 1; j < 5; ++j) C[r][j] = kfma(-f, C[c][j], C[r][j]);
      d[r] = kfma(-f, d[c], d[r]);
    }
  }
  double u[5];
  for (int c = 4; c >= 0; --c) {
    double sacc = d[c];
    for (int j = c + 1; j < 5; ++j) sacc = kfma(-C[c][j], u[j], sacc);
    u[c] = sacc / C[c][c];
  }
  *y = u[3];
  *z = u[4];
  return true;
}

static inline int fivept_stewenius(const double f1[5][3], const double f2[5][3], double E[10][9]) {
  double B[4][9], An[10][20], A[10][20];
  fivept_constraints(f1, f2, B, An);
  for (int r = 0; r < 10; ++r)
    for (int c = 0; c < 20; ++c) A[r][kStewCol[c]] = An[r][c];
  if (!fivept_gauss_jordan(A)) return 0;
  double M[6][10], H[10][10];
  for (int i = 0; i < 10; ++i)
    for (int j = 0; j < 10; ++j) H[i][j] = 0.0;
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 10; ++j) { M[i][j] = -A[i][10 + j]; H[i][j] = M[i][j]; }
  H[6][0] = 1.0; H[7][1] = 1.0; H[8][2] = 1.0; H[9][6] = 1.0;
  double cp[11], rp[11];
  charpoly10(H, cp);
  for (int k = 0; k < 11; ++k) rp[k] = cp[10 - k];
  double xr[20];
  int nroots = 0;
  {
    double r[10];
    int n = roots_unit(cp, 10, r);
    for (int i = 0; i < n; ++i) xr[nroots++] = r[i];
    n = roots_unit(rp, 10, r);
    for (int i = 0; i < n; ++i) {
      if (r[i] == 1.0 || r[i] == 0.0) continue;
      if (nroots < 10) xr[nroots++] = 1.0 / r[i];
    }
  }
  int ns = 0;
  for (int k = 0; k < nroots && ns < 10; ++k) {
    const double x = xr[k];
    double y, z;
    if (!stewenius_yz(M, x, &y, &z)) continue;
    bool ok = true;
    for (int e = 0; e < 9; ++e) {
      const double v = kfma(z, B[2][e], kfma(y, B[1][e], kfma(x, B[0][e], B[3][e])));
      if (!std::isfinite(v)) ok = false;
      E[ns][e] = v;
    }
    if (ok) ++ns;
  }
  return ns;
}

// A.6 computeModelCoefficients for NISTER: sample of 8 (5 solve + 3 extra).
// f1/f2 are the full correspondence arrays [N][3]; model = [R12|t12] 3x4.
static inline bool mono_model(const double* f1, const double* f2, const uint16_t* sample,
                              double* model, int algorithm = 0) {
  double a[5][3], b[5][3];
  for (int k = 0; k < 5; ++k)
    for (int i = 0; i < 3; ++i) {
      a[k][i] = f1[3 * sample[k] + i];
      b[k][i] = f2[3 * sample[k] + i];
    }
  double E[10][9];
  int ne = algorithm == 1 ? fivept_stewenius(a, b, E) : fivept_nister(a, b, E);
  double best = 1000000.0;
  bool found = false;
  for (int e = 0; e < ne; ++e) {
    double U[9], S[3], V[9];
    svd3(E[e], U, S, V);
    // Ra = U W V^T, Rb = U W^T V^T, W = [0 -1 0; 1 0 0; 0 0 1]
    // U W   = [ u1 -u0 u2 ],  U W^T = [ -u1 u0 u2 ]  (columns)
    double Ra[9], Rb[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        double u0 = U[3 * r + 0], u1 = U[3 * r + 1], u2 = U[3 * r + 2];
        double v0 = V[3 * c + 0], v1 = V[3 * c + 1], v2 = V[3 * c + 2];
        Ra[3 * r + c] = kfma(u2, v2, kfma(u1, v0, -(u0 * v1)));
        Rb[3 * r + c] = kfma(u2, v2, kfma(u0, v1, -(u1 * v0)));
      }
    double tt[3] = {S[0] * U[2], S[0] * U[5], S[0] * U[8]};
    for (int cand = 0; cand < 4; ++cand) {
      const double* R = (cand < 2) ? Ra : Rb;
      double sgn = (cand & 1) ? -1.0 : 1.0;
      double M[12];
      for (int r = 0; r < 3; ++r) {
        M[4 * r + 0] = R[3 * r + 0]; M[4 * r + 1] = R[3 * r + 1]; M[4 * r + 2] = R[3 * r + 2];
        M[4 * r + 3] = sgn * tt[r];
      }
      double tinv[3];
      mono_tinv(M, tinv);
      double q = 0.0;
      for (int k = 0; k < 8; ++k)
        q = q + mono_residual(M, tinv, f1 + 3 * sample[k], f2 + 3 * sample[k]);
      if (q < best) {
        best = q;
        found = true;
        for (int i = 0; i < 12; ++i) model[i] = M[i];
      }
    }
  }
  return found;
}

}  // namespace kmo
