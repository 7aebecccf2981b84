// geom.cuh — fp64 minimal solvers and residuals, device side.
//
// Implements the ARITHMETIC CONTRACT of DESIGN.md §4 (operation order, sweep
// counts, tie rules) for:
//   - opengv::relative_pose::fivept_nister + essential-matrix decomposition +
//     triangulate2 + bearing reprojection residual, as used by
//     CentralRelativePoseSacProblem (SURVEY.md A.6;
//     /root/reference/images/kimera-multi.drawio:2589-2592, 2646)
//   - opengv::point_cloud::threept_arun + 3-D residual, as used by
//     PointCloudSacProblem (drawio:2595-2598, 2654)
// Compiled with -fmad=false: every multiply and add below is rounded on its
// own, exactly as written.  fp64 div and sqrt are IEEE on sm_100a.
// The functions are also compilable for the host (KML_HD) so that a CPU test
// harness can single-step them; the product only calls them from kernels.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define KML_HD __host__ __device__ __forceinline__
#else
#define KML_HD inline
#endif

namespace kml {
namespace geom {

struct V3 {
  double x, y, z;
};
KML_HD double dot(const V3& a, const V3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
KML_HD V3 cross(const V3& a, const V3& b) {
  V3 c;
  c.x = a.y * b.z - a.z * b.y;
  c.y = a.z * b.x - a.x * b.z;
  c.z = a.x * b.y - a.y * b.x;
  return c;
}

// 3x3 matrices are passed as 9 doubles, row-major.
struct M3 {
  double m[9];
};

// Hestenes one-sided Jacobi, "proper" SVD: A = U diag(S) V^T, S descending,
// third columns completed by cross products (det U = det V = +1).
KML_HD void jacobi_pair(double* G, double* W, int p, int q, bool& rotated) {
  const double a = (G[p] * G[p] + G[3 + p] * G[3 + p]) + G[6 + p] * G[6 + p];
  const double b = (G[q] * G[q] + G[3 + q] * G[3 + q]) + G[6 + q] * G[6 + q];
  const double g = (G[p] * G[q] + G[3 + p] * G[3 + q]) + G[6 + p] * G[6 + q];
  if (g * g <= 1e-30 * a * b) return;
  rotated = true;
  const double zeta = (b - a) / (2.0 * g);
  const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  const double c = 1.0 / sqrt(1.0 + t * t);
  const double s = c * t;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double gp = G[3 * i + p], gq = G[3 * i + q];
    G[3 * i + p] = c * gp - s * gq;
    G[3 * i + q] = s * gp + c * gq;
    const double wp = W[3 * i + p], wq = W[3 * i + q];
    W[3 * i + p] = c * wp - s * wq;
    W[3 * i + q] = s * wp + c * wq;
  }
}

KML_HD void svd3(const double* A, V3& u0, V3& u1, V3& u2, double* S, V3& v0, V3& v1, V3& v2) {
  double G[9], W[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    G[i] = A[i];
    W[i] = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
  }
  for (int sweep = 0; sweep < 12; ++sweep) {
    bool rotated = false;
    jacobi_pair(G, W, 0, 1, rotated);
    jacobi_pair(G, W, 0, 2, rotated);
    jacobi_pair(G, W, 1, 2, rotated);
    if (!rotated) break;
  }
  const double n0 = sqrt((G[0] * G[0] + G[3] * G[3]) + G[6] * G[6]);
  const double n1 = sqrt((G[1] * G[1] + G[4] * G[4]) + G[7] * G[7]);
  const double n2 = sqrt((G[2] * G[2] + G[5] * G[5]) + G[8] * G[8]);
  // stable descending sort of (n0,n1,n2) carrying the column vectors
  double s0 = n0, s1 = n1, s2 = n2;
  V3 g0 = {G[0], G[3], G[6]}, g1 = {G[1], G[4], G[7]}, g2 = {G[2], G[5], G[8]};
  V3 w0 = {W[0], W[3], W[6]}, w1 = {W[1], W[4], W[7]}, w2 = {W[2], W[5], W[8]};
#define KML_SWAP_COLS(sa, sb, ga, gb, wa, wb) \
  {                                           \
    double ts = sa; sa = sb; sb = ts;         \
    V3 tg = ga; ga = gb; gb = tg;             \
    V3 tw = wa; wa = wb; wb = tw;             \
  }
  if (s1 > s0) KML_SWAP_COLS(s0, s1, g0, g1, w0, w1)
  if (s2 > s1) KML_SWAP_COLS(s1, s2, g1, g2, w1, w2)
  if (s1 > s0) KML_SWAP_COLS(s0, s1, g0, g1, w0, w1)
#undef KML_SWAP_COLS
  S[0] = s0; S[1] = s1; S[2] = s2;
  if (s0 > 0.0) {
    u0.x = g0.x / s0; u0.y = g0.y / s0; u0.z = g0.z / s0;
  } else {
    u0.x = 1.0; u0.y = 0.0; u0.z = 0.0;
  }
  if (s1 > 0.0) {
    u1.x = g1.x / s1; u1.y = g1.y / s1; u1.z = g1.z / s1;
  } else {
    int k = 0;
    double m = fabs(u0.x);
    if (fabs(u0.y) < m) { k = 1; m = fabs(u0.y); }
    if (fabs(u0.z) < m) { k = 2; }
    V3 e = {k == 0 ? 1.0 : 0.0, k == 1 ? 1.0 : 0.0, k == 2 ? 1.0 : 0.0};
    V3 c = cross(u0, e);
    const double nn = sqrt(dot(c, c));
    u1.x = c.x / nn; u1.y = c.y / nn; u1.z = c.z / nn;
  }
  u2 = cross(u0, u1);
  v0 = w0;
  v1 = w1;
  v2 = cross(v0, v1);
}

// ------------------------------------------------------------------ Arun
// p1 = R p2 + t from three correspondences; M = [R|t] row-major 3x4.
KML_HD void arun3(const double* a1, const double* b1, const double* c1, const double* a2,
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
      for (int c = 0; c < 3; ++c) H[3 * r + c] = H[3 * r + c] + d2[r] * d1[c];
  }
  V3 u0, u1, u2, v0, v1, v2;
  double S[3];
  svd3(H, u0, u1, u2, S, v0, v1, v2);
  const double U[9] = {u0.x, u1.x, u2.x, u0.y, u1.y, u2.y, u0.z, u1.z, u2.z};
  const double V[9] = {v0.x, v1.x, v2.x, v0.y, v1.y, v2.y, v0.z, v1.z, v2.z};
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      M[4 * r + c] = (V[3 * r + 0] * U[3 * c + 0] + V[3 * r + 1] * U[3 * c + 1]) +
                     V[3 * r + 2] * U[3 * c + 2];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double rc = (M[4 * r + 0] * m2[0] + M[4 * r + 1] * m2[1]) + M[4 * r + 2] * m2[2];
    M[4 * r + 3] = m1[r] - rc;
  }
}

KML_HD double arun_sqdist(const double* M, double p1x, double p1y, double p1z, double p2x,
                          double p2y, double p2z) {
  const double x = ((M[0] * p2x + M[1] * p2y) + M[2] * p2z) + M[3];
  const double y = ((M[4] * p2x + M[5] * p2y) + M[6] * p2z) + M[7];
  const double z = ((M[8] * p2x + M[9] * p2y) + M[10] * p2z) + M[11];
  const double ex = p1x - x, ey = p1y - y, ez = p1z - z;
  return (ex * ex + ey * ey) + ez * ez;
}

// ------------------------------------------------------- mono residual
KML_HD void mono_tinv(const double* M, double* tinv) {
  tinv[0] = -((M[0] * M[3] + M[4] * M[7]) + M[8] * M[11]);
  tinv[1] = -((M[1] * M[3] + M[5] * M[7]) + M[9] * M[11]);
  tinv[2] = -((M[2] * M[3] + M[6] * M[7]) + M[10] * M[11]);
}
KML_HD double mono_residual(const double* M, const double* tinv, const V3& f1, const V3& f2) {
  const V3 t = {M[3], M[7], M[11]};
  V3 f2u;
  f2u.x = (M[0] * f2.x + M[1] * f2.y) + M[2] * f2.z;
  f2u.y = (M[4] * f2.x + M[5] * f2.y) + M[6] * f2.z;
  f2u.z = (M[8] * f2.x + M[9] * f2.y) + M[10] * f2.z;
  const double b0 = dot(t, f1), b1 = dot(t, f2u);
  const double d12 = dot(f1, f2u);
  const double A00 = dot(f1, f1), A01 = -d12, A10 = d12, A11 = -dot(f2u, f2u);
  const double det = A00 * A11 - A01 * A10;
  const double l0 = (A11 * b0 - A01 * b1) / det;
  const double l1 = (A00 * b1 - A10 * b0) / det;
  V3 p, q;
  p.x = 0.5 * (l0 * f1.x + (t.x + l1 * f2u.x));
  p.y = 0.5 * (l0 * f1.y + (t.y + l1 * f2u.y));
  p.z = 0.5 * (l0 * f1.z + (t.z + l1 * f2u.z));
  q.x = ((M[0] * p.x + M[4] * p.y) + M[8] * p.z) + tinv[0];
  q.y = ((M[1] * p.x + M[5] * p.y) + M[9] * p.z) + tinv[1];
  q.z = ((M[2] * p.x + M[6] * p.y) + M[10] * p.z) + tinv[2];
  const double np = sqrt(dot(p, p)), nq = sqrt(dot(q, q));
  const double e1 = 1.0 - dot(f1, p) / np;
  const double e2 = 1.0 - dot(f2, q) / nq;
  return e1 + e2;
}

// ------------------------------------------------------------ 5-point
// monomial slot tables
//  deg1 [x y z 1]; deg2 [x2 xy xz y2 yz z2 x y z 1];
//  deg3 [x3 y3 x2y xy2 x2z x2 y2z y2 xyz xy | xz2 xz x yz2 yz y z3 z2 z 1]
KML_HD int slot12(int i, int j) {
  const int T[16] = {0, 1, 2, 6, 1, 3, 4, 7, 2, 4, 5, 8, 6, 7, 8, 9};
  return T[4 * i + j];
}
KML_HD int slot23(int i, int j) {
  const int T[40] = {0, 2,  4,  5,  2,  3,  8,  9,  4,  8,  10, 11, 3,  1,  6,  7,  8,  6,  13, 14,
                     10, 13, 16, 17, 5,  9,  11, 12, 9,  7,  14, 15, 11, 14, 17, 18, 12, 15, 18, 19};
  return T[4 * i + j];
}
KML_HD void pmul11(const double* a, const double* b, double* out) {
#pragma unroll
  for (int k = 0; k < 10; ++k) out[k] = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) out[slot12(i, j)] = out[slot12(i, j)] + a[i] * b[j];
}
KML_HD void pmul21_acc(const double* a, const double* b, double* out) {
#pragma unroll
  for (int i = 0; i < 10; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) out[slot23(i, j)] = out[slot23(i, j)] + a[i] * b[j];
}

KML_HD double horner(const double* c, int deg, double x) {
  double r = c[deg];
  for (int i = deg - 1; i >= 0; --i) r = r * x + c[i];
  return r;
}
KML_HD void upmul(const double* a, int da, const double* b, int db, double* out) {
  for (int k = 0; k <= da + db; ++k) out[k] = 0.0;
  for (int i = 0; i <= da; ++i)
    for (int j = 0; j <= db; ++j) out[i + j] = out[i + j] + a[i] * b[j];
}

struct Sturm {
  double c[12][11];
  int deg[12];
  int len;
};
KML_HD void sturm_build(const double* p, int n, Sturm& st) {
  while (n > 0 && p[n] == 0.0) --n;
  for (int i = 0; i <= n; ++i) st.c[0][i] = p[i];
  st.deg[0] = n;
  st.len = 1;
  if (n < 1) return;
  for (int i = 0; i < n; ++i) st.c[1][i] = (double)(i + 1) * p[i + 1];
  st.deg[1] = n - 1;
  st.len = 2;
  while (st.deg[st.len - 1] > 0 && st.len < 12) {
    const double* a = st.c[st.len - 2];
    const double* b = st.c[st.len - 1];
    const int da = st.deg[st.len - 2], db = st.deg[st.len - 1];
    double r[11];
    for (int i = 0; i <= da; ++i) r[i] = a[i];
    for (int d = da; d >= db; --d) {
      const double f = r[d] / b[db];
      for (int i = 0; i < db; ++i) r[d - db + i] = r[d - db + i] - f * b[i];
      r[d] = 0.0;
    }
    int dr = db - 1;
    while (dr >= 0 && r[dr] == 0.0) --dr;
    if (dr < 0) break;
    const double sc = fabs(r[dr]);
    double* o = st.c[st.len];
    for (int i = 0; i <= dr; ++i) o[i] = -(r[i] / sc);
    st.deg[st.len] = dr;
    st.len++;
  }
}
KML_HD int sturm_count(const Sturm& st, double x) {
  int changes = 0, last = 0;
  for (int k = 0; k < st.len; ++k) {
    const double v = horner(st.c[k], st.deg[k], x);
    const int s = (v > 0.0) - (v < 0.0);
    if (s != 0) {
      if (last != 0 && s != last) ++changes;
      last = s;
    }
  }
  return changes;
}

constexpr int kRootGrid = 16;
constexpr int kRootDepth = 40;
constexpr int kRootBisect = 60;

// real roots in (-1,1], ascending; Sturm isolation + 60 sign bisections
KML_HD int roots_unit(const double* p, int n, double* roots) {
  Sturm st;
  sturm_build(p, n, st);
  if (st.deg[0] < 1) return 0;
  int nr = 0;
  int V[kRootGrid + 1];
  for (int i = 0; i <= kRootGrid; ++i) V[i] = sturm_count(st, -1.0 + (double)i * (2.0 / kRootGrid));
  for (int cell = 0; cell < kRootGrid; ++cell) {
    if (V[cell] - V[cell + 1] <= 0) continue;
    double slo[kRootDepth + 2], shi[kRootDepth + 2];
    int svlo[kRootDepth + 2], svhi[kRootDepth + 2], sd[kRootDepth + 2];
    int sp = 1;
    slo[0] = -1.0 + (double)cell * (2.0 / kRootGrid);
    shi[0] = -1.0 + (double)(cell + 1) * (2.0 / kRootGrid);
    svlo[0] = V[cell]; svhi[0] = V[cell + 1]; sd[0] = 0;
    while (sp > 0) {
      --sp;
      double lo = slo[sp], hi = shi[sp];
      const int vlo = svlo[sp], vhi = svhi[sp], d = sd[sp];
      const int r = vlo - vhi;
      if (r <= 0) continue;
      if (r == 1 || d >= kRootDepth) {
        double flo = horner(st.c[0], st.deg[0], lo);
        const double fhi = horner(st.c[0], st.deg[0], hi);
        if (fhi == 0.0) {
          if (nr < 10) roots[nr++] = hi;
          continue;
        }
        if (!((flo < 0.0 && fhi > 0.0) || (flo > 0.0 && fhi < 0.0))) continue;
        for (int it = 0; it < kRootBisect; ++it) {
          const double mid = 0.5 * (lo + hi);
          const double fm = horner(st.c[0], st.deg[0], mid);
          if ((fm < 0.0) == (flo < 0.0)) { lo = mid; flo = fm; } else { hi = mid; }
        }
        if (nr < 10) roots[nr++] = 0.5 * (lo + hi);
        continue;
      }
      const double mid = 0.5 * (lo + hi);
      const int vm = sturm_count(st, mid);
      slo[sp] = mid; shi[sp] = hi; svlo[sp] = vm; svhi[sp] = vhi; sd[sp] = d + 1; ++sp;
      slo[sp] = lo; shi[sp] = mid; svlo[sp] = vlo; svhi[sp] = vm; sd[sp] = d + 1; ++sp;
    }
  }
  return nr;
}

// Null space of the 5x9 epipolar system: Householder QR of its transpose.
KML_HD void nullspace5x9(const double* Q /*[5][9]*/, double* basis /*[4][9]*/) {
  double A[9][5];
  for (int r = 0; r < 5; ++r)
    for (int c = 0; c < 9; ++c) A[c][r] = Q[9 * r + c];
  double v[5][9];
  double vn2[5];
  for (int k = 0; k < 5; ++k) {
    double s2 = 0.0;
    for (int i = k; i < 9; ++i) s2 = s2 + A[i][k] * A[i][k];
    const double nrm = sqrt(s2);
    const double alpha = (A[k][k] >= 0.0) ? -nrm : nrm;
    for (int i = 0; i < 9; ++i) v[k][i] = (i < k) ? 0.0 : A[i][k];
    v[k][k] = v[k][k] - alpha;
    double n2 = 0.0;
    for (int i = k; i < 9; ++i) n2 = n2 + v[k][i] * v[k][i];
    vn2[k] = n2;
    if (n2 > 0.0) {
      for (int j = k; j < 5; ++j) {
        double d = 0.0;
        for (int i = k; i < 9; ++i) d = d + v[k][i] * A[i][j];
        const double f = (2.0 * d) / n2;
        for (int i = k; i < 9; ++i) A[i][j] = A[i][j] - f * v[k][i];
      }
    }
  }
  for (int b = 0; b < 4; ++b) {
    double x[9];
    for (int i = 0; i < 9; ++i) x[i] = (i == 5 + b) ? 1.0 : 0.0;
    for (int k = 4; k >= 0; --k) {
      if (!(vn2[k] > 0.0)) continue;
      double d = 0.0;
      for (int i = k; i < 9; ++i) d = d + v[k][i] * x[i];
      const double f = (2.0 * d) / vn2[k];
      for (int i = k; i < 9; ++i) x[i] = x[i] - f * v[k][i];
    }
    for (int i = 0; i < 9; ++i) basis[9 * b + i] = x[i];
  }
}

// f1^T E f2 = 0, E = [t12]x R12; up to 10 solutions, row-major in E[10*9]
KML_HD int fivept_nister(const V3* f1, const V3* f2, double* E) {
  double Q[45];
  for (int k = 0; k < 5; ++k) {
    const double a[3] = {f1[k].x, f1[k].y, f1[k].z};
    const double b[3] = {f2[k].x, f2[k].y, f2[k].z};
    for (int j = 0; j < 3; ++j)
      for (int i = 0; i < 3; ++i) Q[9 * k + 3 * j + i] = b[i] * a[j];
  }
  double B[36];
  nullspace5x9(Q, B);
  double Ep[9][4];
  for (int e = 0; e < 9; ++e)
    for (int b = 0; b < 4; ++b) Ep[e][b] = B[9 * b + e];
  double A[10][20];
  for (int r = 0; r < 10; ++r)
    for (int c = 0; c < 20; ++c) A[r][c] = 0.0;
  {
    double m[10], n[10], d[10];
    pmul11(Ep[4], Ep[8], m); pmul11(Ep[5], Ep[7], n);
    for (int k = 0; k < 10; ++k) d[k] = m[k] - n[k];
    pmul21_acc(d, Ep[0], A[0]);
    pmul11(Ep[5], Ep[6], m); pmul11(Ep[3], Ep[8], n);
    for (int k = 0; k < 10; ++k) d[k] = m[k] - n[k];
    pmul21_acc(d, Ep[1], A[0]);
    pmul11(Ep[3], Ep[7], m); pmul11(Ep[4], Ep[6], n);
    for (int k = 0; k < 10; ++k) d[k] = m[k] - n[k];
    pmul21_acc(d, Ep[2], A[0]);
  }
  {
    double EEt[6][10];  // (0,0) (0,1) (0,2) (1,1) (1,2) (2,2)
    int idx = 0;
    for (int i = 0; i < 3; ++i)
      for (int j = i; j < 3; ++j) {
        double a[10], b[10], c[10];
        pmul11(Ep[3 * i + 0], Ep[3 * j + 0], a);
        pmul11(Ep[3 * i + 1], Ep[3 * j + 1], b);
        pmul11(Ep[3 * i + 2], Ep[3 * j + 2], c);
        for (int k = 0; k < 10; ++k) EEt[idx][k] = (a[k] + b[k]) + c[k];
        ++idx;
      }
    for (int k = 0; k < 10; ++k) {
      const double htr = 0.5 * ((EEt[0][k] + EEt[3][k]) + EEt[5][k]);
      EEt[0][k] = EEt[0][k] - htr;
      EEt[3][k] = EEt[3][k] - htr;
      EEt[5][k] = EEt[5][k] - htr;
    }
    const int sym[9] = {0, 1, 2, 1, 3, 4, 2, 4, 5};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double* row = A[1 + 3 * i + j];
        for (int k = 0; k < 3; ++k) pmul21_acc(EEt[sym[3 * i + k]], Ep[3 * k + j], row);
      }
  }
  for (int c = 0; c < 10; ++c) {
    int pr = c;
    double pv = fabs(A[c][c]);
    for (int r = c + 1; r < 10; ++r) {
      const double v = fabs(A[r][c]);
      if (v > pv) { pv = v; pr = r; }
    }
    if (!(pv > 0.0)) return 0;
    if (pr != c)
      for (int j = 0; j < 20; ++j) { const double t = A[c][j]; A[c][j] = A[pr][j]; A[pr][j] = t; }
    const double piv = A[c][c];
    for (int j = 0; j < 20; ++j) A[c][j] = A[c][j] / piv;
    for (int r = 0; r < 10; ++r) {
      if (r == c) continue;
      const double f = A[r][c];
      for (int j = 0; j < 20; ++j) A[r][j] = A[r][j] - f * A[c][j];
    }
  }
  double Bz[3][3][5];
  for (int r = 0; r < 3; ++r) {
    const double* e = A[4 + 2 * r];
    const double* f = A[5 + 2 * r];
    for (int col = 0; col < 2; ++col) {
      const int o = 10 + 3 * col;
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
  double t1[9], t2[9], p1[8], p2[8], p3[7];
  upmul(Bz[0][1], 3, Bz[1][2], 4, t1); upmul(Bz[0][2], 4, Bz[1][1], 3, t2);
  for (int k = 0; k < 8; ++k) p1[k] = t1[k] - t2[k];
  upmul(Bz[0][2], 4, Bz[1][0], 3, t1); upmul(Bz[0][0], 3, Bz[1][2], 4, t2);
  for (int k = 0; k < 8; ++k) p2[k] = t1[k] - t2[k];
  upmul(Bz[0][0], 3, Bz[1][1], 3, t1); upmul(Bz[0][1], 3, Bz[1][0], 3, t2);
  for (int k = 0; k < 7; ++k) p3[k] = t1[k] - t2[k];
  double n1[11], n2[11], n3[11], nz[11], rz[11];
  upmul(p1, 7, Bz[2][0], 3, n1);
  upmul(p2, 7, Bz[2][1], 3, n2);
  upmul(p3, 6, Bz[2][2], 4, n3);
  for (int k = 0; k < 11; ++k) nz[k] = (n1[k] + n2[k]) + n3[k];
  for (int k = 0; k < 11; ++k) rz[k] = nz[10 - k];
  double zr[20];
  int nroots = 0;
  {
    double r[10];
    int n = roots_unit(nz, 10, r);
    for (int i = 0; i < n; ++i) zr[nroots++] = r[i];
    n = roots_unit(rz, 10, r);
    for (int i = 0; i < n; ++i) {
      if (r[i] == 1.0 || r[i] == 0.0) continue;
      zr[nroots++] = 1.0 / r[i];
    }
  }
  int ns = 0;
  for (int k = 0; k < nroots && ns < 10; ++k) {
    const double z = zr[k];
    const double d = horner(p3, 6, z);
    const double x = horner(p1, 7, z) / d;
    const double y = horner(p2, 7, z) / d;
    bool ok = true;
    for (int e = 0; e < 9; ++e) {
      const double v = ((x * B[e] + y * B[9 + e]) + z * B[18 + e]) + B[27 + e];
      if (!isfinite(v)) ok = false;
      E[9 * ns + e] = v;
    }
    if (ok) ++ns;
  }
  return ns;
}

// computeModelCoefficients (NISTER): 8-point sample, 5 solve + all 8 score.
// f1/f2: the 8 sampled bearings.  Returns false if no model.
KML_HD bool mono_model(const V3* f1, const V3* f2, double* model) {
  double E[90];
  const int ne = fivept_nister(f1, f2, E);
  double best = 1000000.0;
  bool found = false;
  for (int e = 0; e < ne; ++e) {
    V3 u0, u1, u2, v0, v1, v2;
    double S[3];
    svd3(E + 9 * e, u0, u1, u2, S, v0, v1, v2);
    const double U[9] = {u0.x, u1.x, u2.x, u0.y, u1.y, u2.y, u0.z, u1.z, u2.z};
    const double V[9] = {v0.x, v1.x, v2.x, v0.y, v1.y, v2.y, v0.z, v1.z, v2.z};
    double Ra[9], Rb[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        const double a0 = U[3 * r + 0], a1 = U[3 * r + 1], a2 = U[3 * r + 2];
        const double b0 = V[3 * c + 0], b1 = V[3 * c + 1], b2 = V[3 * c + 2];
        Ra[3 * r + c] = (a1 * b0 - a0 * b1) + a2 * b2;
        Rb[3 * r + c] = (a0 * b1 - a1 * b0) + a2 * b2;
      }
    const double tt[3] = {S[0] * U[2], S[0] * U[5], S[0] * U[8]};
    for (int cand = 0; cand < 4; ++cand) {
      const double* R = (cand < 2) ? Ra : Rb;
      const double sgn = (cand & 1) ? -1.0 : 1.0;
      double M[12];
      for (int r = 0; r < 3; ++r) {
        M[4 * r + 0] = R[3 * r + 0];
        M[4 * r + 1] = R[3 * r + 1];
        M[4 * r + 2] = R[3 * r + 2];
        M[4 * r + 3] = sgn * tt[r];
      }
      double tinv[3];
      mono_tinv(M, tinv);
      double q = 0.0;
      for (int k = 0; k < 8; ++k) q = q + mono_residual(M, tinv, f1[k], f2[k]);
      if (q < best) {
        best = q;
        found = true;
        for (int i = 0; i < 12; ++i) model[i] = M[i];
      }
    }
  }
  return found;
}

}  // namespace geom
}  // namespace kml
