// fivept_thread.cuh — the mono minimal solver of one RANSAC round, staged for
// the GPU (opengv fivept_nister + essential decomposition + 8-point
// disambiguation; /root/reference/images/kimera-multi.drawio:
// 2589-2592, 2646).  Every stage executes the operation sequence of the
// ARITHMETIC CONTRACT; the stages only differ in what one
// thread owns:
//   stage 1  mono_front_thread    thread = draw.  Null space (Householder, in
//            registers), 10x20 constraint build (generated straight-line code),
//            Gauss-Jordan (one templated step per pivot column), B(z) cofactors in
//            registers -> n(z), p1, p2, p3, basis (70 doubles).  Only the 10x20
//            system lives in SHARED memory, as S(i) = sm[i*STRIDE + thread]: any
//            per-thread index (pivot row!) is bank-conflict free.  200 slots.
//   stage 2  mono_isolate_thread  thread = draw.  Sturm chains of n(z) and of
//            the reversed polynomial, root counts, 32-cell sign grid (or
//            bisection on the Sturm count when two roots share a cell, deferred
//            to its own compacted launch) -> one isolating bracket per real
//            root.  Generic-position fast path in registers; the generic
//            variable-degree code keeps 88 scratch slots:
//              scratch 0..10 | nz 11..21 | chain (triangular) 22..87
//   stage 3  mono_item            thread = (draw, root) ITEM: refine the root,
//            E, SVD, four (R,t) candidates scored on the 8 sample points.
// Division and sqrt are not inlined in the big kernels and the front CTA's warps
// pass phase barriers: the SM fetches one compact instruction stream (the first
// fully inlined build stalled 63 % of its warp time on instruction fetch).
#pragma once
#include "geom.cuh"

namespace kml {
namespace geom {

constexpr int kTphSlots = 200;    // stage 1
constexpr int kIsoSlots = 88;     // stage 2
constexpr int kFrontOut = 70;     // NISTER: nz[11] p1[8] p2[8] p3[7] basis[36]
// STEWENIUS (row f4): charpoly[11] (unused 23) basis[36] at the same offset 34, then the six
// non-trivial rows of the action matrix [6][10]
constexpr int kFrontOutStew = 130;
constexpr int kStewRowsOff = 70;
// column of Nister's monomial order [x3 y3 x2y xy2 x2z x2 y2z y2 xyz xy | xz2 xz x yz2 yz y z3 z2 z 1]
// in the degree order [x3 x2y x2z xy2 xyz xz2 y3 y2z yz2 z3 | x2 xy xz y2 yz z2 x y z 1]
__host__ __device__ constexpr int stew_slot(int i) {
  constexpr int col[20] = {0, 6, 1, 3, 2, 10, 7, 13, 4, 11, 5, 12, 16, 8, 14, 17, 9, 15, 18, 19};
  return (i / 20) * 20 + col[i % 20];
}
constexpr int kMaxBrackets = 20;  // <= 10 per chain
constexpr int kTRootGrid = 32;
constexpr int kTRootDepth = 48;
constexpr int kTRootBisect = 6;
constexpr int kTRootNewton = 6;

// Horner on a strided shared-memory polynomial (ascending coefficients)
template <int STRIDE>
__device__ __noinline__ double horner_s(const double* c, int deg, double x) {
  double r = c[deg * STRIDE];
  for (int i = deg - 1; i >= 0; --i) r = kfma(r, x, c[i * STRIDE]);
  return r;
}
// coefficient k of a*b, strided operands
template <int STRIDE>
__device__ __noinline__ double conv_s(const double* a, int da, const double* b, int db, int k) {
  double r = 0.0;
  const int i0 = max(0, k - db), i1 = min(da, k);
  for (int i = i0; i <= i1; ++i) r = kfma(a[i * STRIDE], b[(k - i) * STRIDE], r);
  return r;
}

// coefficient k of a*b for register polynomials of degrees DA, DB; with k a compile-time
// constant after unrolling the guards fold away (terms in ascending i, like conv_s)
template <int DA, int DB>
__device__ __forceinline__ double conv_r(const double* a, const double* b, int k) {
  double r = 0.0;
#pragma unroll
  for (int i = 0; i <= DA; ++i)
    if (i >= k - DB && i <= k) r = kfma(a[i], b[k - i], r);
  return r;
}

// One Gauss-Jordan pivot step (column C) of the 10x20 constraint matrix in strided shared
// memory, partial pivoting.  C is a compile-time constant: the scaled pivot row sits in
// registers and every row update is a straight line of loads / stores at immediate offsets.
// Column C itself is never read again, and rows 0..3 are neither pivot candidates nor
// outputs once C >= 4 (only rows 4..9 feed B(z)): those updates are skipped; every value
// that is used later is produced by the same operations as in a full sweep.
// FULL: every row is updated (the Stewenius variant reads rows 0..5 of the reduced system).
template <int STRIDE, int C, bool FULL = false>
__device__ __forceinline__ void gj_step(double* sm, bool& failed) {
#define S(i) sm[(i) * STRIDE]
  int pr = C;
  double pv = fabs(S(C * 20 + C));
#pragma unroll
  for (int r = C + 1; r < 10; ++r) {
    const double v = fabs(S(r * 20 + C));
    if (v > pv) { pv = v; pr = r; }
  }
  if (!(pv > 0.0)) failed = true;
  if (pr != C) {
    double* rowp = sm + pr * 20 * STRIDE;
#pragma unroll
    for (int j = C; j < 20; ++j) {
      const double t = S(C * 20 + j);
      S(C * 20 + j) = rowp[j * STRIDE];
      rowp[j * STRIDE] = t;
    }
  }
  const double inv = kdiv(1.0, S(C * 20 + C));  // pivot row scaled by the reciprocal
  double prow[20];
#pragma unroll
  for (int j = C + 1; j < 20; ++j) {
    prow[j] = S(C * 20 + j) * inv;
    S(C * 20 + j) = prow[j];
  }
#pragma unroll 1
  for (int r = ((C >= 4 && !FULL) ? 4 : 0); r < 10; ++r) {
    if (r == C) continue;
    double* row = sm + r * 20 * STRIDE;
    const double f = row[C * STRIDE];
#pragma unroll
    for (int j = C + 1; j < 20; ++j) row[j * STRIDE] = kfma(-f, prow[j], row[j * STRIDE]);
  }
#undef S
}

// Characteristic polynomial det(lambda I - H) of the 10x10 matrix H[r][c] = S(r*20 + 10 + c)
// (row f4; contract: DESIGN.md §4.12): elimination to Hessenberg form with first-maximum
// pivoting, then the recurrence over the leading minors, whose polynomials P[k] (k <= 9, k + 1
// coefficients, monic) live in the dead left halves S(k*20 + 0..9).  c[0..10] ascending, c[10] = 1.
template <int STRIDE>
__device__ __noinline__ void charpoly10_s(double* sm, double* c) {
#define H(r, cc) sm[((r) * 20 + 10 + (cc)) * STRIDE]
#define P(k, j) sm[((k) * 20 + (j)) * STRIDE]
  const int n = 10;
  for (int m = 1; m < n - 1; ++m) {
    double x = 0.0;
    int i = m;
    for (int j = m; j < n; ++j)
      if (fabs(H(j, m - 1)) > fabs(x)) { x = H(j, m - 1); i = j; }
    if (i != m) {
      for (int j = m - 1; j < n; ++j) { const double t = H(i, j); H(i, j) = H(m, j); H(m, j) = t; }
      for (int j = 0; j < n; ++j) { const double t = H(j, i); H(j, i) = H(j, m); H(j, m) = t; }
    }
    if (x != 0.0) {
      for (int i2 = m + 1; i2 < n; ++i2) {
        double y = H(i2, m - 1);
        if (y != 0.0) {
          y = kdiv(y, x);
          H(i2, m - 1) = y;
          for (int j = m; j < n; ++j) H(i2, j) = kfma(-y, H(m, j), H(i2, j));
          for (int j = 0; j < n; ++j) H(j, m) = kfma(y, H(j, i2), H(j, m));
        }
      }
    }
  }
  P(0, 0) = 1.0;
  for (int k = 1; k <= n; ++k) {
    const int col = k - 1;
    const double h = H(col, col);
    // the new polynomial goes to c[] first: P(k-1, .) is still read below
    for (int j = 0; j <= k; ++j) c[j] = j <= k - 1 ? kfma(-h, P(k - 1, j), j >= 1 ? P(k - 1, j - 1) : 0.0) : P(k - 1, j - 1);
    double prod = 1.0;
    for (int i = 1; i <= k - 1; ++i) {
      const int row = col - i;
      prod = prod * H(row + 1, row);
      const double sc = H(row, col) * prod;
      for (int j = 0; j <= k - 1 - i; ++j) c[j] = kfma(-sc, P(k - 1 - i, j), c[j]);
    }
    if (k < n)
      for (int j = 0; j <= k; ++j) P(k, j) = c[j];
  }
#undef H
#undef P
}

// ============================================================== stage 1
// sm: this thread's slot 0.  ga/gb: the problem's correspondences (query /
// match bearings, [N][3]); smp: the 8 sample indices.  Writes the 70 doubles
// of kFrontOut to `out` (NaN basis if the constraint system was singular).
// `alive` = false makes the thread a passenger (barriers only, no output).
// ALG 0 = NISTER (kFrontOut doubles out), 1 = STEWENIUS (kFrontOutStew doubles out).
template <int STRIDE, bool SYNC, int ALG = 0>
__device__ void mono_front_thread(double* sm, const double* __restrict__ ga, const double* __restrict__ gb,
                                  const uint16_t* __restrict__ smp, bool alive, double* __restrict__ out) {
#define S(i) sm[(i) * STRIDE]
#define KML_PHASE() do { if (SYNC) __syncthreads(); } while (0)
  bool failed = !alive;
  int sidx[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) sidx[k] = alive ? (int)smp[k] : 0;
  // ------------------------------------------------ phase 1: null space
  // A9[i][j] = Q[j][i] = f_m[j][i%3] * f_q[j][i/3]; five Householder reflections, then the last
  // four columns of Q are the basis.  Everything is unrolled with static indices, so the 9x5
  // matrix, the reflectors and the basis live in registers.
  double B[36];
  {
    double A[45], V[45], N2[5], BETA[5];  // BETA = 2 / |v|^2: one division per reflector
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const double* fq = ga + 3 * sidx[j];
      const double* fm = gb + 3 * sidx[j];
      const double q0 = fq[0], q1 = fq[1], q2 = fq[2], m0 = fm[0], m1 = fm[1], m2 = fm[2];
      A[0 * 5 + j] = m0 * q0; A[1 * 5 + j] = m1 * q0; A[2 * 5 + j] = m2 * q0;
      A[3 * 5 + j] = m0 * q1; A[4 * 5 + j] = m1 * q1; A[5 * 5 + j] = m2 * q1;
      A[6 * 5 + j] = m0 * q2; A[7 * 5 + j] = m1 * q2; A[8 * 5 + j] = m2 * q2;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      double s2 = 0.0;
#pragma unroll
      for (int i = k; i < 9; ++i) s2 = kfma(A[i * 5 + k], A[i * 5 + k], s2);
      const double nrm = ksqrt(s2);
      const double alpha = (A[k * 5 + k] >= 0.0) ? -nrm : nrm;
#pragma unroll
      for (int i = k; i < 9; ++i) V[k * 9 + i] = A[i * 5 + k];
      V[k * 9 + k] = V[k * 9 + k] - alpha;
      double n2 = 0.0;
#pragma unroll
      for (int i = k; i < 9; ++i) n2 = kfma(V[k * 9 + i], V[k * 9 + i], n2);
      N2[k] = n2;
      BETA[k] = kdiv(2.0, n2);
      if (n2 > 0.0) {
#pragma unroll
        for (int j = k; j < 5; ++j) {
          double d = 0.0;
#pragma unroll
          for (int i = k; i < 9; ++i) d = kfma(V[k * 9 + i], A[i * 5 + j], d);
          const double f = d * BETA[k];
#pragma unroll
          for (int i = k; i < 9; ++i) A[i * 5 + j] = kfma(-f, V[k * 9 + i], A[i * 5 + j]);
        }
      }
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      double e[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) e[i] = (i == 5 + b) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 4; k >= 0; --k) {
        const double n2 = N2[k];
        if (!(n2 > 0.0)) continue;
        double d = 0.0;
#pragma unroll
        for (int i = k; i < 9; ++i) d = kfma(V[k * 9 + i], e[i], d);
        const double f = d * BETA[k];
#pragma unroll
        for (int i = k; i < 9; ++i) e[i] = kfma(-f, V[k * 9 + i], e[i]);
      }
#pragma unroll
      for (int i = 0; i < 9; ++i) B[b * 9 + i] = e[i];
    }
  }
  // the basis goes out now (E(z) is rebuilt from it downstream); a singular constraint system,
  // known only after the elimination, overwrites it with NaN
  if (alive) {
#pragma unroll
    for (int i = 0; i < 36; ++i) out[34 + i] = B[i];
  }
  // ------------------------------------- phase 2: constraint matrix 10x20
  if constexpr (ALG == 1) {  // the same ten constraints, columns in the degree order of the Stewenius solver
#define SA(i) S(stew_slot(i))
#include "fivept_build.inc"
#undef SA
  } else {
#define SA(i) S(i)
#include "fivept_build.inc"
#undef SA
  }
  KML_PHASE();
  // ----------------------------------------------- phase 3: Gauss-Jordan
  constexpr bool FULL = ALG == 1;
  gj_step<STRIDE, 0, FULL>(sm, failed); gj_step<STRIDE, 1, FULL>(sm, failed); gj_step<STRIDE, 2, FULL>(sm, failed);
  gj_step<STRIDE, 3, FULL>(sm, failed); gj_step<STRIDE, 4, FULL>(sm, failed); gj_step<STRIDE, 5, FULL>(sm, failed);
  gj_step<STRIDE, 6, FULL>(sm, failed); gj_step<STRIDE, 7, FULL>(sm, failed); gj_step<STRIDE, 8, FULL>(sm, failed);
  gj_step<STRIDE, 9, FULL>(sm, failed);
  KML_PHASE();
  if (alive && failed) {
#pragma unroll 1
    for (int i = 0; i < 36; ++i) out[34 + i] = nan("");
  }
  if constexpr (ALG == 1) {
    // ------------------- phase 4 (STEWENIUS): action matrix of x, characteristic polynomial
    // rows 0..5 = -B[0..5] of the reduced system [I | B]; rows 6..9 = unit rows (x*x = x2, x*y = xy,
    // x*z = xz, x*1 = x) — written over the dead rows 6..9 of B
#pragma unroll 1
    for (int i = 0; i < 6; ++i)
#pragma unroll 1
      for (int j = 0; j < 10; ++j) {
        const double v = -S(i * 20 + 10 + j);
        S(i * 20 + 10 + j) = v;
        if (alive) out[kStewRowsOff + i * 10 + j] = v;
      }
#pragma unroll 1
    for (int i = 6; i < 10; ++i)
#pragma unroll 1
      for (int j = 0; j < 10; ++j) S(i * 20 + 10 + j) = 0.0;
    S(6 * 20 + 10 + 0) = 1.0; S(7 * 20 + 10 + 1) = 1.0; S(8 * 20 + 10 + 2) = 1.0; S(9 * 20 + 10 + 6) = 1.0;
    double cp[11];
    charpoly10_s<STRIDE>(sm, cp);
    if (alive) {
#pragma unroll 1
      for (int k = 0; k < 11; ++k) out[k] = cp[k];
    }
  } else {
    // --------------------------------- phase 4: B(z), cofactors, n(z)
    // B(z) entries as polynomials in z (ascending), rows 4..9 of the reduced system:
    // c0[r], c1[r] of degree 3, c2[r] of degree 4 (rows 0..3 of A are dead)
    double c0[3][4], c1[3][4], c2[3][5];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int e = (4 + 2 * r) * 20, f = (5 + 2 * r) * 20;
      c0[r][0] = S(e + 12);
      c0[r][1] = S(e + 11) - S(f + 12);
      c0[r][2] = S(e + 10) - S(f + 11);
      c0[r][3] = -S(f + 10);
      c1[r][0] = S(e + 15);
      c1[r][1] = S(e + 14) - S(f + 15);
      c1[r][2] = S(e + 13) - S(f + 14);
      c1[r][3] = -S(f + 13);
      c2[r][0] = S(e + 19);
      c2[r][1] = S(e + 18) - S(f + 19);
      c2[r][2] = S(e + 17) - S(f + 18);
      c2[r][3] = S(e + 16) - S(f + 17);
      c2[r][4] = -S(f + 16);
    }
    // cofactors of the third row and n(z) = det B(z); the sums keep the ascending term order
    double p1[8], p2[8], p3[7];
#pragma unroll
    for (int k = 0; k < 8; ++k) p1[k] = conv_r<3, 4>(c1[0], c2[1], k) - conv_r<4, 3>(c2[0], c1[1], k);
#pragma unroll
    for (int k = 0; k < 8; ++k) p2[k] = conv_r<4, 3>(c2[0], c0[1], k) - conv_r<3, 4>(c0[0], c2[1], k);
#pragma unroll
    for (int k = 0; k < 7; ++k) p3[k] = conv_r<3, 3>(c0[0], c1[1], k) - conv_r<3, 3>(c1[0], c0[1], k);
    if (alive) {
#pragma unroll
      for (int k = 0; k < 11; ++k)
        out[k] = (conv_r<7, 3>(p1, c0[2], k) + conv_r<7, 3>(p2, c1[2], k)) + conv_r<6, 4>(p3, c2[2], k);
#pragma unroll
      for (int k = 0; k < 8; ++k) out[11 + k] = p1[k];
#pragma unroll
      for (int k = 0; k < 8; ++k) out[19 + k] = p2[k];
#pragma unroll
      for (int k = 0; k < 7; ++k) out[27 + k] = p3[k];
    }
  }
#undef KML_PHASE
#undef S
}

// Fixed-degree Horner on register coefficients.  Callers pad with the polynomial's own
// zero leading coefficients: the recurrence then reaches the true leading coefficient
// exactly (0*x + c = c for finite x), so the value equals the variable-degree recurrence
// up to the sign of a zero, which none of the uses below can see.
template <int DEG>
__device__ __forceinline__ double horner_r(const double* c, double x) {
  double r = c[DEG];
#pragma unroll
  for (int i = DEG - 1; i >= 0; --i) r = kfma(r, x, c[i]);
  return r;
}

// ============================================================== stage 2
__device__ __forceinline__ int tri_off(int k) { return 22 + 11 * k - (k * (k - 1)) / 2; }

// Sturm chain of the polynomial in slots 11..21 (or its reversal) into the
// triangular chain area; degrees packed 4 bits each; returns the length.
template <int STRIDE>
__device__ __noinline__ int sturm_build_s(double* sm, bool reversed, unsigned long long* degs_out) {
#define SS(i) sm[(i) * STRIDE]
  int n = 10;
  while (n > 0 && SS(11 + (reversed ? 10 - n : n)) == 0.0) --n;
  for (int i = 0; i <= n; ++i) SS(tri_off(0) + i) = SS(11 + (reversed ? 10 - i : i));
  unsigned long long degs = (unsigned long long)n;
  int len = 1;
  if (n >= 1) {
    for (int i = 0; i < n; ++i) SS(tri_off(1) + i) = (double)(i + 1) * SS(11 + (reversed ? 10 - (i + 1) : (i + 1)));
    degs |= (unsigned long long)(n - 1) << 4;
    len = 2;
    while (len < 12) {
      const int db = (int)((degs >> (4 * (len - 1))) & 15u);
      if (db <= 0) break;
      const int da = (int)((degs >> (4 * (len - 2))) & 15u);
      const int oa = tri_off(len - 2), ob = tri_off(len - 1), oc = tri_off(len);
      for (int i = 0; i <= da; ++i) SS(i) = SS(oa + i);  // remainder r[0..da] in scratch 0..10
      const bool unit_lc = len - 1 >= 2;  // a remainder's leading coefficient is exactly +-1
      for (int d = da; d >= db; --d) {
        const double f = unit_lc ? SS(d) * SS(ob + db) : kdiv(SS(d), SS(ob + db));
        for (int i = 0; i < db; ++i) SS(d - db + i) = kfma(-f, SS(ob + i), SS(d - db + i));
        SS(d) = 0.0;
      }
      int dr = db - 1;
      while (dr >= 0 && SS(dr) == 0.0) --dr;
      if (dr < 0) break;  // exact gcd reached
      const double inv = kdiv(1.0, fabs(SS(dr)));
      for (int i = 0; i < dr; ++i) SS(oc + i) = -(SS(i) * inv);
      SS(oc + dr) = SS(dr) > 0.0 ? -1.0 : 1.0;
      degs |= (unsigned long long)dr << (4 * len);
      ++len;
    }
  }
  *degs_out = degs;
  return len;
#undef SS
}

template <int STRIDE>
__device__ __noinline__ int sturm_count_s(const double* sm, unsigned long long degs, int len, double x) {
  int changes = 0, last = 0;
  for (int k = 0; k < len; ++k) {
    const double v = horner_s<STRIDE>(sm + tri_off(k) * STRIDE, (int)((degs >> (4 * k)) & 15u), x);
    const int s = (v > 0.0) - (v < 0.0);
    if (s != 0) {
      if (last != 0 && s != last) ++changes;
      last = s;
    }
  }
  return changes;
}

// Isolating brackets (lo, hi] of the real roots in (-1,1] of the chain's first
// polynomial, ascending; returns their number R (<= 10).
// mode 0: grid pass only — if the grid does not isolate the roots, *deferred is set and no
// bracket is written (the rare, expensive Sturm bisection runs in a separate, compacted launch
// so that it does not stall the other 31 lanes of every warp); mode 1: bisection on the count.
template <int STRIDE>
__device__ __noinline__ int isolate_unit_s(double* sm, unsigned long long degs, int len, double* __restrict__ brk,
                                           int mode, bool* deferred, int only_j = -1) {
  const int d0 = (int)(degs & 15u);
  if (d0 < 1) return 0;
  const double* c0 = sm + tri_off(0) * STRIDE;
  const int vm1 = sturm_count_s<STRIDE>(sm, degs, len, -1.0), vp1 = sturm_count_s<STRIDE>(sm, degs, len, 1.0);
  int R = vm1 - vp1;
  if (R > 10) R = 10;
  if (R <= 0) return 0;
  // grid pass: 32 sign-test cells (x_{i-1}, x_i], x_i = -1 + i/16; if the number of bracketing
  // cells equals the Sturm count they are the isolating brackets, else bisect on the count
  unsigned cells = 0u;
  int nb = 0;
  // (mode 1 is only entered for chains the grid pass could not separate: no need to repeat it)
  if (mode == 0) {
    // The polynomial goes to registers, zero-padded to degree 10: a Horner recurrence that
    // starts from padding zeros reaches c[d0] exactly (0*x + c = c for finite x), so the
    // values equal those of the degree-d0 recurrence up to the sign of a zero, which the
    // comparisons below do not see.  Sign bits of the 33 grid values, then the cells.
    double cr[11];
#pragma unroll
    for (int i = 0; i <= 10; ++i) cr[i] = (i <= d0) ? c0[i * STRIDE] : 0.0;
    unsigned long long neg = 0ull, pos = 0ull, zer = 0ull;
#pragma unroll 3
    for (int i = 0; i <= kTRootGrid; ++i) {
      const double x = -1.0 + (double)i * (2.0 / kTRootGrid);
      double r = cr[10];
#pragma unroll
      for (int k = 9; k >= 0; --k) r = kfma(r, x, cr[k]);
      neg |= (unsigned long long)(r < 0.0) << i;
      pos |= (unsigned long long)(r > 0.0) << i;
      zer |= (unsigned long long)(r == 0.0) << i;
    }
    // cell i-1 = (x_{i-1}, x_i] brackets a root: sign change across it, or f(x_i) == 0
    cells = (unsigned)((((neg << 1) & pos) | ((pos << 1) & neg) | zer) >> 1);
    nb = __popc(cells);
  }
  const bool grid_ok = (mode == 0) && (nb == R);
  if (mode == 0 && !grid_ok) {
    *deferred = true;
    return R;
  }
  int cell = -1;
  // mode 1: the bisections of the R roots are independent, the caller asks for one of them
  for (int j = (mode == 1 ? only_j : 0); j < (mode == 1 ? min(only_j + 1, R) : R); ++j) {
    double lo = -1.0, hi = 1.0;
    if (grid_ok) {
      do { ++cell; } while (!((cells >> cell) & 1u));
      lo = -1.0 + (double)cell * (2.0 / kTRootGrid);
      hi = -1.0 + (double)(cell + 1) * (2.0 / kTRootGrid);
    } else {
      int vlo = vm1, vhi = vp1, jj = j;
      for (int depth = 0; depth < kTRootDepth; ++depth) {
        if (vlo - vhi == 1) break;
        const double mid = 0.5 * (lo + hi);
        const int vm = sturm_count_s<STRIDE>(sm, degs, len, mid);
        const int left = vlo - vm;
        if (jj < left) { hi = mid; vhi = vm; } else { jj -= left; lo = mid; vlo = vm; }
      }
    }
    brk[2 * j] = lo;
    brk[2 * j + 1] = hi;
  }
  return R;
}

// ---- generic-position fast path of stage 2, all in registers ----------------------------
// For a polynomial in generic position every Sturm remainder loses exactly one degree
// (10, 9, ..., 0).  Then the chain can be built with compile-time degrees, only the last two
// members are alive at any time, and the signs at -1 and +1 are taken as each member appears:
// the same divisions, products and sums in the same order as sturm_build_s / sturm_count_s.
// Any exact zero where a leading coefficient is expected hands the draw to the generic code.
struct SturmSigns {
  int last_m = 0, last_p = 0, ch_m = 0, ch_p = 0;
  __device__ __forceinline__ void add(double vm, double vp) {
    const int sm_ = (vm > 0.0) - (vm < 0.0), sp_ = (vp > 0.0) - (vp < 0.0);
    if (sm_ != 0) { if (last_m != 0 && sm_ != last_m) ++ch_m; last_m = sm_; }
    if (sp_ != 0) { if (last_p != 0 && sp_ != last_p) ++ch_p; last_p = sp_; }
  }
};
// remainder of pa (degree DB+1) by pb (degree DB), negated and scaled by 1/|leading|: degree DB-1.
// UNIT: pb is itself a remainder, whose leading coefficient is exactly +-1 (dividing = multiplying).
template <int DB, bool UNIT>
__device__ __forceinline__ bool sturm_step(const double* pa, const double* pb, double* pc) {
  double r[DB + 2];
#pragma unroll
  for (int i = 0; i <= DB + 1; ++i) r[i] = pa[i];
  {
    const double f = UNIT ? r[DB + 1] * pb[DB] : kdiv(r[DB + 1], pb[DB]);
#pragma unroll
    for (int i = 0; i < DB; ++i) r[1 + i] = kfma(-f, pb[i], r[1 + i]);
  }
  {
    const double f = UNIT ? r[DB] * pb[DB] : kdiv(r[DB], pb[DB]);
#pragma unroll
    for (int i = 0; i < DB; ++i) r[i] = kfma(-f, pb[i], r[i]);
  }
  if (r[DB - 1] == 0.0) return false;  // the degree drops by more than one
  const double inv = kdiv(1.0, fabs(r[DB - 1]));
#pragma unroll
  for (int i = 0; i < DB - 1; ++i) pc[i] = -(r[i] * inv);
  pc[DB - 1] = r[DB - 1] > 0.0 ? -1.0 : 1.0;
  return true;
}
template <int DB>
__device__ __forceinline__ bool sturm_tail(const double* pa, const double* pb, SturmSigns& sg) {
  double pc[DB];
  if (!sturm_step<DB, (DB < 9)>(pa, pb, pc)) return false;  // DB = 9: the divisor is the derivative, not a remainder
  sg.add(horner_r<DB - 1>(pc, -1.0), horner_r<DB - 1>(pc, 1.0));
  if constexpr (DB - 1 >= 1) return sturm_tail<DB - 1>(pb, pc, sg);
  return true;
}
constexpr int kTRootGridFast = 32;  // = kTRootGrid (defined above)
// a[0..10]: the chain's first polynomial.  Returns the root count R in (-1, 1]; brackets are
// written when the sign grid separates the roots, else *deferred is set (the bisection runs in
// mono_isolate_deferred_kernel); *degenerate asks for the generic path.
__device__ __forceinline__ int isolate_chain_fast(const double* a, double* __restrict__ brk, bool* deferred,
                                                  bool* degenerate) {
  if (a[10] == 0.0) { *degenerate = true; return 0; }
  double p1[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) p1[i] = (double)(i + 1) * a[i + 1];
  SturmSigns sg;
  sg.add(horner_r<10>(a, -1.0), horner_r<10>(a, 1.0));
  sg.add(horner_r<9>(p1, -1.0), horner_r<9>(p1, 1.0));
  if (!sturm_tail<9>(a, p1, sg)) { *degenerate = true; return 0; }
  int R = sg.ch_m - sg.ch_p;
  if (R > 10) R = 10;
  if (R <= 0) return 0;
  unsigned long long neg = 0ull, pos = 0ull, zer = 0ull;
#pragma unroll 3
  for (int i = 0; i <= kTRootGridFast; ++i) {
    const double x = -1.0 + (double)i * (2.0 / kTRootGridFast);
    const double r = horner_r<10>(a, x);
    neg |= (unsigned long long)(r < 0.0) << i;
    pos |= (unsigned long long)(r > 0.0) << i;
    zer |= (unsigned long long)(r == 0.0) << i;
  }
  unsigned cells = (unsigned)((((neg << 1) & pos) | ((pos << 1) & neg) | zer) >> 1);
  if (__popc(cells) != R) { *deferred = true; return R; }
  for (int j = 0; j < R; ++j) {
    const int cell = __ffs(cells) - 1;
    cells &= cells - 1u;
    brk[2 * j] = -1.0 + (double)cell * (2.0 / kTRootGridFast);
    brk[2 * j + 1] = -1.0 + (double)(cell + 1) * (2.0 / kTRootGridFast);
  }
  return R;
}

// fo: this draw's stage-1 output.  Writes up to 10 + 10 brackets to brk and
// returns R0 | R1 << 8 | deferred0 << 16 | deferred1 << 17 (0 if the draw has
// no usable polynomial).
// (the rare generic path keeps its 88 scratch slots in local memory: no shared memory, so the
// kernel's occupancy is set by registers alone)
__device__ __noinline__ int mono_isolate_generic(const double* __restrict__ fo, double* __restrict__ brk);
__device__ __forceinline__ int mono_isolate_thread(const double* __restrict__ fo, double* __restrict__ brk,
                                                   bool force_generic) {
  if (!(fo[34] == fo[34])) return 0;  // singular constraint system (NaN basis)
  bool def0 = false, def1 = false;
  if (!force_generic) {
    double a[11];
    bool degenerate = false;
#pragma unroll
    for (int k = 0; k < 11; ++k) a[k] = fo[k];
    const int R0f = isolate_chain_fast(a, brk, &def0, &degenerate);
    if (!degenerate) {
#pragma unroll
      for (int k = 0; k < 11; ++k) a[k] = fo[10 - k];
      const int R1f = isolate_chain_fast(a, brk + 2 * R0f, &def1, &degenerate);
      if (!degenerate) return R0f | (R1f << 8) | ((int)def0 << 16) | ((int)def1 << 17);
    }
  }
  // a leading coefficient or a remainder's leading term is exactly zero
  return mono_isolate_generic(fo, brk);
}
__device__ __noinline__ int mono_isolate_generic(const double* __restrict__ fo, double* __restrict__ brk) {
  double sm[kIsoSlots];
#define S(i) sm[(i)]
#pragma unroll 1
  for (int k = 0; k < 11; ++k) S(11 + k) = fo[k];
  unsigned long long degs;
  bool def0 = false, def1 = false;
  int len = sturm_build_s<1>(sm, false, &degs);
  const int R0 = isolate_unit_s<1>(sm, degs, len, brk, 0, &def0);
  len = sturm_build_s<1>(sm, true, &degs);
  const int R1 = isolate_unit_s<1>(sm, degs, len, brk + 2 * R0, 0, &def1);
  return R0 | (R1 << 8) | ((int)def0 << 16) | ((int)def1 << 17);
#undef S
}
// Second grid of the deferred chains (oracle: kRootGrid2): 256 cells on (-1, 1], the rule of the first
// grid (sign change across the cell, or p(x_i) == 0).  If the number of bracketing cells equals the
// chain's root count R they are the isolating brackets and root j's is returned; four independent
// Horner recurrences are in flight at a time.
constexpr int kTRootGrid2 = 256;
__device__ __noinline__ bool isolate_grid2(const double* cr, int R, int j, double* lo, double* hi) {
  int nb = 0, cell_j = -1;
  double fprev = horner_r<10>(cr, -1.0);
#pragma unroll 1
  for (int i0 = 1; i0 <= kTRootGrid2; i0 += 4) {
    double f[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) f[u] = horner_r<10>(cr, -1.0 + (double)(i0 + u) * (2.0 / kTRootGrid2));
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const double fi = f[u];
      const bool hit = (fprev < 0.0 && fi > 0.0) || (fprev > 0.0 && fi < 0.0) || fi == 0.0;
      if (hit) {
        if (nb == j) cell_j = i0 + u - 1;
        ++nb;
      }
      fprev = fi;
    }
  }
  if (nb != R || cell_j < 0) return false;
  *lo = -1.0 + (double)cell_j * (2.0 / kTRootGrid2);
  *hi = -1.0 + (double)(cell_j + 1) * (2.0 / kTRootGrid2);
  return true;
}
// ---- warp-cooperative Sturm fallback ------------------------------------------------------------
// For the chains the second grid cannot separate either (roots closer than 1/128, numerically
// multiple roots) the serial thread below spends ~70 us on one chain: ~120 IEEE divisions of the
// remainder sequence and, per bisection step, eleven Horner recurrences one after the other.  The
// warp does the same operations side by side: lane i owns coefficient i of a remainder update,
// lane k evaluates chain member k, the sign changes are counted from two ballots.  Every value is
// produced by the same operations in the same order as in sturm_build_s / sturm_count_s /
// isolate_unit_s(mode 1).  w: 80 doubles of shared memory per warp (the triangular chain, offsets
// wtri(k), and the remainder in w[66..76]).
constexpr int kWarpSturmSlots = 80;
__device__ __forceinline__ int wtri(int k) { return 11 * k - (k * (k - 1)) / 2; }
__device__ __noinline__ int sturm_build_w(double* w, const double* __restrict__ fo, bool reversed, int lane,
                                          unsigned long long* degs_out) {
  int n = 10;
  while (n > 0 && fo[reversed ? 10 - n : n] == 0.0) --n;
  if (lane <= n) w[lane] = fo[reversed ? 10 - lane : lane];
  unsigned long long degs = (unsigned long long)n;
  int len = 1;
  if (n >= 1) {
    if (lane < n) w[wtri(1) + lane] = (double)(lane + 1) * fo[reversed ? 10 - (lane + 1) : (lane + 1)];
    degs |= (unsigned long long)(n - 1) << 4;
    len = 2;
    __syncwarp();
    double* r = w + 66;
    while (len < 12) {
      const int db = (int)((degs >> (4 * (len - 1))) & 15u);
      if (db <= 0) break;
      const int da = (int)((degs >> (4 * (len - 2))) & 15u);
      const int oa = wtri(len - 2), ob = wtri(len - 1), oc = wtri(len);
      if (lane <= da) r[lane] = w[oa + lane];
      __syncwarp();
      const bool unit_lc = len - 1 >= 2;  // a remainder's leading coefficient is exactly +-1
      for (int d = da; d >= db; --d) {
        const double f = unit_lc ? r[d] * w[ob + db] : kdiv(r[d], w[ob + db]);
        __syncwarp();
        if (lane < db) r[d - db + lane] = kfma(-f, w[ob + lane], r[d - db + lane]);
        if (lane == 31) r[d] = 0.0;
        __syncwarp();
      }
      int dr = db - 1;
      while (dr >= 0 && r[dr] == 0.0) --dr;
      if (dr < 0) break;  // exact gcd reached
      const double inv = kdiv(1.0, fabs(r[dr]));
      if (lane < dr) w[oc + lane] = -(r[lane] * inv);
      if (lane == dr) w[oc + lane] = r[dr] > 0.0 ? -1.0 : 1.0;
      degs |= (unsigned long long)dr << (4 * len);
      ++len;
      __syncwarp();
    }
  }
  __syncwarp();
  *degs_out = degs;
  return len;
}
// sign changes of a chain of `len` members from the sign ballots of its values (sturm_count_s)
__device__ __forceinline__ int sturm_changes(unsigned pos, unsigned neg, int len) {
  int changes = 0, last = 0;
  for (int k = 0; k < len; ++k) {
    const int s = (int)((pos >> k) & 1u) - (int)((neg >> k) & 1u);
    if (s != 0) {
      if (last != 0 && s != last) ++changes;
      last = s;
    }
  }
  return changes;
}
// value of chain member k at x
__device__ __forceinline__ double sturm_member_w(const double* w, unsigned long long degs, int k, double x) {
  const int dg = (int)((degs >> (4 * k)) & 15u);
  const double* c = w + wtri(k);
  double v = c[dg];
  for (int i = dg - 1; i >= 0; --i) v = kfma(v, x, c[i]);
  return v;
}
// Brackets of all real roots in (-1, 1] of the chain in w by bisection on the count
// (isolate_unit_s, mode 1, for every root).  Two roots at a time: each half-warp evaluates the (at
// most 11) members at its own midpoint, one pair of ballots serves both.
__device__ __noinline__ void isolate_roots_w(const double* w, unsigned long long degs, int len,
                                             double* __restrict__ brk, int lane) {
  if ((int)(degs & 15u) < 1) return;
  const int half = lane >> 4, hl = lane & 15;
  int vm1, vp1;
  {  // counts at -1 (lower half-warp) and +1 (upper half-warp)
    const double v = hl < len ? sturm_member_w(w, degs, hl, half ? 1.0 : -1.0) : 0.0;
    const unsigned pos = __ballot_sync(0xFFFFFFFFu, v > 0.0), neg = __ballot_sync(0xFFFFFFFFu, v < 0.0);
    vm1 = sturm_changes(pos & 0xFFFFu, neg & 0xFFFFu, len);
    vp1 = sturm_changes(pos >> 16, neg >> 16, len);
  }
  int R = vm1 - vp1;
  if (R > 10) R = 10;
  for (int j0 = 0; j0 < R; j0 += 2) {
    const int j = j0 + half;
    const bool active = j < R;
    double lo = -1.0, hi = 1.0;
    int vlo = vm1, vhi = vp1, jj = j;
    for (int depth = 0; depth < kTRootDepth; ++depth) {
      const bool done = !active || (vlo - vhi == 1);
      if (__all_sync(0xFFFFFFFFu, done)) break;
      const double mid = 0.5 * (lo + hi);
      const double v = hl < len ? sturm_member_w(w, degs, hl, mid) : 0.0;
      const unsigned pos = __ballot_sync(0xFFFFFFFFu, v > 0.0), neg = __ballot_sync(0xFFFFFFFFu, v < 0.0);
      const int vm = sturm_changes((pos >> (16 * half)) & 0xFFFFu, (neg >> (16 * half)) & 0xFFFFu, len);
      if (!done) {
        const int left = vlo - vm;
        if (jj < left) { hi = mid; vhi = vm; } else { jj -= left; lo = mid; vlo = vm; }
      }
    }
    if (active && hl == 0) {
      brk[2 * j] = lo;
      brk[2 * j + 1] = hi;
    }
  }
}
// The second grid (isolate_grid2) of one chain by the whole warp: lane l evaluates x_i for i = l, l + 32, ...;
// true = the bracketing cells are the R isolating brackets (written), false = the Sturm fallback has to do it.
__device__ __noinline__ bool isolate_grid2_w(const double* __restrict__ fo, bool reversed, int R, double* __restrict__ brk,
                                             int lane) {
  double cr[11];
#pragma unroll
  for (int k = 0; k < 11; ++k) cr[k] = reversed ? fo[10 - k] : fo[k];
  unsigned neg[9], pos[9], zer[9];
#pragma unroll
  for (int b = 0; b < 9; ++b) {
    const int i = 32 * b + lane;
    const double f = horner_r<10>(cr, -1.0 + (double)i * (2.0 / kTRootGrid2));
    const bool in = i <= kTRootGrid2;
    neg[b] = __ballot_sync(0xFFFFFFFFu, in && f < 0.0);
    pos[b] = __ballot_sync(0xFFFFFFFFu, in && f > 0.0);
    zer[b] = __ballot_sync(0xFFFFFFFFu, in && f == 0.0);
  }
  // cell c = (x_c, x_{c+1}]: sign change across it, or p(x_{c+1}) == 0
  unsigned hit[8];
  int nb = 0;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const unsigned negs = (neg[b] >> 1) | (neg[b + 1] << 31), poss = (pos[b] >> 1) | (pos[b + 1] << 31);
    const unsigned zers = (zer[b] >> 1) | (zer[b + 1] << 31);
    hit[b] = (neg[b] & poss) | (pos[b] & negs) | zers;
    nb += __popc(hit[b]);
  }
  if (nb != R) return false;
  if (lane < R) {  // lane j takes the j-th bracketing cell
    int skip = lane, cell = -1;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int c = __popc(hit[b]);
      if (cell < 0) {
        if (skip < c) {
          unsigned m = hit[b];
          for (int t = 0; t < skip; ++t) m &= m - 1u;
          cell = 32 * b + __ffs(m) - 1;
        } else {
          skip -= c;
        }
      }
    }
    brk[2 * lane] = -1.0 + (double)cell * (2.0 / kTRootGrid2);
    brk[2 * lane + 1] = -1.0 + (double)(cell + 1) * (2.0 / kTRootGrid2);
  }
  return true;
}

// the deferred case of one (draw, chain, root): the finer grid first, else bisection on the Sturm count
// (serial form, one thread per root: tests/device_math_host.cpp; the kernel runs the warp forms above)
template <int STRIDE>
__device__ void mono_isolate_deferred_thread(double* sm, const double* __restrict__ fo, int chain, int root, int R,
                                             double* __restrict__ brk) {
  {
    double cr[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) cr[k] = chain ? fo[10 - k] : fo[k];
    double lo, hi;
    if (isolate_grid2(cr, R, root, &lo, &hi)) {
      brk[2 * root] = lo;
      brk[2 * root + 1] = hi;
      return;
    }
  }
#define S(i) sm[(i) * STRIDE]
#pragma unroll 1
  for (int k = 0; k < 11; ++k) S(11 + k) = fo[k];
  unsigned long long degs;
  bool dummy = false;
  const int len = sturm_build_s<STRIDE>(sm, chain != 0, &degs);
  isolate_unit_s<STRIDE>(sm, degs, len, brk, 1, &dummy, root);
#undef S
}

// ============================================================== stage 3

// Refines the root of bracket (lo, hi] of n(z) (chain 0) or of the reversed
// polynomial (chain 1) and maps it to z.  fo = stage-1 output of the draw.
__device__ __noinline__ bool refine_root(const double* __restrict__ fo, int chain, double lo, double hi,
                                         double* z_out) {
  double c0[11], c1[10];
#pragma unroll
  for (int i = 0; i <= 10; ++i) c0[i] = chain ? fo[10 - i] : fo[i];
#pragma unroll
  for (int i = 0; i < 10; ++i) c1[i] = (double)(i + 1) * c0[i + 1];
  double flo = horner_r<10>(c0, lo);
  const double fhi = horner_r<10>(c0, hi);
  double root;
  if (fhi == 0.0) {
    root = hi;
  } else {
    if (!((flo < 0.0 && fhi > 0.0) || (flo > 0.0 && fhi < 0.0))) return false;
#pragma unroll 1
    for (int it = 0; it < kTRootBisect; ++it) {
      const double mid = 0.5 * (lo + hi);
      const double fm = horner_r<10>(c0, mid);
      if ((fm < 0.0) == (flo < 0.0)) { lo = mid; flo = fm; } else { hi = mid; }
    }
    double x = 0.5 * (lo + hi);
#pragma unroll 1
    for (int it = 0; it < kTRootNewton; ++it) {
      const double fx = horner_r<10>(c0, x);
      const double dfx = horner_r<9>(c1, x);
      if ((fx < 0.0) == (flo < 0.0)) { lo = x; flo = fx; } else { hi = x; }
      double xn = x - kdiv(fx, dfx);
      if (!(xn >= lo && xn <= hi)) xn = 0.5 * (lo + hi);
      x = xn;
    }
    root = x;
  }
  if (chain) {
    if (root == 1.0 || root == 0.0) return false;  // z = 1 belongs to chain 0; u = 0 is z = inf
    root = kdiv(1.0, root);
  }
  *z_out = root;
  return true;
}

// E(z) = x X + y Y + z Z + W with x = p1(z)/p3(z), y = p2(z)/p3(z); false if not finite.
__device__ __noinline__ bool essential_from_root(const double* __restrict__ fo, double z, double* E) {
  double c[8];
#pragma unroll
  for (int i = 0; i < 7; ++i) c[i] = fo[27 + i];
  const double rd = kdiv(1.0, horner_r<6>(c, z));
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i] = fo[11 + i];
  const double x = horner_r<7>(c, z) * rd;
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i] = fo[19 + i];
  const double y = horner_r<7>(c, z) * rd;
  bool ok = true;
#pragma unroll 1
  for (int e = 0; e < 9; ++e) {
    const double v = kfma(z, fo[34 + 18 + e], kfma(y, fo[34 + 9 + e], kfma(x, fo[34 + e], fo[34 + 27 + e])));
    if (!isfinite(v)) ok = false;
    E[e] = v;
  }
  return ok;
}

// STEWENIUS (row f4): E from the eigenvalue x of the action matrix whose six non-trivial rows are
// fo[kStewRowsOff + 10 j + k]: the eigenvector v = [x2, xy, xz, v3, v4, v5, x, y, z, 1] solves
// rows 0..5 of (M - x I) v = 0, six equations in (v3, v4, v5, y, z), Gaussian elimination with
// first-maximum row pivoting (contract: DESIGN.md §4.12); then E = ((x X + y Y) + z Z) + W.
__device__ __noinline__ bool essential_from_root_stew(const double* __restrict__ fo, double x, double* E) {
  const double x2 = x * x;
  double C[6][5], d[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const double* a = fo + kStewRowsOff + 10 * j;
    C[j][0] = a[3]; C[j][1] = a[4]; C[j][2] = a[5];
    C[j][3] = kfma(a[1], x, a[7]);
    C[j][4] = kfma(a[2], x, a[8]);
    d[j] = -kfma(a[0], x2, kfma(a[6], x, a[9]));
  }
  d[0] = kfma(x, x2, d[0]);
  C[1][3] = C[1][3] - x2;
  C[2][4] = C[2][4] - x2;
  C[3][0] = C[3][0] - x;
  C[4][1] = C[4][1] - x;
  C[5][2] = C[5][2] - x;
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    int pr = c;
    double pv = fabs(C[c][c]);
#pragma unroll
    for (int r = c + 1; r < 6; ++r) {
      const double v = fabs(C[r][c]);
      if (v > pv) { pv = v; pr = r; }
    }
    if (!(pv > 0.0)) return false;
    // swap rows c and pr without dynamic register indexing
#pragma unroll
    for (int r = c + 1; r < 6; ++r)
      if (r == pr) {
#pragma unroll
        for (int j = 0; j < 5; ++j) { const double t = C[c][j]; C[c][j] = C[r][j]; C[r][j] = t; }
        const double t = d[c]; d[c] = d[r]; d[r] = t;
      }
#pragma unroll
    for (int r = c + 1; r < 6; ++r) {
      const double f = kdiv(C[r][c], C[c][c]);
#pragma unroll
      for (int j = c + 1; j < 5; ++j) C[r][j] = kfma(-f, C[c][j], C[r][j]);
      d[r] = kfma(-f, d[c], d[r]);
    }
  }
  double u[5];
#pragma unroll
  for (int c = 4; c >= 0; --c) {
    double sacc = d[c];
#pragma unroll
    for (int j = c + 1; j < 5; ++j) sacc = kfma(-C[c][j], u[j], sacc);
    u[c] = kdiv(sacc, C[c][c]);
  }
  const double y = u[3], z = u[4];
  bool ok = true;
#pragma unroll 1
  for (int e = 0; e < 9; ++e) {
    const double v = kfma(z, fo[34 + 18 + e], kfma(y, fo[34 + 9 + e], kfma(x, fo[34 + e], fo[34 + 27 + e])));
    if (!isfinite(v)) ok = false;
    E[e] = v;
  }
  return ok;
}

// Decomposition of one essential matrix: Ra = U W V^T, Rb = U W^T V^T, t = s0 u2.
__device__ __forceinline__ void essential_candidates(const double* E, double* Ra, double* Rb, double* tt) {
  double U[9], Sv[3], V[9];
  svd3_r(E, U, Sv, V);
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double a0 = U[3 * r + 0], a1 = U[3 * r + 1], a2 = U[3 * r + 2];
      const double b0 = V[3 * c + 0], b1 = V[3 * c + 1], b2 = V[3 * c + 2];
      Ra[3 * r + c] = kfma(a2, b2, kfma(a1, b0, -(a0 * b1)));
      Rb[3 * r + c] = kfma(a2, b2, kfma(a0, b1, -(a1 * b0)));
    }
  tt[0] = Sv[0] * U[2];
  tt[1] = Sv[0] * U[5];
  tt[2] = Sv[0] * U[8];
}
// candidate `cand` of (Ra,t),(Ra,-t),(Rb,t),(Rb,-t) as [R12 | t12]
__device__ __forceinline__ void candidate_model(const double* Ra, const double* Rb, const double* tt, int cand,
                                                double* M) {
  const double sgn = (cand & 1) ? -1.0 : 1.0;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    M[4 * r + 0] = (cand < 2) ? Ra[3 * r + 0] : Rb[3 * r + 0];
    M[4 * r + 1] = (cand < 2) ? Ra[3 * r + 1] : Rb[3 * r + 1];
    M[4 * r + 2] = (cand < 2) ? Ra[3 * r + 2] : Rb[3 * r + 2];
    M[4 * r + 3] = sgn * tt[r];
  }
}
// Residuals of the candidate pair (R, t) and (R, -t) for one correspondence.
// Negating t negates b0, b1, l0, l1, p, tinv and q exactly (every IEEE operation is
// sign-symmetric), so with x1 = f1.p/|p| and x2 = f2.q/|q| of the +t candidate the -t
// candidate's residual is (1 + x1) + (1 + x2): two candidates for one evaluation, bit
// for bit what mono_residual returns for each of them.
__device__ __forceinline__ void mono_residual_pair(const double* R /*3x3*/, const V3& t, const double* tinv,
                                                   const V3& f1, const V3& f2, double* r_pos, double* r_neg) {
  V3 f2u;
  f2u.x = kfma(R[2], f2.z, kfma(R[1], f2.y, R[0] * f2.x));
  f2u.y = kfma(R[5], f2.z, kfma(R[4], f2.y, R[3] * f2.x));
  f2u.z = kfma(R[8], f2.z, kfma(R[7], f2.y, R[6] * f2.x));
  const double b0 = dot(t, f1), b1 = dot(t, f2u);
  const double d12 = dot(f1, f2u);
  const double A00 = dot(f1, f1), A01 = -d12, A10 = d12, A11 = -dot(f2u, f2u);
  const double det = kfma(A00, A11, -(A01 * A10));
  const double rdet = 1.0 / det;
  const double l0 = kfma(A11, b0, -(A01 * b1)) * rdet;
  const double l1 = kfma(A00, b1, -(A10 * b0)) * rdet;
  V3 p, q;
  p.x = 0.5 * kfma(l0, f1.x, kfma(l1, f2u.x, t.x));
  p.y = 0.5 * kfma(l0, f1.y, kfma(l1, f2u.y, t.y));
  p.z = 0.5 * kfma(l0, f1.z, kfma(l1, f2u.z, t.z));
  q.x = kfma(R[6], p.z, kfma(R[3], p.y, kfma(R[0], p.x, tinv[0])));
  q.y = kfma(R[7], p.z, kfma(R[4], p.y, kfma(R[1], p.x, tinv[1])));
  q.z = kfma(R[8], p.z, kfma(R[5], p.y, kfma(R[2], p.x, tinv[2])));
  const double x1 = dot(f1, p) * krsqrt(dot(p, p));
  const double x2 = dot(f2, q) * krsqrt(dot(q, q));
  *r_pos = (1.0 - x1) + (1.0 - x2);
  *r_neg = (1.0 + x1) + (1.0 + x2);
}

// One (draw, root) item: refine the bracketed root, build E, decompose it, score the four
// (R, t) candidates on the 8 sample points (sequential sums) and keep the first
// smallest one below the reference's initial quality 1e6.
// returns 0 root not refined, 1 refined but nothing usable, 2 *q_out / M_out written
template <int ALG = 0>
__device__ __forceinline__ int mono_item(const double* __restrict__ fo, int chain, double lo, double hi,
                                         const double* __restrict__ ga, const double* __restrict__ gb,
                                         const uint16_t* __restrict__ smp, double* q_out, double* M_out) {
  double z, E[9];
  if (!refine_root(fo, chain, lo, hi, &z)) return 0;  // NISTER: root z of n(z); STEWENIUS: eigenvalue x
  if (ALG == 1 ? !essential_from_root_stew(fo, z, E) : !essential_from_root(fo, z, E)) return 1;
  double Ra[9], Rb[9], tt[3];
  essential_candidates(E, Ra, Rb, tt);
  const V3 t = {tt[0], tt[1], tt[2]};
  double ta[3], tb[3];  // -R^T t of the +t candidates (mono_tinv)
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    ta[c] = -kfma(Ra[6 + c], tt[2], kfma(Ra[3 + c], tt[1], Ra[c] * tt[0]));
    tb[c] = -kfma(Rb[6 + c], tt[2], kfma(Rb[3 + c], tt[1], Rb[c] * tt[0]));
  }
  double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
#pragma unroll 1
  for (int k = 0; k < 8; ++k) {
    const int idx = smp[k];
    const V3 f1 = {ga[3 * idx], ga[3 * idx + 1], ga[3 * idx + 2]};
    const V3 f2 = {gb[3 * idx], gb[3 * idx + 1], gb[3 * idx + 2]};
    double rp, rn;
    mono_residual_pair(Ra, t, ta, f1, f2, &rp, &rn);
    q0 = q0 + rp;
    q1 = q1 + rn;
    mono_residual_pair(Rb, t, tb, f1, f2, &rp, &rn);
    q2 = q2 + rp;
    q3 = q3 + rn;
  }
  double best = 1000000.0;
  int bc = -1;
  if (q0 < best) { best = q0; bc = 0; }
  if (q1 < best) { best = q1; bc = 1; }
  if (q2 < best) { best = q2; bc = 2; }
  if (q3 < best) { best = q3; bc = 3; }
  if (bc < 0) return 1;
  *q_out = best;
  candidate_model(Ra, Rb, tt, bc, M_out);
  return 2;
}

}  // namespace geom
}  // namespace kml
