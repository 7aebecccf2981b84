// fivept_warp.cuh — one RANSAC hypothesis per WARP: the mono minimal solver
// (opengv fivept_nister + essential decomposition + 8-point disambiguation,
// SURVEY.md A.6) with its phases spread over the 32 lanes and a per-warp
// shared-memory workspace.  Every output element is produced by exactly the
// operation sequence of the ARITHMETIC CONTRACT — the lanes
// only decide WHO computes an element, never the order of the additions that
// make it — so results are bit-identical to the scalar CPU oracle.
//   phase 1  null space (Householder QR of Q^T)      lanes = columns / basis vectors
//   phase 2  10x20 cubic-constraint matrix           lanes = output coefficients
//   phase 3  Gauss-Jordan, partial pivoting          lanes = columns
//   phase 4  B(z), cofactors, degree-10 n(z)         lanes = coefficients
//   phase 5  two Sturm chains (n(z), reversed n)     half-warp per chain
//   phase 6  root isolation + refinement             lanes = roots
//   phase 7  E assembly, SVD, R/t candidates         lanes = solutions
//   phase 8  8-point scoring of 4 candidates         lanes = (candidate, point)
#pragma once
#include "geom.cuh"

namespace kml {
namespace geom {

struct MonoWs {
  double f1[8][3], f2[8][3];
  double A9[45];   // Householder work matrix [9][5]
  double Hv[45];   // reflectors v[k][i]
  double Hn[5];    // |v_k|^2
  double B[36];    // null-space basis X,Y,Z,W (row-major 3x3 each)
  double EEt[60];  // (0,0) (0,1) (0,2) (1,1) (1,2) (2,2) x 10 coefficients
  double D[30];    // the three 2x2 minors of det(E)
  double A[200];   // constraint matrix [10][20]
  double Bz[45];   // [3][3][5]
  double p1[8], p2[8], p3[7], nz[11], rz[11];
  double ch[2][12][11];
  double zr[10];
  double E[10][9];
  double Ra[10][9], Rb[10][9], tt[10][3];
  int deg[2][12];
  int len[2];
};

// contribution tables (pair = i*4+j, lexicographic order, -1 = none)
__constant__ signed char kT12[20] = {0, -1, 1, 4, 2, 8, 5, -1, 6, 9, 10, -1, 3, 12, 7, 13, 11, 14, 15, -1};
__constant__ signed char kT23[60] = {0,  -1, -1, 13, -1, -1, 1,  4,  -1, 5,  12, -1, 2,  8,  -1,
                                     3,  24, -1, 14, 17, -1, 15, 29, -1, 6,  9,  16, 7,  25, 28,
                                     10, 20, -1, 11, 26, 32, 27, 36, -1, 18, 21, -1, 19, 30, 33,
                                     31, 37, -1, 22, -1, -1, 23, 34, -1, 35, 38, -1, 39, -1, -1};

constexpr unsigned kFull = 0xFFFFFFFFu;

struct MonoTables {
  signed char t12[20];
  signed char t23[60];
};
__device__ __forceinline__ void load_mono_tables(MonoTables* t, int tid, int nthreads) {
  for (int i = tid; i < 20; i += nthreads) t->t12[i] = kT12[i];
  for (int i = tid; i < 60; i += nthreads) t->t23[i] = kT23[i];
}

// coefficient t of (deg1 poly of E entry ea) * (deg1 poly of E entry eb)
__device__ __noinline__ double pm11(const double* B, const MonoTables& T, int ea, int eb, int t) {
  const int c0 = T.t12[2 * t], c1 = T.t12[2 * t + 1];
  double r = 0.0 + B[(c0 >> 2) * 9 + ea] * B[(c0 & 3) * 9 + eb];
  if (c1 >= 0) r = r + B[(c1 >> 2) * 9 + ea] * B[(c1 & 3) * 9 + eb];
  return r;
}
// acc += coefficient s of (deg2 poly a) * (deg1 poly of E entry e)
__device__ __forceinline__ double pm21_acc(double acc, const double* a, const double* B,
                                           const MonoTables& T, int e, int s) {
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    const int c = T.t23[3 * s + u];
    if (c >= 0) acc = acc + a[c >> 2] * B[(c & 3) * 9 + e];
  }
  return acc;
}
// coefficient k of a*b (univariate, ascending coefficients)
__device__ __noinline__ double conv(const double* a, int da, const double* b, int db, int k) {
  double r = 0.0;
  const int i0 = max(0, k - db), i1 = min(da, k);
  for (int i = i0; i <= i1; ++i) r = r + a[i] * b[k - i];
  return r;
}

// Sturm sign-change count at x, evaluated by a half-warp: lane hl evaluates
// chain polynomial hl (same Horner recurrence as the scalar contract), the
// sign sequence is then compared through ballots.  x must be uniform within
// the half-warp; all 32 lanes must call (the two halves may pass different x).
__device__ __noinline__ int sturm_count_half(const MonoWs& ws, int h, int hl, int len, double x) {
  int sg = 0;
  if (hl < len) {
    const double v = horner(ws.ch[h][hl], ws.deg[h][hl], x);
    sg = (v > 0.0) - (v < 0.0);
  }
  const unsigned sh = h * 16;
  const unsigned pos = (__ballot_sync(kFull, sg > 0) >> sh) & 0xFFFFu;
  const unsigned neg = (__ballot_sync(kFull, sg < 0) >> sh) & 0xFFFFu;
  const unsigned nzm = pos | neg;
  bool changed = false;
  if (sg != 0) {
    const unsigned below = nzm & ((1u << hl) - 1u);
    if (below) {
      const int pk = 31 - __clz(below);  // previous non-zero sign in the sequence
      changed = (((pos >> pk) ^ (pos >> hl)) & 1u) != 0u;
    }
  }
  return __popc((__ballot_sync(kFull, changed) >> sh) & 0xFFFFu);
}

constexpr int kWRootDepth = 48;
constexpr int kWRootBisect = 10;
constexpr int kWRootNewton = 8;

// ws.f1 / ws.f2 hold the 8 sampled bearings.  Returns (warp-uniform) whether a
// model exists; M (12 doubles, identical on all lanes) = [R12 | t12].
template <bool SYNC>
__device__ bool mono_model_warp(MonoWs& ws, const MonoTables& T, int lane, double* M) {
  bool failed = false;  // singular constraint system: keep going (phase barriers), discard at the end
  // ------------------------------------------------ phase 1: null space
  for (int t = lane; t < 45; t += 32) {
    const int i = t / 5, j = t % 5;  // A9[i][j] = Q[j][i] = f2[j][i%3] * f1[j][i/3]
    ws.A9[t] = ws.f2[j][i % 3] * ws.f1[j][i / 3];
  }
  __syncwarp();
#pragma unroll 1
  for (int k = 0; k < 5; ++k) {
    if (lane == 0) {
      double s2 = 0.0;
      for (int i = k; i < 9; ++i) s2 = s2 + ws.A9[i * 5 + k] * ws.A9[i * 5 + k];
      const double nrm = ksqrt(s2);
      const double alpha = (ws.A9[k * 5 + k] >= 0.0) ? -nrm : nrm;
      for (int i = 0; i < 9; ++i) ws.Hv[k * 9 + i] = (i < k) ? 0.0 : ws.A9[i * 5 + k];
      ws.Hv[k * 9 + k] = ws.Hv[k * 9 + k] - alpha;
      double n2 = 0.0;
      for (int i = k; i < 9; ++i) n2 = n2 + ws.Hv[k * 9 + i] * ws.Hv[k * 9 + i];
      ws.Hn[k] = n2;
    }
    __syncwarp();
    const double n2 = ws.Hn[k];
    if (n2 > 0.0 && lane >= k && lane < 5) {
      const int j = lane;
      double d = 0.0;
      for (int i = k; i < 9; ++i) d = d + ws.Hv[k * 9 + i] * ws.A9[i * 5 + j];
      const double f = kdiv(2.0 * d, n2);
      for (int i = k; i < 9; ++i) ws.A9[i * 5 + j] = ws.A9[i * 5 + j] - f * ws.Hv[k * 9 + i];
    }
    __syncwarp();
  }
  if (lane < 4) {
    double* x = ws.B + lane * 9;
#pragma unroll 1
    for (int i = 0; i < 9; ++i) x[i] = (i == 5 + lane) ? 1.0 : 0.0;
#pragma unroll 1
    for (int k = 4; k >= 0; --k) {
      const double n2 = ws.Hn[k];
      if (n2 > 0.0) {
        double d = 0.0;
#pragma unroll 1
        for (int i = k; i < 9; ++i) d = d + ws.Hv[k * 9 + i] * x[i];
        const double f = kdiv(2.0 * d, n2);
#pragma unroll 1
        for (int i = k; i < 9; ++i) x[i] = x[i] - f * ws.Hv[k * 9 + i];
      }
    }
  }
  __syncwarp();
  if (SYNC) __syncthreads();  // keep the CTA's warps in the same phase (one instruction stream)
  // ------------------------------------- phase 2: constraint matrix 10x20
  const double* B = ws.B;
  for (int o = lane; o < 60; o += 32) {
    const int idx = o / 10, t = o % 10;
    const int i = (idx < 3) ? 0 : ((idx < 5) ? 1 : 2);
    const int j = (idx < 3) ? idx : ((idx < 5) ? idx - 2 : 2);
    ws.EEt[o] = (pm11(B, T, 3 * i + 0, 3 * j + 0, t) + pm11(B, T, 3 * i + 1, 3 * j + 1, t)) +
                pm11(B, T, 3 * i + 2, 3 * j + 2, t);
  }
  if (lane < 30) {
    const int k = lane / 10, t = lane % 10;
    double v;
    if (k == 0) v = pm11(B, T, 4, 8, t) - pm11(B, T, 5, 7, t);
    else if (k == 1) v = pm11(B, T, 5, 6, t) - pm11(B, T, 3, 8, t);
    else v = pm11(B, T, 3, 7, t) - pm11(B, T, 4, 6, t);
    ws.D[lane] = v;
  }
  __syncwarp();
  if (lane < 10) {
    const double htr = 0.5 * ((ws.EEt[lane] + ws.EEt[30 + lane]) + ws.EEt[50 + lane]);
    ws.EEt[lane] = ws.EEt[lane] - htr;
    ws.EEt[30 + lane] = ws.EEt[30 + lane] - htr;
    ws.EEt[50 + lane] = ws.EEt[50 + lane] - htr;
  }
  __syncwarp();
  for (int o = lane; o < 200; o += 32) {
    const int r = o / 20, s = o % 20;
    double acc = 0.0;
    if (r == 0) {
#pragma unroll 1
      for (int k = 0; k < 3; ++k) acc = pm21_acc(acc, ws.D + 10 * k, B, T, k, s);
    } else {
      const int ii = (r - 1) / 3, jj = (r - 1) % 3;
#pragma unroll 1
      for (int k = 0; k < 3; ++k) {
        // symmetric index of (ii,k) among (0,0)(0,1)(0,2)(1,1)(1,2)(2,2)
        const int lo = min(ii, k), hi = max(ii, k);
        const int sym = (lo == 0) ? hi : ((lo == 1) ? 2 + hi : 5);
        acc = pm21_acc(acc, ws.EEt + 10 * sym, B, T, 3 * k + jj, s);
      }
    }
    ws.A[o] = acc;
  }
  __syncwarp();
  if (SYNC) __syncthreads();  // keep the CTA's warps in the same phase (one instruction stream)
  // ----------------------------------------------- phase 3: Gauss-Jordan
#pragma unroll 1
  for (int c = 0; c < 10; ++c) {
    int pr = c;
    double pv = 0.0;
    if (lane == c) {
      pv = fabs(ws.A[c * 20 + c]);
      for (int r = c + 1; r < 10; ++r) {
        const double v = fabs(ws.A[r * 20 + c]);
        if (v > pv) { pv = v; pr = r; }
      }
    }
    pr = __shfl_sync(kFull, pr, c);
    pv = __shfl_sync(kFull, pv, c);
    if (!(pv > 0.0)) failed = true;
    if (lane < 20 && pr != c) {
      const double t = ws.A[c * 20 + lane];
      ws.A[c * 20 + lane] = ws.A[pr * 20 + lane];
      ws.A[pr * 20 + lane] = t;
    }
    __syncwarp();
    const double piv = ws.A[c * 20 + c];
    double f[10];
#pragma unroll
    for (int r = 0; r < 10; ++r) f[r] = ws.A[r * 20 + c];
    __syncwarp();
    if (lane < 20) {
      const double acj = kdiv(ws.A[c * 20 + lane], piv);
      ws.A[c * 20 + lane] = acj;
#pragma unroll
      for (int r = 0; r < 10; ++r)
        if (r != c) ws.A[r * 20 + lane] = ws.A[r * 20 + lane] - f[r] * acj;
    }
    __syncwarp();
  }
  if (SYNC) __syncthreads();  // keep the CTA's warps in the same phase (one instruction stream)
  // --------------------------------- phase 4: B(z), cofactors, n(z), reversed
  for (int o = lane; o < 45; o += 32) {
    const int r = o / 15, col = (o / 5) % 3, m = o % 5;
    const double* e = ws.A + (4 + 2 * r) * 20;
    const double* f = ws.A + (5 + 2 * r) * 20;
    double v;
    if (col < 2) {
      const int ob = 10 + 3 * col;
      if (m == 0) v = e[ob + 2];
      else if (m == 1) v = e[ob + 1] - f[ob + 2];
      else if (m == 2) v = e[ob + 0] - f[ob + 1];
      else if (m == 3) v = -f[ob + 0];
      else v = 0.0;
    } else {
      if (m == 0) v = e[19];
      else if (m == 1) v = e[18] - f[19];
      else if (m == 2) v = e[17] - f[18];
      else if (m == 3) v = e[16] - f[17];
      else v = -f[16];
    }
    ws.Bz[o] = v;
  }
  __syncwarp();
  {
    const double* b00 = ws.Bz + 0, *b01 = ws.Bz + 5, *b02 = ws.Bz + 10;
    const double* b10 = ws.Bz + 15, *b11 = ws.Bz + 20, *b12 = ws.Bz + 25;
    if (lane < 8) ws.p1[lane] = conv(b01, 3, b12, 4, lane) - conv(b02, 4, b11, 3, lane);
    else if (lane < 16) ws.p2[lane - 8] = conv(b02, 4, b10, 3, lane - 8) - conv(b00, 3, b12, 4, lane - 8);
    else if (lane < 23) ws.p3[lane - 16] = conv(b00, 3, b11, 3, lane - 16) - conv(b01, 3, b10, 3, lane - 16);
  }
  __syncwarp();
  if (lane < 11) {
    const double v = (conv(ws.p1, 7, ws.Bz + 30, 3, lane) + conv(ws.p2, 7, ws.Bz + 35, 3, lane)) +
                     conv(ws.p3, 6, ws.Bz + 40, 4, lane);
    ws.nz[lane] = v;
    ws.rz[10 - lane] = v;
  }
  __syncwarp();
  if (SYNC) __syncthreads();  // keep the CTA's warps in the same phase (one instruction stream)
  // ------------------------------------------ phase 5: two Sturm chains
  const int h = lane >> 4, hl = lane & 15;
  const double* poly = h ? ws.rz : ws.nz;
  int len;
  {
    int n = 10;
    while (n > 0 && poly[n] == 0.0) --n;
    if (hl <= n) ws.ch[h][0][hl] = poly[hl];
    len = 1;
    if (hl == 0) ws.deg[h][0] = n;
    if (n >= 1) {
      if (hl < n) ws.ch[h][1][hl] = (double)(hl + 1) * poly[hl + 1];
      if (hl == 0) ws.deg[h][1] = n - 1;
      len = 2;
    }
  }
  __syncwarp();
  bool stopped = false;
#pragma unroll 1
  for (int step = 0; step < 10; ++step) {
    const bool act = !stopped && len >= 2 && len < 12 && ws.deg[h][len - 1] > 0;
    if (__ballot_sync(kFull, act) == 0u) break;
    const int da = act ? ws.deg[h][len - 2] : 0, db = act ? ws.deg[h][len - 1] : 1;
    const double* bp = ws.ch[h][act ? len - 1 : 0];
    double r = (act && hl <= da) ? ws.ch[h][len - 2][hl] : 0.0;
    const int nsteps = act ? (da - db + 1) : 0;
    const int maxsteps = max(nsteps, __shfl_xor_sync(kFull, nsteps, 16));
    for (int s = 0; s < maxsteps; ++s) {
      const bool go = act && s < nsteps;
      const int d = da - s;
      const double rd = __shfl_sync(kFull, r, go ? d : 0, 16);
      if (go) {
        const double f = kdiv(rd, bp[db]);
        const int t = hl - (d - db);
        if (t >= 0 && t < db) r = r - f * bp[t];
        if (hl == d) r = 0.0;
      }
    }
    const unsigned bal = __ballot_sync(kFull, act && hl < db && r != 0.0);
    const unsigned mine = (bal >> (h * 16)) & 0xFFFFu;
    const int dr = mine ? (31 - __clz(mine)) : -1;
    const double rdr = __shfl_sync(kFull, r, dr < 0 ? 0 : dr, 16);
    if (act) {
      if (dr < 0) {
        stopped = true;  // exact gcd reached
      } else {
        const double sc = fabs(rdr);
        if (hl <= dr) ws.ch[h][len][hl] = -kdiv(r, sc);
        if (hl == 0) ws.deg[h][len] = dr;
        ++len;
      }
    }
    __syncwarp();
  }
  if (SYNC) __syncthreads();  // keep the CTA's warps in the same phase (one instruction stream)
  // ------------------------------- phase 6: roots in (-1,1] of each chain
  bool rvalid = false;
  double rz = 0.0;
  {
    const int d0 = ws.deg[h][0];
    // the counts are cooperative (all lanes participate); a degenerate chain
    // (degree < 1) yields R = 0 below
    const int vm1 = sturm_count_half(ws, h, hl, len, -1.0);
    const int vp1 = sturm_count_half(ws, h, hl, len, 1.0);
    const int R = (d0 >= 1) ? min(vm1 - vp1, 10) : 0;
    const int Rmax = max(R, __shfl_xor_sync(kFull, R, 16));
    // isolate root j of each chain by bisection on the count; lane hl == j keeps the bracket
    double lo = -1.0, hi = 1.0;
#pragma unroll 1
    for (int j = 0; j < Rmax; ++j) {
      double l = -1.0, u = 1.0;
      int vlo = vm1, vhi = vp1, jj = j;
      const bool mine = j < R;
#pragma unroll 1
      for (int depth = 0; depth < kWRootDepth; ++depth) {
        const bool go = mine && (vlo - vhi != 1);
        if (__ballot_sync(kFull, go) == 0u) break;
        const double mid = 0.5 * (l + u);
        const int vm = sturm_count_half(ws, h, hl, len, mid);
        if (go) {
          const int left = vlo - vm;
          if (jj < left) { u = mid; vhi = vm; } else { jj -= left; l = mid; vlo = vm; }
        }
      }
      if (hl == j) { lo = l; hi = u; }
    }
    if (hl < R) {
      const double* c0 = ws.ch[h][0];
      const double* c1 = ws.ch[h][1];
      const int d1 = ws.deg[h][1];
      double flo = horner(c0, d0, lo);
      const double fhi = horner(c0, d0, hi);
      if (fhi == 0.0) {
        rvalid = true;
        rz = hi;
      } else if ((flo < 0.0 && fhi > 0.0) || (flo > 0.0 && fhi < 0.0)) {
        for (int it = 0; it < kWRootBisect; ++it) {
          const double mid = 0.5 * (lo + hi);
          const double fm = horner(c0, d0, mid);
          if ((fm < 0.0) == (flo < 0.0)) { lo = mid; flo = fm; } else { hi = mid; }
        }
        double x = 0.5 * (lo + hi);
        for (int it = 0; it < kWRootNewton; ++it) {
          const double fx = horner(c0, d0, x);
          const double dfx = horner(c1, d1, x);
          if ((fx < 0.0) == (flo < 0.0)) { lo = x; flo = fx; } else { hi = x; }
          double xn = x - kdiv(fx, dfx);
          if (!(xn >= lo && xn <= hi)) xn = 0.5 * (lo + hi);
          x = xn;
        }
        rvalid = true;
        rz = x;
      }
      if (h == 1 && rvalid) {
        if (rz == 1.0 || rz == 0.0) rvalid = false;  // z = 1 belongs to chain 0; u = 0 is z = inf
        else rz = kdiv(1.0, rz);
      }
    }
  }
  __syncwarp();
  int nroots;
  {
    const unsigned bal = __ballot_sync(kFull, rvalid);
    const int pos = __popc(bal & ((1u << lane) - 1u));
    if (rvalid && pos < 10) ws.zr[pos] = rz;
    nroots = min(__popc(bal), 10);
  }
  __syncwarp();
  if (SYNC) __syncthreads();  // keep the CTA's warps in the same phase (one instruction stream)
  // --------------------------- phase 7: E per root, SVD, rotation candidates
  int ns;
  {
    bool ok = false;
    double Ev[9];
    if (lane < nroots) {
      const double z = ws.zr[lane];
      const double d = horner(ws.p3, 6, z);
      const double x = kdiv(horner(ws.p1, 7, z), d);
      const double y = kdiv(horner(ws.p2, 7, z), d);
      ok = true;
#pragma unroll
      for (int e = 0; e < 9; ++e) {
        const double v = ((x * B[e] + y * B[9 + e]) + z * B[18 + e]) + B[27 + e];
        if (!isfinite(v)) ok = false;
        Ev[e] = v;
      }
    }
    const unsigned bal = __ballot_sync(kFull, ok);
    ns = __popc(bal);
    if (ok) {
      const int pos = __popc(bal & ((1u << lane) - 1u));
#pragma unroll
      for (int e = 0; e < 9; ++e) ws.E[pos][e] = Ev[e];
    }
  }
  __syncwarp();
  if (lane < ns) {
    double U[9], S[3], V[9];
    svd3(ws.E[lane], U, S, V);
#pragma unroll 1
    for (int r = 0; r < 3; ++r)
#pragma unroll 1
      for (int c = 0; c < 3; ++c) {
        const double a0 = U[3 * r + 0], a1 = U[3 * r + 1], a2 = U[3 * r + 2];
        const double b0 = V[3 * c + 0], b1 = V[3 * c + 1], b2 = V[3 * c + 2];
        ws.Ra[lane][3 * r + c] = (a1 * b0 - a0 * b1) + a2 * b2;
        ws.Rb[lane][3 * r + c] = (a0 * b1 - a1 * b0) + a2 * b2;
      }
    ws.tt[lane][0] = S[0] * U[2];
    ws.tt[lane][1] = S[0] * U[5];
    ws.tt[lane][2] = S[0] * U[8];
  }
  __syncwarp();
  if (SYNC) __syncthreads();  // keep the CTA's warps in the same phase (one instruction stream)
  // ------------------------- phase 8: score (solution, candidate) on 8 points
  double best = 1000000.0;
  bool found = false;
  int be = 0, bc = 0;
  {
    const int c = lane >> 3, k = lane & 7;
    const V3 g1 = {ws.f1[k][0], ws.f1[k][1], ws.f1[k][2]};
    const V3 g2 = {ws.f2[k][0], ws.f2[k][1], ws.f2[k][2]};
    const double sgn = (c & 1) ? -1.0 : 1.0;
#pragma unroll 1
    for (int e = 0; e < ns; ++e) {
      const double* R = (c < 2) ? ws.Ra[e] : ws.Rb[e];
      double Mc[12], tinv[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        Mc[4 * r + 0] = R[3 * r + 0];
        Mc[4 * r + 1] = R[3 * r + 1];
        Mc[4 * r + 2] = R[3 * r + 2];
        Mc[4 * r + 3] = sgn * ws.tt[e][r];
      }
      mono_tinv(Mc, tinv);
      const double res = mono_residual(Mc, tinv, g1, g2);
      double q = 0.0;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) q = q + __shfl_sync(kFull, res, (lane & 24) + kk);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const double qc = __shfl_sync(kFull, q, cc * 8);
        if (qc < best) { best = qc; found = true; be = e; bc = cc; }
      }
    }
  }
  if (found) {
    const double* R = (bc < 2) ? ws.Ra[be] : ws.Rb[be];
    const double sgn = (bc & 1) ? -1.0 : 1.0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      M[4 * r + 0] = R[3 * r + 0];
      M[4 * r + 1] = R[3 * r + 1];
      M[4 * r + 2] = R[3 * r + 2];
      M[4 * r + 3] = sgn * ws.tt[be][r];
    }
  }
  return found && !failed;
}

}  // namespace geom
}  // namespace kml
