// ransac.cu — batched opengv::sac::Ransac on sm_100a (compiled -fmad=false).
//
// Replaces Ransac<CentralRelativePoseSacProblem>(NISTER)::computeModel as
// called by LoopClosureDetector::geometricVerificationNister and
// Ransac<PointCloudSacProblem>::computeModel as called by recoverPose
// (/root/reference/images/kimera-multi.drawio:2589-2598,
// 2646, 2654; thresholds of the same family:
// /root/reference/params/D455/LcdParams.yaml:51-56,64-66).
//
// The reference loop is sequential (adaptive stop).  Here every problem
// (candidate keyframe pair) advances in ROUNDS of `chunk` draws of the
// pre-drawn sample stream:
//   chunk kernel  : one CTA per problem; phase 1 one hypothesis per thread
//                   (minimal solver), phase 2 one hypothesis per warp-pass
//                   (lanes stride over all correspondences, ballot+popc
//                   inlier count); correspondences live in shared memory.
//   replay kernel : one warp per problem; lane 0 replays the reference's
//                   serial scan over the chunk (skip rule, strict-improvement
//                   rule, k update from a host-computed table, iteration
//                   cap), then draws the next chunk of samples from the
//                   persistent partial Fisher-Yates state.
// A finished problem's CTAs exit immediately, so the host can enqueue all
// rounds without synchronising.
#include <limits.h>

#include "common.cuh"
#include "geom.cuh"
#include "fivept_thread.cuh"
#include "kernels.h"

namespace kml {

using geom::V3;

// --------------------------------------------------------------- sampling
// SampleConsensusProblem::drawIndexSample: for i<S swap(shuffled[i],
// shuffled[i + rnd() % (N-i)]); rnd() = mt19937()>>1 (pre-drawn, host).
// Warp version: the modulo of every (draw, i) is independent of the shuffle state, so all lanes
// compute them (jrel_s, at most kRoundCap*S entries) and lane 0 is left with the swaps.
// out[slot][S]: the samples of the round's draws [first, last), slot = draw - first.
template <int S>
__device__ void draw_samples_warp(uint16_t* perm_s, uint16_t* jrel_s, int N, const uint32_t* __restrict__ raw,
                                  int first, int last, uint16_t* out, int lane) {
  const int n = (last - first) * S;
  for (int idx = lane; idx < n; idx += 32) {
    const int i = idx % S;
    jrel_s[idx] = (uint16_t)(raw[first * S + idx] % (uint32_t)(N - i));
  }
  __syncwarp();
  if (lane == 0) {
    for (int d = 0; d < last - first; ++d) {
#pragma unroll
      for (int i = 0; i < S; ++i) {
        const int j = i + (int)jrel_s[d * S + i];
        const uint16_t t = perm_s[i];
        perm_s[i] = perm_s[j];
        perm_s[j] = t;
      }
#pragma unroll
      for (int i = 0; i < S; ++i) out[(size_t)d * S + i] = perm_s[i];  // slot d of the round
    }
  }
  __syncwarp();
}
// dynamic shared memory of sac_init / sac_replay: perm[stride] u16 | vc[kRoundCap] i32 | jrel[kRoundCap*S] u16
__host__ __device__ inline size_t sac_perm_bytes(int stride) { return ((size_t)stride * 2 + 15) / 16 * 16; }
template <int S>
__host__ __device__ inline size_t sac_warp_smem(int stride) {
  return sac_perm_bytes(stride) + sizeof(int32_t) * kRoundCap + sizeof(uint16_t) * kRoundCap * S;
}

// samples of draw `gd` of problem p (N correspondences): from the per-N table, or from the
// round's per-problem buffer (slot = gd - r_begin)
template <int S>
__device__ __forceinline__ const uint16_t* sac_sample(const SacArgs& a, int p, int N, int gd, int slot) {
  if (a.sample_tab) return a.sample_tab + ((size_t)N * a.cap_draws + gd) * S;
  return a.samples + ((size_t)p * kRoundCap + slot) * S;
}

// tab[N][cap_draws][S]: one warp per N replays drawIndexSample over the whole stream
template <int S>
__global__ void __launch_bounds__(32) sample_table_kernel(const uint32_t* __restrict__ raw, int cap_draws, int nmax,
                                                          uint16_t* __restrict__ tab) {
  KML_DYN_SMEM(uint16_t, perm_s);
  const int N = blockIdx.x, lane = threadIdx.x;
  if (N < S || N > nmax) return;
  uint16_t* jrel_s = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(perm_s) + sac_perm_bytes(nmax));
  for (int i = lane; i < N; i += 32) perm_s[i] = (uint16_t)i;
  __syncwarp();
  for (int first = 0; first < cap_draws; first += kRoundCap) {
    const int last = min(cap_draws, first + kRoundCap);
    draw_samples_warp<S>(perm_s, jrel_s, N, raw, first, last, tab + ((size_t)N * cap_draws + first) * S, lane);
  }
}

template <int S, int CHUNK>
__global__ void __launch_bounds__(32) sac_init_kernel(SacArgs a) {
  KML_DYN_SMEM(uint16_t, perm_s);
  const int p = blockIdx.x;
  const int lane = threadIdx.x;
  const int N = a.N[p];
  SacState* st = &a.st[p];
  const bool tab = a.sample_tab != nullptr;
  if (!tab)
    for (int i = lane; i < N; i += 32) perm_s[i] = (uint16_t)i;
  bool unit = true;
  if (S == 8) {  // unit bearings are the precondition of the fast inlier filter (geom::mono_inlier_fast)
    const double* ga = a.a + (size_t)p * a.stride * 3;
    const double* gb = a.b + (size_t)p * a.stride * 3;
    for (int i = lane; i < N; i += 32) {
      const double na = (ga[3 * i] * ga[3 * i] + ga[3 * i + 1] * ga[3 * i + 1]) + ga[3 * i + 2] * ga[3 * i + 2];
      const double nb = (gb[3 * i] * gb[3 * i] + gb[3 * i + 1] * gb[3 * i + 1]) + gb[3 * i + 2] * gb[3 * i + 2];
      unit = unit && fabs(na - 1.0) <= 1e-12 && fabs(nb - 1.0) <= 1e-12;
    }
  }
  unit = __all_sync(0xFFFFFFFFu, unit);
  __syncwarp();
  if (lane == 0) {
    st->unit_bearings = unit ? 1 : 0;
    st->iterations = 0;
    st->skipped = 0;
    st->draws = 0;
    st->best = -INT_MAX;
    st->best_draw = -1;
    st->exhausted = 0;
    st->k = 1.0;
    st->done = (N < S) ? 1 : 0;  // getSamples(): N < sample_size => loop exits, no model
    st->r_begin = 0;
    st->r_end = (N < S) ? 0 : min(sac_round_draws(0, a.first), a.cap_draws);
  }
  __syncwarp();
  if (!tab) {
    if (N >= S) {
      uint16_t* jrel_s = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(perm_s) + sac_perm_bytes(a.stride) +
                                                     sizeof(int32_t) * kRoundCap);
      draw_samples_warp<S>(perm_s, jrel_s, N, a.raw, 0, min(sac_round_draws(0, a.first), a.cap_draws),
                           a.samples + (size_t)p * kRoundCap * S, lane);
    }
    for (int i = lane; i < N; i += 32) a.perm[(size_t)p * a.stride + i] = perm_s[i];
  }
  if (a.n_inliers && lane == 0) a.n_inliers[p] = 0;
  if (lane == 0 && N >= S) a.active[atomicAdd(&a.n_active[0], 1u)] = p;  // round 0 works on every solvable problem
}

// `(double)iterations < k` for a non-negative integer count, as an integer limit: iterations < ceil(k)
// (NaN and k <= 0 never hold; beyond INT_MAX always does for any count the loop can reach)
__device__ __forceinline__ int sac_k_limit(double k) {
  if (!(k > 0.0)) return 0;
  if (k >= 2147483647.0) return INT_MAX;
  return (int)ceil(k);
}
// Ransac::computeModel control flow, replayed over the draws of one round.
// k never increases, so after a round the number of trials still required is
// known exactly (up to skipped samples): the next round covers all of them.
template <int S, int CHUNK>
__device__ void sac_replay_body(const SacArgs& a, int round, int p) {
  KML_DYN_SMEM(uint16_t, perm_s);
  const int lane = threadIdx.x;
  SacState* st = &a.st[p];
  if (st->done) return;
  const int N = a.N[p];
  int32_t* vc_s = reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(perm_s) + sac_perm_bytes(a.stride));
  uint16_t* jrel_s = reinterpret_cast<uint16_t*>(vc_s + kRoundCap);
  const bool tab = a.sample_tab != nullptr;
  if (!tab)
    for (int i = lane; i < N; i += 32) perm_s[i] = a.perm[(size_t)p * a.stride + i];
  {  // the round's (valid, count) results, fetched by the whole warp: count, or -1 for "no model"
    const int rb = st->r_begin, re = st->r_end;
    const int32_t* gv = a.valid + (size_t)p * kRoundCap;
    const int32_t* gc = a.counts + (size_t)p * kRoundCap;
    for (int i = lane; i < re - rb; i += 32) vc_s[i] = gv[i] ? gc[i] : -1;
  }
  __syncwarp();
  int nb_w = 0, ne_w = 0, done_w = 1;
  if (lane == 0) {
    int iterations = st->iterations, skipped = st->skipped, draws = st->draws;
    int best = st->best, best_draw = st->best_draw, done = 0, exhausted = 0;
    double k = st->k;
    const int max_skip = a.max_iterations * 10;
    const int r_begin = st->r_begin, r_end = st->r_end;
    int best_slot = -1;
    // the loop condition in integers (one lane walks up to 512 draws: every cycle of this chain is latency
    // of the round); klim changes only when a better model shows up
    int klim = a.full ? INT_MAX : sac_k_limit(k);
#pragma unroll 4
    for (int gd = r_begin; gd < r_end; ++gd) {
      if (!(iterations < klim && skipped < max_skip)) { done = 1; break; }
      ++draws;
      const int n = vc_s[gd - r_begin];
      if (n < 0) { ++skipped; continue; }
      if (n > best) {
        best = n;
        best_draw = gd;
        best_slot = gd - r_begin;
        k = a.ktable[(size_t)N * a.ktable_n + n];
        if (!a.full) klim = sac_k_limit(k);
      }
      ++iterations;
      if (iterations > a.max_iterations) { done = 1; break; }
    }
    if (!done && !((a.full || (double)iterations < k) && skipped < max_skip)) done = 1;
    int nb = r_end, ne = r_end;
    if (!done) {
      // trials still required if no further sample is skipped
      const int cap_it = a.max_iterations + 1;
      int target = cap_it;
      if (!a.full && k < (double)cap_it) target = (int)ceil(k);
      int rem = target - iterations;
      if (rem < 1) rem = 1;
      // k is only an upper bound (it shrinks whenever a better model shows
      // up), so the next round evaluates at most as many new draws as have
      // been evaluated so far (doubling, capped at kRoundCap).  The host enqueues
      // kSacRounds rounds blindly and keeps adding rounds while a problem is
      // pending, so every draw the reference loop can consume (max_iterations + 1
      // counted trials + max_skip skipped samples = cap_draws) is reachable.
      const int grow = sac_round_draws(round + 1, a.first);
      ne = min(a.cap_draws, nb + min(rem + 16, grow));
      if (nb >= a.cap_draws) { done = 1; exhausted = 1; }  // unreachable: the loop ends by its own limits first
    }
    if (best_slot >= 0) {  // model_coefficients_ = model
      const double* m = a.models + ((size_t)p * kRoundCap + best_slot) * 12;
      for (int i = 0; i < 12; ++i) a.best_model[(size_t)p * 12 + i] = m[i];
    }
    st->iterations = iterations;
    st->skipped = skipped;
    st->draws = draws;
    st->best = best;
    st->best_draw = best_draw;
    st->k = k;
    st->exhausted = exhausted;
    st->done = done;
    st->r_begin = nb;
    st->r_end = ne;
    nb_w = nb; ne_w = ne; done_w = done;
  }
  nb_w = __shfl_sync(0xFFFFFFFFu, nb_w, 0);
  ne_w = __shfl_sync(0xFFFFFFFFu, ne_w, 0);
  done_w = __shfl_sync(0xFFFFFFFFu, done_w, 0);
  if (!done_w && !tab) {
    draw_samples_warp<S>(perm_s, jrel_s, N, a.raw, nb_w, ne_w, a.samples + (size_t)p * kRoundCap * S, lane);
    for (int i = lane; i < N; i += 32) a.perm[(size_t)p * a.stride + i] = perm_s[i];
  }
  // the problems the next round still works on (order is irrelevant: every buffer is indexed by p)
  if (!done_w && lane == 0) {
    const int nx = (round + 1) & 1;
    a.active[(size_t)nx * a.P + atomicAdd(&a.n_active[nx], 1u)] = p;
  }
}
// Every per-round kernel walks the round's ACTIVE LIST (written by sac_init for round 0 and by the
// previous round's replay afterwards) with a grid that shrinks with the round: a late round
// launches a few hundred CTAs for the handful of problems still running instead of one CTA per
// problem slot that finds `done` and leaves.
#define KML_ACTIVE_LOOP(body_call)                                                    \
  const int32_t* list_ = a.active + (size_t)(round & 1) * a.P;                        \
  const int n_act_ = (int)a.n_active[round & 1];                                      \
  for (int ai_ = blockIdx.x; ai_ < n_act_; ai_ += gridDim.x) {                        \
    const int p = list_[ai_];                                                         \
    body_call;                                                                        \
    __syncthreads(); /* shared memory of the body is reused by the next problem */    \
  }
template <int S, int CHUNK>
__global__ void __launch_bounds__(32) sac_replay_kernel(SacArgs a, int round) {
  KML_ACTIVE_LOOP((sac_replay_body<S, CHUNK>(a, round, p)))
}

// ------------------------------------------------------------ mono round
// Six kernels per round.  CTA (p, blk) of the grid kernels owns the 64 draws
// r_begin + blk*64 + [0,64) of problem p; they hand over through per-round
// global buffers indexed by slot = p*kRoundCap + blk*64 + t.
//   mono_front_kernel    thread = draw: null space, constraint build, Gauss-
//                        Jordan, cofactor polynomials (fivept_thread.cuh stage 1);
//                        200 shared-memory slots per thread -> 4 warps per SM.
//   mono_isolate_kernel  thread = draw: Sturm chains + isolating brackets of the
//                        real roots (stage 2, registers only); reserves the
//                        draw's range in the round's item list.
//   mono_isolate_deferred_kernel  thread = (draw, chain, root) of the compacted
//                        list of chains whose sign grid did not separate the roots.
//   mono_item_kernel     thread = (draw, root) ITEM, grid-stride over the whole
//                        round's list: refine root, E, SVD, the four (R,t)
//                        candidates scored on the 8 sample points — every item
//                        costs the same, no barrier, no tail.
//   mono_count_kernel    thread = draw: winner in the reference's order (strict <,
//                        first 10 refined roots); warp = draw: lanes stride over
//                        the correspondences, inlier count by ballot + popc with
//                        the fast inlier filter, stopping once the draw cannot
//                        beat the best count of the draws before it.
//   sac_replay_kernel    lane 0 per problem replays Ransac::computeModel.
// The front kernel's CTA is ONE warp of draws (its 200 shared-memory slots per draw set the
// occupancy: 51 KB per CTA, four CTAs per SM): the first rounds evaluate 32 draws per problem, and a
// 64-thread CTA would hold a second warp's worth of shared memory for nothing.
constexpr int kFrontChunk = 32;
template <int ALG>
__device__ void mono_front_body(const SacArgs& a, int p) {
  KML_DYN_SMEM(double, smem_d);
  const SacState st = a.st[p];
  if (st.done) return;
  const int d0 = st.r_begin + blockIdx.y * kFrontChunk;
  if (d0 >= st.r_end) return;
  const int tid = threadIdx.x;
  const double* ga = a.a + (size_t)p * a.stride * 3;
  const double* gb = a.b + (size_t)p * a.stride * 3;
  const int nh = min(kFrontChunk, st.r_end - d0);
  const bool live = tid < nh;
  const size_t slot = (size_t)p * kRoundCap + blockIdx.y * kFrontChunk + tid;
  const int js = live ? blockIdx.y * kFrontChunk + tid : 0;
  const uint16_t* smp = sac_sample<8>(a, p, a.N[p], st.r_begin + js, js);
  geom::mono_front_thread<kFrontChunk, true, ALG>(smem_d + tid, ga, gb, smp, live, a.fsol + slot * a.fo_stride);
}
template <int ALG>
__global__ void __launch_bounds__(kFrontChunk, 4) mono_front_kernel(SacArgs a, int round) {
  KML_ACTIVE_LOOP(mono_front_body<ALG>(a, p))
}

// One warp of draws per CTA, like the front kernel: the first rounds evaluate 32 draws per problem, and the
// second warp of a 64-thread CTA would only hold registers (9 instead of 18 working warps per SM).
constexpr int kIsoChunk = 32;
__device__ void mono_isolate_body(const SacArgs& a, int p) {
  const SacState st = a.st[p];
  if (st.done) return;
  const int d0 = st.r_begin + blockIdx.y * kIsoChunk;
  if (d0 >= st.r_end) return;
  const int tid = threadIdx.x, lane = tid & 31;
  const int nh = min(kIsoChunk, st.r_end - d0);
  const size_t slot = (size_t)p * kRoundCap + blockIdx.y * kIsoChunk + tid;
  int nr = 0;
  if (tid < nh) {
    nr = geom::mono_isolate_thread(a.fsol + slot * a.fo_stride, a.brk + slot * 2 * geom::kMaxBrackets,
                                   (a.force_generic & 1) != 0);
    for (int chain = 0; chain < 2; ++chain)
      if ((nr >> (16 + chain)) & 1) {  // one deferred item per chain
        const unsigned at = atomicAdd(a.fb_count, 1u);
        if (at + 1u > a.item_cap) { *a.overflow = 1u; nr = 0; break; }  // the host re-runs the batch with larger lists
        a.fb_list[at] = (uint32_t)(slot * 2 + chain);
      }
  }
  // item ranges: one reservation per warp, lanes take consecutive sub-ranges
  int n = (nr & 255) + ((nr >> 8) & 255);
  int incl = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += v;
  }
  const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
  unsigned base = 0;
  if (lane == 0 && total > 0) base = atomicAdd(a.item_count, (unsigned)total);
  const unsigned wbase = __shfl_sync(0xFFFFFFFFu, base, 0);
  base = wbase + (unsigned)(incl - n);
  if (wbase + (unsigned)total > a.item_cap) {  // list full: no items for this warp's draws, flag the batch
    if (lane == 0) *a.overflow = 1u;
    n = 0;
    nr = 0;
  }
  if (tid < nh) {
    a.nroot[slot] = nr & 0xFFFF;
    a.item_base[slot] = base;
    for (int r = 0; r < n; ++r) a.item_list[base + r] = (uint32_t)(slot * 32 + r);
  }
}
__global__ void __launch_bounds__(kIsoChunk) mono_isolate_kernel(SacArgs a, int round) {
  KML_ACTIVE_LOOP(mono_isolate_body(a, p))
}

// Deferred root isolations of the round ((draw, chain) pairs whose 32-cell grid did not
// separate the roots, ~8 % of the chains): compacted work list of chains.
// One WARP per deferred chain (grid-stride over the compacted list): the 256-cell grid evaluated side by
// side (geom::isolate_grid2_w), and for the chains it does not separate either the Sturm chain and the
// bisections (geom::sturm_build_w / isolate_roots_w).
__global__ void __launch_bounds__(kMonoChunk) mono_isolate_deferred_kernel(SacArgs a) {
  __shared__ double wsm[kMonoChunk / 32][geom::kWarpSturmSlots];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* w = wsm[wib];
  const unsigned n = min(*a.fb_count, a.item_cap);
  constexpr unsigned kWarps = kMonoChunk / 32;
  for (unsigned it = blockIdx.x * kWarps + wib; it < n; it += gridDim.x * kWarps) {
    const uint32_t item = a.fb_list[it];
    const size_t slot = item >> 1;
    const bool chain = item & 1u;
    const int R0 = a.nroot[slot] & 255, R1 = (a.nroot[slot] >> 8) & 255;
    const double* fo = a.fsol + slot * a.fo_stride;
    double* brk = a.brk + slot * 2 * geom::kMaxBrackets + (chain ? 2 * R0 : 0);
    if (!(a.force_generic & 2) &&  // test hook: every deferred chain through the Sturm fallback
        geom::isolate_grid2_w(fo, chain, chain ? R1 : R0, brk, lane))
      continue;
    unsigned long long degs;
    __syncwarp();
    const int len = geom::sturm_build_w(w, fo, chain, lane, &degs);
    geom::isolate_roots_w(w, degs, len, brk, lane);
  }
}

// thread = (draw, root) item over the whole round (grid-stride over the compacted list)
constexpr int kItemThreads = 128;
template <int ALG>
__global__ void __launch_bounds__(kItemThreads, 4) mono_item_kernel(SacArgs a) {
  const unsigned n = min(*a.item_count, a.item_cap);
  for (unsigned it = blockIdx.x * kItemThreads + threadIdx.x; it < n; it += gridDim.x * kItemThreads) {
    const uint32_t code = a.item_list[it];
    const size_t slot = code >> 5;
    const int r = code & 31;
    const int p = (int)(slot / kRoundCap);
    const int R0 = a.nroot[slot] & 255;
    const double* fo = a.fsol + slot * a.fo_stride;
    const double* bk = a.brk + slot * 2 * geom::kMaxBrackets + 2 * r;
    const double* ga = a.a + (size_t)p * a.stride * 3;
    const double* gb = a.b + (size_t)p * a.stride * 3;
    const int js = (int)(slot % kRoundCap);
    const uint16_t* smp = sac_sample<8>(a, p, a.N[p], a.st[p].r_begin + js, js);
    double q = 0.0, M[12];
    const int status = geom::mono_item<ALG>(fo, r >= R0 ? 1 : 0, bk[0], bk[1], ga, gb, smp, &q, M);
    a.item_status[it] = (uint8_t)status;
    if (status == 2) {
      a.item_q[it] = q;
#pragma unroll
      for (int i = 0; i < 12; ++i) a.item_model[(size_t)it * 12 + i] = M[i];
    }
  }
}

#ifdef KML_FILTER_STATS
using geom::g_fstats;
__global__ void fstats_print_kernel() {
  printf("filter: evals %llu rcp_bad %llu rsqrt_bad %llu det %llu n2 %llu pfloor %llu qfloor %llu band %llu insane %llu\n",
         g_fstats[0], g_fstats[1], g_fstats[2], g_fstats[3], g_fstats[4], g_fstats[5], g_fstats[6], g_fstats[7], g_fstats[8]);
}
#endif
// CTA = 64 draws of one problem.  (1) thread = draw: winner among the draw's items in the
// reference's order (roots in order, at most the first 10 refined ones, strict <); the item
// records are fetched as two batches of independent loads.  (2) warp = draw: inlier count of
// the winning model over the problem's correspondences, staged once per CTA in shared memory.
constexpr int kCountThreads = 256;
constexpr int kCountWarps = kCountThreads / 32;
template <bool STAGED>  // STAGED: the problem's bearings fit in shared memory (else read through L1)
__device__ void mono_count_body(const SacArgs& a, int p) {
  KML_DYN_SMEM(double, smem_d);
  const SacState st = a.st[p];
  if (st.done) return;
  const int d0 = st.r_begin + blockIdx.y * kMonoChunk;
  if (d0 >= st.r_end) return;
  const int N = a.N[p];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nh = min(kMonoChunk, st.r_end - d0);
  const size_t slot0 = (size_t)p * kRoundCap + blockIdx.y * kMonoChunk;
  // A draw changes the outcome of the reference loop only if its count exceeds the best count
  // of every EARLIER draw, so counting stops as soon as that is impossible: the bound starts at
  // the best count before this round and rises with the counts of the earlier groups of
  // kCountWarps draws of this chunk (s_cum[g], read as the warps go).
  constexpr int kGroups = kMonoChunk / kCountWarps;
  __shared__ int s_cum[kGroups];
  __shared__ int s_valid[kMonoChunk];
  __shared__ double s_mod[kMonoChunk][12];
  double* s_f = smem_d;  // [6][NP] structure-of-arrays copy of the bearings: f1.xyz, f2.xyz
  const int NP = (N + 31) & ~31;
  const double* ga = a.a + (size_t)p * a.stride * 3;
  const double* gb = a.b + (size_t)p * a.stride * 3;
  if (STAGED) {
    for (int f = tid; f < 3 * N; f += kCountThreads) {
      const int i = f / 3, c = f - 3 * i;
      s_f[c * NP + i] = ga[f];
      s_f[(3 + c) * NP + i] = gb[f];
    }
  }
  const int bound0 = a.full ? -INT_MAX : st.best;
  if (tid < kGroups) s_cum[tid] = bound0;
  if (tid < kMonoChunk) {
    int v = 0;
    if (tid < nh) {
      const size_t slot = slot0 + tid;
      const int nr = a.nroot[slot];
      const int n = (nr & 255) + (nr >> 8);
      const unsigned base = a.item_base[slot];
      int status[geom::kMaxBrackets];
      double q[geom::kMaxBrackets];
#pragma unroll
      for (int r = 0; r < geom::kMaxBrackets; ++r) status[r] = (r < n) ? (int)a.item_status[base + r] : 0;
#pragma unroll
      for (int r = 0; r < geom::kMaxBrackets; ++r) q[r] = (status[r] == 2) ? a.item_q[base + r] : 0.0;
      double best = 1000000.0;
      int br = -1, refined = 0;
#pragma unroll
      for (int r = 0; r < geom::kMaxBrackets; ++r) {
        if (status[r] >= 1 && refined < 10) {
          if (status[r] == 2 && q[r] < best) { best = q[r]; br = r; }
        }
        refined += status[r] >= 1;
      }
      if (br >= 0) {
        v = 1;
        const double* m = a.item_model + (size_t)(base + br) * 12;
#pragma unroll
        for (int i = 0; i < 12; ++i) s_mod[tid][i] = m[i];
      }
    }
    s_valid[tid] = v;
  }
  __syncthreads();
  const bool filter = st.unit_bearings != 0;
  const geom::InlierMargins mg = geom::inlier_margins(a.threshold);
  for (int h = warp; h < nh; h += kCountWarps) {
    const int g = h / kCountWarps;
    const int v = s_valid[h];
    int cnt = 0;
    if (v) {
      double M[12], tinv[3];
#pragma unroll
      for (int i = 0; i < 12; ++i) M[i] = s_mod[h][i];
      geom::mono_tinv(M, tinv);
      for (int i0 = 0; i0 < N; i0 += 32) {
        const int i = i0 + lane;
        bool in = false;
        if (i < N) {
          V3 f1, f2;
          if (STAGED) {
            f1 = {s_f[i], s_f[NP + i], s_f[2 * NP + i]};
            f2 = {s_f[3 * NP + i], s_f[4 * NP + i], s_f[5 * NP + i]};
          } else {
            f1 = {ga[3 * i], ga[3 * i + 1], ga[3 * i + 2]};
            f2 = {gb[3 * i], gb[3 * i + 1], gb[3 * i + 2]};
          }
          const int fast = filter ? geom::mono_inlier_fast(M, tinv, f1, f2, mg) : -1;
          in = fast > 0;
          if (fast < 0) in = geom::mono_residual(M, tinv, f1, f2) < a.threshold;
        }
        cnt += __popc(__ballot_sync(0xFFFFFFFFu, in));
        const int bound = __shfl_sync(0xFFFFFFFFu, *reinterpret_cast<volatile int*>(&s_cum[g]), 0);
        if (cnt + (N - i0 - 32) <= bound) break;  // cannot become the best model any more
      }
      // a (possibly partial) count is a lower bound of the true one: valid bound for later groups
      if (!a.full && lane == 0)
        for (int g2 = g + 1; g2 < kGroups; ++g2) atomicMax(&s_cum[g2], cnt);
    }
    if (lane == 0) {
      a.valid[slot0 + h] = v;
      a.counts[slot0 + h] = cnt;
    }
    if (v && lane < 12) a.models[(slot0 + h) * 12 + lane] = s_mod[h][lane];
  }
}
template <bool STAGED>
__global__ void __launch_bounds__(kCountThreads, 2) mono_count_kernel(SacArgs a, int round) {
  KML_ACTIVE_LOOP(mono_count_body<STAGED>(a, p))
}

// ---------------------------------------------------------- stereo chunk
// ONEPT (row f4): rotation given, one correspondence per draw
template <bool STAGED, bool ONEPT>  // STAGED: the problem's point pairs fit in shared memory (else read through L1)
__device__ void stereo_chunk_body(const SacArgs& a, int p) {
  KML_DYN_SMEM(double, smem_d);
  const SacState st = a.st[p];
  if (st.done) return;
  const int d0 = st.r_begin + blockIdx.y * kStereoChunk;
  if (d0 >= st.r_end) return;
  const int N = a.N[p];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* ga = a.a + (size_t)p * a.stride * 3;
  const double* gb = a.b + (size_t)p * a.stride * 3;
  double* smod = smem_d;  // [kStereoChunk][12]
  const double* s1 = ga;
  const double* s2 = gb;
  if (STAGED) {
    double* w1 = smem_d + 12 * kStereoChunk;
    double* w2 = w1 + 3 * (size_t)N;
    for (int i = tid; i < 3 * N; i += kStereoThreads) {
      w1[i] = ga[i];
      w2[i] = gb[i];
    }
    s1 = w1;
    s2 = w2;
  }
  __syncthreads();
  const int nh = min(kStereoChunk, st.r_end - d0);
  if (tid < nh) {
    double M[12];
    if (ONEPT) {
      const int i0 = *sac_sample<1>(a, p, N, d0 + tid, blockIdx.y * kStereoChunk + tid);
      const double* R = a.prior + (size_t)p * 12;
      const double* pa = s1 + 3 * i0;
      const double* pb = s2 + 3 * i0;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const double r0 = R[4 * r], r1 = R[4 * r + 1], r2 = R[4 * r + 2];
        M[4 * r] = r0; M[4 * r + 1] = r1; M[4 * r + 2] = r2;
        M[4 * r + 3] = pa[r] - ((r0 * pb[0] + r1 * pb[1]) + r2 * pb[2]);
      }
    } else {
      const uint16_t* smp = sac_sample<3>(a, p, N, d0 + tid, blockIdx.y * kStereoChunk + tid);
      const int i0 = smp[0], i1 = smp[1], i2 = smp[2];
      geom::arun3(s1 + 3 * i0, s1 + 3 * i1, s1 + 3 * i2, s2 + 3 * i0, s2 + 3 * i1, s2 + 3 * i2, M);
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) smod[12 * tid + i] = M[i];
  }
  __syncthreads();
  // running bound as in mono_count_kernel: a draw matters only if it beats every earlier one
  constexpr int kWarps = kStereoThreads / 32, kGroups = kStereoChunk / kWarps;
  __shared__ int s_cum[kGroups];
  const int bound0 = a.full ? -INT_MAX : st.best;
  if (tid < kGroups) s_cum[tid] = bound0;
  __syncthreads();
  for (int h = warp; h < nh; h += kWarps) {
    const int g = h / kWarps;
    double M[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) M[i] = smod[12 * h + i];
    int cnt = 0;
    for (int i0 = 0; i0 < N; i0 += 32) {
      const int i = i0 + lane;
      bool in = false;
      if (i < N)
        in = geom::arun_sqdist(M, s1[3 * i], s1[3 * i + 1], s1[3 * i + 2], s2[3 * i],
                               s2[3 * i + 1], s2[3 * i + 2]) < a.sq_crit;
      cnt += __popc(__ballot_sync(0xFFFFFFFFu, in));
      const int bound = __shfl_sync(0xFFFFFFFFu, *reinterpret_cast<volatile int*>(&s_cum[g]), 0);
      if (cnt + (N - i0 - 32) <= bound) break;
    }
    if (!a.full && lane == 0)
      for (int g2 = g + 1; g2 < kGroups; ++g2) atomicMax(&s_cum[g2], cnt);
    if (lane == 0) {
      a.valid[(size_t)p * kRoundCap + blockIdx.y * kStereoChunk + h] = 1;  // threept_arun always yields a model
      a.counts[(size_t)p * kRoundCap + blockIdx.y * kStereoChunk + h] = cnt;
    }
    if (lane < 12) a.models[((size_t)p * kRoundCap + blockIdx.y * kStereoChunk + h) * 12 + lane] = smod[12 * h + lane];
  }
}
template <bool STAGED, bool ONEPT>
__global__ void __launch_bounds__(kStereoThreads) stereo_chunk_kernel(SacArgs a, int round) {
  KML_ACTIVE_LOOP((stereo_chunk_body<STAGED, ONEPT>(a, p)))
}

// ------------------------------------------------ selectWithinDistance
// selectWithinDistance(model_coefficients_, threshold, inliers_) for the winner.
template <bool MONO>
__global__ void __launch_bounds__(128) sac_select_kernel(SacArgs a) {
  const int p = blockIdx.x;
  const SacState st = a.st[p];
  const int tid = threadIdx.x, lane = tid & 31;
  __shared__ int s_cnt;
  __shared__ double s_M[12];
  if (tid == 0) {
    s_cnt = 0;
    if (!st.done && a.pending) atomicAdd(a.pending, 1u);  // the enqueued rounds did not end this problem's loop
  }
  const int N = a.N[p];
  uint32_t* mask = a.inlier_mask + (size_t)p * a.mask_words;
  if (st.best_draw < 0) {  // model_ empty: inliers_.clear(), return false
    for (int w = tid; w < a.mask_words; w += 128) mask[w] = 0u;
    if (tid == 0) a.n_inliers[p] = 0;
    return;
  }
  const double* ga = a.a + (size_t)p * a.stride * 3;
  const double* gb = a.b + (size_t)p * a.stride * 3;
  if (tid < 12) s_M[tid] = a.best_model[(size_t)p * 12 + tid];  // copied by the replay kernel
  __syncthreads();
  double M[12], tinv[3];
#pragma unroll
  for (int i = 0; i < 12; ++i) M[i] = s_M[i];
  if (MONO) geom::mono_tinv(M, tinv);
  int cnt = 0;
  const int nwords = (N + 31) / 32;
  for (int w = tid >> 5; w < a.mask_words; w += 4) {
    bool in = false;
    const int i = w * 32 + lane;
    if (w < nwords && i < N) {
      if (MONO) {
        const V3 f1 = {ga[3 * i], ga[3 * i + 1], ga[3 * i + 2]};
        const V3 f2 = {gb[3 * i], gb[3 * i + 1], gb[3 * i + 2]};
        in = geom::mono_residual(M, tinv, f1, f2) < a.threshold;
      } else {
        in = geom::arun_sqdist(M, ga[3 * i], ga[3 * i + 1], ga[3 * i + 2], gb[3 * i],
                               gb[3 * i + 1], gb[3 * i + 2]) < a.sq_crit;
      }
    }
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, in);
    if (lane == 0) {
      mask[w] = bal;
      cnt += __popc(bal);
    }
  }
  if (lane == 0 && cnt) atomicAdd(&s_cnt, cnt);
  __syncthreads();
  if (tid == 0) a.n_inliers[p] = s_cnt;
}

// ----------------------------------------------------------------- gather
__global__ void __launch_bounds__(128) gather_bearings_kernel(GatherArgs g) {
  const int p = blockIdx.x;
  const PairDesc pd = g.pairs[p];
  if (pd.m_frame < 0) {  // inactive pair slot (no candidate, or its frame is not stored)
    if (threadIdx.x == 0) g.N[p] = 0;
    return;
  }
  const int M = g.M[p];
  const double* qb = g.qb + (size_t)pd.q_slot * g.qF * 3;
  const double* mb = g.sb + (size_t)g.s_off[pd.m_frame] * 3;
  const uint16_t* iq = g.iq + (size_t)p * g.stride;
  const uint16_t* im = g.im + (size_t)p * g.stride;
  double* oa = g.a + (size_t)p * g.stride * 3;
  double* ob = g.b + (size_t)p * g.stride * 3;
  for (int i = threadIdx.x; i < M; i += 128) {
    const int a = iq[i], b = im[i];
    oa[3 * i + 0] = qb[3 * a + 0]; oa[3 * i + 1] = qb[3 * a + 1]; oa[3 * i + 2] = qb[3 * a + 2];
    ob[3 * i + 0] = mb[3 * b + 0]; ob[3 * i + 1] = mb[3 * b + 1]; ob[3 * i + 2] = mb[3 * b + 2];
  }
  if (threadIdx.x == 0) g.N[p] = M;
}

// recoverPose() prologue: keep mono inliers whose 3-D keypoints both have
// norm > 1e-3 (ordered compaction).  N3 = 0 when the mono gate failed.
__global__ void __launch_bounds__(128) gather_points_kernel(StereoGatherArgs sg) {
  const GatherArgs& g = sg.g;
  const int p = blockIdx.x;
  __shared__ int warp_cnt[4];
  __shared__ int base_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) base_s = 0;
  __syncthreads();
  const PairDesc pd = g.pairs[p];
  if (!sg.mono_ok[p] || pd.m_frame < 0) {
    if (tid == 0) g.N[p] = 0;
    return;
  }
  const int M = g.M[p];
  const double* qp = g.qp + (size_t)pd.q_slot * g.qF * 3;
  const double* mp = g.sp + (size_t)g.s_off[pd.m_frame] * 3;
  const uint16_t* iq = g.iq + (size_t)p * g.stride;
  const uint16_t* im = g.im + (size_t)p * g.stride;
  const uint32_t* mask = sg.mono_mask + (size_t)p * sg.mask_words;
  double* oa = g.a + (size_t)p * g.stride * 3;
  double* ob = g.b + (size_t)p * g.stride * 3;
  for (int i0 = 0; i0 < M; i0 += 128) {
    const int i = i0 + tid;
    bool keep = false;
    int qa = 0, mb = 0;
    double ax = 0, ay = 0, az = 0, bx = 0, by = 0, bz = 0;
    if (i < M && ((mask[i >> 5] >> (i & 31)) & 1u)) {
      qa = iq[i];
      mb = im[i];
      ax = qp[3 * qa]; ay = qp[3 * qa + 1]; az = qp[3 * qa + 2];
      bx = mp[3 * mb]; by = mp[3 * mb + 1]; bz = mp[3 * mb + 2];
      const double na = sqrt((ax * ax + ay * ay) + az * az);
      const double nb = sqrt((bx * bx + by * by) + bz * bz);
      keep = na > 1e-3 && nb > 1e-3;
    }
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, keep);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int off = base_s;
    for (int w = 0; w < warp; ++w) off += warp_cnt[w];
    if (keep) {
      const int pos = off + __popc(bal & ((1u << lane) - 1u));
      oa[3 * pos] = ax; oa[3 * pos + 1] = ay; oa[3 * pos + 2] = az;
      ob[3 * pos] = bx; ob[3 * pos + 1] = by; ob[3 * pos + 2] = bz;
      sg.kq[(size_t)p * g.stride + pos] = (uint16_t)qa;
      sg.km[(size_t)p * g.stride + pos] = (uint16_t)mb;
    }
    __syncthreads();
    if (tid == 0) base_s += warp_cnt[0] + warp_cnt[1] + warp_cnt[2] + warp_cnt[3];
    __syncthreads();
  }
  if (tid == 0) g.N[p] = base_s;
}

// acceptance gates of geometricVerificationNister / recoverPose
__global__ void mono_gate_kernel(FinalizeArgs f) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= f.P) return;
  const int inl = f.mono_inl[p];
  const int M = f.M[p];
  int ok = f.mono_st[p].best_draw >= 0;
  if (ok && inl < f.min_inliers) ok = 0;
  if (ok && (double)inl / (double)M < f.min_ratio_mono) ok = 0;
  f.mono_ok[p] = ok;
}
// One thread per pair slot: the acceptance gates of recoverPose and the pair's kml_result record
// (VLCEdge fields + the counters /root/reference/evaluation/lc_result.py:134-137 reads); the
// batch counters are reduced per warp and added to the BatchStats block.
__global__ void finalize_kernel(FinalizeArgs f) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long hm = 0, hs = 0, rm = 0, rs = 0, em = 0, es = 0, ok_m = 0;
  if (p < f.P && f.pairs[p].m_frame >= 0) {
    const int b = p / f.K, i = p - b * f.K;
    kml_result* r = f.recs + (size_t)b * f.cap + i;
    int status = 1;
    int mono_inl = 0, stereo_inl = 0;
    const SacState sm = f.mono_st[p];
    const SacState s3 = f.st3[p];
    const int N3 = f.N3[p];
    if (f.mono_ok[p]) {
      status = 2;
      mono_inl = f.mono_inl[p];
      for (int rr = 0; rr < 3; ++rr)
        for (int c = 0; c < 3; ++c) r->R_mono[3 * rr + c] = f.mono_model[(size_t)p * 12 + 4 * rr + c];
      const int inl3 = f.inl3[p];
      int ok = N3 >= 3 && s3.best_draw >= 0;
      if (ok && inl3 < f.min_inliers) ok = 0;
      if (ok && (double)inl3 / (double)N3 < f.min_ratio_stereo) ok = 0;
      if (ok) {
        status = 0;
        stereo_inl = inl3;
        for (int k = 0; k < 12; ++k) r->T[k] = f.model3[(size_t)p * 12 + k];
      }
      ok_m = 1;
    }
    r->n_matches = f.M[p];
    r->status = status;
    r->mono_inliers = mono_inl;
    r->stereo_inliers = stereo_inl;
    hm = (unsigned long long)sm.draws;
    rm = hm * (unsigned long long)f.M[p];
    em = (unsigned long long)sm.r_begin;
    if (f.mono_ok[p]) {  // the stereo problem of a pair that failed the mono gate was never set up
      hs = (unsigned long long)s3.draws;
      rs = hs * (unsigned long long)N3;
      es = (unsigned long long)s3.r_begin;
    }
  }
  unsigned long long v[7] = {hm, hs, rm, rs, em, es, ok_m};
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    unsigned long long x = v[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
    v[k] = x;
  }
  if ((threadIdx.x & 31) == 0) {
    if (v[0]) atomicAdd(&f.stats->hyp_m, v[0]);
    if (v[1]) atomicAdd(&f.stats->hyp_s, v[1]);
    if (v[2]) atomicAdd(&f.stats->res_m, v[2]);
    if (v[3]) atomicAdd(&f.stats->res_s, v[3]);
    if (v[4]) atomicAdd(&f.stats->eval_m, v[4]);
    if (v[5]) atomicAdd(&f.stats->eval_s, v[5]);
    if (v[6]) atomicAdd(&f.stats->mono_ok, v[6]);
  }
}

// --------------------------------------------------------------- launchers
static_assert(geom::kFrontOut == 70 && geom::kFrontOutStew == 130, "lcd.cu sizes the stage-1 buffer with these");
static size_t mono_smem() { return sizeof(double) * geom::kTphSlots * kFrontChunk; }
static size_t stereo_smem(int stride) { return sizeof(double) * (6 * (size_t)stride + 12 * kStereoChunk); }

template <class K>
static void ensure_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024)
    KML_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

void launch_sample_table(const uint32_t* raw, int cap_draws, int sample_size, int nmax, uint16_t* tab, cudaStream_t s) {
  const size_t sm = sac_perm_bytes(nmax) + sizeof(uint16_t) * kRoundCap * 8;
  if (sample_size == 8) {
    ensure_smem(sample_table_kernel<8>, sm);
    KML_LAUNCH((sample_table_kernel<8>), nmax + 1, 32, sm, s, raw, cap_draws, nmax, tab);
  } else if (sample_size == 3) {
    ensure_smem(sample_table_kernel<3>, sm);
    KML_LAUNCH((sample_table_kernel<3>), nmax + 1, 32, sm, s, raw, cap_draws, nmax, tab);
  } else {
    ensure_smem(sample_table_kernel<1>, sm);
    KML_LAUNCH((sample_table_kernel<1>), nmax + 1, 32, sm, s, raw, cap_draws, nmax, tab);
  }
}
void launch_sac_init(const SacArgs& a, int sample_size, cudaStream_t s) {
  if (a.P <= 0) return;
  KML_CUDA(cudaMemsetAsync(a.n_active, 0, 2 * sizeof(unsigned int), s));
  const size_t sm = sample_size == 8 ? sac_warp_smem<8>(a.stride) : sac_warp_smem<3>(a.stride);
  if (sample_size == 1) {
    ensure_smem(sac_init_kernel<1, kStereoChunk>, sm);
    ensure_smem(sac_replay_kernel<1, kStereoChunk>, sm);
    KML_LAUNCH((sac_init_kernel<1, kStereoChunk>), a.P, 32, sm, s, a);
  } else if (sample_size == 8) {
    ensure_smem(sac_init_kernel<8, kMonoChunk>, sm);
    ensure_smem(sac_replay_kernel<8, kMonoChunk>, sm);
    KML_LAUNCH((sac_init_kernel<8, kMonoChunk>), a.P, 32, sm, s, a);
  } else {
    ensure_smem(sac_init_kernel<3, kStereoChunk>, sm);
    ensure_smem(sac_replay_kernel<3, kStereoChunk>, sm);
    KML_LAUNCH((sac_init_kernel<3, kStereoChunk>), a.P, 32, sm, s, a);
  }
}

// CTAs along x for round r: the active list shrinks as problems finish, and so does the grid
static int active_grid(int P, int round) {
  const int cap = round <= 1 ? kNumSMs * 16 : round == 2 ? kNumSMs * 8 : round == 3 ? kNumSMs * 4 : kNumSMs * 2;
  return max(1, min(P, cap));
}
int launch_mono_round(const SacArgs& a, int round, cudaStream_t s) {
  if (a.P <= 0) return 0;
  const size_t sm = mono_smem();
  const int draws = sac_round_draws(round, a.first);
  const int blocks = (draws + kMonoChunk - 1) / kMonoChunk;
  const int gx = active_grid(a.P, round);
  const int fblocks = (draws + kFrontChunk - 1) / kFrontChunk;
  if (a.alg == 1) {
    ensure_smem(mono_front_kernel<1>, sm);
    KML_LAUNCH((mono_front_kernel<1>), dim3(gx, fblocks), kFrontChunk, sm, s, a, round);
  } else {
    ensure_smem(mono_front_kernel<0>, sm);
    KML_LAUNCH((mono_front_kernel<0>), dim3(gx, fblocks), kFrontChunk, sm, s, a, round);
  }
  KML_CUDA(cudaMemsetAsync(a.fb_count, 0, 2 * sizeof(unsigned int), s));  // fb_count, item_count
  KML_CUDA(cudaMemsetAsync(a.n_active + ((round + 1) & 1), 0, sizeof(unsigned int), s));  // the next round's list
  KML_LAUNCH((mono_isolate_kernel), dim3(gx, (draws + kIsoChunk - 1) / kIsoChunk), kIsoChunk, 0, s, a, round);
  KML_LAUNCH((mono_isolate_deferred_kernel), kNumSMs * 16, kMonoChunk, 0, s, a);
  if (a.alg == 1) KML_LAUNCH((mono_item_kernel<1>), kNumSMs * 16, kItemThreads, 0, s, a);
  else KML_LAUNCH((mono_item_kernel<0>), kNumSMs * 16, kItemThreads, 0, s, a);
  const size_t sm4 = sizeof(double) * 6 * (size_t)((a.stride + 31) & ~31);
  if (sm4 <= 96 * 1024) {
    ensure_smem(mono_count_kernel<true>, sm4);
    KML_LAUNCH((mono_count_kernel<true>), dim3(gx, blocks), kCountThreads, sm4, s, a, round);
  } else {
    KML_LAUNCH((mono_count_kernel<false>), dim3(gx, blocks), kCountThreads, 0, s, a, round);
  }
#ifdef KML_FILTER_STATS
  if (round == a.n_rounds - 1) fstats_print_kernel<<<1, 1, 0, s>>>();
#endif
  KML_LAUNCH((sac_replay_kernel<8, kMonoChunk>), gx, 32, sac_warp_smem<8>(a.stride), s, a, round);
  return 6;
}
int launch_stereo_round(const SacArgs& a, int round, cudaStream_t s) {
  if (a.P <= 0) return 0;
  const size_t sm = stereo_smem(a.stride);
  const int draws = sac_round_draws(round, a.first);
  const int blocks = (draws + kStereoChunk - 1) / kStereoChunk;
  const size_t sm0 = sizeof(double) * 12 * kStereoChunk;
  const int gx = active_grid(a.P, round);
  KML_CUDA(cudaMemsetAsync(a.n_active + ((round + 1) & 1), 0, sizeof(unsigned int), s));  // the next round's list
  if (a.onept) {
    if (sm <= 96 * 1024) {
      ensure_smem(stereo_chunk_kernel<true, true>, sm);
      KML_LAUNCH((stereo_chunk_kernel<true, true>), dim3(gx, blocks), kStereoThreads, sm, s, a, round);
    } else {
      KML_LAUNCH((stereo_chunk_kernel<false, true>), dim3(gx, blocks), kStereoThreads, sm0, s, a, round);
    }
    KML_LAUNCH((sac_replay_kernel<1, kStereoChunk>), gx, 32, sac_warp_smem<3>(a.stride), s, a, round);
    return 2;
  }
  if (sm <= 96 * 1024) {
    ensure_smem(stereo_chunk_kernel<true, false>, sm);
    KML_LAUNCH((stereo_chunk_kernel<true, false>), dim3(gx, blocks), kStereoThreads, sm, s, a, round);
  } else {
    KML_LAUNCH((stereo_chunk_kernel<false, false>), dim3(gx, blocks), kStereoThreads, sm0, s, a, round);
  }
  KML_LAUNCH((sac_replay_kernel<3, kStereoChunk>), gx, 32, sac_warp_smem<3>(a.stride), s, a, round);
  return 2;
}
__global__ void sac_pending_kernel(SacArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const bool pend = p < a.P && !a.st[p].done;
  const unsigned bal = __ballot_sync(0xFFFFFFFFu, pend);
  if ((threadIdx.x & 31) == 0 && bal) atomicAdd(a.pending, (unsigned)__popc(bal));
}
void launch_sac_pending(const SacArgs& a, cudaStream_t s) {
  if (a.P <= 0) return;
  KML_LAUNCH((sac_pending_kernel), (a.P + 127) / 128, 128, 0, s, a);
}
void launch_mono_select(const SacArgs& a, cudaStream_t s) {
  if (a.P <= 0) return;
  KML_LAUNCH((sac_select_kernel<true>), a.P, 128, 0, s, a);
}
void launch_stereo_select(const SacArgs& a, cudaStream_t s) {
  if (a.P <= 0) return;
  KML_LAUNCH((sac_select_kernel<false>), a.P, 128, 0, s, a);
}
void launch_gather_bearings(const GatherArgs& g, cudaStream_t s) {
  if (g.P <= 0) return;
  KML_LAUNCH((gather_bearings_kernel), g.P, 128, 0, s, g);
}
void launch_gather_points(const StereoGatherArgs& g, cudaStream_t s) {
  if (g.g.P <= 0) return;
  KML_LAUNCH((gather_points_kernel), g.g.P, 128, 0, s, g);
}
void launch_mono_gate(const FinalizeArgs& f, cudaStream_t s) {
  if (f.P <= 0) return;
  KML_LAUNCH((mono_gate_kernel), (f.P + 127) / 128, 128, 0, s, f);
}
void launch_finalize(const FinalizeArgs& f, cudaStream_t s) {
  if (f.P <= 0) return;
  KML_LAUNCH((finalize_kernel), (f.P + 127) / 128, 128, 0, s, f);
}

}  // namespace kml
