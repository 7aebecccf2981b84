// bow.cu — DBoW2 L1 inverted-file scoring + top-k on sm_100a.
//
// Replaces DBoW2::TemplatedDatabase::queryL1 and L1Scoring::score as used by
// LoopClosureDetector::detectLoopWithRobot (SURVEY.md A.1-A.3;
// /root/reference/images/kimera-multi.drawio:2571-2580, 2612, 2625; the
// vocabulary/database are created at
// /root/reference/docker/copy/kimera_multi_lcd.patch:36-38).
//
// One CTA per (query, robot database, entry tile):
//   1. stage the query's sparse (word, weight) vector in shared memory;
//   2. look up the CSR row of every query word, block-scan the row lengths;
//   3. stream the touched postings with a flattened (word, posting) index so
//      that consecutive threads read consecutive 8-byte postings of a row;
//   4. accumulate -(|q-d|-|q|-|d|)/2 per database entry into shared-memory
//      accumulators.  The sum is kept in 2^-62 fixed point (u64 atomics), so
//      it is the exactly rounded sum of the exact terms, independent of the
//      accumulation order (DBoW2 sums the same terms in ascending word order
//      in double; the two agree to ~1 ulp and bit-for-bit whenever the double
//      sum is exact, e.g. for float32 wire weights);
//   5. select the max_results best entries: 8-pass MSB radix select on the
//      accumulators, ties at the cut taken in ascending entry id, then a rank
//      sort by (score desc, entry asc).
// Bound: HBM/L2 gather bandwidth on the touched postings.
#include "common.cuh"
#include "kernels.h"

namespace kml {

constexpr int kBowThreads = 256;
constexpr int kWordsPerThread = kBowMaxWords / kBowThreads;  // 4

__device__ __forceinline__ unsigned long long bow_term_fx(double q, double d) {
  // DBoW2: value = fabs(q-d) - fabs(q) - fabs(d)  (<= 0); contribution to the
  // final score is -value/2, held as an integer multiple of 2^-62.
  const double value = fabs(q - d) - fabs(q) - fabs(d);
  const double s = -0.5 * value * kBowScale;
  return s > 0.0 ? (unsigned long long)s : 0ull;
}

__global__ void __launch_bounds__(kBowThreads) bow_score_kernel(BowArgs a) {
  KML_DYN_SMEM(unsigned char, smem_raw);
  unsigned long long* acc = reinterpret_cast<unsigned long long*>(smem_raw);
  __shared__ uint32_t s_ids[kBowMaxWords];
  __shared__ float s_vals[kBowMaxWords];
  __shared__ uint32_t s_row[kBowMaxWords];
  __shared__ uint32_t s_pre[kBowMaxWords + 1];
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_warp[kBowThreads / 32];
  __shared__ unsigned long long s_red[kBowThreads / 32];
  __shared__ unsigned long long s_sel_val[kBowMaxK];
  __shared__ uint32_t s_sel_ent[kBowMaxK];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_need, s_cnt, s_nz;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int bid = blockIdx.x;
  const int tile = bid % a.n_tiles;
  bid /= a.n_tiles;
  const int dbi = bid % a.n_db;
  const int b = bid / a.n_db;
  const BowDb db = a.dbs[dbi];
  const int64_t q0 = a.q_off[b];
  const int nq = min((int)(a.q_off[b + 1] - q0), kBowMaxWords);
  const uint32_t tile_lo = (uint32_t)tile * (uint32_t)a.tile_entries;
  const uint32_t tile_n =
      db.n_entries > tile_lo ? min((uint32_t)a.tile_entries, db.n_entries - tile_lo) : 0u;
  const int max_id = a.max_id ? a.max_id[dbi] : -1;

  for (int i = tid; i < nq; i += kBowThreads) {
    s_ids[i] = a.q_ids[q0 + i];
    s_vals[i] = a.q_vals[q0 + i];
  }
  for (uint32_t e = tid; e < tile_n; e += kBowThreads) acc[e] = 0ull;
  __syncthreads();

  // ---- NSS factor: L1 score of the query against the previous BoW vector
  if (a.nss != nullptr && dbi == 0 && tile == 0) {
    const int64_t p0 = a.p_off[b];
    const int np = (int)(a.p_off[b + 1] - p0);
    unsigned long long part = 0ull;
    for (int i = tid; i < nq; i += kBowThreads) {
      const uint32_t w = s_ids[i];
      int lo = 0, hi = np;  // lower_bound
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a.p_ids[p0 + mid] < w) lo = mid + 1; else hi = mid;
      }
      if (lo < np && a.p_ids[p0 + lo] == w)
        part += bow_term_fx((double)s_vals[i], (double)a.p_vals[p0 + lo]);
    }
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
    if (lane == 0) s_red[warp] = part;
    __syncthreads();
    if (tid == 0) {
      unsigned long long tot = 0ull;
      for (int w = 0; w < kBowThreads / 32; ++w) tot += s_red[w];
      a.nss[b] = (double)tot / kBowScale;
    }
  }

  // ---- row lookup + exclusive scan of the row lengths
  uint32_t len[kWordsPerThread];
  uint32_t tsum = 0;
#pragma unroll
  for (int u = 0; u < kWordsPerThread; ++u) {
    const int i = tid * kWordsPerThread + u;
    uint32_t l = 0;
    if (i < nq) {
      const uint32_t w = s_ids[i];
      if (w < db.W) {
        const uint32_t r0 = __ldg(db.row_ptr + w), r1 = __ldg(db.row_ptr + w + 1);
        s_row[i] = r0;
        l = r1 - r0;
      }
    }
    len[u] = l;
    tsum += l;
  }
  uint32_t incl = tsum;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t woff = 0;
  for (int w = 0; w < warp; ++w) woff += s_warp[w];
  uint32_t run = woff + incl - tsum;
#pragma unroll
  for (int u = 0; u < kWordsPerThread; ++u) {
    const int i = tid * kWordsPerThread + u;
    if (i < kBowMaxWords) s_pre[i] = run;
    run += len[u];
  }
  if (tid == kBowThreads - 1) s_pre[kBowMaxWords] = run;
  __syncthreads();
  const uint32_t T = s_pre[kBowMaxWords];
  if (tid == 0 && tile == 0 && a.postings_touched) atomicAdd(a.postings_touched, (unsigned long long)T);

  // ---- stream the touched postings (flattened index -> coalesced rows)
  for (uint32_t g = tid; g < T; g += kBowThreads) {
    int lo = 0, hi = nq;  // largest i with s_pre[i] <= g
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (s_pre[mid] <= g) lo = mid; else hi = mid;
    }
    const uint2 p = __ldg(db.postings + (size_t)s_row[lo] + (g - s_pre[lo]));
    const uint32_t e = p.x;
    if (e >= tile_lo && e - tile_lo < tile_n && (max_id < 0 || (int)e < max_id)) {
      const unsigned long long fx = bow_term_fx((double)s_vals[lo], (double)__uint_as_float(p.y));
      atomicAdd(&acc[e - tile_lo], fx);
    }
  }
  __syncthreads();

  // ---- top-K: count non-zero accumulators
  {
    uint32_t c = 0;
    for (uint32_t e = tid; e < tile_n; e += kBowThreads) c += (acc[e] != 0ull);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if (lane == 0) s_warp[warp] = c;
    __syncthreads();
    if (tid == 0) {
      uint32_t tot = 0;
      for (int w = 0; w < kBowThreads / 32; ++w) tot += s_warp[w];
      s_nz = (int)tot;
      s_need = min(a.K, (int)tot);
      s_prefix = 0ull;
      s_cnt = 0;
    }
    __syncthreads();
  }
  const int keff = min(a.K, s_nz);
  const size_t obase = ((size_t)(b * a.n_db + dbi) * a.n_tiles + tile);
  if (keff == 0) {
    if (tid == 0) a.out_count[obase] = 0;
    return;
  }
  // 8-pass MSB radix select of the keff-th largest accumulator value
  for (int pass = 7; pass >= 0; --pass) {
    s_hist[tid] = 0;
    __syncthreads();
    const unsigned long long pref = s_prefix;
    for (uint32_t e = tid; e < tile_n; e += kBowThreads) {
      const unsigned long long v = acc[e];
      if (v != 0ull && (pass == 7 || (v >> (8 * (pass + 1))) == pref))
        atomicAdd(&s_hist[(uint32_t)(v >> (8 * pass)) & 255u], 1u);
    }
    __syncthreads();
    if (warp == 0) {
      // bins 255..0, 8 per lane: lane 0 owns the highest 8 bins
      uint32_t h[8], lsum = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) { h[k] = s_hist[255 - (lane * 8 + k)]; lsum += h[k]; }
      uint32_t inc = lsum;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += v;
      }
      const uint32_t need = (uint32_t)s_need;
      uint32_t before = inc - lsum;  // items in strictly higher bins than this lane's
      const bool mine = before < need && inc >= need;
      __syncwarp();
      if (mine) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (before < need && before + h[k] >= need) {
            s_need = (int)(need - before);
            s_prefix = (pref << 8) | (unsigned long long)(255 - (lane * 8 + k));
            before = need;  // stop
          } else if (before < need) {
            before += h[k];
          }
        }
      }
    }
    __syncthreads();
  }
  const unsigned long long vk = s_prefix;  // value of the keff-th best
  const int need_eq = s_need;              // how many entries equal to vk to take
  const int n_gt = keff - need_eq;
  // entries strictly above the cut (unordered), then ties in ascending entry id
  for (uint32_t e = tid; e < tile_n; e += kBowThreads) {
    const unsigned long long v = acc[e];
    if (v > vk) {
      const int pos = atomicAdd(&s_cnt, 1);
      s_sel_val[pos] = v;
      s_sel_ent[pos] = e;
    }
  }
  __syncthreads();
  if (warp == 0) {
    int taken = 0;
    for (uint32_t e0 = 0; e0 < tile_n && taken < need_eq; e0 += 32) {
      const uint32_t e = e0 + lane;
      const bool hit = e < tile_n && acc[e] == vk;
      const unsigned bal = __ballot_sync(0xFFFFFFFFu, hit);
      const int pos = taken + __popc(bal & ((1u << lane) - 1u));
      if (hit && pos < need_eq) {
        s_sel_val[n_gt + pos] = vk;
        s_sel_ent[n_gt + pos] = e;
      }
      taken += __popc(bal);
    }
  }
  __syncthreads();
  // rank sort (score desc, entry asc)
  if (tid < keff) {
    const unsigned long long v = s_sel_val[tid];
    const uint32_t e = s_sel_ent[tid];
    int rank = 0;
    for (int j = 0; j < keff; ++j) {
      const unsigned long long vj = s_sel_val[j];
      const uint32_t ej = s_sel_ent[j];
      rank += (vj > v) || (vj == v && ej < e);
    }
    a.out_entry[obase * a.K + rank] = tile_lo + e;
    a.out_score[obase * a.K + rank] = (double)v / kBowScale;
  }
  if (tid == 0) a.out_count[obase] = keff;
}

void launch_bow(const BowArgs& a, cudaStream_t s) {
  const int grid = a.B * a.n_db * a.n_tiles;
  if (grid <= 0) return;
  const size_t smem = (size_t)a.tile_entries * sizeof(unsigned long long);
  static size_t configured = 0;
  if (smem > configured) {
    KML_CUDA(cudaFuncSetAttribute(bow_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    configured = smem;
  }
  KML_LAUNCH((bow_score_kernel), grid, kBowThreads, smem, s, a);
}

}  // namespace kml
