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
//   2. look up the row {start, len} of every query word (one 8-byte load per word);
//   3. stream the touched postings: a group of 8 lanes takes one row at a time and reads it with
//      128-bit loads (two postings per lane and load; rows start on 16-byte boundaries);
//   4. accumulate -(|q-d|-|q|-|d|)/2 per database entry into shared-memory
//      accumulators.  The sum is kept in 2^-62 fixed point (u64 atomics), so
//      it is the exactly rounded sum of the exact terms, independent of the
//      accumulation order (DBoW2 sums the same terms in ascending word order
//      in double; the two agree to ~1 ulp and bit-for-bit whenever the double
//      sum is exact, e.g. for float32 wire weights).  The thread whose add finds
//      the accumulator at zero appends the entry to the CTA's TOUCHED LIST, so
//      nothing later sweeps the dense tile;
//   5. select the max_results best entries among the touched ones: MSB radix select (8 bits per
//      pass, histogram scanned with warp shuffles) over the list, ties at the cut taken in
//      ascending entry id, then a rank sort by (score desc, entry asc);
// Bound: HBM/L2 gather of the touched rows.
#include "common.cuh"
#include "kernels.h"

namespace kml {

constexpr int kBowThreads = 256;      // databases whose tile leaves room for several CTAs per SM
constexpr int kBowThreadsWide = 512;  // wide tiles (> 96 KB of shared memory): one CTA per SM, twice the threads

__device__ __forceinline__ unsigned long long bow_term_fx(double q, double d) {
  // DBoW2: value = fabs(q-d) - fabs(q) - fabs(d)  (<= 0); contribution to the
  // final score is -value/2, held as an integer multiple of 2^-62.
  const double value = fabs(q - d) - fabs(q) - fabs(d);
  const double s = -0.5 * value * kBowScale;
  return s > 0.0 ? (unsigned long long)s : 0ull;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) bow_score_kernel(BowArgs a) {
  KML_DYN_SMEM(unsigned char, smem_raw);
  // dynamic: accumulators [tile_entries] u64 | touched list [tile_entries] u16
  unsigned long long* acc = reinterpret_cast<unsigned long long*>(smem_raw);
  uint16_t* touched = reinterpret_cast<uint16_t*>(smem_raw + (size_t)a.tile_entries * sizeof(unsigned long long));
  __shared__ uint32_t s_ids[kBowMaxWords];
  __shared__ float s_vals[kBowMaxWords];
  __shared__ uint2 s_row[kBowMaxWords];
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_pids[kBowMaxWords];  // previous BoW vector's word ids (NSS factor)
  __shared__ unsigned long long s_red[THREADS / 32];
  __shared__ unsigned long long s_bnd_val[32];  // boundary candidates of the early-exit select
  __shared__ uint32_t s_bnd_ent[32];
  __shared__ unsigned long long s_sel_val[kBowMaxK];
  __shared__ uint32_t s_sel_ent[kBowMaxK];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_need, s_cnt, s_ntouched, s_eq;

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

  // ---- stage the query, look up its rows, clear the accumulators
  uint32_t tsum = 0;
  for (int i = tid; i < nq; i += THREADS) {
    const uint32_t w = a.q_ids[q0 + i];
    s_ids[i] = w;
    s_vals[i] = a.q_vals[q0 + i];
    uint2 r = make_uint2(0u, 0u);
    if (w < db.W) r = __ldg(db.rows + w);
    s_row[i] = r;
    tsum += r.y;
  }
  for (uint32_t e = tid; e < tile_n; e += THREADS) acc[e] = 0ull;
  if (tid == 0) { s_ntouched = 0; s_cnt = 0; s_eq = 0; }
  __syncthreads();

  // ---- NSS factor: L1 score of the query against the previous BoW vector
  if (a.nss != nullptr && dbi == 0 && tile == 0) {
    const int64_t p0 = a.p_off[b];
    const int np = (int)(a.p_off[b + 1] - p0);
    const int nps = min(np, kBowMaxWords);
    for (int i = tid; i < nps; i += THREADS) s_pids[i] = a.p_ids[p0 + i];
    __syncthreads();
    unsigned long long part = 0ull;
    for (int i = tid; i < nq; i += THREADS) {
      const uint32_t w = s_ids[i];
      int lo = 0, hi = nps;  // lower_bound
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_pids[mid] < w) lo = mid + 1; else hi = mid;
      }
      if (lo < nps && s_pids[lo] == w)
        part += bow_term_fx((double)s_vals[i], (double)a.p_vals[p0 + lo]);
    }
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
    if (lane == 0) s_red[warp] = part;
    __syncthreads();
    if (tid == 0) {
      unsigned long long tot = 0ull;
      for (int w = 0; w < THREADS / 32; ++w) tot += s_red[w];
      a.nss[b] = (double)tot / kBowScale;
    }
  }
  if (tile == 0 && a.postings_touched) {  // algorithmic postings counter: sum of the touched rows' lengths
    for (int o = 16; o > 0; o >>= 1) tsum += __shfl_xor_sync(0xFFFFFFFFu, tsum, o);
    if (lane == 0 && tsum) atomicAdd(a.postings_touched, (unsigned long long)tsum);
  }

  // ---- stream the touched rows: 8 lanes per row, 128-bit loads (2 postings per lane); every
  // group keeps the first load of four rows in flight (rows are short: the loop is latency-bound)
  {
    constexpr int G = THREADS / 8;
    const int grp = tid >> 3, gl = tid & 7;
    const uint4* pool4 = reinterpret_cast<const uint4*>(db.postings);
    auto add2 = [&](const uint4& p, uint32_t k, uint32_t len, double qv) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t e = h ? p.z : p.x, wb = h ? p.w : p.y;
        if (k + h < len && e >= tile_lo && e - tile_lo < tile_n && (max_id < 0 || (int)e < max_id)) {
          const unsigned long long fx = bow_term_fx(qv, (double)__uint_as_float(wb));
          if (fx != 0ull) {
            const unsigned long long old = atomicAdd(&acc[e - tile_lo], fx);
            if (old == 0ull) touched[atomicAdd(&s_ntouched, 1)] = (uint16_t)(e - tile_lo);  // exactly one adder sees zero
          }
        }
      }
    };
    for (int i0 = grp; i0 < nq; i0 += 4 * G) {
      uint2 r[4];
      uint4 p[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * G;
        r[u] = i < nq ? s_row[i] : make_uint2(0u, 0u);
        p[u] = make_uint4(0u, 0u, 0u, 0u);
        if (2u * gl < r[u].y) p[u] = __ldg(pool4 + ((r[u].x + 2u * gl) >> 1));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * G;
        if (i >= nq) continue;
        const double qv = (double)s_vals[i];
        add2(p[u], 2u * gl, r[u].y, qv);
        for (uint32_t k = 2u * gl + 16u; k < r[u].y; k += 16u) add2(__ldg(pool4 + ((r[u].x + k) >> 1)), k, r[u].y, qv);
      }
    }
  }
  __syncthreads();

  // ---- top-K among the touched entries
  const int nt = s_ntouched;
  const int keff = min(a.K, nt);
  const size_t obase = ((size_t)(b * a.n_db + dbi) * a.n_tiles + tile);
  if (keff == 0) {
    if (tid == 0) a.out_count[obase] = 0;
    return;
  }
  // The keff best by (value desc, entry asc).  MSB radix select over the touched list, 8 bits per
  // pass, histogram scanned by warp 0 with shuffles.  After a pass every entry whose prefix is above
  // the selected bin is in, every entry below it is out, and only the bin itself is undecided: as
  // soon as the bin holds at most 32 entries the passes stop and one warp ranks those directly.
  // Otherwise (many equal values) all 8 passes run and the cut falls inside a group of equal values,
  // which is taken in ascending entry id.
  int shift = 64;          // bits of the value below the decided prefix
  int need_eq = keff;      // entries to take from the undecided set
  int n_bnd = 0;           // size of the undecided set when the passes stopped early
  if (nt > a.K) {
    if (tid == 0) { s_need = keff; s_prefix = 0ull; s_eq = nt; }
    for (int pass = 7; pass >= 0; --pass) {
      if (tid < 256) s_hist[tid] = 0;
      __syncthreads();
      if (s_eq <= 32) break;  // uniform: s_eq is only written between barriers
      const unsigned long long pref = s_prefix;
      for (int i = tid; i < nt; i += THREADS) {
        const unsigned long long v = acc[touched[i]];
        if (pass == 7 || (v >> (8 * (pass + 1))) == pref)
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
              s_eq = (int)h[k];  // entries in the undecided bin
              before = need;     // stop
            } else if (before < need) {
              before += h[k];
            }
          }
        }
      }
      shift = 8 * pass;
      __syncthreads();
    }
    need_eq = s_need;
    n_bnd = s_eq;
  }
  const unsigned long long pref = (nt > a.K) ? s_prefix : 0ull;
  const int n_gt = keff - need_eq;
  // decided entries (prefix above the bin; everything when nt <= K) in any order, the undecided
  // ones to the boundary buffer (early exit) — the rank sort below orders the output
  if (nt <= a.K) {
    for (int i = tid; i < nt; i += THREADS) {
      const uint32_t e = touched[i];
      s_sel_val[i] = acc[e];
      s_sel_ent[i] = e;
    }
  } else {
    if (tid == 0) s_ntouched = 0;  // reused as the boundary counter
    __syncthreads();
    for (int i = tid; i < nt; i += THREADS) {
      const uint32_t e = touched[i];
      const unsigned long long v = acc[e];
      const unsigned long long hi = shift >= 64 ? 0ull : (v >> shift);
      if (hi > pref) {
        const int pos = atomicAdd(&s_cnt, 1);
        s_sel_val[pos] = v;
        s_sel_ent[pos] = e;
      } else if (hi == pref && n_bnd <= 32) {
        const int pos = atomicAdd(&s_ntouched, 1);
        s_bnd_val[pos] = v;
        s_bnd_ent[pos] = e;
      }
    }
    __syncthreads();
    if (n_bnd <= 32) {
      // one warp ranks the undecided entries by (value desc, entry asc) and takes the first need_eq
      if (warp == 0) {
        const bool have = lane < n_bnd;
        const unsigned long long v = have ? s_bnd_val[lane] : 0ull;
        const uint32_t e = have ? s_bnd_ent[lane] : 0xFFFFFFFFu;
        int rank = 0;
        for (int j = 0; j < n_bnd; ++j) {
          const unsigned long long vj = s_bnd_val[j];
          const uint32_t ej = s_bnd_ent[j];
          rank += (vj > v) || (vj == v && ej < e);
        }
        if (have && rank < need_eq) {
          s_sel_val[n_gt + rank] = v;
          s_sel_ent[n_gt + rank] = e;
        }
      }
    } else if (warp == 0) {
      // all 8 passes ran: the undecided entries all hold exactly the value `pref`, more than 32 of
      // them (rare).  Ties are taken in ascending entry id, found by a sweep of the dense
      // accumulators in entry order.
      int taken = 0;
      for (uint32_t e0 = 0; e0 < tile_n && taken < need_eq; e0 += 32) {
        const uint32_t e = e0 + lane;
        const bool hit = e < tile_n && acc[e] == pref;
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, hit);
        const int pos = taken + __popc(bal & ((1u << lane) - 1u));
        if (hit && pos < need_eq) {
          s_sel_val[n_gt + pos] = pref;
          s_sel_ent[n_gt + pos] = e;
        }
        taken += __popc(bal);
      }
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

// Applies one incremental update of an inverted file (bow_merge.h BowInvFile::plan_append): thread
// i copies relocated row i, writes new posting i and sets row-table entry i; the three target
// disjoint slots, so no ordering between them is needed.
__global__ void bow_append_kernel(uint2* rows, uint2* pool, const uint4* __restrict__ copies, int n_copies,
                                  const uint4* __restrict__ posts, int n_posts, const uint4* __restrict__ row_cmds,
                                  int n_rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_copies) {
    const uint4 c = copies[i];  // {src, dst, n}
    for (uint32_t k = 0; k < c.z; ++k) pool[c.y + k] = pool[c.x + k];
  }
  if (i < n_posts) {
    const uint4 p = posts[i];   // {dst, entry, weight bits}
    pool[p.x] = make_uint2(p.y, p.z);
  }
  if (i < n_rows) {
    const uint4 r = row_cmds[i];  // {row, start, len}
    rows[r.x] = make_uint2(r.y, r.z);
  }
}

void launch_bow(const BowArgs& a, cudaStream_t s) {
  const int grid = a.B * a.n_db * a.n_tiles;
  if (grid <= 0) return;
  const size_t smem = (size_t)a.tile_entries * (sizeof(unsigned long long) + sizeof(uint16_t));
  // the opt-in is per device and costs microseconds: set it on every launch that needs it instead of
  // caching it in a process-wide static (a second GPU in the same process would never get it)
  if (smem > 96 * 1024) {
    KML_CUDA(cudaFuncSetAttribute(bow_score_kernel<kBowThreadsWide>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KML_LAUNCH((bow_score_kernel<kBowThreadsWide>), grid, kBowThreadsWide, smem, s, a);
  } else {
    if (smem > 48 * 1024)
      KML_CUDA(cudaFuncSetAttribute(bow_score_kernel<kBowThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KML_LAUNCH((bow_score_kernel<kBowThreads>), grid, kBowThreads, smem, s, a);
  }
}

void launch_bow_append(uint2* rows, uint2* pool, const uint4* copies, int n_copies, const uint4* posts, int n_posts,
                       const uint4* row_cmds, int n_rows, cudaStream_t s) {
  const int n = max(n_copies, max(n_posts, n_rows));
  if (n <= 0) return;
  KML_LAUNCH((bow_append_kernel), (n + 255) / 256, 256, 0, s, rows, pool, copies, n_copies, posts, n_posts, row_cmds, n_rows);
}

}  // namespace kml
