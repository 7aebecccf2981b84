// select.cu — the control steps of detectLoop / detectLoopWithRobot on the device, and the merge
// of the robot-sharded query.
//
// LoopClosureDetector::detectLoop loops detectLoopWithRobot over every robot database
// (/root/reference/images/kimera-multi.drawio:2571-2580, 2612, 2625):
//   db->query(bow, max_db_results)            -> bow_score_kernel (bow.cu), one list per entry tile
//   cut at the first Score < alpha * nss      -> here
//   inter_robot_only, |pose_q - pose| < dist_local window (intra-robot)   -> here
//   score / nss, (pose, score) appended to the caller's candidate list    -> here
// kimera_distributed::detectLoop then queues the candidates for verifyLoopSpin (drawio:2638-2654);
// this build verifies the top_k_verify best by (score desc, robot asc, pose asc) per query.
// select_candidates_kernel does all of that for one query per CTA and writes the query's
// kml_result records, the candidate pair descriptors and the matcher's job list directly, so the
// batch runs from the BoW scan to the last RANSAC round without a host round trip.
//
// merge_shards_kernel is the tail of the sharded query: after the all-gather of
// the per-rank record blocks, one warp per query keeps the best `cap` of the union.
#include "common.cuh"
#include "kernels.h"

namespace kml {

constexpr int kSelThreads = 128;

// candidates being ranked, structure of arrays in shared memory
struct SelBuf {
  double sc[2 * kSelMaxK];
  uint64_t k1[2 * kSelMaxK], k2[2 * kSelMaxK];
  int32_t pl[2 * kSelMaxK];
};

// Keeps the best `keep` of X[0, n) by (sc desc, k1 asc, k2 asc) — a total order: (k1, k2) is
// unique — in rank order at X[0, min(n, keep)).  Y is scratch.
__device__ void rank_keep(SelBuf& X, SelBuf& Y, int n, int keep) {
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += kSelThreads) {
    const double s = X.sc[i];
    const uint64_t a = X.k1[i], b = X.k2[i];
    int r = 0;
    for (int j = 0; j < n; ++j) {
      const double sj = X.sc[j];
      const uint64_t aj = X.k1[j], bj = X.k2[j];
      r += (sj > s) || (sj == s && (aj < a || (aj == a && bj < b)));
    }
    if (r < keep) {
      Y.sc[r] = s; Y.k1[r] = a; Y.k2[r] = b; Y.pl[r] = X.pl[i];
    }
  }
  __syncthreads();
  const int m = min(n, keep);
  for (int i = threadIdx.x; i < m; i += kSelThreads) {
    X.sc[i] = Y.sc[i]; X.k1[i] = Y.k1[i]; X.k2[i] = Y.k2[i]; X.pl[i] = Y.pl[i];
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kSelThreads) select_candidates_kernel(SelectArgs a) {
  __shared__ SelBuf G, W, Y;  // G: best K candidates so far; W: one database's result list; Y: scratch
  __shared__ int s_n, s_active;
  const int b = blockIdx.x, tid = threadIdx.x;
  const uint64_t qr = a.q_robot[b], qp = a.q_pose[b];
  const double nss = a.nss[b];
  int nG = 0;
  unsigned survivors = 0;
  if (tid == 0) s_active = 0;
  if (!(nss < a.min_nss)) {  // detectLoopWithRobot: false if nss < min_nss_factor
    const double cut = a.alpha * nss;
    for (int d = 0; d < a.n_db; ++d) {
      const BowDb db = a.dbs[d];
      if (a.inter_robot_only && qr == db.robot) continue;
      // TemplatedDatabase::query(bow, max_db_results): the best Kdb of the union of the entry tiles'
      // lists by (score desc, entry asc); a single tile's list is that set already
      int nL = 0;
      for (int t = 0; t < a.n_tiles; ++t) {
        const size_t l = ((size_t)b * a.n_db + d) * a.n_tiles + t;
        const int c = min(a.bow_count[l], a.Kdb);
        for (int i = tid; i < c; i += kSelThreads) {
          W.sc[nL + i] = a.bow_score[l * a.Kdb + i];
          W.k1[nL + i] = a.bow_entry[l * a.Kdb + i];
          W.k2[nL + i] = 0;
          W.pl[nL + i] = 0;
        }
        if (t > 0) {
          rank_keep(W, Y, nL + c, a.Kdb);
          nL = min(nL + c, a.Kdb);
        } else {
          nL = c;
        }
      }
      if (tid == 0) s_n = nG;
      __syncthreads();
      for (int i = tid; i < nL; i += kSelThreads) {
        const double s = W.sc[i];
        if (s < cut) continue;  // lower_bound(Result::geq) + resize on a best-first list
        const uint32_t e = (uint32_t)W.k1[i];
        const uint64_t pose = db.entry_pose[e];
        if (qr == db.robot) {
          const uint64_t dd = qp > pose ? qp - pose : pose - qp;
          if (dd < (uint64_t)a.dist_local) continue;
        }
        const int pos = atomicAdd(&s_n, 1);
        G.sc[pos] = s / nss;
        G.k1[pos] = db.robot;
        G.k2[pos] = pose;
        G.pl[pos] = db.entry_frame[e];
      }
      __syncthreads();
      const int n = s_n;
      survivors += (unsigned)(n - nG);
      if (n > a.K) {
        rank_keep(G, Y, n, a.K);
        nG = a.K;
      } else {
        nG = n;
      }
    }
    rank_keep(G, Y, nG, a.K);
  }
  __syncthreads();
  const int nv = min(nG, a.K);
  for (int i = tid; i < a.K; i += kSelThreads) {
    const int p = b * a.K + i;
    PairDesc pd;
    pd.q_slot = b;
    pd.m_frame = -1;
    HamJob job;
    job.q = nullptr; job.t = nullptr; job.nq = 0; job.nt = 0;
    job.keys = a.keys + (size_t)p * a.key_stride * 2;
    int nq = 0;
    if (i < nv) {
      kml_result* r = a.recs + (size_t)b * a.cap + i;  // zeroed by the host before the launch
      r->q_robot = qr; r->q_pose = qp;
      r->m_robot = G.k1[i]; r->m_pose = G.k2[i];
      r->norm_bow_score = G.sc[i];
      const int fr = G.pl[i];
      if (fr >= 0) {
        pd.m_frame = fr;
        nq = a.qF;
        job.q = a.q_desc + (size_t)b * a.qF * 32;
        job.nq = a.qF;
        job.t = a.s_desc + (size_t)a.s_off[fr] * 32;
        job.nt = a.s_F[fr];
        r->status = 1;
        atomicAdd(&s_active, 1);
      } else {
        r->status = 3;  // frameExists() false: the candidate waits for its VLC frame
      }
    }
    a.pairs[p] = pd;
    a.jobs[p] = job;
    a.nq[p] = nq;
  }
  __syncthreads();
  if (tid == 0) {
    a.counts[b] = nv;
    if (survivors) atomicAdd(&a.stats->survivors, (unsigned long long)survivors);
    if (s_active) atomicAdd(&a.stats->pairs, (unsigned long long)s_active);
  }
}

// One warp per query.  Every rank's list is ranked by the same key, but the merge does not rely
// on it: each record's rank in the union is counted directly (<= nranks * cap_in records), exact
// ties of the key (the same keyframe held by two ranks) go to the lower rank, then position.
__global__ void __launch_bounds__(128) merge_shards_kernel(MergeArgs a) {
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= a.B) return;
  int total = 0;
  for (int r = 0; r < a.nranks; ++r) {
    const int32_t* cnt = reinterpret_cast<const int32_t*>(a.base + a.blk_stride * r + a.counts_off);
    total += min(max(cnt[b], 0), a.cap_in);
  }
  for (int i = lane; i < a.nranks * a.cap_in; i += 32) {
    const int r = i / a.cap_in, k = i - r * a.cap_in;
    const uint8_t* blk = a.base + a.blk_stride * r;
    const int c = min(max(reinterpret_cast<const int32_t*>(blk + a.counts_off)[b], 0), a.cap_in);
    if (k >= c) continue;
    const kml_result* x = reinterpret_cast<const kml_result*>(blk) + (size_t)b * a.cap_in + k;
    const double s = x->norm_bow_score;
    const uint64_t mr = x->m_robot, mp = x->m_pose;
    int rank = 0;
    for (int r2 = 0; r2 < a.nranks; ++r2) {
      const uint8_t* blk2 = a.base + a.blk_stride * r2;
      const int c2 = min(max(reinterpret_cast<const int32_t*>(blk2 + a.counts_off)[b], 0), a.cap_in);
      const kml_result* y = reinterpret_cast<const kml_result*>(blk2) + (size_t)b * a.cap_in;
      for (int k2 = 0; k2 < c2; ++k2) {
        const double s2 = y[k2].norm_bow_score;
        const uint64_t mr2 = y[k2].m_robot, mp2 = y[k2].m_pose;
        const bool before = (s2 > s) || (s2 == s && (mr2 < mr || (mr2 == mr && mp2 < mp)));
        const bool same = (s2 == s && mr2 == mr && mp2 == mp);
        rank += before || (same && (r2 < r || (r2 == r && k2 < k)));
      }
    }
    if (rank < a.cap) {
      const uint64_t* src = reinterpret_cast<const uint64_t*>(x);
      uint64_t* dst = reinterpret_cast<uint64_t*>(a.out + (size_t)b * a.cap + rank);
#pragma unroll
      for (int w = 0; w < (int)(sizeof(kml_result) / 8); ++w) dst[w] = src[w];
    }
  }
  if (lane == 0) a.counts[b] = min(total, a.cap);
  if (b == 0 && lane == 0) {
    int32_t e = 0;
    for (int r = 0; r < a.nranks; ++r)
      e = max(e, *reinterpret_cast<const int32_t*>(a.base + a.blk_stride * r + a.err_off));
    *a.err_out = e;
  }
}

__global__ void block_flags_kernel(BatchStats* st, const unsigned int* overflow, int32_t* flag) {
  const unsigned int ov = *overflow;
  st->item_overflow = ov;
  *flag = (ov || st->pending_m || st->pending_s) ? 1 : 0;
}

void launch_select(const SelectArgs& a, cudaStream_t s) {
  if (a.B <= 0) return;
  KML_LAUNCH((select_candidates_kernel), a.B, kSelThreads, 0, s, a);
}
void launch_block_flags(BatchStats* stats, const unsigned int* overflow, int32_t* flag, cudaStream_t s) {
  KML_LAUNCH((block_flags_kernel), 1, 1, 0, s, stats, overflow, flag);
}
void launch_merge(const MergeArgs& a, cudaStream_t s) {
  if (a.B <= 0) return;
  KML_LAUNCH((merge_shards_kernel), (a.B + 3) / 4, 128, 0, s, a);
}

}  // namespace kml
