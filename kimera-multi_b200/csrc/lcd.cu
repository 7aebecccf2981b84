// lcd.cu — host side of the LoopClosureDetector replacement: database and
// frame store residency in HBM, candidate selection (detectLoopWithRobot
// post-processing), and the device verification pipeline
// (computeMatchedIndices -> geometricVerificationNister -> recoverPose,
// SURVEY.md A.4-A.8; call order of kimera_distributed's verifyLoopSpin,
// /root/reference/images/kimera-multi.drawio:2638-2654).
#include <algorithm>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "handle.h"
#include "bow_merge.h"
#include "sac_host.h"

using namespace kml;

namespace kml {
int comm_nranks(const kml_handle* h);
int comm_rank(const kml_handle* h);
int comm_allgather(kml_handle* h, const void* d_send, void* d_recv, size_t bytes);
int comm_wait(kml_handle* h);
}  // namespace kml

#define KML_API_BEGIN(h)                              \
  if (!(h)) return KML_ERR_ARG;                       \
  try {                                               \
    KML_CUDA(cudaSetDevice((h)->device));
#define KML_API_END(h)                                \
  }                                                   \
  catch (const kml::CudaError& e) {                   \
    (h)->err = e.what();                              \
    cudaGetLastError();                               \
    return KML_ERR_CUDA;                              \
  }                                                   \
  catch (const std::exception& e) {                   \
    (h)->err = e.what();                              \
    return KML_ERR_ARG;                               \
  }

static int fail(kml_handle* h, int code, const char* msg) {
  h->err = msg;
  return code;
}

// ===================================================================== DB
static RobotDb* get_db(kml_handle* h, uint64_t robot, bool create) {
  auto it = h->sh->dbs.find(robot);
  if (it != h->sh->dbs.end()) return it->second.get();
  if (!create) return nullptr;
  auto db = std::unique_ptr<RobotDb>(new RobotDb());
  db->robot = robot;
  RobotDb* p = db.get();
  h->sh->dbs[robot] = std::move(db);
  return p;
}

static int add_bow_host(kml_handle* h, RobotDb* db, uint64_t pose, const uint32_t* ids,
                        const float* vals, int n) {
  if (n < 0 || (n > 0 && (!ids || !vals))) return fail(h, KML_ERR_ARG, "add_bow: bad argument");
  if (n > kBowMaxWords) return fail(h, KML_ERR_CAPACITY, "add_bow: more than 1024 words");
  for (int i = 1; i < n; ++i)
    if (ids[i] <= ids[i - 1]) return fail(h, KML_ERR_ARG, "add_bow: word ids must be strictly ascending");
  if (db->pose_to_entry.count(pose)) return KML_OK;  // addBowVector: already present -> no-op
  const uint32_t entry = db->n_entries();
  db->pose_to_entry[pose] = entry;
  db->entry_to_pose.push_back(pose);
  {  // frameExists(robot, pose) already?  then the entry's dense frame index is known now
    auto fit = h->sh->frames.find(RobotPoseId(db->robot, pose));
    db->entry_to_frame.push_back(fit == h->sh->frames.end() ? -1 : fit->second.index);
  }
  db->ids.insert(db->ids.end(), ids, ids + n);
  db->vals.insert(db->vals.end(), vals, vals + n);
  db->off.push_back((int64_t)db->ids.size());
  if (!db->dirty) {  // the inverted file exists: queue this vector's postings for the in-place append
    for (int i = 0; i < n; ++i) {
      uint32_t bits;
      memcpy(&bits, &vals[i], 4);
      db->pend_word.push_back(ids[i]);
      db->pend_entry.push_back(entry);
      db->pend_wbits.push_back(bits);
    }
    // a bulk load is cheaper as one counting sort than as in-place appends
    if (db->pend_word.size() > (size_t)(1u << 18) && db->pend_word.size() > db->inv.live / 2) db->dirty = true;
  }
  {
    std::lock_guard<std::mutex> lk(h->sh->mu);
    h->sh->version++;
  }
  return KML_OK;
}

// Full (re)build of a robot's inverted file from its insertion log: counting sort by word id (rows
// come out ascending in entry id because entries are visited in order), one upload.  Used for the
// first query, after a bulk load, and when an incremental append does not fit (word beyond the row
// table, pool full, more garbage than live postings).
static void rebuild_invfile(kml_handle* h, RobotDb* db) {
  static_assert(sizeof(BowPosting) == sizeof(uint2) && sizeof(BowRow) == sizeof(uint2), "inverted-file layout");
  std::vector<BowRow> rows;
  std::vector<BowPosting> pool;
  db->inv.build(db->off, db->ids, db->vals, db->n_entries(), &rows, &pool);
  if (db->rows.cap < rows.size() || db->postings.cap < db->inv.pool_cap) {
    db->rows.scratch(rows.size());
    db->postings.scratch((size_t)db->inv.pool_cap);
  } else {
    db->inv.pool_cap = db->postings.cap;  // keep the larger allocation
  }
  if (!rows.empty())
    KML_CUDA(cudaMemcpyAsync(db->rows.p, rows.data(), rows.size() * sizeof(uint2), cudaMemcpyHostToDevice, h->stream));
  if (!pool.empty())
    KML_CUDA(cudaMemcpyAsync(db->postings.p, pool.data(), pool.size() * sizeof(uint2), cudaMemcpyHostToDevice, h->stream));
  KML_CUDA(cudaStreamSynchronize(h->stream));
  db->pend_word.clear(); db->pend_entry.clear(); db->pend_wbits.clear();
  db->dirty = false;
}

// Appends the postings of the vectors added since the last query in place (BowInvFile::plan_append
// + bow_append_kernel): O(new postings), independent of the database size — the deployment pattern
// is one addBowVector per keyframe followed by a query
// (/root/reference/launch/kimera_vio_jackal.launch:13-15).
static void flush_appends(kml_handle* h, RobotDb* db) {
  const size_t n = db->pend_word.size();
  if (n == 0) return;
  // (word, entry) order: stable sort by word keeps the entries of a word ascending
  std::vector<uint32_t> order(n);
  for (size_t i = 0; i < n; ++i) order[i] = (uint32_t)i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return db->pend_word[x] < db->pend_word[y]; });
  std::vector<uint32_t> w(n), e(n), b(n);
  for (size_t i = 0; i < n; ++i) { w[i] = db->pend_word[order[i]]; e[i] = db->pend_entry[order[i]]; b[i] = db->pend_wbits[order[i]]; }
  BowUpdate up;
  if (!db->inv.plan_append(w.data(), e.data(), b.data(), n, &up)) {
    rebuild_invfile(h, db);
    return;
  }
  const size_t nc = up.copies.size(), np = up.posts.size(), nr = up.rows.size();
  db->d_cmds.scratch(nc + np + nr);
  cudaStream_t s = h->stream;
  if (nc) KML_CUDA(cudaMemcpyAsync(db->d_cmds.p, up.copies.data(), nc * sizeof(uint4), cudaMemcpyHostToDevice, s));
  KML_CUDA(cudaMemcpyAsync(db->d_cmds.p + nc, up.posts.data(), np * sizeof(uint4), cudaMemcpyHostToDevice, s));
  KML_CUDA(cudaMemcpyAsync(db->d_cmds.p + nc + np, up.rows.data(), nr * sizeof(uint4), cudaMemcpyHostToDevice, s));
  launch_bow_append(db->rows.p, db->postings.p, db->d_cmds.p, (int)nc, db->d_cmds.p + nc, (int)np, db->d_cmds.p + nc + np, (int)nr, s);
  h->stats.kernel_launches += 1;
  KML_CUDA(cudaGetLastError());
  KML_CUDA(cudaStreamSynchronize(s));  // `up` leaves scope; queries of other lanes follow
  db->pend_word.clear(); db->pend_entry.clear(); db->pend_wbits.clear();
}

static void sync_db_maps(kml_handle* h, RobotDb* db);
static BowDb db_view(kml_handle* h, RobotDb* db) {
  {
    std::lock_guard<std::mutex> lk(h->sh->mu);
    if (db->n_entries() > 0) {
      if (db->dirty) rebuild_invfile(h, db);
      else flush_appends(h, db);
    }
    sync_db_maps(h, db);
  }
  BowDb v;
  v.rows = db->rows.p;
  v.postings = db->postings.p;
  v.W = db->n_entries() > 0 ? db->inv.W : 0u;
  v.n_entries = db->n_entries();
  v.entry_pose = db->d_entry_pose.p;
  v.entry_frame = db->d_entry_frame.p;
  v.robot = db->robot;
  return v;
}

// Host-visible result of one BoW launch
struct BowOut {
  int B = 0, n_db = 0, K = 0;
  std::vector<uint32_t> entry;  // [B][n_db][K]
  std::vector<double> score;
  std::vector<int32_t> count;   // [B][n_db]
  std::vector<double> nss;      // [B]
};

// Score B query vectors (device CSR form) against the given databases.
static void run_bow(kml_handle* h, const std::vector<RobotDb*>& dbs, int B, const int64_t* d_qoff,
                    const uint32_t* d_qids, const float* d_qvals, const int64_t* d_poff,
                    const uint32_t* d_pids, const float* d_pvals, int K,
                    const std::vector<int32_t>* max_id, BowOut* out) {
  const int n_db = (int)dbs.size();
  out->B = B; out->n_db = n_db; out->K = K;
  out->entry.assign((size_t)B * n_db * K, 0);
  out->score.assign((size_t)B * n_db * K, 0.0);
  out->count.assign((size_t)B * n_db, 0);
  out->nss.assign(B, 0.0);
  if (B <= 0 || n_db <= 0) return;
  std::vector<BowDb> views(n_db);
  uint32_t max_entries = 1;
  for (int i = 0; i < n_db; ++i) {
    views[i] = db_view(h, dbs[i]);
    max_entries = std::max(max_entries, views[i].n_entries);
  }
  int tile = 0, n_tiles = 0;
  bow_tiling(max_entries, &tile, &n_tiles);
  h->d_dbs.scratch(n_db);
  KML_CUDA(cudaMemcpyAsync(h->d_dbs.p, views.data(), sizeof(BowDb) * n_db, cudaMemcpyHostToDevice, h->stream));
  const int32_t* d_maxid = nullptr;
  if (max_id) {
    h->d_maxid.scratch(n_db);
    KML_CUDA(cudaMemcpyAsync(h->d_maxid.p, max_id->data(), sizeof(int32_t) * n_db, cudaMemcpyHostToDevice, h->stream));
    d_maxid = h->d_maxid.p;
  }
  const size_t nlist = (size_t)B * n_db * n_tiles;
  h->d_bow_entry.scratch(nlist * K);
  h->d_bow_score.scratch(nlist * K);
  h->d_bow_count.scratch(nlist);
  h->d_nss.scratch(B);
  h->d_postings.scratch(1);
  KML_CUDA(cudaMemsetAsync(h->d_postings.p, 0, sizeof(unsigned long long), h->stream));
  KML_CUDA(cudaMemsetAsync(h->d_nss.p, 0, sizeof(double) * B, h->stream));
  BowArgs a;
  a.dbs = h->d_dbs.p; a.n_db = n_db; a.B = B;
  a.q_off = d_qoff; a.q_ids = d_qids; a.q_vals = d_qvals;
  a.p_off = d_poff; a.p_ids = d_pids; a.p_vals = d_pvals;
  a.K = K; a.max_id = d_maxid; a.tile_entries = tile; a.n_tiles = n_tiles;
  a.out_entry = h->d_bow_entry.p; a.out_score = h->d_bow_score.p; a.out_count = h->d_bow_count.p;
  a.nss = d_poff ? h->d_nss.p : nullptr;
  a.postings_touched = h->d_postings.p;
  KML_CUDA(cudaEventRecord(h->ev[0], h->stream));
  launch_bow(a, h->stream);
  KML_CUDA(cudaGetLastError());
  KML_CUDA(cudaEventRecord(h->ev[1], h->stream));
  h->stats.kernel_launches += 1;
  h->h_bow_entry.scratch(nlist * K);
  h->h_bow_score.scratch(nlist * K);
  h->h_bow_count.scratch(nlist);
  h->h_nss.scratch(B + 1);
  KML_CUDA(cudaMemcpyAsync(h->h_bow_entry.p, h->d_bow_entry.p, nlist * K * 4, cudaMemcpyDeviceToHost, h->stream));
  KML_CUDA(cudaMemcpyAsync(h->h_bow_score.p, h->d_bow_score.p, nlist * K * 8, cudaMemcpyDeviceToHost, h->stream));
  KML_CUDA(cudaMemcpyAsync(h->h_bow_count.p, h->d_bow_count.p, nlist * 4, cudaMemcpyDeviceToHost, h->stream));
  KML_CUDA(cudaMemcpyAsync(h->h_nss.p, h->d_nss.p, B * 8, cudaMemcpyDeviceToHost, h->stream));
  unsigned long long postings = 0;
  KML_CUDA(cudaMemcpyAsync(&postings, h->d_postings.p, 8, cudaMemcpyDeviceToHost, h->stream));
  h->wait_stream();
  KML_CUDA(cudaEventElapsedTime(&h->stats.ms_bow, h->ev[0], h->ev[1]));
  h->stats.bow_postings_last = postings;
  for (int b = 0; b < B; ++b) out->nss[b] = h->h_nss.p[b];
  // merge entry tiles (each tile's list is already sorted best-first)
  merge_bow_tiles(h->h_bow_entry.p, h->h_bow_score.p, h->h_bow_count.p, B, n_db, n_tiles, K,
                  out->entry.data(), out->score.data(), out->count.data());
}

// upload one host BoW vector (and optionally a second one) in CSR form
static void upload_single_bow(kml_handle* h, const uint32_t* ids, const float* vals, int n,
                              const uint32_t* pids, const float* pvals, int pn) {
  int64_t off[2] = {0, n}, poff[2] = {0, pn};
  h->d_qoff.scratch(2); h->d_poff.scratch(2);
  h->d_qids.scratch(std::max(n, 1)); h->d_qvals.scratch(std::max(n, 1));
  h->d_pids.scratch(std::max(pn, 1)); h->d_pvals.scratch(std::max(pn, 1));
  KML_CUDA(cudaMemcpyAsync(h->d_qoff.p, off, 16, cudaMemcpyHostToDevice, h->stream));
  KML_CUDA(cudaMemcpyAsync(h->d_poff.p, poff, 16, cudaMemcpyHostToDevice, h->stream));
  if (n) {
    KML_CUDA(cudaMemcpyAsync(h->d_qids.p, ids, 4 * n, cudaMemcpyHostToDevice, h->stream));
    KML_CUDA(cudaMemcpyAsync(h->d_qvals.p, vals, 4 * n, cudaMemcpyHostToDevice, h->stream));
  }
  if (pn) {
    KML_CUDA(cudaMemcpyAsync(h->d_pids.p, pids, 4 * pn, cudaMemcpyHostToDevice, h->stream));
    KML_CUDA(cudaMemcpyAsync(h->d_pvals.p, pvals, 4 * pn, cudaMemcpyHostToDevice, h->stream));
  }
  KML_CUDA(cudaStreamSynchronize(h->stream));  // host arrays `off` go out of scope
}

static int check_bow(kml_handle* h, const uint32_t* ids, const float* vals, int n) {
  if (n < 0 || (n > 0 && (!ids || !vals))) return fail(h, KML_ERR_ARG, "bad BoW vector");
  if (n > kBowMaxWords) return fail(h, KML_ERR_CAPACITY, "BoW vector has more than 1024 words");
  for (int i = 1; i < n; ++i)
    if (ids[i] <= ids[i - 1]) return fail(h, KML_ERR_ARG, "BoW word ids must be strictly ascending");
  return KML_OK;
}

// detectLoopWithRobot steps 3-7 on one (query, db) result list
struct Cand {
  double score;
  uint64_t robot, pose;
  RobotDb* db;
  uint32_t entry;
};
static void select_candidates(const kml_handle* h, const RobotDb* db, uint64_t q_robot,
                              uint64_t q_pose, double nss, const uint32_t* entry,
                              const double* score, int count, std::vector<Cand>* out) {
  const kml_params& P = h->prm;
  if (P.inter_robot_only && q_robot == db->robot) return;
  if (nss < P.min_nss_factor) return;
  const double cut = P.alpha * nss;
  for (int i = 0; i < count; ++i) {
    if (score[i] < cut) break;  // list is best-first: lower_bound(Result::geq) + resize
    const uint64_t pose = db->entry_to_pose[entry[i]];
    if (q_robot == db->robot) {
      const uint64_t d = q_pose > pose ? q_pose - pose : pose - q_pose;
      if (d < (uint64_t)P.dist_local) continue;
    }
    out->push_back({score[i] / nss, db->robot, pose, const_cast<RobotDb*>(db), entry[i]});
  }
}

// ============================================================ frame store
static int add_frames_host(kml_handle* h, uint64_t robot, const uint64_t* poses, int count,
                           const uint8_t* desc, const double* bearings, const double* points,
                           int F) {
  if (count < 0 || F < 0 || F > 65535 || (count > 0 && F > 0 && (!desc || !bearings || !points)) ||
      (count > 0 && !poses))
    return fail(h, KML_ERR_ARG, "add_frame: bad argument (F must be < 65536)");
  const size_t total = (size_t)count * F;
  const int64_t base = h->sh->n_feat;
  h->sh->s_desc.reserve((size_t)(base + total) * 32 + 32, (size_t)base * 32, h->stream);
  h->sh->s_bear.reserve((size_t)(base + total) * 3 + 3, (size_t)base * 3, h->stream);
  h->sh->s_pts.reserve((size_t)(base + total) * 3 + 3, (size_t)base * 3, h->stream);
  if (total) {
    KML_CUDA(cudaMemcpyAsync(h->sh->s_desc.p + (size_t)base * 32, desc, total * 32, cudaMemcpyHostToDevice, h->stream));
    KML_CUDA(cudaMemcpyAsync(h->sh->s_bear.p + (size_t)base * 3, bearings, total * 24, cudaMemcpyHostToDevice, h->stream));
    KML_CUDA(cudaMemcpyAsync(h->sh->s_pts.p + (size_t)base * 3, points, total * 24, cudaMemcpyHostToDevice, h->stream));
  }
  for (int i = 0; i < count; ++i) {
    const RobotPoseId id(robot, poses[i]);
    FrameRec rec;
    rec.feat_off = base + (int64_t)i * F;
    rec.F = F;
    rec.index = (int32_t)h->sh->frame_off_h.size();
    h->sh->frame_off_h.push_back(rec.feat_off);
    h->sh->frame_F_h.push_back(F);
    h->sh->frames[id] = rec;  // vlc_frames_[id] = frame (overwrite keeps the newest copy)
    // the BoW entry of this keyframe (if it has one) now points at the stored frame
    auto dit = h->sh->dbs.find(robot);
    if (dit != h->sh->dbs.end()) {
      auto eit = dit->second->pose_to_entry.find(poses[i]);
      if (eit != dit->second->pose_to_entry.end()) {
        dit->second->entry_to_frame[eit->second] = rec.index;
        if (eit->second < dit->second->synced) dit->second->frame_patches.push_back(eit->second);
      }
    }
  }
  h->sh->n_feat = base + (int64_t)total;
  {
    std::lock_guard<std::mutex> lk(h->sh->mu);
    h->sh->version++;
  }
  KML_CUDA(cudaStreamSynchronize(h->stream));  // caller may free its buffers on return
  return KML_OK;
}

// ====================================================== RANSAC host helpers
// k as a function of (N, best inlier count): the exact expressions of
// opengv::sac::Ransac::computeModel evaluated with the host libm, so that the
// device replay takes the same branches as a CPU run.
static void ensure_ktable(kml_handle* h, int Nmax, int sample_size, double prob, DevBuf<double>* buf,
                          int* cur_n) {
  std::lock_guard<std::mutex> lk(h->sh->mu);
  const int need = Nmax + 1;
  if (*cur_n >= need) return;
  int n1 = std::max(need, 512);
  std::vector<double> tab;
  fill_ktable(n1, sample_size, prob, &tab);
  if (buf->p) {  // another lane's kernels may still read the old table: retire it instead of freeing
    h->sh->retired.push_back(buf->p);
    buf->p = nullptr;
    buf->cap = 0;
  }
  buf->scratch(tab.size());
  KML_CUDA(cudaMemcpyAsync(buf->p, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, h->stream));
  KML_CUDA(cudaStreamSynchronize(h->stream));
  *cur_n = n1;
}

// sample_tab[N][cap_draws][S] for N <= nmax (SacArgs::sample_tab), built on first use; nullptr
// for problems beyond kSampleTabMaxN correspondences (the per-problem sampler runs instead)
static const uint16_t* ensure_sample_table(kml_handle* h, int S, int nmax, int cap_draws, int* tab_nmax) {
  if (nmax > kSampleTabMaxN || getenv("KML_NO_SAMPLE_TABLE")) return nullptr;
  const int slot = S == 8 ? 0 : S == 3 ? 1 : 2;
  std::lock_guard<std::mutex> lk(h->sh->mu);
  kml_shared& sh = *h->sh;
  if (sh.samptab_nmax[slot] < nmax) {
    const int n1 = std::min(kSampleTabMaxN, std::max(nmax, 512));
    if (sh.d_samptab[slot].p) {  // another lane's kernels may still read the old table
      sh.retired.push_back(sh.d_samptab[slot].p);
      sh.d_samptab[slot].p = nullptr;
      sh.d_samptab[slot].cap = 0;
    }
    sh.d_samptab[slot].scratch((size_t)(n1 + 1) * cap_draws * S);
    launch_sample_table(sh.d_raw.p, cap_draws, S, n1, sh.d_samptab[slot].p, h->stream);
    h->stats.kernel_launches += 1;
    KML_CUDA(cudaGetLastError());
    KML_CUDA(cudaStreamSynchronize(h->stream));
    sh.samptab_nmax[slot] = n1;
  }
  *tab_nmax = sh.samptab_nmax[slot];
  return sh.d_samptab[slot].p;
}

struct SacBufs {
  DevBuf<SacState>* st;
  DevBuf<double>* best;
  DevBuf<uint32_t>* mask;
  DevBuf<int32_t>* inl;
};

// Buffers and arguments of one batched RANSAC (mono: S=8, stereo: S=3) over P problems whose
// correspondences are already gathered in d_a/d_b.  `pending` (nullable) receives the number of
// problems whose loop has not ended when the select kernel runs.
static SacArgs prepare_sac(kml_handle* h, bool mono, int P, const double* d_a, const double* d_b,
                           const int32_t* d_N, int stride, int full, SacBufs out, int mask_words,
                           unsigned int* pending, const double* d_prior = nullptr) {
  const size_t Pa = (size_t)std::max(P, 1);
  const kml_params& prm = h->prm;
  const bool onept = !mono && d_prior != nullptr;  // row f4: rotation given, 1-point samples
  const int S = mono ? 8 : (onept ? 1 : 3);
  const int max_it = mono ? prm.max_ransac_iterations_mono : prm.max_ransac_iterations;
  if (mono)
    ensure_ktable(h, stride, 8, prm.ransac_probability_mono, &h->sh->d_ktable_mono, &h->sh->ktable_n_mono);
  else if (onept)
    ensure_ktable(h, stride, 1, prm.ransac_probability, &h->sh->d_ktable_stereo1, &h->sh->ktable_n_stereo1);
  else
    ensure_ktable(h, stride, 3, prm.ransac_probability, &h->sh->d_ktable_stereo, &h->sh->ktable_n_stereo);
  const int raw_len = (int)h->sh->raw_h.size();
  const int cap_draws = sac_cap_draws(raw_len, S, max_it);
  h->d_perm.scratch(Pa * stride);
  h->d_samples.scratch(Pa * kRoundCap * 8);
  h->d_models.scratch(Pa * kRoundCap * 12);
  h->d_fb_list.scratch(8);
  if (mono) {
    h->d_nsol.scratch(Pa * kRoundCap);
    h->d_esol.scratch(Pa * kRoundCap * (prm.mono_algorithm == 1 ? 130 : 70));
    h->d_brk.scratch(Pa * kRoundCap * 40);
    // (draw, root) items of a round: a draw has at most 10 real roots, a round at most kRoundCap
    // draws per problem, but late rounds are large only for the few problems still running — the
    // lists are sized for 2 items per (problem, slot) and the batch is re-run with larger lists
    // if a round ever needs more (BatchStats::item_overflow / SacArgs::overflow)
    const size_t worst = Pa * kRoundCap * 10;
    size_t want = std::max<size_t>(h->item_cap, std::min<size_t>(worst, std::max<size_t>(Pa * kRoundCap * 2, 65536)));
    want = std::min(want, worst);
    h->item_cap = want;
    h->d_fb_list.scratch(want + 8);
    h->d_item_base.scratch(Pa * kRoundCap);
    h->d_item_list.scratch(want);
    h->d_item_q.scratch(want);
    h->d_item_model.scratch(want * 12);
    h->d_item_status.scratch(want);
  }
  h->d_valid.scratch(Pa * kRoundCap);
  h->d_counts.scratch(Pa * kRoundCap);
  h->d_active.scratch(2 * Pa + 2);
  out.st->scratch(Pa);
  out.best->scratch(Pa * 12);
  out.mask->scratch(Pa * mask_words);
  out.inl->scratch(Pa);
  SacArgs a;
  a.P = P; a.a = d_a; a.b = d_b; a.N = d_N; a.stride = stride;
  a.raw = h->sh->d_raw.p; a.raw_len = raw_len; a.cap_draws = cap_draws;
  a.perm = h->d_perm.p; a.samples = h->d_samples.p; a.models = h->d_models.p;
  a.fsol = h->d_esol.p; a.nroot = h->d_nsol.p; a.brk = h->d_brk.p;
  // d_fb_list: [0] deferred counter, [1] item counter, [2] item-list overflow flag of the mono rounds,
  // [3] pending counter of finish_sac, [4] the stereo problem's (never raised) overflow word, [8..] deferred list
  a.fb_list = h->d_fb_list.p + 8; a.fb_count = h->d_fb_list.p; a.item_count = h->d_fb_list.p + 1;
  a.overflow = h->d_fb_list.p + (mono ? 2 : 4); a.pending = pending;
  a.item_cap = (unsigned int)std::min<size_t>(h->item_cap, 0xFFFFFFF0u);
  a.item_base = h->d_item_base.p; a.item_list = h->d_item_list.p; a.item_q = h->d_item_q.p;
  a.item_model = h->d_item_model.p; a.item_status = h->d_item_status.p;
  a.valid = h->d_valid.p; a.counts = h->d_counts.p; a.st = out.st->p; a.best_model = out.best->p;
  a.n_active = reinterpret_cast<unsigned int*>(h->d_active.p); a.active = h->d_active.p + 2;
  a.ktable = mono ? h->sh->d_ktable_mono.p : onept ? h->sh->d_ktable_stereo1.p : h->sh->d_ktable_stereo.p;
  a.ktable_n = mono ? h->sh->ktable_n_mono : onept ? h->sh->ktable_n_stereo1 : h->sh->ktable_n_stereo;
  a.onept = onept ? 1 : 0;
  a.prior = d_prior;
  // latency schedule for small problem sets (a single query's 16 candidate pairs, the single-pair entry
  // points).  (Tried and dropped: the latency schedule for every stereo problem — 128-draw rounds made the
  // stereo stage of a C2 batch slower, 0.48 -> 0.60 ms: the extra draws' inlier counts cost more than the
  // three rounds they save.)
  const bool latency = P <= 256, single = P <= kSacSingleMaxP;
  a.first = single ? kSacFirstSingle : latency ? kSacFirstLatency : kSacFirstThroughput;
  a.n_rounds = single ? kSacRoundsSingle : latency ? kSacRoundsLatency : kSacRoundsThroughput;
  a.alg = (mono && prm.mono_algorithm == 1) ? 1 : 0;
  a.fo_stride = a.alg == 1 ? 130 : 70;  // geom::kFrontOutStew / kFrontOut (static_assert in ransac.cu)
  a.tab_nmax = 0;
  a.sample_tab = ensure_sample_table(h, S, stride, cap_draws, &a.tab_nmax);
  a.threshold = mono ? prm.ransac_threshold_mono : prm.ransac_threshold;
  a.sq_crit = sq_crit_of(prm.ransac_threshold);
  a.max_iterations = max_it; a.full = full;
  a.force_generic = (getenv("KML_FORCE_GENERIC_ISOLATE") ? 1 : 0) | (getenv("KML_NO_ROOT_GRID2") ? 2 : 0);
  a.inlier_mask = out.mask->p; a.mask_words = mask_words; a.n_inliers = out.inl->p;
  return a;
}

// init + the blind rounds + selectWithinDistance of the winners; everything is enqueued on the
// handle's stream, no host synchronisation.  Round 0 evaluates one chunk; each later round covers
// every trial the reference loop can still need (k never increases) up to the doubling schedule.
static void enqueue_sac(kml_handle* h, bool mono, const SacArgs& a, SacBufs out) {
  if (a.P <= 0) return;
  cudaStream_t s = h->stream;
  KML_CUDA(cudaMemsetAsync(out.best->p, 0, sizeof(double) * 12 * a.P, s));
  KML_CUDA(cudaMemsetAsync(a.overflow, 0, sizeof(unsigned int), s));
  launch_sac_init(a, mono ? 8 : (a.onept ? 1 : 3), s);
  h->stats.kernel_launches += 1;
  for (int r = 0; r < a.n_rounds; ++r)
    h->stats.kernel_launches += mono ? launch_mono_round(a, r, s) : launch_stereo_round(a, r, s);
  if (mono) launch_mono_select(a, s); else launch_stereo_select(a, s);
  h->stats.kernel_launches += 1;
  KML_CUDA(cudaGetLastError());
}

// Host-driven continuation for the problems the blind rounds did not finish (many skipped
// samples, or max_ransac_iterations beyond the blind schedule): rounds are added until every
// problem's loop has ended by its own limits, then the winners are selected again.  Synchronous.
// Returns false if the item lists overflowed (the caller re-runs with h->item_cap raised).
static bool finish_sac(kml_handle* h, bool mono, const SacArgs& a0) {
  if (a0.P <= 0) return true;
  cudaStream_t s = h->stream;
  SacArgs a = a0;
  a.pending = h->d_fb_list.p + 3;
  unsigned int flags[2] = {0, 0};  // overflow, pending
  bool again = false;
  for (int r = a.n_rounds;; ++r) {
    KML_CUDA(cudaMemsetAsync(a.pending, 0, sizeof(unsigned int), s));
    launch_sac_pending(a, s);
    h->stats.kernel_launches += 1;
    KML_CUDA(cudaMemcpyAsync(&flags[0], a.overflow, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    KML_CUDA(cudaMemcpyAsync(&flags[1], a.pending, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    KML_CUDA(cudaStreamSynchronize(s));
    if (flags[0]) return false;
    if (!flags[1]) break;
    again = true;
    h->stats.kernel_launches += mono ? launch_mono_round(a, r, s) : launch_stereo_round(a, r, s);
    KML_CUDA(cudaGetLastError());
  }
  if (again) {
    a.pending = nullptr;
    if (mono) launch_mono_select(a, s); else launch_stereo_select(a, s);
    h->stats.kernel_launches += 1;
  }
  return true;
}

// prepare + enqueue + finish for the single-problem entry points (synchronous)
static SacArgs run_sac_sync(kml_handle* h, bool mono, int P, const double* d_a, const double* d_b,
                            const int32_t* d_N, int stride, int full, SacBufs out, int mask_words,
                            const double* d_prior = nullptr) {
  for (int attempt = 0;; ++attempt) {
    SacArgs a = prepare_sac(h, mono, P, d_a, d_b, d_N, stride, full, out, mask_words, nullptr, d_prior);
    enqueue_sac(h, mono, a, out);
    if (finish_sac(h, mono, a)) return a;
    if (attempt >= 6) throw std::runtime_error("RANSAC item lists keep overflowing");
    h->item_cap = std::max<size_t>(h->item_cap * 4, 65536);  // larger lists, same deterministic run
  }
}

// =========================================================== batch query
// device views of the resident databases, the entry -> pose / frame maps and the frame table
static void sync_db_maps(kml_handle* h, RobotDb* db) {
  const size_t n = db->n_entries();
  if (db->synced == n && db->frame_patches.empty()) return;
  cudaStream_t s = h->stream;
  db->d_entry_pose.reserve(std::max<size_t>(n, 1), db->synced, s);
  db->d_entry_frame.reserve(std::max<size_t>(n, 1), db->synced, s);
  if (n > db->synced) {
    KML_CUDA(cudaMemcpyAsync(db->d_entry_pose.p + db->synced, db->entry_to_pose.data() + db->synced,
                             8 * (n - db->synced), cudaMemcpyHostToDevice, s));
    KML_CUDA(cudaMemcpyAsync(db->d_entry_frame.p + db->synced, db->entry_to_frame.data() + db->synced,
                             4 * (n - db->synced), cudaMemcpyHostToDevice, s));
  }
  if (db->frame_patches.size() > 64) {
    if (db->synced)
      KML_CUDA(cudaMemcpyAsync(db->d_entry_frame.p, db->entry_to_frame.data(), 4 * db->synced, cudaMemcpyHostToDevice, s));
  } else {
    for (uint32_t e : db->frame_patches)
      if (e < db->synced)
        KML_CUDA(cudaMemcpyAsync(db->d_entry_frame.p + e, db->entry_to_frame.data() + e, 4, cudaMemcpyHostToDevice, s));
  }
  KML_CUDA(cudaStreamSynchronize(s));
  db->synced = n;
  db->frame_patches.clear();
}

static void ensure_frame_offsets(kml_handle* h) {
  std::lock_guard<std::mutex> lk(h->sh->mu);
  kml_shared& sh = *h->sh;
  const size_t n = sh.frame_off_h.size();
  if (sh.frames_synced == n && sh.s_off.p) return;
  cudaStream_t s = h->stream;
  sh.s_off.reserve(std::max<size_t>(n, 1), sh.frames_synced, s);
  sh.s_F.reserve(std::max<size_t>(n, 1), sh.frames_synced, s);
  if (n > sh.frames_synced) {
    KML_CUDA(cudaMemcpyAsync(sh.s_off.p + sh.frames_synced, sh.frame_off_h.data() + sh.frames_synced,
                             8 * (n - sh.frames_synced), cudaMemcpyHostToDevice, s));
    KML_CUDA(cudaMemcpyAsync(sh.s_F.p + sh.frames_synced, sh.frame_F_h.data() + sh.frames_synced,
                             4 * (n - sh.frames_synced), cudaMemcpyHostToDevice, s));
  }
  KML_CUDA(cudaStreamSynchronize(s));
  sh.frames_synced = n;
}

// d_dbs: this handle's device array of database views, rebuilt when any add* moved the version
static void ensure_views(kml_handle* h) {
  kml_shared& sh = *h->sh;
  uint64_t ver;
  {
    std::lock_guard<std::mutex> lk(sh.mu);
    ver = sh.version;
  }
  if (h->views_version == ver && h->d_dbs.p) return;
  std::vector<BowDb> views;
  uint32_t max_entries = 1;
  for (auto& kv : sh.dbs) {
    views.push_back(db_view(h, kv.second.get()));
    max_entries = std::max(max_entries, views.back().n_entries);
  }
  h->views_n_db = (int)views.size();
  bow_tiling(max_entries, &h->views_tile, &h->views_n_tiles);
  h->d_dbs.scratch(std::max<size_t>(views.size(), 1));
  if (!views.empty()) {
    KML_CUDA(cudaMemcpyAsync(h->d_dbs.p, views.data(), sizeof(BowDb) * views.size(), cudaMemcpyHostToDevice, h->stream));
    KML_CUDA(cudaStreamSynchronize(h->stream));
  }
  h->views_version = ver;
}

// Device pointers of the uploaded query batch inside kml_handle::d_in (fixed layout per batch
// shape (B, F): a steady stream of equal-shaped batches never moves them)
struct BatchIn {
  uint64_t *q_robot, *q_pose;
  int64_t *qoff, *poff;
  uint32_t *qids, *pids;
  float *qvals, *pvals;
  uint8_t* desc;
  double *bear, *pts;
  size_t bytes;
};
static BatchIn batch_layout(uint8_t* base, int B, int F) {
  BatchIn L;
  size_t o = 0;
  auto take = [&](size_t n) { const size_t at = o; o += (n + 255) / 256 * 256; return base + at; };
  const size_t w_cap = (size_t)B * kBowMaxWords;  // check_bow caps a vector at kBowMaxWords
  L.q_robot = reinterpret_cast<uint64_t*>(take(8 * (size_t)B));
  L.q_pose = reinterpret_cast<uint64_t*>(take(8 * (size_t)B));
  L.qoff = reinterpret_cast<int64_t*>(take(8 * ((size_t)B + 1)));
  L.poff = reinterpret_cast<int64_t*>(take(8 * ((size_t)B + 1)));
  L.qids = reinterpret_cast<uint32_t*>(take(4 * w_cap));
  L.pids = reinterpret_cast<uint32_t*>(take(4 * w_cap));
  L.qvals = reinterpret_cast<float*>(take(4 * w_cap));
  L.pvals = reinterpret_cast<float*>(take(4 * w_cap));
  L.desc = take((size_t)B * F * 32 + 32);
  L.bear = reinterpret_cast<double*>(take((size_t)B * F * 24 + 24));
  L.pts = reinterpret_cast<double*>(take((size_t)B * F * 24 + 24));
  L.bytes = o;
  return L;
}

static bool is_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

static int batch_upload(kml_handle* h, int B, const uint64_t* q_robot, const uint64_t* q_pose,
                        const int64_t* bow_off, const uint32_t* ids, const float* vals,
                        const int64_t* prev_off, const uint32_t* prev_ids, const float* prev_vals,
                        const uint8_t* desc, const double* bearings, const double* points, int F,
                        bool wait, bool stage_only = false) {
  if (B < 0 || F < 0 || F > 65535) return fail(h, KML_ERR_ARG, "query_batch: bad B or F");
  if (B > 0 && (!q_robot || !q_pose || !bow_off || !prev_off || !desc || !bearings || !points))
    return fail(h, KML_ERR_ARG, "query_batch: null argument");
  for (int b = 0; b < B; ++b) {
    int rc = check_bow(h, ids + bow_off[b], vals + bow_off[b], (int)(bow_off[b + 1] - bow_off[b]));
    if (rc != KML_OK) return rc;
    rc = check_bow(h, prev_ids + prev_off[b], prev_vals + prev_off[b], (int)(prev_off[b + 1] - prev_off[b]));
    if (rc != KML_OK) return rc;
  }
  h->B = B; h->qF = F;
  if (B == 0) return KML_OK;
  const size_t nw = (size_t)bow_off[B], np = (size_t)prev_off[B];
  cudaStream_t s = h->stream;
  // The batch lives in ONE device buffer with a fixed layout per (B, F): sized for the largest
  // batch of this shape, not for this batch's word count, so a steady stream of B-query batches
  // never reallocates (cudaFree waits for the whole device, i.e. for every other lane's kernels
  // and pending collectives).  Arrays the caller holds in pinned memory (kml_host_alloc) are
  // copied from where they are; pageable ones are staged through the handle's pinned buffer, so
  // every H2D copy is asynchronous to the host and to the other lanes.
  const BatchIn L0 = batch_layout(nullptr, B, F);
  h->d_in.scratch(L0.bytes);
  const BatchIn D = batch_layout(h->d_in.p, B, F);
  h->h_in.scratch(L0.bytes);
  const BatchIn H = batch_layout(h->h_in.p, B, F);
  auto put = [&](void* dev, void* stage, const void* src, size_t bytes) {
    if (!bytes) return;
    if (stage_only) {  // small batches: everything goes to the staging buffer, ONE copy (or the graph's) moves it
      memcpy(stage, src, bytes);
      return;
    }
    const void* from = src;
    if (!is_pinned(src)) {
      memcpy(stage, src, bytes);
      from = stage;
    }
    KML_CUDA(cudaMemcpyAsync(dev, from, bytes, cudaMemcpyHostToDevice, s));
  };
  // the staging buffer may still feed the previous upload of this lane
  KML_CUDA(cudaStreamSynchronize(s));
  put(D.q_robot, H.q_robot, q_robot, 8 * (size_t)B);
  put(D.q_pose, H.q_pose, q_pose, 8 * (size_t)B);
  put(D.qoff, H.qoff, bow_off, 8 * ((size_t)B + 1));
  put(D.poff, H.poff, prev_off, 8 * ((size_t)B + 1));
  put(D.qids, H.qids, ids, 4 * nw);
  put(D.qvals, H.qvals, vals, 4 * nw);
  put(D.pids, H.pids, prev_ids, 4 * np);
  put(D.pvals, H.pvals, prev_vals, 4 * np);
  put(D.desc, H.desc, desc, (size_t)B * F * 32);
  put(D.bear, H.bear, bearings, (size_t)B * F * 24);
  put(D.pts, H.pts, points, (size_t)B * F * 24);
  if (wait) KML_CUDA(cudaStreamSynchronize(s));  // the caller may reuse its buffers on return
  return KML_OK;
}

// layout of one rank's record block: kml_result[B][cap] | int32 counts[B] | int32 error | pad
struct BlockLayout {
  size_t rec_bytes, counts_off, err_off, blk;
};
static BlockLayout block_layout(int B, int cap) {
  BlockLayout L;
  L.rec_bytes = sizeof(kml_result) * (size_t)B * cap;
  L.counts_off = L.rec_bytes;
  L.err_off = L.counts_off + sizeof(int32_t) * (size_t)B;
  L.blk = (L.err_off + sizeof(int32_t) + 255) / 256 * 256;
  return L;
}

// The whole query pipeline of the uploaded batch, enqueued on the handle's stream without a host
// synchronisation: BoW scoring -> candidate selection -> kNN + Lowe -> mono 5-pt RANSAC -> stereo
// Arun RANSAC -> records written to `blk` (this rank's block of the all-gather buffer).
struct BatchPlan {
  int B, K, cap, P, stride, mask_words;
  SacArgs mono, stereo;
  FinalizeArgs fin;
  StereoGatherArgs sg;
};

static void enqueue_tail(kml_handle* h, BatchPlan* pl, bool from_mono_select);
// stage-timing events are recorded in eager mode only: inside a stream capture they would become
// graph nodes that cudaEventElapsedTime cannot read
static inline void rec_event(kml_handle* h, cudaEvent_t e) {
  if (!h->capturing) KML_CUDA(cudaEventRecord(e, h->stream));
}

static void batch_enqueue(kml_handle* h, int cap, uint8_t* blk, BatchPlan* pl) {
  const int B = h->B;
  const int K = std::max(0, std::min(h->prm.top_k_verify, cap));
  const int P = B * K;
  const BlockLayout bl = block_layout(B, cap);
  cudaStream_t s = h->stream;
  pl->B = B; pl->K = K; pl->cap = cap; pl->P = P;
  ensure_frame_offsets(h);
  ensure_views(h);
  const int n_db = h->views_n_db;
  const BatchIn D = batch_layout(h->d_in.p, B, h->qF);
  kml_result* d_recs = reinterpret_cast<kml_result*>(blk);
  int32_t* d_counts = reinterpret_cast<int32_t*>(blk + bl.counts_off);
  h->d_stats.scratch(1);
  rec_event(h, h->ev[6]);
  KML_CUDA(cudaMemsetAsync(blk, 0, bl.blk, s));
  KML_CUDA(cudaMemsetAsync(h->d_stats.p, 0, sizeof(BatchStats), s));
  // ---- detectLoop: NSS factor + DBoW2 query against every resident database
  const int n_tiles = h->views_n_tiles, Kdb = h->prm.max_db_results;
  const size_t nlist = (size_t)B * std::max(n_db, 1) * n_tiles;
  h->d_bow_entry.scratch(nlist * Kdb);
  h->d_bow_score.scratch(nlist * Kdb);
  h->d_bow_count.scratch(nlist);
  h->d_nss.scratch(B);
  KML_CUDA(cudaMemsetAsync(h->d_nss.p, 0, sizeof(double) * B, s));
  rec_event(h, h->ev[0]);
  if (n_db > 0) {
    BowArgs a;
    a.dbs = h->d_dbs.p; a.n_db = n_db; a.B = B;
    a.q_off = D.qoff; a.q_ids = D.qids; a.q_vals = D.qvals;
    a.p_off = D.poff; a.p_ids = D.pids; a.p_vals = D.pvals;
    a.K = Kdb; a.max_id = nullptr; a.tile_entries = h->views_tile; a.n_tiles = n_tiles;
    a.out_entry = h->d_bow_entry.p; a.out_score = h->d_bow_score.p; a.out_count = h->d_bow_count.p;
    a.nss = h->d_nss.p;
    a.postings_touched = &h->d_stats.p->postings;
    launch_bow(a, s);
    h->stats.kernel_launches += 1;
  }
  rec_event(h, h->ev[1]);
  // ---- candidate selection, records, pair slots p = b*K + i
  const int stride = std::max(h->qF, 8);
  const int mask_words = (stride + 31) / 32;
  pl->stride = stride; pl->mask_words = mask_words;
  const size_t Pa = (size_t)std::max(P, 1);
  h->d_keys.scratch(Pa * stride * 2);
  h->d_jobs.scratch(Pa); h->d_pairs.scratch(Pa); h->d_nq.scratch(Pa);
  h->d_iq.scratch(Pa * stride); h->d_im.scratch(Pa * stride);
  h->d_kq.scratch(Pa * stride); h->d_km.scratch(Pa * stride);
  h->d_M.scratch(Pa); h->d_N3.scratch(Pa); h->d_mono_ok.scratch(Pa);
  h->d_a.scratch(Pa * stride * 3); h->d_b.scratch(Pa * stride * 3);
  SelectArgs sa;
  sa.dbs = h->d_dbs.p; sa.n_db = n_db; sa.B = B; sa.n_tiles = n_tiles; sa.Kdb = Kdb;
  sa.bow_entry = h->d_bow_entry.p; sa.bow_score = h->d_bow_score.p; sa.bow_count = h->d_bow_count.p;
  sa.nss = h->d_nss.p; sa.q_robot = D.q_robot; sa.q_pose = D.q_pose;
  sa.alpha = h->prm.alpha; sa.min_nss = h->prm.min_nss_factor;
  sa.inter_robot_only = h->prm.inter_robot_only; sa.dist_local = h->prm.dist_local;
  sa.K = K; sa.cap = cap;
  sa.s_desc = h->sh->s_desc.p; sa.s_off = h->sh->s_off.p; sa.s_F = h->sh->s_F.p;
  sa.q_desc = D.desc; sa.qF = h->qF;
  sa.recs = d_recs; sa.counts = d_counts;
  sa.pairs = h->d_pairs.p; sa.jobs = h->d_jobs.p; sa.nq = h->d_nq.p; sa.keys = h->d_keys.p; sa.key_stride = stride;
  sa.stats = h->d_stats.p;
  launch_select(sa, s);
  h->stats.kernel_launches += 1;
  if (P == 0) {
    for (int e = 2; e <= 5; ++e) rec_event(h, h->ev[e]);
    rec_event(h, h->ev[7]);
    KML_CUDA(cudaGetLastError());
    return;
  }
  // ---- computeMatchedIndices
  rec_event(h, h->ev[2]);
  if (h->prm.matcher_engine == 1) launch_hamming_jobs_tc(h->d_jobs.p, P, s);
  else launch_hamming_jobs(h->d_jobs.p, P, h->prm.matcher_norm, s);
  launch_lowe_compact(h->d_keys.p, h->d_nq.p, stride, h->prm.lowe_ratio, h->prm.matcher_norm, h->d_iq.p, h->d_im.p, h->d_M.p, P, s);
  h->stats.kernel_launches += 2;
  rec_event(h, h->ev[3]);
  // ---- geometricVerificationNister
  GatherArgs g;
  g.P = P; g.pairs = h->d_pairs.p; g.qb = D.bear; g.qp = D.pts; g.qF = h->qF;
  g.sb = h->sh->s_bear.p; g.sp = h->sh->s_pts.p; g.s_off = h->sh->s_off.p;
  g.iq = h->d_iq.p; g.im = h->d_im.p; g.M = h->d_M.p; g.stride = stride;
  g.a = h->d_a.p; g.b = h->d_b.p; g.N = h->d_N3.p;  // N3 reused as "N" of the mono stage
  launch_gather_bearings(g, s);
  h->stats.kernel_launches += 1;
  SacBufs mono{&h->d_st_mono, &h->d_best_mono, &h->d_mask_mono, &h->d_inl_mono};
  pl->mono = prepare_sac(h, true, P, h->d_a.p, h->d_b.p, h->d_M.p, stride, 0, mono, mask_words, &h->d_stats.p->pending_m);
  SacBufs st3{&h->d_st_stereo, &h->d_best_stereo, &h->d_mask_stereo, &h->d_inl_stereo};
  pl->stereo = prepare_sac(h, false, P, h->d_a.p, h->d_b.p, h->d_N3.p, stride, 0, st3, mask_words, &h->d_stats.p->pending_s,
                           h->prm.ransac_use_1point_3d3d ? h->d_best_mono.p : nullptr);
  enqueue_sac(h, true, pl->mono, mono);
  FinalizeArgs& f = pl->fin;
  f.P = P; f.mono_st = h->d_st_mono.p; f.mono_inl = h->d_inl_mono.p; f.M = h->d_M.p;
  f.mono_model = h->d_best_mono.p;
  f.min_inliers = h->prm.geometric_verification_min_inlier_count;
  f.min_ratio_mono = h->prm.ransac_inlier_percentage_mono;
  f.min_ratio_stereo = h->prm.geometric_verification_min_inlier_percentage;
  f.mono_ok = h->d_mono_ok.p;
  f.st3 = h->d_st_stereo.p; f.inl3 = h->d_inl_stereo.p; f.N3 = h->d_N3.p; f.model3 = h->d_best_stereo.p;
  f.pairs = h->d_pairs.p; f.K = K; f.cap = cap; f.recs = d_recs; f.stats = h->d_stats.p;
  pl->sg.g = g;
  pl->sg.mono_mask = h->d_mask_mono.p; pl->sg.mask_words = mask_words; pl->sg.mono_ok = h->d_mono_ok.p;
  pl->sg.kq = h->d_kq.p; pl->sg.km = h->d_km.p;
  enqueue_tail(h, pl, false);
}

// mono acceptance gate -> recoverPose -> records.  Also the re-run after a host-driven
// continuation of the mono rounds (from_mono_select: the winners' inlier sets are selected again).
static void enqueue_tail(kml_handle* h, BatchPlan* pl, bool redo) {
  cudaStream_t s = h->stream;
  if (redo)  // the counters the first pass added for these stages, and its pending flags
    KML_CUDA(cudaMemsetAsync(&h->d_stats.p->mono_ok, 0, sizeof(BatchStats) - offsetof(BatchStats, mono_ok), s));
  launch_mono_gate(pl->fin, s);
  h->stats.kernel_launches += 1;
  rec_event(h, h->ev[4]);
  launch_gather_points(pl->sg, s);
  h->stats.kernel_launches += 1;
  SacBufs st3{&h->d_st_stereo, &h->d_best_stereo, &h->d_mask_stereo, &h->d_inl_stereo};
  enqueue_sac(h, false, pl->stereo, st3);
  if (redo) finish_sac(h, false, pl->stereo);
  launch_finalize(pl->fin, s);
  h->stats.kernel_launches += 1;
  rec_event(h, h->ev[5]);
  rec_event(h, h->ev[7]);
  KML_CUDA(cudaGetLastError());
}

static void batch_stats_to_host(kml_handle* h, const BatchStats& st, bool graph = false) {
  if (graph) {  // a replayed graph carries no stage events: only the total around the launch
    h->stats.ms_bow = h->stats.ms_match = h->stats.ms_mono = h->stats.ms_stereo = 0.f;
  } else {
    KML_CUDA(cudaEventElapsedTime(&h->stats.ms_bow, h->ev[0], h->ev[1]));
    KML_CUDA(cudaEventElapsedTime(&h->stats.ms_match, h->ev[2], h->ev[3]));
    KML_CUDA(cudaEventElapsedTime(&h->stats.ms_mono, h->ev[3], h->ev[4]));
    KML_CUDA(cudaEventElapsedTime(&h->stats.ms_stereo, h->ev[4], h->ev[5]));
  }
  KML_CUDA(cudaEventElapsedTime(&h->stats.ms_total, h->ev[6], h->ev[7]));
  h->stats.bow_postings_last = st.postings;
  h->stats.pairs_last = st.pairs;
  h->stats.mono_hypotheses_last = st.hyp_m;
  h->stats.stereo_hypotheses_last = st.hyp_s;
  h->stats.mono_residuals_last = st.res_m;
  h->stats.stereo_residuals_last = st.res_s;
  h->stats.total_bow_matches += st.survivors;
  h->stats.total_geom_verifications_mono += st.pairs;
  h->stats.total_geometric_verifications += st.mono_ok;
  if (getenv("KML_DEBUG_TIMING"))
    fprintf(stderr, "[kml] pairs %llu; mono draws consumed %llu evaluated %llu; stereo consumed %llu evaluated %llu; "
            "ms bow %.3f match %.3f mono %.3f stereo %.3f total %.3f\n",
            st.pairs, st.hyp_m, st.eval_m, st.hyp_s, st.eval_s, h->stats.ms_bow, h->stats.ms_match,
            h->stats.ms_mono, h->stats.ms_stereo, h->stats.ms_total);
}

struct PoliteScope {  // batches wait without spinning; single queries keep the low-latency spin
  kml_handle* h;
  PoliteScope(kml_handle* hh, bool on) : h(hh) { h->polite_wait = on; }
  ~PoliteScope() { h->polite_wait = false; }
};

// flag values of a record block (BlockLayout::err_off)
enum { kBlkOk = 0, kBlkRedo = 1, kBlkFailed = 2 };
// replays the captured local pipeline of this batch shape if there is one (defined with the graph code below)
extern "C" {
static bool batch_enqueue_graph(kml_handle* h, int cap, uint8_t* blk, BatchPlan* pl);
}

// Runs the uploaded batch.  Common case: ONE enqueue of the whole pipeline (plus, when sharded,
// the in-place ncclAllGather of the record blocks and the device merge), one D2H of the final
// records into pinned memory, one host wait.  Every rank's block carries a flag word, so all
// ranks take the same decision after the exchange: kBlkRedo = some rank has RANSAC problems the
// blind rounds did not finish (or overflowed item lists) — it continues them from the host and
// the exchange is repeated; kBlkFailed = some rank hit an error before the exchange — every
// rank still enters the collective (nobody is left waiting) and every rank returns the error.
static int batch_run(kml_handle* h, int cap, bool sharded, kml_result* out, int32_t* counts,
                     const uint64_t* seq = nullptr) {
  const int B = h->B;
  if (B == 0) return KML_OK;
  PoliteScope scope(h, B >= 16);
  const int nr = sharded ? comm_nranks(h) : 1;
  const int rank = sharded ? comm_rank(h) : 0;
  const BlockLayout bl = block_layout(B, cap);
  cudaStream_t s = h->stream;
  h->d_blocks.scratch(bl.blk * nr);
  if (nr > 1) h->d_merged.scratch(bl.blk);
  h->h_recs.scratch(bl.blk);
  h->h_stats.scratch(1);
  h->d_stats.scratch(1);
  uint8_t* blk = h->d_blocks.p + bl.blk * rank;
  const bool dbg = getenv("KML_DEBUG_TIMING") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  BatchPlan pl;
  std::string local_err;
  int local_rc = KML_OK;
  bool enqueued = false;
  for (int attempt = 0;; ++attempt) {
    // ---- local pipeline (first pass) or the continuation the previous exchange asked for
    if (local_rc == KML_OK) {
      try {
        if (!enqueued) {
          h->graph_replayed = batch_enqueue_graph(h, cap, blk, &pl);
          if (!h->graph_replayed) batch_enqueue(h, cap, blk, &pl);
          enqueued = true;
        } else {
          const BatchStats st = *h->h_stats.p;
          unsigned int overflow = st.item_overflow;
          if (!overflow && (st.pending_m || st.pending_s)) {
            if (!finish_sac(h, true, pl.mono)) overflow = 1;
            else enqueue_tail(h, &pl, true);
          }
          if (overflow) {  // larger item lists, same deterministic run from the top
            if (attempt >= 6) throw std::runtime_error("query_batch: RANSAC item lists keep overflowing");
            h->item_cap = std::max<size_t>(h->item_cap * 4, 65536);
            batch_enqueue(h, cap, blk, &pl);
          }
        }
        if (pl.P > 0) launch_block_flags(h->d_stats.p, pl.mono.overflow, reinterpret_cast<int32_t*>(blk + bl.err_off), s);
        h->stats.kernel_launches += pl.P > 0;
        KML_CUDA(cudaGetLastError());
      } catch (const kml::CudaError& e) {
        local_err = e.what(); local_rc = KML_ERR_CUDA; cudaGetLastError();
      } catch (const std::exception& e) {
        local_err = e.what(); local_rc = KML_ERR_ARG;
      }
    }
    if (local_rc != KML_OK) {  // tell the peers; best effort if the device itself is gone
      const int32_t flag = kBlkFailed;
      cudaMemcpyAsync(blk + bl.err_off, &flag, sizeof(flag), cudaMemcpyHostToDevice, s);
      cudaStreamSynchronize(s);
      cudaGetLastError();
      if (nr == 1) { h->err = local_err; return local_rc; }
    }
    // ---- exchange + merge (sharded), final records to pinned memory
    const uint8_t* final_blk = blk;
    if (nr > 1) {
      int rc;
      if (seq && attempt == 0) {  // this batch's turn among the lanes' collectives (same order on every rank)
        std::unique_lock<std::mutex> lk(h->sh->seq_mu);
        h->sh->seq_cv.wait(lk, [&] { return h->sh->seq_next >= *seq; });
        rc = comm_allgather(h, blk, h->d_blocks.p, bl.blk);
        if (h->sh->seq_next == *seq) h->sh->seq_next = *seq + 1;
        lk.unlock();
        h->sh->seq_cv.notify_all();
      } else {
        rc = comm_allgather(h, blk, h->d_blocks.p, bl.blk);
      }
      if (rc != KML_OK) return rc;
      MergeArgs m;
      m.base = h->d_blocks.p; m.blk_stride = bl.blk; m.counts_off = bl.counts_off; m.err_off = bl.err_off;
      m.nranks = nr; m.B = B; m.cap_in = cap; m.cap = cap;
      m.out = reinterpret_cast<kml_result*>(h->d_merged.p);
      m.counts = reinterpret_cast<int32_t*>(h->d_merged.p + bl.counts_off);
      m.err_out = reinterpret_cast<int32_t*>(h->d_merged.p + bl.err_off);
      KML_CUDA(cudaMemsetAsync(h->d_merged.p, 0, bl.blk, s));
      launch_merge(m, s);
      h->stats.kernel_launches += 1;
      final_blk = h->d_merged.p;
    }
    KML_CUDA(cudaMemcpyAsync(h->h_recs.p, final_blk, bl.err_off + sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    KML_CUDA(cudaMemcpyAsync(h->h_stats.p, h->d_stats.p, sizeof(BatchStats), cudaMemcpyDeviceToHost, s));
    const int wrc = nr > 1 ? comm_wait(h) : (h->wait_stream(), KML_OK);
    if (wrc != KML_OK) return wrc;
    const int32_t flag = *reinterpret_cast<const int32_t*>(h->h_recs.p + bl.err_off);
    if (flag == kBlkOk) break;
    if (flag >= kBlkFailed || local_rc != KML_OK) {
      h->err = local_rc != KML_OK ? local_err : "query_batch_sharded: another rank failed before the exchange";
      return local_rc != KML_OK ? local_rc : KML_ERR_NCCL;
    }
    if (attempt >= 40) return fail(h, KML_ERR_CUDA, "query_batch: RANSAC continuation did not converge");
  }
  batch_stats_to_host(h, *h->h_stats.p, h->graph_replayed);
  memcpy(out, h->h_recs.p, bl.rec_bytes);
  memcpy(counts, h->h_recs.p + bl.counts_off, sizeof(int32_t) * (size_t)B);
  if (dbg) {
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "[kml r%d] batch of %d: host wall %.3f ms (device total %.3f ms)\n", rank, B, ms, h->stats.ms_total);
  }
  return KML_OK;
}

// Query-side frame arrays of a single-pair verification call
struct QuerySide {
  const uint8_t* desc; const double* bear; const double* pts; int F;  // [slots][F][...]
};

// single stored frame as the "query side"
static bool stored_query_side(kml_handle* h, uint64_t robot, uint64_t pose, QuerySide* qs) {
  auto it = h->sh->frames.find(RobotPoseId(robot, pose));
  if (it == h->sh->frames.end()) return false;
  const FrameRec& r = it->second;
  qs->desc = h->sh->s_desc.p + (size_t)r.feat_off * 32;
  qs->bear = h->sh->s_bear.p + (size_t)r.feat_off * 3;
  qs->pts = h->sh->s_pts.p + (size_t)r.feat_off * 3;
  qs->F = r.F;
  return true;
}

// gather + RANSAC on caller-provided index lists for ONE stored pair
static int sac_on_lists(kml_handle* h, bool mono, uint64_t qr, uint64_t qp, uint64_t mr,
                        uint64_t mp, uint32_t* inl_q, uint32_t* inl_m, int* count, double* model12,
                        int* n_valid_out, const double* R_prior = nullptr) {
  if (!inl_q || !inl_m || !count || *count < 0) return fail(h, KML_ERR_ARG, "bad index lists");
  QuerySide qs;
  if (!stored_query_side(h, qr, qp, &qs)) return KML_NO_FRAME;
  auto mit = h->sh->frames.find(RobotPoseId(mr, mp));
  if (mit == h->sh->frames.end()) return KML_NO_FRAME;
  const int M = *count;
  for (int i = 0; i < M; ++i)
    if (inl_q[i] >= (uint32_t)qs.F || inl_m[i] >= (uint32_t)mit->second.F)
      return fail(h, KML_ERR_ARG, "feature index out of range");
  ensure_frame_offsets(h);
  const int stride = std::max(M, 8);
  const int mask_words = (stride + 31) / 32;
  std::vector<uint16_t> iq(stride, 0), im(stride, 0);
  for (int i = 0; i < M; ++i) { iq[i] = (uint16_t)inl_q[i]; im[i] = (uint16_t)inl_m[i]; }
  PairDesc pd{0, mit->second.index};
  int32_t one = 1;
  cudaStream_t s = h->stream;
  h->d_pairs.scratch(1); h->d_iq.scratch(stride); h->d_im.scratch(stride);
  h->d_kq.scratch(stride); h->d_km.scratch(stride); h->d_M.scratch(1); h->d_N3.scratch(1);
  h->d_mono_ok.scratch(1); h->d_a.scratch((size_t)stride * 3); h->d_b.scratch((size_t)stride * 3);
  KML_CUDA(cudaMemcpyAsync(h->d_pairs.p, &pd, sizeof(pd), cudaMemcpyHostToDevice, s));
  KML_CUDA(cudaMemcpyAsync(h->d_iq.p, iq.data(), 2 * stride, cudaMemcpyHostToDevice, s));
  KML_CUDA(cudaMemcpyAsync(h->d_im.p, im.data(), 2 * stride, cudaMemcpyHostToDevice, s));
  KML_CUDA(cudaMemcpyAsync(h->d_M.p, &M, 4, cudaMemcpyHostToDevice, s));
  KML_CUDA(cudaMemcpyAsync(h->d_mono_ok.p, &one, 4, cudaMemcpyHostToDevice, s));
  GatherArgs g;
  g.P = 1; g.pairs = h->d_pairs.p; g.qb = qs.bear; g.qp = qs.pts; g.qF = qs.F;
  g.sb = h->sh->s_bear.p; g.sp = h->sh->s_pts.p; g.s_off = h->sh->s_off.p;
  g.iq = h->d_iq.p; g.im = h->d_im.p; g.M = h->d_M.p; g.stride = stride;
  g.a = h->d_a.p; g.b = h->d_b.p; g.N = h->d_N3.p;
  SacBufs bufs = mono ? SacBufs{&h->d_st_mono, &h->d_best_mono, &h->d_mask_mono, &h->d_inl_mono}
                      : SacBufs{&h->d_st_stereo, &h->d_best_stereo, &h->d_mask_stereo, &h->d_inl_stereo};
  std::vector<uint16_t> kq(stride), km(stride);
  if (mono) {
    launch_gather_bearings(g, s);
  } else {
    // recoverPose filters on the 3-D keypoints of every listed pair: all-ones mask
    std::vector<uint32_t> ones(mask_words, 0xFFFFFFFFu);
    h->d_mask_mono.scratch(mask_words);
    KML_CUDA(cudaMemcpyAsync(h->d_mask_mono.p, ones.data(), 4 * mask_words, cudaMemcpyHostToDevice, s));
    StereoGatherArgs sg;
    sg.g = g; sg.mono_mask = h->d_mask_mono.p; sg.mask_words = mask_words;
    sg.mono_ok = h->d_mono_ok.p; sg.kq = h->d_kq.p; sg.km = h->d_km.p;
    launch_gather_points(sg, s);
  }
  h->stats.kernel_launches += 1;
  const double* d_prior = nullptr;
  if (!mono && R_prior && h->prm.ransac_use_1point_3d3d) {  // row f4: the rotation is given
    double pr[12] = {R_prior[0], R_prior[1], R_prior[2], 0.0, R_prior[3], R_prior[4], R_prior[5], 0.0,
                     R_prior[6], R_prior[7], R_prior[8], 0.0};
    h->d_prior.scratch(12);
    KML_CUDA(cudaMemcpyAsync(h->d_prior.p, pr, sizeof(pr), cudaMemcpyHostToDevice, s));
    KML_CUDA(cudaStreamSynchronize(s));
    d_prior = h->d_prior.p;
  }
  run_sac_sync(h, mono, 1, h->d_a.p, h->d_b.p, h->d_N3.p, stride, 0, bufs, mask_words, d_prior);
  SacState st;
  int32_t N = 0, ninl = 0;
  std::vector<uint32_t> mask(mask_words);
  KML_CUDA(cudaMemcpyAsync(&st, bufs.st->p, sizeof(st), cudaMemcpyDeviceToHost, s));
  KML_CUDA(cudaMemcpyAsync(&N, h->d_N3.p, 4, cudaMemcpyDeviceToHost, s));
  KML_CUDA(cudaMemcpyAsync(&ninl, bufs.inl->p, 4, cudaMemcpyDeviceToHost, s));
  KML_CUDA(cudaMemcpyAsync(mask.data(), bufs.mask->p, 4 * mask_words, cudaMemcpyDeviceToHost, s));
  KML_CUDA(cudaMemcpyAsync(model12, bufs.best->p, 96, cudaMemcpyDeviceToHost, s));
  if (!mono) {
    KML_CUDA(cudaMemcpyAsync(kq.data(), h->d_kq.p, 2 * stride, cudaMemcpyDeviceToHost, s));
    KML_CUDA(cudaMemcpyAsync(km.data(), h->d_km.p, 2 * stride, cudaMemcpyDeviceToHost, s));
  }
  KML_CUDA(cudaStreamSynchronize(s));
  if (st.exhausted) return fail(h, KML_ERR_STREAM_EXHAUSTED, "pre-drawn sample stream exhausted");  // cannot happen: the stream covers the loop's own limits
  *n_valid_out = N;
  if (N < (mono ? 8 : 3)) return KML_TOO_FEW_POINTS;  // recoverPose needs three points in either stereo mode
  if (st.best_draw < 0) return KML_RANSAC_FAIL;
  const kml_params& P = h->prm;
  if (ninl < P.geometric_verification_min_inlier_count) return KML_TOO_FEW_INLIERS;
  const double ratio = mono ? P.ransac_inlier_percentage_mono : P.geometric_verification_min_inlier_percentage;
  if ((double)ninl / (double)N < ratio) return KML_TOO_FEW_INLIERS;
  // rewrite the index lists to the inlier subset (ascending)
  int c = 0;
  for (int i = 0; i < N; ++i)
    if ((mask[i >> 5] >> (i & 31)) & 1u) {
      const uint32_t a = mono ? (uint32_t)iq[i] : (uint32_t)kq[i];
      const uint32_t b = mono ? (uint32_t)im[i] : (uint32_t)km[i];
      inl_q[c] = a; inl_m[c] = b; ++c;
    }
  *count = c;
  return KML_OK;
}

extern "C" {

int kml_add_bow(kml_handle* h, uint64_t robot, uint64_t pose, const uint32_t* ids,
                const float* vals, int n) {
  KML_API_BEGIN(h)
  return add_bow_host(h, get_db(h, robot, true), pose, ids, vals, n);
  KML_API_END(h)
}

int kml_add_bow_bulk(kml_handle* h, uint64_t robot, const uint64_t* poses, int count,
                     const int64_t* off, const uint32_t* ids, const float* vals) {
  KML_API_BEGIN(h)
  if (count < 0 || (count > 0 && (!poses || !off))) return fail(h, KML_ERR_ARG, "add_bow_bulk: bad argument");
  RobotDb* db = get_db(h, robot, true);
  for (int i = 0; i < count; ++i) {
    int rc = add_bow_host(h, db, poses[i], ids + off[i], vals + off[i], (int)(off[i + 1] - off[i]));
    if (rc != KML_OK) return rc;
  }
  return KML_OK;
  KML_API_END(h)
}

int kml_add_frame(kml_handle* h, uint64_t robot, uint64_t pose, const uint8_t* desc,
                  const double* bearings, const double* points, int F) {
  KML_API_BEGIN(h)
  return add_frames_host(h, robot, &pose, 1, desc, bearings, points, F);
  KML_API_END(h)
}

int kml_add_frames_bulk(kml_handle* h, uint64_t robot, const uint64_t* poses, int count,
                        const uint8_t* desc, const double* bearings, const double* points, int F) {
  KML_API_BEGIN(h)
  return add_frames_host(h, robot, poses, count, desc, bearings, points, F);
  KML_API_END(h)
}

int kml_frame_exists(kml_handle* h, uint64_t robot, uint64_t pose) {
  if (!h) return KML_ERR_ARG;
  return h->sh->frames.count(RobotPoseId(robot, pose)) ? 1 : 0;
}
int kml_bow_exists(kml_handle* h, uint64_t robot, uint64_t pose) {
  if (!h) return KML_ERR_ARG;
  RobotDb* db = get_db(h, robot, false);
  return (db && db->pose_to_entry.count(pose)) ? 1 : 0;
}
int kml_num_bow_for_robot(kml_handle* h, uint64_t robot) {
  if (!h) return KML_ERR_ARG;
  RobotDb* db = get_db(h, robot, false);
  return db ? (int)db->n_entries() : 0;
}
int kml_get_bow_vector(kml_handle* h, uint64_t robot, uint64_t pose, uint32_t* ids, float* vals,
                       int cap, int* count) {
  if (!h || !count) return KML_ERR_ARG;
  RobotDb* db = get_db(h, robot, false);
  if (!db) return KML_NO_DB;
  auto it = db->pose_to_entry.find(pose);
  if (it == db->pose_to_entry.end()) return KML_NO_PREV_BOW;
  const int64_t o = db->off[it->second];
  const int n = (int)(db->off[it->second + 1] - o);
  if (n > cap) return fail(h, KML_ERR_CAPACITY, "get_bow_vector: capacity");
  if (n) {
    memcpy(ids, &db->ids[o], 4 * n);
    memcpy(vals, &db->vals[o], 4 * n);
  }
  *count = n;
  return KML_OK;
}

int kml_db_query(kml_handle* h, uint64_t robot, const uint32_t* ids, const float* vals, int n,
                 int max_results, int max_id, uint32_t* out_entry, double* out_score, int cap,
                 int* count) {
  KML_API_BEGIN(h)
  if (!count || !out_entry || !out_score) return fail(h, KML_ERR_ARG, "db_query: null output");
  *count = 0;
  int rc = check_bow(h, ids, vals, n);
  if (rc != KML_OK) return rc;
  RobotDb* db = get_db(h, robot, false);
  if (!db) return KML_NO_DB;
  if (max_results <= 0 || max_results > kBowMaxK)
    return fail(h, KML_ERR_CAPACITY, "db_query: max_results must be in [1,128]");
  upload_single_bow(h, ids, vals, n, nullptr, nullptr, 0);
  BowOut bo;
  std::vector<int32_t> mid{max_id};
  run_bow(h, {db}, 1, h->d_qoff.p, h->d_qids.p, h->d_qvals.p, nullptr, nullptr, nullptr,
          max_results, &mid, &bo);
  const int c = std::min(bo.count[0], cap);
  memcpy(out_entry, bo.entry.data(), 4 * c);
  memcpy(out_score, bo.score.data(), 8 * c);
  *count = c;
  return KML_OK;
  KML_API_END(h)
}

int kml_bow_score(kml_handle* h, const uint32_t* ids1, const float* vals1, int n1,
                  const uint32_t* ids2, const float* vals2, int n2, double* out) {
  KML_API_BEGIN(h)
  if (!out) return KML_ERR_ARG;
  int rc = check_bow(h, ids1, vals1, n1);
  if (rc != KML_OK) return rc;
  rc = check_bow(h, ids2, vals2, n2);
  if (rc != KML_OK) return rc;
  // the NSS path of the scorer kernel against an empty one-entry database view
  upload_single_bow(h, ids1, vals1, n1, ids2, vals2, n2);
  RobotDb empty;
  empty.robot = ~0ull;
  BowOut bo;
  run_bow(h, {&empty}, 1, h->d_qoff.p, h->d_qids.p, h->d_qvals.p, h->d_poff.p, h->d_pids.p,
          h->d_pvals.p, 1, nullptr, &bo);
  *out = bo.nss[0];
  return KML_OK;
  KML_API_END(h)
}

static int detect_impl(kml_handle* h, const std::vector<RobotDb*>& dbs, uint64_t q_robot,
                       uint64_t q_pose, const uint32_t* ids, const float* vals, int n,
                       uint64_t* out_robot, uint64_t* out_pose, double* out_score, int cap,
                       int* count) {
  *count = 0;
  int rc = check_bow(h, ids, vals, n);
  if (rc != KML_OK) return rc;
  if (dbs.empty()) return KML_NO_DB;
  // findPreviousBoWVector(query, max_nrFrames_between_queries)
  RobotDb* own = get_db(h, q_robot, false);
  int64_t po = 0;
  int pn = -1;
  if (own)
    for (int i = 1; i <= h->prm.max_nrFrames_between_queries; ++i) {
      if (q_pose < (uint64_t)i) break;
      auto it = own->pose_to_entry.find(q_pose - i);
      if (it != own->pose_to_entry.end()) {
        po = own->off[it->second];
        pn = (int)(own->off[it->second + 1] - po);
        break;
      }
    }
  if (pn < 0) return KML_NO_PREV_BOW;
  upload_single_bow(h, ids, vals, n, own->ids.data() + po, own->vals.data() + po, pn);
  BowOut bo;
  run_bow(h, dbs, 1, h->d_qoff.p, h->d_qids.p, h->d_qvals.p, h->d_poff.p, h->d_pids.p,
          h->d_pvals.p, h->prm.max_db_results, nullptr, &bo);
  if (bo.nss[0] < h->prm.min_nss_factor) return KML_NSS_TOO_LOW;
  std::vector<Cand> cands;
  for (size_t d = 0; d < dbs.size(); ++d)
    select_candidates(h, dbs[d], q_robot, q_pose, bo.nss[0], &bo.entry[d * bo.K],
                      &bo.score[d * bo.K], bo.count[d], &cands);
  h->stats.total_bow_matches += cands.size();
  if ((int)cands.size() > cap) return fail(h, KML_ERR_CAPACITY, "detect_loop: output capacity");
  for (size_t i = 0; i < cands.size(); ++i) {
    out_robot[i] = cands[i].robot; out_pose[i] = cands[i].pose; out_score[i] = cands[i].score;
  }
  *count = (int)cands.size();
  return cands.empty() ? KML_NO_MATCH : KML_OK;
}

int kml_detect_loop_with_robot(kml_handle* h, uint64_t robot, uint64_t q_robot, uint64_t q_pose,
                               const uint32_t* ids, const float* vals, int n, uint64_t* out_robot,
                               uint64_t* out_pose, double* out_score, int cap, int* count) {
  KML_API_BEGIN(h)
  if (!count || !out_robot || !out_pose || !out_score) return fail(h, KML_ERR_ARG, "null output");
  *count = 0;
  RobotDb* db = get_db(h, robot, false);
  if (!db) return KML_NO_DB;
  if (h->prm.inter_robot_only && robot == q_robot) return KML_INTER_ROBOT_ONLY;
  return detect_impl(h, {db}, q_robot, q_pose, ids, vals, n, out_robot, out_pose, out_score, cap, count);
  KML_API_END(h)
}

int kml_detect_loop(kml_handle* h, uint64_t q_robot, uint64_t q_pose, const uint32_t* ids,
                    const float* vals, int n, uint64_t* out_robot, uint64_t* out_pose,
                    double* out_score, int cap, int* count) {
  KML_API_BEGIN(h)
  if (!count || !out_robot || !out_pose || !out_score) return fail(h, KML_ERR_ARG, "null output");
  std::vector<RobotDb*> dbs;
  for (auto& kv : h->sh->dbs) dbs.push_back(kv.second.get());
  return detect_impl(h, dbs, q_robot, q_pose, ids, vals, n, out_robot, out_pose, out_score, cap, count);
  KML_API_END(h)
}

int kml_compute_matched_indices(kml_handle* h, uint64_t qr, uint64_t qp, uint64_t mr, uint64_t mp,
                                uint32_t* i_query, uint32_t* i_match, int cap, int* count) {
  KML_API_BEGIN(h)
  if (!count || !i_query || !i_match) return fail(h, KML_ERR_ARG, "null output");
  *count = 0;
  QuerySide qs;
  if (!stored_query_side(h, qr, qp, &qs)) return KML_NO_FRAME;
  auto mit = h->sh->frames.find(RobotPoseId(mr, mp));
  if (mit == h->sh->frames.end()) return KML_NO_FRAME;
  if (qs.F == 0) return KML_OK;
  const int stride = qs.F;
  cudaStream_t s = h->stream;
  h->d_keys.scratch((size_t)stride * 2); h->d_jobs.scratch(1); h->d_nq.scratch(1);
  h->d_iq.scratch(stride); h->d_im.scratch(stride); h->d_M.scratch(1);
  HamJob job;
  job.q = qs.desc; job.nq = qs.F;
  job.t = h->sh->s_desc.p + (size_t)mit->second.feat_off * 32; job.nt = mit->second.F;
  job.keys = h->d_keys.p;
  KML_CUDA(cudaMemcpyAsync(h->d_jobs.p, &job, sizeof(job), cudaMemcpyHostToDevice, s));
  KML_CUDA(cudaMemcpyAsync(h->d_nq.p, &qs.F, 4, cudaMemcpyHostToDevice, s));
  if (h->prm.matcher_engine == 1) launch_hamming_jobs_tc(h->d_jobs.p, 1, s);
  else launch_hamming_jobs(h->d_jobs.p, 1, h->prm.matcher_norm, s);
  launch_lowe_compact(h->d_keys.p, h->d_nq.p, stride, h->prm.lowe_ratio, h->prm.matcher_norm, h->d_iq.p, h->d_im.p, h->d_M.p, 1, s);
  h->stats.kernel_launches += 2;
  KML_CUDA(cudaGetLastError());
  int32_t M = 0;
  std::vector<uint16_t> iq(stride), im(stride);
  KML_CUDA(cudaMemcpyAsync(&M, h->d_M.p, 4, cudaMemcpyDeviceToHost, s));
  KML_CUDA(cudaMemcpyAsync(iq.data(), h->d_iq.p, 2 * stride, cudaMemcpyDeviceToHost, s));
  KML_CUDA(cudaMemcpyAsync(im.data(), h->d_im.p, 2 * stride, cudaMemcpyDeviceToHost, s));
  KML_CUDA(cudaStreamSynchronize(s));
  if (M > cap) return fail(h, KML_ERR_CAPACITY, "compute_matched_indices: output capacity");
  for (int i = 0; i < M; ++i) { i_query[i] = iq[i]; i_match[i] = im[i]; }
  *count = M;
  return KML_OK;
  KML_API_END(h)
}

int kml_geometric_verification_nister(kml_handle* h, uint64_t qr, uint64_t qp, uint64_t mr,
                                      uint64_t mp, uint32_t* inl_q, uint32_t* inl_m, int* count,
                                      double* R) {
  KML_API_BEGIN(h)
  if (!R) return fail(h, KML_ERR_ARG, "null output");
  h->stats.total_geom_verifications_mono += 1;
  double M[12];
  int nvalid = 0;
  int rc = sac_on_lists(h, true, qr, qp, mr, mp, inl_q, inl_m, count, M, &nvalid);
  if (rc != KML_OK) return rc;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R[3 * r + c] = M[4 * r + c];
  return KML_OK;
  KML_API_END(h)
}

int kml_recover_pose(kml_handle* h, uint64_t qr, uint64_t qp, uint64_t mr, uint64_t mp,
                     uint32_t* inl_q, uint32_t* inl_m, int* count, const double* R_prior,
                     double* T) {
  KML_API_BEGIN(h)
  // adapter.setR12(prior) is not read by threept_arun; with ransac_use_1point_3d3d the
  // prior IS the rotation and one point pair per hypothesis gives the translation (row f4)
  if (!T) return fail(h, KML_ERR_ARG, "null output");
  h->stats.total_geometric_verifications += 1;
  double M[12];
  int nvalid = 0;
  int rc = sac_on_lists(h, false, qr, qp, mr, mp, inl_q, inl_m, count, M, &nvalid, R_prior);
  if (rc != KML_OK) return rc;
  memcpy(T, M, 96);
  return KML_OK;
  KML_API_END(h)
}

#ifndef KML_HOST_EMULATION
// ---------------------------------------------------------------- CUDA graph for small batches
// A batch is ~75 kernel launches, ~25 memsets and a dozen copies.  For the batches kimera_distributed
// actually sends one at a time (a single query, BASELINE.json configs[0]) the enqueue cost and the
// launch gaps ARE the latency, so batches of fewer than 16 queries replay a captured graph: staged
// H2D copies of the (fixed-layout) batch buffer, the whole pipeline, the flag word, the D2H of the
// records and counters — one cudaGraphLaunch, one wait.  The graph is keyed by everything that is
// baked into kernel arguments (batch shape, database tiling, every buffer address); it is captured
// the second time a key is seen (the first, eager, run allocates everything), re-captured when the
// key moves, and any capture failure switches the handle back to the eager path for good.
struct GraphCache {
  cudaGraphExec_t exec = nullptr;
  uint64_t key = 0, seen = 0;
  bool disabled = false;
};
static uint64_t mix64(uint64_t hsh, uint64_t v) {
  hsh ^= v + 0x9E3779B97F4A7C15ull + (hsh << 6) + (hsh >> 2);
  return hsh;
}
static uint64_t graph_key(kml_handle* h, int cap) {
  kml_shared& sh = *h->sh;
  uint64_t k = 0x5EEDull;
  const void* ptrs[] = {h->d_in.p, h->d_blocks.p, h->d_stats.p, h->d_dbs.p, h->d_bow_entry.p, h->d_bow_score.p, h->d_bow_count.p,
                        h->d_nss.p, h->d_keys.p, h->d_jobs.p, h->d_pairs.p, h->d_nq.p, h->d_iq.p, h->d_im.p, h->d_kq.p,
                        h->d_km.p, h->d_M.p, h->d_N3.p, h->d_mono_ok.p, h->d_a.p, h->d_b.p, h->d_perm.p, h->d_samples.p,
                        h->d_models.p, h->d_fb_list.p, h->d_nsol.p, h->d_esol.p, h->d_brk.p, h->d_item_base.p,
                        h->d_item_list.p, h->d_item_q.p, h->d_item_model.p, h->d_item_status.p, h->d_valid.p, h->d_counts.p,
                        h->d_active.p, h->d_st_mono.p, h->d_st_stereo.p, h->d_best_mono.p, h->d_best_stereo.p,
                        h->d_mask_mono.p, h->d_mask_stereo.p, h->d_inl_mono.p, h->d_inl_stereo.p, h->h_in.p, h->h_recs.p,
                        h->h_stats.p, sh.s_desc.p, sh.s_bear.p, sh.s_pts.p, sh.s_off.p, sh.s_F.p, sh.d_raw.p,
                        sh.d_ktable_mono.p, sh.d_ktable_stereo.p, sh.d_ktable_stereo1.p, sh.d_samptab[0].p,
                        sh.d_samptab[1].p, sh.d_samptab[2].p};
  for (const void* p : ptrs) k = mix64(k, (uint64_t)(uintptr_t)p);
  const uint64_t vals[] = {(uint64_t)h->B, (uint64_t)h->qF, (uint64_t)cap, (uint64_t)h->views_n_db, (uint64_t)h->views_tile,
                           (uint64_t)h->views_n_tiles, (uint64_t)h->item_cap, (uint64_t)sh.ktable_n_mono,
                           (uint64_t)sh.ktable_n_stereo, (uint64_t)sh.ktable_n_stereo1, (uint64_t)sh.samptab_nmax[0],
                           (uint64_t)sh.samptab_nmax[1], (uint64_t)sh.samptab_nmax[2]};
  for (uint64_t v : vals) k = mix64(k, v);
  return k ? k : 1;
}
// The local pipeline of a throughput batch (B >= 16: BoW scoring ... records in this rank's block, ~75
// kernels and ~25 memsets) as ONE graph launch.  Several lanes enqueue batches concurrently, and with
// eager launches the host side (a few hundred driver calls per batch, all lanes behind one context
// lock) is what the GPU waits for between the short RANSAC kernels.  The copies, the exchange of a
// sharded query and the host-driven continuation stay outside: batch_run goes on exactly as after an
// eager enqueue, with the BatchPlan remembered from the capture.  Same keying as the small-batch graph.
struct EnqueueGraph {
  cudaGraphExec_t exec = nullptr;
  uint64_t key = 0, seen = 0;
  bool disabled = false;
  BatchPlan plan;
  unsigned long long launches = 0;
};
static bool batch_enqueue_graph(kml_handle* h, int cap, uint8_t* blk, BatchPlan* pl) {
  if (h->B < 16 || getenv("KML_NO_GRAPH")) return false;
  if (!h->enqueue_graph) h->enqueue_graph = new EnqueueGraph();
  EnqueueGraph* g = static_cast<EnqueueGraph*>(h->enqueue_graph);
  if (g->disabled) return false;
  if (!h->d_in.p || !h->d_stats.p) { g->seen = 0; return false; }
  ensure_frame_offsets(h);
  ensure_views(h);
  const uint64_t key = mix64(graph_key(h, cap), (uint64_t)(uintptr_t)blk);
  cudaStream_t s = h->stream;
  if (!(g->exec && g->key == key)) {
    if (g->seen != key) {  // first sight of this shape: the eager run allocates and builds every table
      g->seen = key;
      return false;
    }
    if (g->exec) { cudaGraphExecDestroy(g->exec); g->exec = nullptr; }
    cudaGraph_t graph = nullptr;
    const unsigned long long l0 = h->stats.kernel_launches;
    bool ok = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
      h->capturing = true;
      try {
        batch_enqueue(h, cap, blk, &g->plan);
      } catch (const std::exception&) {
        ok = false;
      }
      h->capturing = false;
      if (cudaStreamEndCapture(s, &graph) != cudaSuccess || !graph) ok = false;
    }
    g->launches = h->stats.kernel_launches - l0;
    h->stats.kernel_launches = l0;
    if (ok && cudaGraphInstantiate(&g->exec, graph, 0) != cudaSuccess) ok = false;
    if (graph) cudaGraphDestroy(graph);
    if (!ok || mix64(graph_key(h, cap), (uint64_t)(uintptr_t)blk) != key) {  // capture failed, or something moved under it
      cudaGetLastError();
      if (g->exec) { cudaGraphExecDestroy(g->exec); g->exec = nullptr; }
      g->disabled = !ok;
      g->seen = 0;
      return false;
    }
    g->key = key;
  }
  KML_CUDA(cudaEventRecord(h->ev[6], s));
  KML_CUDA(cudaGraphLaunch(g->exec, s));
  KML_CUDA(cudaEventRecord(h->ev[7], s));
  h->stats.kernel_launches += g->launches;
  *pl = g->plan;
  return true;
}
void kml_graph_destroy_internal(kml_handle* h) {
  if (EnqueueGraph* e = static_cast<EnqueueGraph*>(h->enqueue_graph)) {
    if (e->exec) cudaGraphExecDestroy(e->exec);
    delete e;
    h->enqueue_graph = nullptr;
  }
  GraphCache* g = static_cast<GraphCache*>(h->graph_cache);
  if (!g) return;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  delete g;
  h->graph_cache = nullptr;
}
// Tries the graph path for the batch staged in h->h_in.  Returns 1 if the records were produced by
// a graph replay (out / counts filled), 0 if the caller must run the eager path, < 0 on error.
static int batch_run_graph(kml_handle* h, int cap, kml_result* out, int32_t* counts) {
  static const bool off = getenv("KML_NO_GRAPH") != nullptr;
  if (off || h->B <= 0 || h->B >= 16) return 0;
  if (!h->graph_cache) h->graph_cache = new GraphCache();
  GraphCache* g = static_cast<GraphCache*>(h->graph_cache);
  if (g->disabled) return 0;
  const int B = h->B;
  const BlockLayout bl = block_layout(B, cap);
  if (!h->d_in.p || !h->d_blocks.p || h->d_blocks.cap < bl.blk || !h->h_recs.p || !h->d_stats.p) { g->seen = 0; return 0; }
  ensure_frame_offsets(h);
  ensure_views(h);
  const uint64_t key = graph_key(h, cap);
  cudaStream_t s = h->stream;
  if (!(g->exec && g->key == key)) {
    if (g->seen != key) {  // first sight of this shape: the eager run allocates and builds every table
      g->seen = key;
      return 0;
    }
    if (g->exec) { cudaGraphExecDestroy(g->exec); g->exec = nullptr; }
    const BatchIn L0 = batch_layout(nullptr, B, h->qF);
    cudaGraph_t graph = nullptr;
    bool ok = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
      h->capturing = true;
      try {
        KML_CUDA(cudaMemcpyAsync(h->d_in.p, h->h_in.p, L0.bytes, cudaMemcpyHostToDevice, s));
        BatchPlan pl;
        batch_enqueue(h, cap, h->d_blocks.p, &pl);
        if (pl.P > 0) launch_block_flags(h->d_stats.p, pl.mono.overflow, reinterpret_cast<int32_t*>(h->d_blocks.p + bl.err_off), s);
        KML_CUDA(cudaMemcpyAsync(h->h_recs.p, h->d_blocks.p, bl.err_off + sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        KML_CUDA(cudaMemcpyAsync(h->h_stats.p, h->d_stats.p, sizeof(BatchStats), cudaMemcpyDeviceToHost, s));
      } catch (const std::exception&) {
        ok = false;
      }
      h->capturing = false;
      if (cudaStreamEndCapture(s, &graph) != cudaSuccess || !graph) ok = false;
    }
    if (ok && cudaGraphInstantiate(&g->exec, graph, 0) != cudaSuccess) ok = false;
    if (graph) cudaGraphDestroy(graph);
    if (!ok || graph_key(h, cap) != key) {  // capture failed, or something was (re)allocated under it
      cudaGetLastError();
      if (g->exec) { cudaGraphExecDestroy(g->exec); g->exec = nullptr; }
      g->disabled = !ok;
      g->seen = 0;
      return 0;
    }
    g->key = key;
  }
  KML_CUDA(cudaEventRecord(h->ev[6], s));
  KML_CUDA(cudaGraphLaunch(g->exec, s));
  KML_CUDA(cudaEventRecord(h->ev[7], s));
  h->stats.kernel_launches += 75;  // the pipeline's kernels, replayed
  h->wait_stream();
  const int32_t flag = *reinterpret_cast<const int32_t*>(h->h_recs.p + bl.err_off);
  if (flag != kBlkOk) return 0;  // pending RANSAC problems / item overflow (rare): the eager path owns the continuation
  batch_stats_to_host(h, *h->h_stats.p, true);
  memcpy(out, h->h_recs.p, bl.rec_bytes);
  memcpy(counts, h->h_recs.p + bl.counts_off, sizeof(int32_t) * (size_t)B);
  return 1;
}
#else
void kml_graph_destroy_internal(kml_handle*) {}
static bool batch_enqueue_graph(kml_handle*, int, uint8_t*, BatchPlan*) { return false; }
#endif

int kml_query_batch_upload(kml_handle* h, int B, const uint64_t* q_robot, const uint64_t* q_pose,
                           const int64_t* bow_off, const uint32_t* ids, const float* vals,
                           const int64_t* prev_off, const uint32_t* prev_ids,
                           const float* prev_vals, const uint8_t* desc, const double* bearings,
                           const double* points, int F) {
  KML_API_BEGIN(h)
  return batch_upload(h, B, q_robot, q_pose, bow_off, ids, vals, prev_off, prev_ids, prev_vals,
                      desc, bearings, points, F, true);
  KML_API_END(h)
}

int kml_query_batch_run(kml_handle* h, kml_result* out, int cap, int32_t* counts) {
  KML_API_BEGIN(h)
  if (cap <= 0 || (h->B > 0 && (!out || !counts))) return fail(h, KML_ERR_ARG, "query_batch_run: bad output");
  return batch_run(h, cap, false, out, counts);
  KML_API_END(h)
}

int kml_query_batch(kml_handle* h, int B, const uint64_t* q_robot, const uint64_t* q_pose,
                    const int64_t* bow_off, const uint32_t* ids, const float* vals,
                    const int64_t* prev_off, const uint32_t* prev_ids, const float* prev_vals,
                    const uint8_t* desc, const double* bearings, const double* points, int F,
                    kml_result* out, int cap, int32_t* counts) {
  KML_API_BEGIN(h)
  if (cap <= 0 || (B > 0 && (!out || !counts))) return fail(h, KML_ERR_ARG, "query_batch: bad output");
  // the copies are only enqueued here: batch_run's single wait at the end covers them, and the
  // caller's buffers are not read after this function returns
#ifndef KML_HOST_EMULATION
  const bool small = B > 0 && B < 16;  // latency path: staged batch + captured graph
#else
  const bool small = false;
#endif
  int rc = batch_upload(h, B, q_robot, q_pose, bow_off, ids, vals, prev_off, prev_ids, prev_vals,
                        desc, bearings, points, F, false, /*stage_only=*/small);
  if (rc != KML_OK) return rc;
#ifndef KML_HOST_EMULATION
  if (small) {  // latency path: replay the captured graph of this batch shape (or fall through)
    const int g = batch_run_graph(h, cap, out, counts);
    if (g < 0) return g;
    if (g == 1) return KML_OK;
    // eager: the batch is staged in h_in, copy it over as batch_upload would have
    const BatchIn L0 = batch_layout(nullptr, B, F);
    KML_CUDA(cudaMemcpyAsync(h->d_in.p, h->h_in.p, L0.bytes, cudaMemcpyHostToDevice, h->stream));
  }
#endif
  rc = batch_run(h, cap, false, out, counts);
  if (rc != KML_OK) cudaStreamSynchronize(h->stream);  // an early error return must not leave copies in flight
  return rc;
  KML_API_END(h)
}

// Merge rule of the sharded query: per query, the union of every rank's records re-ranked by
// (score desc, match robot asc, match pose asc), best `cap` kept.  Block r starts at
// base + r*blk_stride: [B][cap_in] records, then (at rec_bytes) B int32 counts.  Pointers are
// sorted, not the 200-byte records, and nothing is allocated per query.
static inline bool rec_before(const kml_result* x, const kml_result* y) {
  if (x->norm_bow_score != y->norm_bow_score) return x->norm_bow_score > y->norm_bow_score;
  if (x->m_robot != y->m_robot) return x->m_robot < y->m_robot;
  return x->m_pose < y->m_pose;
}
static void merge_rank_blocks(const uint8_t* base, size_t blk_stride, size_t rec_bytes, int nr, int B,
                              int cap_in, int cap, kml_result* out, int32_t* counts) {
  std::vector<const kml_result*> ptr, head(nr), end(nr);
  ptr.reserve((size_t)nr * cap_in);
  for (int b = 0; b < B; ++b) {
    // every rank emits its list already ranked by the same key: a k-way merge of the heads needs
    // cap * nranks comparisons; a list found out of order sends the query to the full sort
    bool sorted = true;
    for (int r = 0; r < nr; ++r) {
      const uint8_t* blk = base + blk_stride * r;
      const kml_result* recs = reinterpret_cast<const kml_result*>(blk) + (size_t)b * cap_in;
      const int32_t c = reinterpret_cast<const int32_t*>(blk + rec_bytes)[b];
      head[r] = recs;
      end[r] = recs + c;
      for (int i = 1; i < c && sorted; ++i) sorted = !rec_before(recs + i, recs + i - 1);
    }
    int c = 0;
    if (sorted) {
      while (c < cap) {
        int br = -1;
        for (int r = 0; r < nr; ++r)  // strict "before" keeps the lower rank on exact ties
          if (head[r] != end[r] && (br < 0 || rec_before(head[r], head[br]))) br = r;
        if (br < 0) break;
        out[(size_t)b * cap + c++] = *head[br]++;
      }
    } else {
      ptr.clear();
      for (int r = 0; r < nr; ++r)
        for (const kml_result* p = head[r]; p != end[r]; ++p) ptr.push_back(p);
      // (rank, position) breaks exact ties of the key, which keeps the order of a stable sort
      std::sort(ptr.begin(), ptr.end(), [](const kml_result* x, const kml_result* y) {
        if (rec_before(x, y)) return true;
        if (rec_before(y, x)) return false;
        return x < y;
      });
      c = (int)std::min<size_t>(ptr.size(), (size_t)cap);
      for (int i = 0; i < c; ++i) out[(size_t)b * cap + i] = *ptr[i];
    }
    counts[b] = c;
  }
}

// Host-side entry to the same merge (tests, callers that move the blocks themselves): `blocks`
// holds nranks consecutive blocks of [B][cap_in] records followed by B int32 counts each.
int kml_merge_shard_records(const void* blocks, int nranks, int B, int cap_in, int cap, kml_result* out,
                            int32_t* counts) {
  if (!blocks || nranks < 1 || B < 0 || cap_in < 1 || cap < 1 || (B > 0 && (!out || !counts))) return KML_ERR_ARG;
  const size_t rec_bytes = sizeof(kml_result) * (size_t)B * cap_in;
  const size_t blk = rec_bytes + sizeof(int32_t) * (size_t)B;
  merge_rank_blocks(static_cast<const uint8_t*>(blocks), blk, rec_bytes, nranks, B, cap_in, cap, out, counts);
  return KML_OK;
}

// Sharded query: every rank runs the (identical) uploaded batch against its own robot databases
// and writes its records into its block of the all-gather buffer; ONE in-place ncclAllGather
// exchanges the blocks and a device kernel re-ranks the union per query (batch_run).
int kml_query_batch_sharded(kml_handle* h, kml_result* out, int cap, int32_t* counts) {
  KML_API_BEGIN(h)
  if (cap <= 0 || (h->B > 0 && (!out || !counts))) return fail(h, KML_ERR_ARG, "bad output");
  return batch_run(h, cap, true, out, counts);
  KML_API_END(h)
}

int kml_query_batch_sharded_seq(kml_handle* h, uint64_t seq, kml_result* out, int cap, int32_t* counts) {
  KML_API_BEGIN(h)
  if (cap <= 0 || (h->B > 0 && (!out || !counts))) return fail(h, KML_ERR_ARG, "bad output");
  if (h->B == 0) {  // nothing to exchange, but the sequence must still advance
    {
      std::unique_lock<std::mutex> lk(h->sh->seq_mu);
      h->sh->seq_cv.wait(lk, [&] { return h->sh->seq_next >= seq; });
      if (h->sh->seq_next == seq) h->sh->seq_next = seq + 1;
    }
    h->sh->seq_cv.notify_all();
    return KML_OK;
  }
  return batch_run(h, cap, true, out, counts, &seq);
  KML_API_END(h)
}
int kml_query_batch_sharded_host(kml_handle* h, int use_seq, uint64_t seq, int B, const uint64_t* q_robot,
                                 const uint64_t* q_pose, const int64_t* bow_off, const uint32_t* ids,
                                 const float* vals, const int64_t* prev_off, const uint32_t* prev_ids,
                                 const float* prev_vals, const uint8_t* desc, const double* bearings,
                                 const double* points, int F, kml_result* out, int cap, int32_t* counts) {
  KML_API_BEGIN(h)
  if (cap <= 0 || (B > 0 && (!out || !counts))) return fail(h, KML_ERR_ARG, "query_batch_sharded_host: bad output");
  int rc = batch_upload(h, B, q_robot, q_pose, bow_off, ids, vals, prev_off, prev_ids, prev_vals,
                        desc, bearings, points, F, false);
  if (rc != KML_OK) return rc;  // argument errors are the same on every rank (same batch): nobody enters the collective
  if (B == 0) return use_seq ? kml_query_batch_sharded_seq(h, seq, out, cap, counts) : KML_OK;
  rc = batch_run(h, cap, true, out, counts, use_seq ? &seq : nullptr);
  if (rc != KML_OK) cudaStreamSynchronize(h->stream);
  return rc;
  KML_API_END(h)
}
int kml_comm_seq_reset(kml_handle* h, uint64_t next_seq) {
  if (!h) return KML_ERR_ARG;
  {
    std::lock_guard<std::mutex> lk(h->sh->seq_mu);
    h->sh->seq_next = next_seq;
  }
  h->sh->seq_cv.notify_all();
  return KML_OK;
}

// merge_shards_kernel on host blocks (the device tail of the sharded query, callable without NCCL)
int kml_merge_shard_records_device(kml_handle* h, const void* blocks, int nranks, int B, int cap_in, int cap,
                                   kml_result* out, int32_t* counts) {
  KML_API_BEGIN(h)
  if (!blocks || nranks < 1 || B < 0 || cap_in < 1 || cap < 1 || (B > 0 && (!out || !counts)))
    return fail(h, KML_ERR_ARG, "merge_shard_records_device: bad argument");
  if (B == 0) return KML_OK;
  const size_t rec_in = sizeof(kml_result) * (size_t)B * cap_in;
  const size_t blk = rec_in + sizeof(int32_t) * (size_t)B;       // host layout: records | counts
  const size_t dblk = (blk + sizeof(int32_t) + 255) / 256 * 256;  // device layout adds the flag word
  const BlockLayout ol = block_layout(B, cap);
  cudaStream_t s = h->stream;
  h->d_blocks.scratch(dblk * nranks);
  h->d_merged.scratch(ol.blk);
  KML_CUDA(cudaMemsetAsync(h->d_blocks.p, 0, dblk * nranks, s));
  for (int r = 0; r < nranks; ++r)
    KML_CUDA(cudaMemcpyAsync(h->d_blocks.p + dblk * r, static_cast<const uint8_t*>(blocks) + blk * r, blk,
                             cudaMemcpyHostToDevice, s));
  MergeArgs m;
  m.base = h->d_blocks.p; m.blk_stride = dblk; m.counts_off = rec_in; m.err_off = blk;
  m.nranks = nranks; m.B = B; m.cap_in = cap_in; m.cap = cap;
  m.out = reinterpret_cast<kml_result*>(h->d_merged.p);
  m.counts = reinterpret_cast<int32_t*>(h->d_merged.p + ol.counts_off);
  m.err_out = reinterpret_cast<int32_t*>(h->d_merged.p + ol.err_off);
  KML_CUDA(cudaMemsetAsync(h->d_merged.p, 0, ol.blk, s));
  launch_merge(m, s);
  h->stats.kernel_launches += 1;
  KML_CUDA(cudaGetLastError());
  KML_CUDA(cudaMemcpyAsync(out, h->d_merged.p, ol.rec_bytes, cudaMemcpyDeviceToHost, s));
  KML_CUDA(cudaMemcpyAsync(counts, h->d_merged.p + ol.counts_off, sizeof(int32_t) * (size_t)B, cudaMemcpyDeviceToHost, s));
  KML_CUDA(cudaStreamSynchronize(s));
  return KML_OK;
  KML_API_END(h)
}

// Pinned host memory for the caller's batch buffers (and result block): arrays allocated here are
// copied to the device from where they are, without the staging copy pageable memory needs.
int kml_host_alloc(size_t bytes, void** out) {
  if (!out) return KML_ERR_ARG;
  *out = nullptr;
  if (cudaMallocHost(out, bytes ? bytes : 1) != cudaSuccess) {
    cudaGetLastError();
    return KML_ERR_CUDA;
  }
  return KML_OK;
}
int kml_host_free(void* p) {
  if (p && cudaFreeHost(p) != cudaSuccess) {
    cudaGetLastError();
    return KML_ERR_CUDA;
  }
  return KML_OK;
}

// ------------------------------------------------- batched RANSAC (C4 etc.)
static int ransac_batch(kml_handle* h, bool mono, int P, int N, const double* a, const double* b,
                        int full, double* models, int32_t* n_inliers, int32_t* iterations,
                        int32_t* best_draw, uint32_t* inlier_mask, float* ms_kernel, const double* R9 = nullptr) {
  if (P < 0 || N < 0 || N > 65535 || (P > 0 && N > 0 && (!a || !b))) return fail(h, KML_ERR_ARG, "ransac_batch: bad argument");
  if (P == 0) return KML_OK;
  const int stride = std::max(N, 8);
  const int mask_words = (stride + 31) / 32;
  cudaStream_t s = h->stream;
  h->d_a.scratch((size_t)P * stride * 3); h->d_b.scratch((size_t)P * stride * 3);
  h->d_N3.scratch(P);
  std::vector<int32_t> Ns(P, N);
  KML_CUDA(cudaMemcpyAsync(h->d_N3.p, Ns.data(), 4 * P, cudaMemcpyHostToDevice, s));
  if (N) {
    KML_CUDA(cudaMemcpy2DAsync(h->d_a.p, (size_t)stride * 24, a, (size_t)N * 24, (size_t)N * 24, P, cudaMemcpyHostToDevice, s));
    KML_CUDA(cudaMemcpy2DAsync(h->d_b.p, (size_t)stride * 24, b, (size_t)N * 24, (size_t)N * 24, P, cudaMemcpyHostToDevice, s));
  }
  SacBufs bufs = mono ? SacBufs{&h->d_st_mono, &h->d_best_mono, &h->d_mask_mono, &h->d_inl_mono}
                      : SacBufs{&h->d_st_stereo, &h->d_best_stereo, &h->d_mask_stereo, &h->d_inl_stereo};
  const double* d_prior = nullptr;
  if (R9 && !mono) {  // 1-point problem: per-problem rotation as the left block of a 3x4
    std::vector<double> pr((size_t)P * 12, 0.0);
    for (int p = 0; p < P; ++p)
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) pr[(size_t)p * 12 + 4 * r + c] = R9[(size_t)p * 9 + 3 * r + c];
    h->d_prior.scratch(pr.size());
    KML_CUDA(cudaMemcpyAsync(h->d_prior.p, pr.data(), pr.size() * 8, cudaMemcpyHostToDevice, s));
    KML_CUDA(cudaStreamSynchronize(s));
    d_prior = h->d_prior.p;
  }
  KML_CUDA(cudaEventRecord(h->ev[2], s));
  run_sac_sync(h, mono, P, h->d_a.p, h->d_b.p, h->d_N3.p, stride, full, bufs, mask_words, d_prior);
  KML_CUDA(cudaEventRecord(h->ev[3], s));
  std::vector<SacState> st(P);
  std::vector<uint32_t> mask((size_t)P * mask_words);
  KML_CUDA(cudaMemcpyAsync(st.data(), bufs.st->p, sizeof(SacState) * P, cudaMemcpyDeviceToHost, s));
  KML_CUDA(cudaMemcpyAsync(mask.data(), bufs.mask->p, 4 * mask.size(), cudaMemcpyDeviceToHost, s));
  if (models) KML_CUDA(cudaMemcpyAsync(models, bufs.best->p, 96 * (size_t)P, cudaMemcpyDeviceToHost, s));
  if (n_inliers) KML_CUDA(cudaMemcpyAsync(n_inliers, bufs.inl->p, 4 * P, cudaMemcpyDeviceToHost, s));
  KML_CUDA(cudaStreamSynchronize(s));
  float ms = 0;
  KML_CUDA(cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]));
  if (ms_kernel) *ms_kernel = ms;
  uint64_t draws = 0;
  const int words_out = (N + 31) / 32;
  for (int p = 0; p < P; ++p) {
    if (st[p].exhausted) return fail(h, KML_ERR_STREAM_EXHAUSTED, "pre-drawn sample stream exhausted");
    if (iterations) iterations[p] = st[p].iterations;
    if (best_draw) best_draw[p] = st[p].best_draw;
    draws += st[p].draws;
    if (inlier_mask)
      for (int w = 0; w < std::max(words_out, 1); ++w)
        inlier_mask[(size_t)p * std::max(words_out, 1) + w] = w < words_out ? mask[(size_t)p * mask_words + w] : 0u;
  }
  if (mono) { h->stats.mono_hypotheses_last = draws; h->stats.mono_residuals_last = draws * (uint64_t)N; }
  else { h->stats.stereo_hypotheses_last = draws; h->stats.stereo_residuals_last = draws * (uint64_t)N; }
  h->stats.pairs_last = P;
  return KML_OK;
}

int kml_ransac_arun_batch(kml_handle* h, int P, int N, const double* p1, const double* p2,
                          int full, double* models, int32_t* n_inliers, int32_t* iterations,
                          int32_t* best_draw, uint32_t* inlier_mask, float* ms_kernel) {
  KML_API_BEGIN(h)
  return ransac_batch(h, false, P, N, p1, p2, full, models, n_inliers, iterations, best_draw, inlier_mask, ms_kernel);
  KML_API_END(h)
}
int kml_ransac_onepoint_batch(kml_handle* h, int P, int N, const double* p1, const double* p2, const double* R,
                              int full, double* models, int32_t* n_inliers, int32_t* iterations,
                              int32_t* best_draw, uint32_t* inlier_mask, float* ms_kernel) {
  KML_API_BEGIN(h)
  if (P > 0 && !R) return fail(h, KML_ERR_ARG, "ransac_onepoint_batch: null rotation");
  return ransac_batch(h, false, P, N, p1, p2, full, models, n_inliers, iterations, best_draw, inlier_mask, ms_kernel, R);
  KML_API_END(h)
}
int kml_ransac_nister_batch(kml_handle* h, int P, int N, const double* f1, const double* f2,
                            int full, double* models, int32_t* n_inliers, int32_t* iterations,
                            int32_t* best_draw, uint32_t* inlier_mask, float* ms_kernel) {
  KML_API_BEGIN(h)
  return ransac_batch(h, true, P, N, f1, f2, full, models, n_inliers, iterations, best_draw, inlier_mask, ms_kernel);
  KML_API_END(h)
}

}  // extern "C"
