// oracle/src/oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE; see oracle/kmo.h).
// Scalar restatement of SURVEY.md Appendix A.1-A.8.  The upstream sources
// (DBoW2 TemplatedDatabase.h / ScoringObject.cpp, OpenCV matchers.cpp,
// OpenGV sac/Ransac.hpp + SampleConsensusProblem.hpp, Kimera-Multi-LCD
// loop_closure_detector.cpp) are named by /root/reference/kimera_multi.repos
// but not vendored; the only literal reference lines for this path are
// /root/reference/docker/copy/kimera_multi_lcd.patch:30-38.
#include "../kmo.h"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <random>
#include <unordered_map>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "geom.hpp"

using namespace kmo;

// =========================================================================
// A.2  L1Scoring::score  (DBoW2/src/ScoringObject.cpp)
// =========================================================================
static double l1_score(const uint32_t* ids1, const float* v1, int n1, const uint32_t* ids2,
                       const float* v2, int n2) {
  double score = 0.0;
  int i = 0, j = 0;
  while (i < n1 && j < n2) {
    if (ids1[i] == ids2[j]) {
      double vi = (double)v1[i], wi = (double)v2[j];
      score += std::fabs(vi - wi) - std::fabs(vi) - std::fabs(wi);
      ++i; ++j;
    } else if (ids1[i] < ids2[j]) {
      ++i;
    } else {
      ++j;
    }
  }
  return -score / 2.0;
}

// =========================================================================
// A.1  TemplatedDatabase::add / queryL1  (DBoW2/include/DBoW2/TemplatedDatabase.h)
// =========================================================================
struct kmo_db {
  struct IFPair { uint32_t entry; double weight; };
  std::unordered_map<uint32_t, std::vector<IFPair>> ifile;  // word -> row (ascending entry)
  uint32_t nentries = 0;
};

static int db_query(const kmo_db* db, const uint32_t* ids, const float* vals, int n,
                    int max_results, int max_id, uint32_t* out_entry, double* out_score,
                    int cap) {
  std::map<uint32_t, double> pairs;
  for (int k = 0; k < n; ++k) {  // query words ascending (caller guarantees sorted ids)
    auto it = db->ifile.find(ids[k]);
    if (it == db->ifile.end()) continue;
    const double q = (double)vals[k];
    for (const auto& p : it->second) {
      if (max_id == -1 || (int)p.entry < max_id) {
        double value = std::fabs(q - p.weight) - std::fabs(q) - std::fabs(p.weight);
        auto pit = pairs.lower_bound(p.entry);
        if (pit != pairs.end() && !(pairs.key_comp()(p.entry, pit->first)))
          pit->second += value;
        else
          pairs.insert(pit, std::make_pair(p.entry, value));
      }
    }
  }
  std::vector<std::pair<double, uint32_t>> ret;
  ret.reserve(pairs.size());
  for (auto& kv : pairs) ret.emplace_back(kv.second, kv.first);
  // ascending raw score (most negative = best); exact ties -> lower entry id
  std::sort(ret.begin(), ret.end());
  if (max_results > 0 && (int)ret.size() > max_results) ret.resize(max_results);
  int cnt = 0;
  for (auto& r : ret) {
    if (cnt >= cap) break;
    out_entry[cnt] = r.second;
    out_score[cnt] = -r.first / 2.0;
    ++cnt;
  }
  return cnt;
}

// =========================================================================
// A.4 / B.1  BFMatcher(NORM_HAMMING).knnMatch(k=2) + Lowe ratio
// =========================================================================
static inline int hamming256(const uint8_t* a, const uint8_t* b) {
  // cv::hal::normHamming on 32 bytes (4 x 64-bit popcount)
  uint64_t x[4], y[4];
  memcpy(x, a, 32);
  memcpy(y, b, 32);
  return __builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) +
         __builtin_popcountll(x[2] ^ y[2]) + __builtin_popcountll(x[3] ^ y[3]);
}

static inline int l1_32(const uint8_t* a, const uint8_t* b) {  // cv::hal::normL1 on 32 bytes
  int d = 0;
  for (int i = 0; i < 32; ++i) d += std::abs((int)a[i] - (int)b[i]);
  return d;
}

static void knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, uint32_t* idx,
                 uint16_t* dist, int norm = 0) {
  for (int i = 0; i < nq; ++i) {
    int d0 = INT_MAX, d1 = INT_MAX;
    uint32_t i0 = 0xFFFFFFFFu, i1 = 0xFFFFFFFFu;
    for (int j = 0; j < nt; ++j) {
      int d = norm == 1 ? l1_32(q + 32 * i, t + 32 * j) : hamming256(q + 32 * i, t + 32 * j);
      if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = (uint32_t)j; }
      else if (d < d1) { d1 = d; i1 = (uint32_t)j; }
    }
    idx[2 * i] = i0; idx[2 * i + 1] = i1;
    dist[2 * i] = (d0 == INT_MAX) ? 0xFFFF : (uint16_t)d0;
    dist[2 * i + 1] = (d1 == INT_MAX) ? 0xFFFF : (uint16_t)d1;
  }
}

static int match_lowe(const uint8_t* q, int nq, const uint8_t* t, int nt, double lowe,
                      uint32_t* iq, uint32_t* im, int norm = 0) {
  std::vector<uint32_t> idx(2 * (size_t)std::max(nq, 1));
  std::vector<uint16_t> dist(2 * (size_t)std::max(nq, 1));
  knn2(q, nq, t, nt, idx.data(), dist.data(), norm);
  int cnt = 0;
  for (int i = 0; i < nq; ++i) {
    if (idx[2 * i + 1] == 0xFFFFFFFFu) continue;  // fewer than 2 neighbours
    // float distance promoted to double, strict <
    if ((double)(float)dist[2 * i] < lowe * (double)(float)dist[2 * i + 1]) {
      iq[cnt] = (uint32_t)i;
      im[cnt] = idx[2 * i];
      ++cnt;
    }
  }
  return cnt;
}

// =========================================================================
// A.7  SampleConsensusProblem sampling (portable stream)
// =========================================================================
struct Sampler {
  std::mt19937 rng;
  std::vector<int> shuffled;
  int s;
  Sampler(int N, int s_, uint32_t seed) : rng(seed), shuffled(N), s(s_) {
    for (int i = 0; i < N; ++i) shuffled[i] = i;
  }
  void draw(uint16_t* out) {
    const int N = (int)shuffled.size();
    for (int i = 0; i < s; ++i) {
      uint32_t r = (uint32_t)rng() >> 1;  // == uniform_int_distribution<int>(0,INT_MAX) on libstdc++ 13
      int j = i + (int)(r % (uint32_t)(N - i));
      std::swap(shuffled[i], shuffled[j]);
    }
    for (int i = 0; i < s; ++i) out[i] = (uint16_t)shuffled[i];
  }
};

// =========================================================================
// A.5  sac::Ransac::computeModel — generic over the two problems
// =========================================================================
template <class Problem>
static void ransac(Problem& prob, int N, int sample_size, double threshold, double probability,
                   int max_iterations, uint32_t seed, kmo_ransac_result* res,
                   uint32_t* inliers) {
  memset(res, 0, sizeof(*res));
  res->best_draw = -1;
  int iterations = 0, best = -INT_MAX, skipped = 0, draws = 0;
  double k = 1.0;
  const int max_skip = max_iterations * 10;
  bool have_model = false;
  double model[12], best_model[12];
  uint16_t sample[16];
  if (N >= sample_size) {
    Sampler sampler(N, sample_size, seed);
    while ((double)iterations < k && skipped < max_skip) {
      sampler.draw(sample);
      const int this_draw = draws++;
      if (!prob.compute(sample, model)) { ++skipped; continue; }
      int n = prob.count(model, threshold);
      if (n > best) {
        best = n;
        have_model = true;
        res->best_draw = this_draw;
        memcpy(best_model, model, sizeof(model));
        double w = (double)n / (double)N;
        double p_no = 1.0 - std::pow(w, (double)sample_size);
        p_no = std::max(DBL_EPSILON, p_no);
        p_no = std::min(1.0 - DBL_EPSILON, p_no);
        k = std::log(1.0 - probability) / std::log(p_no);
      }
      ++iterations;
      if (iterations > max_iterations) break;
    }
  }
  res->iterations = iterations;
  res->skipped = skipped;
  res->draws = draws;
  res->success = have_model ? 1 : 0;
  if (!have_model) return;
  memcpy(res->model, best_model, sizeof(best_model));
  res->n_inliers = prob.select(best_model, threshold, inliers);
}

struct ArunProblem {
  const double* p1; const double* p2; int N;
  bool compute(const uint16_t* s, double* model) const {
    arun3(p1 + 3 * s[0], p1 + 3 * s[1], p1 + 3 * s[2], p2 + 3 * s[0], p2 + 3 * s[1],
          p2 + 3 * s[2], model);
    return true;
  }
  int count(const double* m, double thr) const {
    int c = 0;
    for (int i = 0; i < N; ++i) c += (std::sqrt(arun_sqdist(m, p1 + 3 * i, p2 + 3 * i)) < thr);
    return c;
  }
  int select(const double* m, double thr, uint32_t* out) const {
    int c = 0;
    for (int i = 0; i < N; ++i)
      if (std::sqrt(arun_sqdist(m, p1 + 3 * i, p2 + 3 * i)) < thr) out[c++] = (uint32_t)i;
    return c;
  }
};

// Row f4 (/root/reference/params/D455/LcdParams.yaml:58 ransac_use_1point_3d3d): the point-cloud
// problem with the rotation given (recoverPose's R_prior = the mono rotation).  One correspondence
// fixes the translation, t = p1_i - R p2_i, so the sample size is 1; residual and thresholds are
// PointCloudSacProblem's (A.8).
struct OnePointProblem {
  const double* p1; const double* p2; int N; double R[9];
  bool compute(const uint16_t* s, double* model) const {
    const double* a = p1 + 3 * s[0];
    const double* b = p2 + 3 * s[0];
    for (int r = 0; r < 3; ++r) {
      model[4 * r + 0] = R[3 * r + 0]; model[4 * r + 1] = R[3 * r + 1]; model[4 * r + 2] = R[3 * r + 2];
      model[4 * r + 3] = a[r] - ((R[3 * r + 0] * b[0] + R[3 * r + 1] * b[1]) + R[3 * r + 2] * b[2]);
    }
    return true;
  }
  int count(const double* m, double thr) const {
    int c = 0;
    for (int i = 0; i < N; ++i) c += (std::sqrt(arun_sqdist(m, p1 + 3 * i, p2 + 3 * i)) < thr);
    return c;
  }
  int select(const double* m, double thr, uint32_t* out) const {
    int c = 0;
    for (int i = 0; i < N; ++i)
      if (std::sqrt(arun_sqdist(m, p1 + 3 * i, p2 + 3 * i)) < thr) out[c++] = (uint32_t)i;
    return c;
  }
};

struct NisterProblem {  // CentralRelativePoseSacProblem: algorithm 0 = NISTER, 1 = STEWENIUS (row f4)
  const double* f1; const double* f2; int N; int algorithm = 0;
  bool compute(const uint16_t* s, double* model) const { return mono_model(f1, f2, s, model, algorithm); }
  int count(const double* m, double thr) const {
    double tinv[3];
    mono_tinv(m, tinv);
    int c = 0;
    for (int i = 0; i < N; ++i) c += (mono_residual(m, tinv, f1 + 3 * i, f2 + 3 * i) < thr);
    return c;
  }
  int select(const double* m, double thr, uint32_t* out) const {
    double tinv[3];
    mono_tinv(m, tinv);
    int c = 0;
    for (int i = 0; i < N; ++i)
      if (mono_residual(m, tinv, f1 + 3 * i, f2 + 3 * i) < thr) out[c++] = (uint32_t)i;
    return c;
  }
};

// =========================================================================
// A.3-A.8  LoopClosureDetector (Kimera-Multi-LCD src/loop_closure_detector.cpp)
// =========================================================================
struct Frame {
  int F = 0;
  std::vector<uint8_t> desc;
  std::vector<double> bearings, points;
};
struct Bow {
  std::vector<uint32_t> ids;
  std::vector<float> vals;
};
typedef std::pair<uint64_t, uint64_t> RobotPoseId;

struct kmo_lcd {
  kmo_params prm;
  std::map<uint64_t, std::unique_ptr<kmo_db>> db_BoW_;  // ordered: detectLoop visits robots ascending
  std::map<uint64_t, std::vector<uint64_t>> db_EntryId_to_PoseId_;
  std::map<uint64_t, std::map<uint64_t, Bow>> bow_vectors_;
  std::map<RobotPoseId, Frame> vlc_frames_;
};

static bool find_prev_bow(const kmo_lcd* L, uint64_t robot, uint64_t pose, const Bow** prev) {
  auto rit = L->bow_vectors_.find(robot);
  if (rit == L->bow_vectors_.end()) return false;
  for (int i = 1; i <= L->prm.max_nrFrames_between_queries; ++i) {
    if (pose < (uint64_t)i) break;
    auto it = rit->second.find(pose - i);
    if (it != rit->second.end()) { *prev = &it->second; return true; }
  }
  return false;
}

// steps 1,3-7 of A.3 detectLoopWithRobot given the nss factor
static int detect_with_robot_nss(const kmo_lcd* L, uint64_t robot, uint64_t q_robot,
                                 uint64_t q_pose, const uint32_t* ids, const float* vals, int n,
                                 double nss, uint64_t* out_robot, uint64_t* out_pose,
                                 double* out_score, int cap) {
  auto dit = L->db_BoW_.find(robot);
  if (dit == L->db_BoW_.end()) return 0;
  if (L->prm.inter_robot_only && q_robot == robot) return 0;
  if (nss < L->prm.min_nss_factor) return 0;
  const int K = L->prm.max_db_results;
  std::vector<uint32_t> e(K > 0 ? K : 1);
  std::vector<double> s(K > 0 ? K : 1);
  int m = db_query(dit->second.get(), ids, vals, n, K, -1, e.data(), s.data(), K);
  const double cut = L->prm.alpha * nss;
  int keep = 0;
  while (keep < m && !(s[keep] < cut)) ++keep;  // lower_bound with Result::geq
  const auto& e2p = L->db_EntryId_to_PoseId_.at(robot);
  int cnt = 0;
  for (int i = 0; i < keep && cnt < cap; ++i) {
    uint64_t pose = e2p[e[i]];
    if (q_robot == robot) {
      uint64_t d = q_pose > pose ? q_pose - pose : pose - q_pose;
      if (d < (uint64_t)L->prm.dist_local) continue;
    }
    out_robot[cnt] = robot;
    out_pose[cnt] = pose;
    out_score[cnt] = s[i] / nss;
    ++cnt;
  }
  return cnt;
}

struct Cand { double score; uint64_t robot, pose; };

static int verify_pair(const kmo_lcd* L, const Frame& fq, const Frame& fm, kmo_result* r,
                       std::vector<uint32_t>& iq, std::vector<uint32_t>& im) {
  const kmo_params& P = L->prm;
  // computeMatchedIndices (A.4)
  iq.resize(std::max(fq.F, 1)); im.resize(std::max(fq.F, 1));
  int M = match_lowe(fq.desc.data(), fq.F, fm.desc.data(), fm.F, P.lowe_ratio, iq.data(), im.data(), P.matcher_norm);
  r->n_matches = M;
  r->mono_inliers = 0; r->stereo_inliers = 0; r->status = 1;
  // geometricVerificationNister (A.6)
  std::vector<double> b1(3 * (size_t)std::max(M, 1)), b2(3 * (size_t)std::max(M, 1));
  for (int i = 0; i < M; ++i)
    for (int c = 0; c < 3; ++c) {
      b1[3 * i + c] = fq.bearings[3 * iq[i] + c];
      b2[3 * i + c] = fm.bearings[3 * im[i] + c];
    }
  kmo_ransac_result rr;
  std::vector<uint32_t> inl(std::max(M, 1));
  NisterProblem np{b1.data(), b2.data(), M, P.mono_algorithm};
  ransac(np, M, 8, P.ransac_threshold_mono, P.ransac_probability_mono,
         P.max_ransac_iterations_mono, P.ransac_seed, &rr, inl.data());
  if (!rr.success) return 1;
  if (rr.n_inliers < P.geometric_verification_min_inlier_count) return 1;
  if ((double)rr.n_inliers / (double)M < P.ransac_inlier_percentage_mono) return 1;
  r->mono_inliers = rr.n_inliers;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r->R_mono[3 * i + j] = rr.model[4 * i + j];
  std::vector<uint32_t> jq(rr.n_inliers), jm(rr.n_inliers);
  for (int i = 0; i < rr.n_inliers; ++i) { jq[i] = iq[inl[i]]; jm[i] = im[inl[i]]; }
  // recoverPose (A.8)
  r->status = 2;
  std::vector<double> p1, p2;
  std::vector<uint32_t> kq, km;
  for (int i = 0; i < rr.n_inliers; ++i) {
    const double* a = &fq.points[3 * jq[i]];
    const double* b = &fm.points[3 * jm[i]];
    double na = std::sqrt(dot3(a, a)), nb = std::sqrt(dot3(b, b));
    if (na > 1e-3 && nb > 1e-3) {
      for (int c = 0; c < 3; ++c) { p1.push_back(a[c]); p2.push_back(b[c]); }
      kq.push_back(jq[i]); km.push_back(jm[i]);
    }
  }
  int N3 = (int)kq.size();
  if (N3 < 3) return 2;
  kmo_ransac_result r3;
  std::vector<uint32_t> inl3(N3);
  if (P.ransac_use_1point_3d3d) {
    OnePointProblem op{p1.data(), p2.data(), N3, {0}};
    memcpy(op.R, r->R_mono, sizeof(op.R));
    ransac(op, N3, 1, P.ransac_threshold, P.ransac_probability, P.max_ransac_iterations, P.ransac_seed, &r3, inl3.data());
  } else {
    ArunProblem ap{p1.data(), p2.data(), N3};
    ransac(ap, N3, 3, P.ransac_threshold, P.ransac_probability, P.max_ransac_iterations,
           P.ransac_seed, &r3, inl3.data());
  }
  if (!r3.success) return 2;
  if (r3.n_inliers < P.geometric_verification_min_inlier_count) return 2;
  if ((double)r3.n_inliers / (double)N3 < P.geometric_verification_min_inlier_percentage) return 2;
  r->stereo_inliers = r3.n_inliers;
  memcpy(r->T, r3.model, sizeof(r3.model));
  r->status = 0;
  iq.resize(r3.n_inliers); im.resize(r3.n_inliers);
  for (int i = 0; i < r3.n_inliers; ++i) { iq[i] = kq[inl3[i]]; im[i] = km[inl3[i]]; }
  return 0;
}

static int lcd_query(const kmo_lcd* L, uint64_t q_robot, uint64_t q_pose, const uint32_t* ids,
                     const float* vals, int n, const uint32_t* pids, const float* pvals, int pn,
                     const Frame& fq, kmo_result* out, int cap) {
  double nss = l1_score(ids, vals, n, pids, pvals, pn);
  std::vector<Cand> cands;
  const int K = std::max(L->prm.max_db_results, 1);
  std::vector<uint64_t> rr(K), pp(K);
  std::vector<double> ss(K);
  for (auto& kv : L->db_BoW_) {
    int m = detect_with_robot_nss(L, kv.first, q_robot, q_pose, ids, vals, n, nss, rr.data(),
                                  pp.data(), ss.data(), K);
    for (int i = 0; i < m; ++i) cands.push_back({ss[i], rr[i], pp[i]});
  }
  std::stable_sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) {
    if (a.score != b.score) return a.score > b.score;
    if (a.robot != b.robot) return a.robot < b.robot;
    return a.pose < b.pose;
  });
  int nv = std::min((int)cands.size(), std::min(L->prm.top_k_verify, cap));
  std::vector<uint32_t> iq, im;
  for (int i = 0; i < nv; ++i) {
    kmo_result* r = &out[i];
    memset(r, 0, sizeof(*r));
    r->q_robot = q_robot; r->q_pose = q_pose;
    r->m_robot = cands[i].robot; r->m_pose = cands[i].pose;
    r->norm_bow_score = cands[i].score;
    auto fit = L->vlc_frames_.find(RobotPoseId(cands[i].robot, cands[i].pose));
    if (fit == L->vlc_frames_.end()) { r->status = 3; continue; }
    verify_pair(L, fq, fit->second, r, iq, im);
  }
  return nv;
}

// =========================================================================
// C interface
// =========================================================================
extern "C" {

void kmo_default_params(kmo_params* p) {
  p->inter_robot_only = 0;
  p->alpha = 0.5;
  p->dist_local = 90;
  p->max_db_results = 50;
  p->min_nss_factor = 0.05;
  p->max_nrFrames_between_queries = 2;
  p->lowe_ratio = 0.9;
  p->ransac_threshold_mono = 1e-6;
  p->ransac_inlier_percentage_mono = 0.01;
  p->max_ransac_iterations_mono = 1000;
  p->ransac_probability_mono = 0.995;
  p->ransac_threshold = 0.5;
  p->max_ransac_iterations = 1000;
  p->ransac_probability = 0.995;
  p->geometric_verification_min_inlier_count = 5;
  p->geometric_verification_min_inlier_percentage = 0.0;
  p->ransac_seed = 12345u;
  p->top_k_verify = 16;
  p->matcher_norm = 0;
  p->mono_algorithm = 0;
  p->ransac_use_1point_3d3d = 0;
}

double kmo_bow_score(const uint32_t* ids1, const float* vals1, int n1, const uint32_t* ids2,
                     const float* vals2, int n2) {
  return l1_score(ids1, vals1, n1, ids2, vals2, n2);
}

kmo_db* kmo_db_create(void) { return new kmo_db(); }
void kmo_db_destroy(kmo_db* db) { delete db; }
uint32_t kmo_db_size(const kmo_db* db) { return db->nentries; }
uint32_t kmo_db_add(kmo_db* db, const uint32_t* ids, const float* vals, int n) {
  uint32_t id = db->nentries++;
  for (int i = 0; i < n; ++i) db->ifile[ids[i]].push_back({id, (double)vals[i]});
  return id;
}
int kmo_db_query(const kmo_db* db, const uint32_t* ids, const float* vals, int n,
                 int max_results, int max_id, uint32_t* out_entry, double* out_score, int cap) {
  return db_query(db, ids, vals, n, max_results, max_id, out_entry, out_score, cap);
}

void kmo_hamming_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, uint32_t* idx,
                      uint16_t* dist) {
  knn2(q, nq, t, nt, idx, dist);
}
int kmo_match_lowe(const uint8_t* q, int nq, const uint8_t* t, int nt, double lowe_ratio,
                   uint32_t* i_query, uint32_t* i_match) {
  return match_lowe(q, nq, t, nt, lowe_ratio, i_query, i_match);
}

void kmo_sample_stream(int N, int s, uint32_t seed, int n_draws, uint16_t* out) {
  if (N < s) return;
  Sampler sm(N, s, seed);
  for (int d = 0; d < n_draws; ++d) sm.draw(out + (size_t)d * s);
}

void kmo_arun3(const double* p1, const double* p2, double* model) {
  arun3(p1, p1 + 3, p1 + 6, p2, p2 + 3, p2 + 6, model);
}
int kmo_fivept_nister(const double* f1, const double* f2, double* E) {
  double a[5][3], b[5][3], Es[10][9];
  memcpy(a, f1, sizeof(a)); memcpy(b, f2, sizeof(b));
  int n = fivept_nister(a, b, Es);
  memcpy(E, Es, sizeof(double) * 9 * n);
  return n;
}
void kmo_svd3(const double* A, double* U, double* S, double* V) { svd3(A, U, S, V); }
void kmo_krsqrt(const double* x, int n, double* out) { for (int i = 0; i < n; ++i) out[i] = krsqrt(x[i]); }
int kmo_mono_model(const double* f1, const double* f2, const uint16_t* sample8, double* model) {
  return mono_model(f1, f2, sample8, model) ? 1 : 0;
}
double kmo_mono_residual(const double* model, const double* f1, const double* f2) {
  double tinv[3];
  mono_tinv(model, tinv);
  return mono_residual(model, tinv, f1, f2);
}
double kmo_arun_residual(const double* model, const double* p1, const double* p2) {
  return std::sqrt(arun_sqdist(model, p1, p2));
}

void kmo_ransac_arun(const double* p1, const double* p2, int N, double thr, double prob,
                     int max_iter, uint32_t seed, kmo_ransac_result* res, uint32_t* inliers) {
  ArunProblem ap{p1, p2, N};
  ransac(ap, N, 3, thr, prob, max_iter, seed, res, inliers);
}
void kmo_ransac_nister(const double* f1, const double* f2, int N, double thr, double prob,
                       int max_iter, uint32_t seed, kmo_ransac_result* res, uint32_t* inliers) {
  NisterProblem np{f1, f2, N, 0};
  ransac(np, N, 8, thr, prob, max_iter, seed, res, inliers);
}
void kmo_debug_root_grid2(int on) { kmo::g_root_grid2 = on; }
void kmo_ransac_stewenius(const double* f1, const double* f2, int N, double thr, double prob,
                          int max_iter, uint32_t seed, kmo_ransac_result* res, uint32_t* inliers) {
  NisterProblem np{f1, f2, N, 1};
  ransac(np, N, 8, thr, prob, max_iter, seed, res, inliers);
}
int kmo_fivept_stewenius(const double* f1, const double* f2, double* E) {
  double a[5][3], b[5][3], Es[10][9];
  memcpy(a, f1, sizeof(a)); memcpy(b, f2, sizeof(b));
  int n = fivept_stewenius(a, b, Es);
  memcpy(E, Es, sizeof(double) * 9 * n);
  return n;
}
int kmo_mono_model_alg(const double* f1, const double* f2, const uint16_t* sample8, double* model, int algorithm) {
  return mono_model(f1, f2, sample8, model, algorithm) ? 1 : 0;
}

kmo_lcd* kmo_lcd_create(const kmo_params* p) {
  kmo_lcd* L = new kmo_lcd();
  if (p) L->prm = *p; else kmo_default_params(&L->prm);
  return L;
}
void kmo_lcd_destroy(kmo_lcd* L) { delete L; }

void kmo_lcd_add_bow(kmo_lcd* L, uint64_t robot, uint64_t pose, const uint32_t* ids,
                     const float* vals, int n) {
  auto& bv = L->bow_vectors_[robot];
  if (bv.count(pose)) return;
  auto& db = L->db_BoW_[robot];
  if (!db) db.reset(new kmo_db());
  uint32_t entry = kmo_db_add(db.get(), ids, vals, n);
  auto& e2p = L->db_EntryId_to_PoseId_[robot];
  if (e2p.size() <= entry) e2p.resize(entry + 1);
  e2p[entry] = pose;
  Bow b;
  b.ids.assign(ids, ids + n);
  b.vals.assign(vals, vals + n);
  bv[pose] = std::move(b);
}

void kmo_lcd_add_frame(kmo_lcd* L, uint64_t robot, uint64_t pose, const uint8_t* desc,
                       const double* bearings, const double* points, int F) {
  Frame f;
  f.F = F;
  f.desc.assign(desc, desc + 32 * (size_t)F);
  f.bearings.assign(bearings, bearings + 3 * (size_t)F);
  f.points.assign(points, points + 3 * (size_t)F);
  L->vlc_frames_[RobotPoseId(robot, pose)] = std::move(f);
}

int kmo_lcd_detect_loop_with_robot(kmo_lcd* L, uint64_t robot, uint64_t q_robot, uint64_t q_pose,
                                   const uint32_t* ids, const float* vals, int n,
                                   uint64_t* out_robot, uint64_t* out_pose, double* out_score,
                                   int cap) {
  if (!L->db_BoW_.count(robot)) return 0;
  const Bow* prev = nullptr;
  if (!find_prev_bow(L, q_robot, q_pose, &prev)) return 0;
  double nss = l1_score(ids, vals, n, prev->ids.data(), prev->vals.data(), (int)prev->ids.size());
  return detect_with_robot_nss(L, robot, q_robot, q_pose, ids, vals, n, nss, out_robot,
                               out_pose, out_score, cap);
}

int kmo_lcd_detect_loop(kmo_lcd* L, uint64_t q_robot, uint64_t q_pose, const uint32_t* ids,
                        const float* vals, int n, uint64_t* out_robot, uint64_t* out_pose,
                        double* out_score, int cap) {
  int cnt = 0;
  for (auto& kv : L->db_BoW_)
    cnt += kmo_lcd_detect_loop_with_robot(L, kv.first, q_robot, q_pose, ids, vals, n,
                                          out_robot + cnt, out_pose + cnt, out_score + cnt,
                                          cap - cnt);
  return cnt;
}

int kmo_lcd_compute_matched_indices(kmo_lcd* L, uint64_t qr, uint64_t qp, uint64_t mr,
                                    uint64_t mp, uint32_t* i_query, uint32_t* i_match, int cap) {
  auto q = L->vlc_frames_.find(RobotPoseId(qr, qp));
  auto m = L->vlc_frames_.find(RobotPoseId(mr, mp));
  if (q == L->vlc_frames_.end() || m == L->vlc_frames_.end()) return 0;
  std::vector<uint32_t> iq(std::max(q->second.F, 1)), im(std::max(q->second.F, 1));
  int M = match_lowe(q->second.desc.data(), q->second.F, m->second.desc.data(), m->second.F,
                     L->prm.lowe_ratio, iq.data(), im.data(), L->prm.matcher_norm);
  M = std::min(M, cap);
  memcpy(i_query, iq.data(), sizeof(uint32_t) * M);
  memcpy(i_match, im.data(), sizeof(uint32_t) * M);
  return M;
}

int kmo_lcd_geometric_verification_nister(kmo_lcd* L, uint64_t qr, uint64_t qp, uint64_t mr,
                                          uint64_t mp, uint32_t* inl_q, uint32_t* inl_m,
                                          int* count, double* R) {
  auto q = L->vlc_frames_.find(RobotPoseId(qr, qp));
  auto m = L->vlc_frames_.find(RobotPoseId(mr, mp));
  if (q == L->vlc_frames_.end() || m == L->vlc_frames_.end()) return 0;
  const int M = *count;
  std::vector<double> b1(3 * (size_t)std::max(M, 1)), b2(3 * (size_t)std::max(M, 1));
  for (int i = 0; i < M; ++i)
    for (int c = 0; c < 3; ++c) {
      b1[3 * i + c] = q->second.bearings[3 * inl_q[i] + c];
      b2[3 * i + c] = m->second.bearings[3 * inl_m[i] + c];
    }
  kmo_ransac_result rr;
  std::vector<uint32_t> inl(std::max(M, 1));
  const kmo_params& P = L->prm;
  NisterProblem np{b1.data(), b2.data(), M, P.mono_algorithm};
  ransac(np, M, 8, P.ransac_threshold_mono, P.ransac_probability_mono,
         P.max_ransac_iterations_mono, P.ransac_seed, &rr, inl.data());
  if (!rr.success) return 0;
  if (rr.n_inliers < P.geometric_verification_min_inlier_count) return 0;
  if ((double)rr.n_inliers / (double)M < P.ransac_inlier_percentage_mono) return 0;
  std::vector<uint32_t> jq(rr.n_inliers), jm(rr.n_inliers);
  for (int i = 0; i < rr.n_inliers; ++i) { jq[i] = inl_q[inl[i]]; jm[i] = inl_m[inl[i]]; }
  memcpy(inl_q, jq.data(), sizeof(uint32_t) * rr.n_inliers);
  memcpy(inl_m, jm.data(), sizeof(uint32_t) * rr.n_inliers);
  *count = rr.n_inliers;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[3 * i + j] = rr.model[4 * i + j];
  return 1;
}

int kmo_lcd_recover_pose(kmo_lcd* L, uint64_t qr, uint64_t qp, uint64_t mr, uint64_t mp,
                         uint32_t* inl_q, uint32_t* inl_m, int* count, double* T) {
  return kmo_lcd_recover_pose_prior(L, qr, qp, mr, mp, inl_q, inl_m, count, nullptr, T);
}
void kmo_ransac_onepoint(const double* p1, const double* p2, int N, const double* R, double thr, double prob,
                         int max_iter, uint32_t seed, kmo_ransac_result* res, uint32_t* inliers) {
  OnePointProblem op{p1, p2, N, {0}};
  memcpy(op.R, R, sizeof(op.R));
  ransac(op, N, 1, thr, prob, max_iter, seed, res, inliers);
}
int kmo_lcd_recover_pose_prior(kmo_lcd* L, uint64_t qr, uint64_t qp, uint64_t mr, uint64_t mp,
                               uint32_t* inl_q, uint32_t* inl_m, int* count, const double* R_prior, double* T) {
  auto q = L->vlc_frames_.find(RobotPoseId(qr, qp));
  auto m = L->vlc_frames_.find(RobotPoseId(mr, mp));
  if (q == L->vlc_frames_.end() || m == L->vlc_frames_.end()) return 0;
  const int M = *count;
  std::vector<double> p1, p2;
  std::vector<uint32_t> kq, km;
  for (int i = 0; i < M; ++i) {
    const double* a = &q->second.points[3 * inl_q[i]];
    const double* b = &m->second.points[3 * inl_m[i]];
    double na = std::sqrt(dot3(a, a)), nb = std::sqrt(dot3(b, b));
    if (na > 1e-3 && nb > 1e-3) {
      for (int c = 0; c < 3; ++c) { p1.push_back(a[c]); p2.push_back(b[c]); }
      kq.push_back(inl_q[i]); km.push_back(inl_m[i]);
    }
  }
  int N3 = (int)kq.size();
  if (N3 < 3) return 0;
  kmo_ransac_result r3;
  std::vector<uint32_t> inl3(N3);
  const kmo_params& P = L->prm;
  if (P.ransac_use_1point_3d3d && R_prior) {
    OnePointProblem op{p1.data(), p2.data(), N3, {0}};
    memcpy(op.R, R_prior, sizeof(op.R));
    ransac(op, N3, 1, P.ransac_threshold, P.ransac_probability, P.max_ransac_iterations, P.ransac_seed, &r3, inl3.data());
  } else {
    ArunProblem ap{p1.data(), p2.data(), N3};
    ransac(ap, N3, 3, P.ransac_threshold, P.ransac_probability, P.max_ransac_iterations,
           P.ransac_seed, &r3, inl3.data());
  }
  if (!r3.success) return 0;
  if (r3.n_inliers < P.geometric_verification_min_inlier_count) return 0;
  if ((double)r3.n_inliers / (double)N3 < P.geometric_verification_min_inlier_percentage) return 0;
  for (int i = 0; i < r3.n_inliers; ++i) { inl_q[i] = kq[inl3[i]]; inl_m[i] = km[inl3[i]]; }
  *count = r3.n_inliers;
  memcpy(T, r3.model, sizeof(double) * 12);
  return 1;
}

int kmo_lcd_query(kmo_lcd* L, uint64_t q_robot, uint64_t q_pose, const uint32_t* ids,
                  const float* vals, int n, const uint32_t* prev_ids, const float* prev_vals,
                  int prev_n, const uint8_t* desc, const double* bearings, const double* points,
                  int F, kmo_result* out, int cap) {
  Frame fq;
  fq.F = F;
  fq.desc.assign(desc, desc + 32 * (size_t)F);
  fq.bearings.assign(bearings, bearings + 3 * (size_t)F);
  fq.points.assign(points, points + 3 * (size_t)F);
  return lcd_query(L, q_robot, q_pose, ids, vals, n, prev_ids, prev_vals, prev_n, fq, out, cap);
}

int kmo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int kmo_lcd_query_batch(kmo_lcd* L, int B, const uint64_t* q_robot, const uint64_t* q_pose,
                        const int64_t* bow_off, const uint32_t* ids, const float* vals,
                        const int64_t* prev_off, const uint32_t* prev_ids,
                        const float* prev_vals, const uint8_t* desc, const double* bearings,
                        const double* points, int F, kmo_result* out, int cap_per_query,
                        int32_t* counts, int threads) {
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
  (void)threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (int b = 0; b < B; ++b) {
    counts[b] = kmo_lcd_query(
        L, q_robot[b], q_pose[b], ids + bow_off[b], vals + bow_off[b],
        (int)(bow_off[b + 1] - bow_off[b]), prev_ids + prev_off[b], prev_vals + prev_off[b],
        (int)(prev_off[b + 1] - prev_off[b]), desc + (size_t)b * F * 32,
        bearings + (size_t)b * F * 3, points + (size_t)b * F * 3, F,
        out + (size_t)b * cap_per_query, cap_per_query);
  }
  int total = 0;
  for (int b = 0; b < B; ++b) total += counts[b];
  return total;
}

}  // extern "C"

// exposed for tests: real roots of a polynomial (ascending coefficients) in (-1,1]
extern "C" int kmo_roots_unit(const double* p, int n, double* roots) { return roots_unit(p, n, roots); }

// =========================================================================
// A.9  TemplatedVocabulary::transform (DBoW2/include/DBoW2/TemplatedVocabulary.h,
//      BowVector::addWeight / normalize in DBoW2/src/BowVector.cpp) — row f1
// =========================================================================
struct kmo_vocab {
  int k, L;
  std::vector<uint8_t> nodes;    // breadth-first, level 1 first
  std::vector<double> weights;   // per leaf
};
extern "C" kmo_vocab* kmo_vocab_create(int k, int L, const uint8_t* node_desc, const double* w) {
  kmo_vocab* v = new kmo_vocab();
  v->k = k; v->L = L;
  size_t nodes = 0, lvl = 1;
  for (int l = 0; l < L; ++l) { lvl *= (size_t)k; nodes += lvl; }
  v->nodes.assign(node_desc, node_desc + nodes * 32);
  v->weights.assign(w, w + lvl);
  return v;
}
extern "C" void kmo_vocab_destroy(kmo_vocab* v) { delete v; }
extern "C" int kmo_vocab_transform(const kmo_vocab* voc, const uint8_t* desc, int F, uint32_t* ids, double* vals) {
  std::map<uint32_t, double> v;
  for (int f = 0; f < F; ++f) {
    size_t idx = 0, level_off = 0, level_n = 1;
    for (int l = 0; l < voc->L; ++l) {
      level_n *= (size_t)voc->k;
      const uint8_t* children = voc->nodes.data() + 32 * (level_off + idx * (size_t)voc->k);
      int best = hamming256(desc + 32 * f, children), bc = 0;
      for (int c = 1; c < voc->k; ++c) {
        const int d = hamming256(desc + 32 * f, children + 32 * c);
        if (d < best) { best = d; bc = c; }
      }
      idx = idx * (size_t)voc->k + (size_t)bc;
      level_off += level_n;
    }
    const double w = voc->weights[idx];
    if (w > 0) {  // BowVector::addWeight
      auto it = v.lower_bound((uint32_t)idx);
      if (it != v.end() && !(v.key_comp()((uint32_t)idx, it->first))) it->second += w;
      else v.insert(it, std::make_pair((uint32_t)idx, w));
    }
  }
  double norm = 0.0;  // BowVector::normalize(L1)
  for (auto& kv : v) norm += std::fabs(kv.second);
  int n = 0;
  for (auto& kv : v) {
    ids[n] = kv.first;
    vals[n] = norm > 0.0 ? kv.second / norm : kv.second;
    ++n;
  }
  return n;
}

extern "C" void kmo_l1_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, uint32_t* idx, uint16_t* dist) {
  knn2(q, nq, t, nt, idx, dist, 1);
}

// ===================================================== row f3 post filters
// LcdThirdPartyWrapper::computeIslands / checkTemporalConstraint as Kimera-VIO runs them after the
// alpha*nss cut.  Walks the results by ascending id exactly like the
// upstream loop: first/last entry of the running island, best score so far, island closed when
// the next id is max_intraisland_gap or more away.
extern "C" int kmo_compute_islands(const uint64_t* ids, const double* scores, int n, int max_gap,
                                   int min_len, kmo_island* out, int cap) {
  struct Res { uint64_t id; double score; };
  std::vector<Res> q(n);
  for (int i = 0; i < n; ++i) q[i] = Res{ids[i], scores[i]};
  std::vector<kmo_island> islands;
  if (q.size() == 1) {
    islands.push_back(kmo_island{q[0].id, q[0].id, q[0].id, q[0].score, q[0].score});
  } else if (!q.empty()) {
    std::sort(q.begin(), q.end(), [](const Res& a, const Res& b) { return a.id < b.id; });
    long long first_entry = (long long)q[0].id, last_entry = (long long)q[0].id;
    size_t i_first = 0, i_last = 0;
    double best_score = q[0].score;
    uint64_t best_entry = q[0].id;
    auto island_score = [&](size_t a, size_t b) {
      double sum = 0.0;
      for (size_t i = a; i <= b; ++i) sum += q[i].score;
      return sum;
    };
    for (size_t idx = 1; idx < q.size(); ++idx) {
      if ((long long)q[idx].id - last_entry < (long long)max_gap) {
        last_entry = (long long)q[idx].id;
        i_last = idx;
        if (q[idx].score > best_score) { best_score = q[idx].score; best_entry = q[idx].id; }
      } else {
        if (last_entry - first_entry + 1 >= (long long)min_len)
          islands.push_back(kmo_island{(uint64_t)first_entry, (uint64_t)last_entry, best_entry,
                                       island_score(i_first, i_last), best_score});
        first_entry = last_entry = (long long)q[idx].id;
        i_first = i_last = idx;
        best_score = q[idx].score;
        best_entry = q[idx].id;
      }
    }
    if (last_entry - first_entry + 1 >= (long long)min_len)
      islands.push_back(kmo_island{(uint64_t)first_entry, (uint64_t)last_entry, best_entry,
                                   island_score(i_first, i_last), best_score});
  }
  const int m = std::min<int>((int)islands.size(), cap);
  for (int i = 0; i < m; ++i) out[i] = islands[i];
  return (int)islands.size();
}

extern "C" int kmo_check_temporal_constraint(kmo_temporal_state* st, uint64_t id, const kmo_island* island,
                                             int max_between_queries, int max_between_islands,
                                             int min_temporal_matches) {
  if (st->temporal_entries == 0 || (long long)id - (long long)st->latest_query_id > (long long)max_between_queries) {
    st->temporal_entries = 1;
  } else {
    const long long a1 = (long long)st->latest_island.start_id, a2 = (long long)st->latest_island.end_id;
    const long long b1 = (long long)island->start_id, b2 = (long long)island->end_id;
    const bool overlap = (b1 <= a1 && a1 <= b2) || (a1 <= b1 && b1 <= a2);
    bool gap_is_small = false;
    if (!overlap) {
      const long long d = (a1 > b2) ? a1 - b2 : b1 - a2;
      gap_is_small = d <= (long long)max_between_islands;
    }
    if (overlap || gap_is_small) st->temporal_entries++; else st->temporal_entries = 1;
  }
  st->latest_island = *island;
  st->latest_query_id = id;
  return st->temporal_entries > min_temporal_matches;
}
