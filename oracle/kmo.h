/*
 * kmo.h — C interface of the CPU ORACLE (libkml_oracle.so).
 *
 * TEST INFRASTRUCTURE ONLY.  This library is a scalar CPU restatement of the
 * loop-closure hot path of Kimera-Multi-LCD (DBoW2 L1 database query, OpenCV
 * BFMatcher Hamming kNN + Lowe ratio, OpenGV RANSAC with the 5-point Nister
 * and 3-point Arun minimal solvers).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The
 * product library (libkml.so) never links, loads or calls anything in here.
 *
 * PARITY STATUS: "parity unpinned" against the real DBoW2 / OpenCV-C++ /
 * OpenGV sources: those live in un-vendored repositories named by
 * /root/reference/kimera_multi.repos:14-17,54-57,102-109 (branch names only,
 * no commit pins) and are absent from /root/reference.  The normative spec
 * this file follows is SURVEY.md Appendix A (A.1-A.8).  What pins it instead:
 *   - Hamming kNN  == cv2.BFMatcher(NORM_HAMMING).knnMatch   (bit exact)
 *   - RNG stream   == SURVEY.md Appendix B.2 known answers
 *   - 5-pt solver  ⊇ cv2.findEssentialMat solution set, invariants
 *   - Arun         == numpy SVD Kabsch (1e-12)
 *   - BoW L1       == dense numpy 1 - 0.5*|v-w|_1 (1e-12)
 * (see tests/test_oracle.py; the same external implementations wrote the
 * committed fixture tests/golden/golden_v1.npz checked by tests/test_golden.py).
 *
 * All fp64 arithmetic is compiled with -ffp-contract=off so that the operation
 * order written in the sources is the operation order executed; the fused
 * multiply-adds of the contract are explicit kfma() calls (src/geom.hpp).
 */
#ifndef KMO_H_
#define KMO_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- parameters (mirror of kml_params in include/kml.h; kept separate on
 * purpose so the oracle does not include product headers) ---------------- */
typedef struct kmo_params {
  int32_t inter_robot_only;
  double alpha;
  int32_t dist_local;
  int32_t max_db_results;
  double min_nss_factor;
  int32_t max_nrFrames_between_queries;
  double lowe_ratio;
  double ransac_threshold_mono;
  double ransac_inlier_percentage_mono;
  int32_t max_ransac_iterations_mono;
  double ransac_probability_mono;
  double ransac_threshold;
  int32_t max_ransac_iterations;
  double ransac_probability;
  int32_t geometric_verification_min_inlier_count;
  double geometric_verification_min_inlier_percentage;
  uint32_t ransac_seed;
  int32_t top_k_verify;
  int32_t matcher_norm; /* 0 NORM_HAMMING, 1 NORM_L1 (upstream's create(3)) */
  int32_t mono_algorithm;          /* 0 NISTER, 1 STEWENIUS (/root/reference/params/D455/LcdParams.yaml:68-73) */
  int32_t ransac_use_1point_3d3d;  /* LcdParams.yaml:58: recoverPose with the mono rotation given, 1-point samples */
} kmo_params;

void kmo_default_params(kmo_params* p);

/* ---- A.2  L1Scoring::score ------------------------------------------- */
double kmo_bow_score(const uint32_t* ids1, const float* vals1, int n1,
                     const uint32_t* ids2, const float* vals2, int n2);

/* ---- A.1  TemplatedDatabase add / queryL1 ----------------------------- */
typedef struct kmo_db kmo_db;
kmo_db* kmo_db_create(void);
void kmo_db_destroy(kmo_db*);
uint32_t kmo_db_add(kmo_db*, const uint32_t* ids, const float* vals, int n);
uint32_t kmo_db_size(const kmo_db*);
/* results: best first (score descending, ties entry ascending) */
int kmo_db_query(const kmo_db*, const uint32_t* ids, const float* vals, int n,
                 int max_results, int max_id, uint32_t* out_entry,
                 double* out_score, int cap);

/* ---- A.4 / B.1  BFMatcher(NORM_HAMMING).knnMatch(k=2) + Lowe ---------- */
/* idx/dist are [nq][2]; missing neighbours are idx=0xFFFFFFFF dist=0xFFFF */
void kmo_hamming_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt,
                      uint32_t* idx, uint16_t* dist);
void kmo_l1_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, uint32_t* idx, uint16_t* dist);
int kmo_match_lowe(const uint8_t* q, int nq, const uint8_t* t, int nt,
                   double lowe_ratio, uint32_t* i_query, uint32_t* i_match);

/* ---- A.7  sample stream ------------------------------------------------ */
/* out[n_draws][s]: persistent partial Fisher-Yates over [0,N) */
void kmo_sample_stream(int N, int s, uint32_t seed, int n_draws, uint16_t* out);

/* ---- minimal solvers (exposed for cross-checks) ------------------------ */
/* p1,p2: 3 points each, row-major [3][3]; model: [R|t] row-major 3x4,
 * p1 = R p2 + t */
void kmo_arun3(const double* p1, const double* p2, double* model);
/* f1,f2: 5 unit bearings each [5][3]; returns number of essential matrices
 * written to E[10][9] (row-major, f1^T E f2 = 0) */
int kmo_fivept_nister(const double* f1, const double* f2, double* E);
/* row f4: the Stewenius variant (action matrix; real solutions), same conventions */
int kmo_fivept_stewenius(const double* f1, const double* f2, double* E);
int kmo_mono_model_alg(const double* f1, const double* f2, const uint16_t* sample8, double* model, int algorithm);

/* proper SVD used by both solvers: A = U diag(S) V^T, det U = det V = +1 */
void kmo_svd3(const double* A, double* U, double* S, double* V);
/* the contract's reciprocal square root (geom.hpp krsqrt) on n values */
void kmo_krsqrt(const double* x, int n, double* out);
/* mono model from an 8-point sample (5 solve + 3 disambiguate); returns 0
 * if no valid model */
int kmo_mono_model(const double* f1, const double* f2, const uint16_t* sample8,
                   double* model);
double kmo_mono_residual(const double* model, const double* f1, const double* f2);
double kmo_arun_residual(const double* model, const double* p1, const double* p2);

/* ---- A.5  sac::Ransac::computeModel ------------------------------------ */
typedef struct kmo_ransac_result {
  int32_t success;        /* model_ non-empty */
  int32_t iterations;     /* counted trials */
  int32_t skipped;
  int32_t draws;          /* loop passes = sample-stream rows consumed */
  int32_t best_draw;      /* stream row of the winning sample */
  int32_t n_inliers;
  double model[12];
} kmo_ransac_result;

/* stereo: p1 = query points, p2 = match points, [N][3] */
void kmo_ransac_arun(const double* p1, const double* p2, int N, double thr,
                     double prob, int max_iter, uint32_t seed,
                     kmo_ransac_result* res, uint32_t* inliers);
/* mono: f1 = query bearings, f2 = match bearings, [N][3] */
void kmo_ransac_nister(const double* f1, const double* f2, int N, double thr,
                       double prob, int max_iter, uint32_t seed,
                       kmo_ransac_result* res, uint32_t* inliers);
void kmo_ransac_stewenius(const double* f1, const double* f2, int N, double thr, double prob, int max_iter,
                          uint32_t seed, kmo_ransac_result* res, uint32_t* inliers);
/* test knob: 0 = root isolation skips the 256-cell grid (the library's KML_NO_ROOT_GRID2), so that every
 * chain the 32-cell grid does not separate goes through the Sturm bisection on both sides */
void kmo_debug_root_grid2(int on);

/* ---- A.3-A.8  LoopClosureDetector -------------------------------------- */
typedef struct kmo_lcd kmo_lcd;
kmo_lcd* kmo_lcd_create(const kmo_params*);
void kmo_lcd_destroy(kmo_lcd*);
/* ---- A.9 TemplatedVocabulary::transform (row f1) ---- */
typedef struct kmo_vocab kmo_vocab;
kmo_vocab* kmo_vocab_create(int k, int L, const uint8_t* node_desc, const double* word_weights);
void kmo_vocab_destroy(kmo_vocab*);
int kmo_vocab_transform(const kmo_vocab*, const uint8_t* desc, int F, uint32_t* ids, double* vals);
void kmo_lcd_add_bow(kmo_lcd*, uint64_t robot, uint64_t pose,
                     const uint32_t* ids, const float* vals, int n);
void kmo_lcd_add_frame(kmo_lcd*, uint64_t robot, uint64_t pose,
                       const uint8_t* desc, const double* bearings,
                       const double* points, int F);
int kmo_lcd_detect_loop_with_robot(kmo_lcd*, uint64_t robot, uint64_t q_robot,
                                   uint64_t q_pose, const uint32_t* ids,
                                   const float* vals, int n, uint64_t* out_robot,
                                   uint64_t* out_pose, double* out_score, int cap);
int kmo_lcd_detect_loop(kmo_lcd*, uint64_t q_robot, uint64_t q_pose,
                        const uint32_t* ids, const float* vals, int n,
                        uint64_t* out_robot, uint64_t* out_pose,
                        double* out_score, int cap);
int kmo_lcd_compute_matched_indices(kmo_lcd*, uint64_t qr, uint64_t qp,
                                    uint64_t mr, uint64_t mp, uint32_t* i_query,
                                    uint32_t* i_match, int cap);
/* returns 1 on success (reference `true`), 0 otherwise; count is in-out */
int kmo_lcd_geometric_verification_nister(kmo_lcd*, uint64_t qr, uint64_t qp,
                                          uint64_t mr, uint64_t mp,
                                          uint32_t* inl_q, uint32_t* inl_m,
                                          int* count, double* R);
int kmo_lcd_recover_pose(kmo_lcd*, uint64_t qr, uint64_t qp, uint64_t mr,
                         uint64_t mp, uint32_t* inl_q, uint32_t* inl_m,
                         int* count, double* T);
/* the same with a rotation prior (row-major 3x3, nullable): used iff ransac_use_1point_3d3d */
int kmo_lcd_recover_pose_prior(kmo_lcd*, uint64_t qr, uint64_t qp, uint64_t mr,
                               uint64_t mp, uint32_t* inl_q, uint32_t* inl_m,
                               int* count, const double* R_prior, double* T);
/* Ransac over the 1-point problem: rotation R (3x3) given, model = [R | p1_i - R p2_i] */
void kmo_ransac_onepoint(const double* p1, const double* p2, int N, const double* R, double thr,
                         double prob, int max_iter, uint32_t seed, kmo_ransac_result* res, uint32_t* inliers);

/* One verified candidate of a query (the record the product also emits). */
typedef struct kmo_result {
  uint64_t q_robot, q_pose, m_robot, m_pose;
  double norm_bow_score;
  int32_t n_matches, mono_inliers, stereo_inliers;
  int32_t status; /* 0 ok, 1 mono failed, 2 stereo failed */
  double R_mono[9];
  double T[12];
} kmo_result;

/* Full query (the unit of the headline metric): detectLoop over all robot
 * DBs, keep the top_k_verify best normalised scores (score desc, robot asc,
 * pose asc), verify each: matches -> mono RANSAC -> stereo RANSAC.
 * Query frame is passed explicitly (it is not in the store).  prev_* is the
 * previous BoW of the querying robot (NSS).  Returns number of records. */
int kmo_lcd_query(kmo_lcd*, uint64_t q_robot, uint64_t q_pose,
                  const uint32_t* ids, const float* vals, int n,
                  const uint32_t* prev_ids, const float* prev_vals, int prev_n,
                  const uint8_t* desc, const double* bearings,
                  const double* points, int F, kmo_result* out, int cap);
/* batch with OpenMP over queries (threads<=0: all cores). CSR-style inputs. */
int kmo_lcd_query_batch(kmo_lcd*, int B, const uint64_t* q_robot,
                        const uint64_t* q_pose, const int64_t* bow_off,
                        const uint32_t* ids, const float* vals,
                        const int64_t* prev_off, const uint32_t* prev_ids,
                        const float* prev_vals, const uint8_t* desc,
                        const double* bearings, const double* points, int F,
                        kmo_result* out, int cap_per_query, int32_t* counts,
                        int threads);
int kmo_num_threads(void);

/* ---- row f3: islands / temporal constraint ---- */
typedef struct kmo_island {
  uint64_t start_id, end_id, best_id;
  double island_score, best_score;
} kmo_island;
typedef struct kmo_temporal_state {
  int32_t temporal_entries;
  int32_t pad;
  uint64_t latest_query_id;
  kmo_island latest_island;
} kmo_temporal_state;
/* returns the number of islands (may exceed cap; only cap are written) */
int kmo_compute_islands(const uint64_t* ids, const double* scores, int n, int max_gap, int min_len,
                        kmo_island* out, int cap);
int kmo_check_temporal_constraint(kmo_temporal_state* st, uint64_t id, const kmo_island* island,
                                  int max_between_queries, int max_between_islands,
                                  int min_temporal_matches);

#ifdef __cplusplus
}
#endif
#endif
