/*
 * kml.h — C ABI of libkml.so, the B200-native loop-closure front end.
 *
 * Drop-in boundary for the data-parallel hot path of Kimera-Multi's
 * distributed place recognition: kimera_multi_lcd::LoopClosureDetector
 * (detectLoop / detectLoopWithRobot / computeMatchedIndices /
 * geometricVerificationNister / recoverPose) with the DBoW2 database query,
 * OpenCV BFMatcher kNN and OpenGV RANSAC underneath.  The reference keeps
 * that code in un-vendored repositories (/root/reference/kimera_multi.repos:
 * 14-17 dbow2_catkin, 54-57 kimera_multi_lcd, 102-105 opencv3_catkin,
 * 106-109 opengv_catkin); the call graph each entry point replaces is drawn
 * in /root/reference/images/kimera-multi.drawio:2533-2662 and the only
 * literal source lines are /root/reference/docker/copy/kimera_multi_lcd.patch:
 * 30-38 (LoopClosureDetector::loadAndInitialize).  Each declaration below
 * cites the reference interface it stands in for.
 *
 * Conventions: plain C, every function returns int: 0 = ok, >0 = "no result"
 * (the reference's `false` returns), <0 = error (kml_last_error() explains).
 * Caller owns all inputs; the library copies what it keeps.  Outputs are
 * caller-allocated (ptr, capacity, *count).  Arrays are row-major, packed,
 * little-endian.  All host pointers are ordinary host memory.  There is no
 * CPU fallback: every compute entry point launches sm_100a kernels and fails
 * with KML_ERR_CUDA when no B200-class device is usable.
 */
#ifndef KML_H_
#define KML_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define KML_OK 0
/* "no result" codes (reference returns false / empty) */
#define KML_NO_DB 1             /* detectLoopWithRobot: no database for robot */
#define KML_NO_PREV_BOW 2       /* no previous BoW vector for the NSS factor */
#define KML_NSS_TOO_LOW 3       /* nss < min_nss_factor */
#define KML_NO_MATCH 4          /* nothing survived the alpha*nss cut */
#define KML_NO_FRAME 5          /* frameExists() false for one of the ids */
#define KML_TOO_FEW_POINTS 6    /* fewer correspondences than the sample size */
#define KML_RANSAC_FAIL 7       /* computeModel() returned false */
#define KML_TOO_FEW_INLIERS 8   /* inlier count / ratio gate failed */
#define KML_INTER_ROBOT_ONLY 9  /* intra-robot query with inter_robot_only */
/* errors */
#define KML_ERR_ARG (-1)
#define KML_ERR_CUDA (-2)
#define KML_ERR_NCCL (-3)
#define KML_ERR_CAPACITY (-4)
#define KML_ERR_IO (-6)               /* shard file cannot be opened / read / written, or is corrupt */
#define KML_ERR_STREAM_EXHAUSTED (-5) /* pre-drawn sample stream ran out (kept for ABI stability: the stream
                                        * now covers every draw the reference loop can consume) */

/* LcdParams / LcdTpParams (kimera_multi_lcd include/kimera_multi_lcd/types.h;
 * set once in loadAndInitialize, kimera_multi_lcd.patch:30-31).  Defaults:
 * /root/reference/params/D455/LcdParams.yaml and SURVEY.md Appendix D. */
typedef struct kml_params {
  int32_t inter_robot_only;
  double alpha;
  int32_t dist_local;
  int32_t max_db_results;
  double min_nss_factor;
  int32_t max_nrFrames_between_queries;
  int32_t max_nrFrames_between_islands; /* row f3 (islands) — stored, unused */
  int32_t min_temporal_matches;         /* row f3 — stored, unused */
  int32_t max_intraisland_gap;          /* row f3 — stored, unused */
  int32_t min_matches_per_island;       /* row f3 — stored, unused */
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
  int32_t ransac_randomize; /* must be 0: the sample stream is pre-drawn */
  uint32_t ransac_seed;     /* 12345 = OpenGV's fixed seed */
  int32_t top_k_verify;     /* benchmark knob: candidates verified per query */
  int32_t matcher_norm;     /* 0 = NORM_HAMMING (BASELINE.json north_star, default); 1 = NORM_L1 over the 32
                             * descriptor bytes = what upstream's DescriptorMatcher::create(3) really selects
                             * (kimera_multi_lcd.patch:34-35) */
  int32_t matcher_engine;   /* NORM_HAMMING only: 0 = POPC pipe (LOP3 / POPC kernel), 1 = tensor cores (tcgen05.mma
                             * kind::i8 on +-1 expanded descriptors, s32 accumulators in TMEM); same keys, bit for bit */
  int32_t mono_algorithm;   /* ransac_2d2d_algorithm of /root/reference/params/D455/LcdParams.yaml:68-73:
                             * 0 = NISTER (default, what geometricVerificationNister names), 1 = STEWENIUS (action
                             * matrix of the same ten constraints; real solutions) */
  int32_t ransac_use_1point_3d3d; /* /root/reference/params/D455/LcdParams.yaml:58: recoverPose takes the mono
                             * rotation as given and samples ONE point pair per hypothesis (translation only) */
  int32_t reserved0;
} kml_params;

/* One verified candidate (VLCEdge of kimera_distributed + the counters the
 * evaluation scripts read: /root/reference/evaluation/lc_result.py:117-138). */
typedef struct kml_result {
  uint64_t q_robot, q_pose, m_robot, m_pose;
  double norm_bow_score;
  int32_t n_matches, mono_inliers, stereo_inliers;
  int32_t status; /* 0 loop closure, 1 mono verification failed, 2 pose recovery failed, 3 frame missing */
  double R_mono[9]; /* R_query_match of geometricVerificationNister */
  double T[12];     /* T_query_match = [R|t] row-major 3x4, x_q = R x_m + t */
} kml_result;

/* logLcdStat() counters (/root/reference/images/kimera-multi.drawio:86-100)
 * plus device timings (CUDA events on the handle's stream, last batch). */
typedef struct kml_stats {
  uint64_t total_bow_matches;
  uint64_t total_geom_verifications_mono;
  uint64_t total_geometric_verifications;
  uint64_t kernel_launches; /* kernels launched by this handle so far */
  float ms_bow, ms_match, ms_mono, ms_stereo, ms_total;
  uint64_t bow_postings_last; /* inverted-file postings touched by the last batch */
  uint64_t mono_hypotheses_last, stereo_hypotheses_last, pairs_last;
  /* sum over problems of (draws consumed x correspondences): residual evaluations of the reference loop */
  uint64_t mono_residuals_last, stereo_residuals_last;
} kml_stats;

typedef struct kml_handle kml_handle;

void kml_default_params(kml_params* p);
/* LcdParams.yaml -> kml_params: reads the flat "%YAML:1.0" key: value file the reference feeds its
 * loop-closure detector (/root/reference/params/D455/LcdParams.yaml:1-74, located through
 * params_folder, /root/reference/launch/kimera_vio_jackal.launch:37) on top of *p (call
 * kml_default_params first; keys the file does not hold keep their value).  Mapping: alpha,
 * max_db_results, min_nss_factor, min_temporal_matches, min_matches_per_island, max_intraisland_gap,
 * max_nrFrames_between_islands, max_nrFrames_between_queries, lowe_ratio by name;
 * recent_frames_window -> dist_local; ransac_threshold_2d2d / _3d3d -> ransac_threshold_mono /
 * ransac_threshold; ransac_max_iterations and ransac_probability -> both RANSACs;
 * min_nr_3d3d_inliers -> geometric_verification_min_inlier_count; ransac_randomize,
 * ransac_use_1point_3d3d by name; ransac_2d2d_algorithm (OpenGV enum: 0 STEWENIUS, 1 NISTER) ->
 * mono_algorithm (1, 0); matcher_type -> matcher_norm: with literal_matcher_enum != 0 the number is
 * read as cv::DescriptorMatcher::create() reads it (3 = BRUTEFORCE_L1, 4 = BRUTEFORCE_HAMMING,
 * kimera_multi_lcd.patch:34-35), else as the file's own comment table documents it (3 =
 * BRUTEFORCE_HAMMING, LcdParams.yaml:40-46).  Returns KML_OK, or KML_ERR_ARG (unreadable file,
 * unsupported matcher type); *n_mapped (nullable) = how many keys were taken. */
int kml_params_from_yaml(const char* path, int literal_matcher_enum, kml_params* p, int* n_mapped);
/* loadAndInitialize(params) on CUDA device `device` */
int kml_create(const kml_params* p, int device, kml_handle** out);
int kml_destroy(kml_handle* h);
/* A query lane: a second handle that shares `parent`'s databases, frame store and vocabulary
 * (read-only while queries run) but owns its stream and batch buffers, so that two host threads
 * can keep two query batches in flight on one GPU (the tail rounds of one batch's RANSAC overlap
 * the bulk of the other's).  add* / vocab_set on any of them must not overlap queries. */
int kml_create_lane(kml_handle* parent, kml_handle** out);
const char* kml_last_error(const kml_handle* h); /* h may be NULL: last create error */
int kml_get_stats(kml_handle* h, kml_stats* out);
int kml_device_count(void);

/* ---- database side ------------------------------------------------------ */
/* LoopClosureDetector::addBowVector(RobotPoseId, BowVector); word ids must
 * be ascending (DBoW2::BowVector is a std::map).  Weights are float32 as in
 * pose_graph_tools BowVector.msg. */
int kml_add_bow(kml_handle* h, uint64_t robot, uint64_t pose, const uint32_t* word_ids,
                const float* word_vals, int n);
/* bulk variant: `count` vectors in CSR form (off[count+1]) for one robot */
int kml_add_bow_bulk(kml_handle* h, uint64_t robot, const uint64_t* poses, int count,
                     const int64_t* off, const uint32_t* word_ids, const float* word_vals);
/* LoopClosureDetector::addVLCFrame(RobotPoseId, VLCFrame): descriptors_mat_
 * [F][32] CV_8U, versors_ [F][3] unit bearings, keypoints_ [F][3] 3-D points
 * (zero = invalid depth). */
int kml_add_frame(kml_handle* h, uint64_t robot, uint64_t pose, const uint8_t* desc,
                  const double* bearings, const double* points, int F);
int kml_add_frames_bulk(kml_handle* h, uint64_t robot, const uint64_t* poses, int count,
                        const uint8_t* desc, const double* bearings, const double* points,
                        int F);
int kml_frame_exists(kml_handle* h, uint64_t robot, uint64_t pose); /* 1/0 */
int kml_bow_exists(kml_handle* h, uint64_t robot, uint64_t pose);   /* 1/0 */
int kml_num_bow_for_robot(kml_handle* h, uint64_t robot);
int kml_get_bow_vector(kml_handle* h, uint64_t robot, uint64_t pose, uint32_t* ids,
                       float* vals, int cap, int* count);

/* DBoW2::TemplatedDatabase::query(vec, ret, max_results, max_id) on the
 * database of `robot`; results best first, Score in [0,1]. */
int kml_db_query(kml_handle* h, uint64_t robot, const uint32_t* ids, const float* vals, int n,
                 int max_results, int max_id, uint32_t* out_entry, double* out_score, int cap,
                 int* count);
/* TemplatedVocabulary::score == L1Scoring::score */
int kml_bow_score(kml_handle* h, const uint32_t* ids1, const float* vals1, int n1,
                  const uint32_t* ids2, const float* vals2, int n2, double* out);

/* ---- LoopClosureDetector query API -------------------------------------- */
int kml_detect_loop_with_robot(kml_handle* h, uint64_t robot, uint64_t q_robot, uint64_t q_pose,
                               const uint32_t* ids, const float* vals, int n,
                               uint64_t* out_robot, uint64_t* out_pose, double* out_score,
                               int cap, int* count);
int kml_detect_loop(kml_handle* h, uint64_t q_robot, uint64_t q_pose, const uint32_t* ids,
                    const float* vals, int n, uint64_t* out_robot, uint64_t* out_pose,
                    double* out_score, int cap, int* count);
int kml_compute_matched_indices(kml_handle* h, uint64_t q_robot, uint64_t q_pose,
                                uint64_t m_robot, uint64_t m_pose, uint32_t* i_query,
                                uint32_t* i_match, int cap, int* count);
/* inl_q / inl_m / count are in-out, R = R_query_match (row-major 3x3) */
int kml_geometric_verification_nister(kml_handle* h, uint64_t q_robot, uint64_t q_pose,
                                      uint64_t m_robot, uint64_t m_pose, uint32_t* inl_q,
                                      uint32_t* inl_m, int* count, double* R);
/* T = T_query_match row-major 3x4; R_prior may be NULL.  Arun's 3-point solver does not read it;
 * with kml_params.ransac_use_1point_3d3d it IS the rotation and hypotheses are single point pairs */
int kml_recover_pose(kml_handle* h, uint64_t q_robot, uint64_t q_pose, uint64_t m_robot,
                     uint64_t m_pose, uint32_t* inl_q, uint32_t* inl_m, int* count,
                     const double* R_prior, double* T);

/* ---- batched fast paths -------------------------------------------------- */
/* B full loop-closure queries: NSS + BoW scoring against every resident
 * robot DB + alpha*nss cut + top_k_verify candidates + kNN/Lowe + mono 5-pt
 * RANSAC + stereo Arun RANSAC.  BoW vectors in CSR form; query frames
 * [B][F][...] travel with the batch; out[B][cap_per_query], counts[B]. */
int kml_query_batch(kml_handle* h, int B, const uint64_t* q_robot, const uint64_t* q_pose,
                    const int64_t* bow_off, const uint32_t* ids, const float* vals,
                    const int64_t* prev_off, const uint32_t* prev_ids, const float* prev_vals,
                    const uint8_t* desc, const double* bearings, const double* points, int F,
                    kml_result* out, int cap_per_query, int32_t* counts);
/* Same, two-phase for measurement with inputs resident in HBM: upload once,
 * then run the device pipeline any number of times. */
int kml_query_batch_upload(kml_handle* h, int B, const uint64_t* q_robot,
                           const uint64_t* q_pose, const int64_t* bow_off, const uint32_t* ids,
                           const float* vals, const int64_t* prev_off, const uint32_t* prev_ids,
                           const float* prev_vals, const uint8_t* desc, const double* bearings,
                           const double* points, int F);
int kml_query_batch_run(kml_handle* h, kml_result* out, int cap_per_query, int32_t* counts);
/* Page-locked host memory for the caller's batch arrays: an array allocated here is copied to the
 * device from where it lies; ordinary (pageable) arrays are first staged through the handle's own
 * pinned buffer.  Results are identical either way. */
int kml_host_alloc(size_t bytes, void** out);
int kml_host_free(void* p);

/* cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, k=2): idx/dist are [nq][2];
 * a missing neighbour is idx 0xFFFFFFFF, dist 0xFFFF.  ms_kernel (nullable)
 * receives the device time of the matching kernels. */
int kml_hamming_knn2(kml_handle* h, const uint8_t* q, int nq, const uint8_t* t, int64_t nt,
                     uint32_t* idx, uint16_t* dist, float* ms_kernel);
/* cv::BFMatcher(NORM_L1).knnMatch(q, t, k=2) on the descriptor bytes (dist <= 8160) */
int kml_l1_knn2(kml_handle* h, const uint8_t* q, int nq, const uint8_t* t, int64_t nt,
                uint32_t* idx, uint16_t* dist, float* ms_kernel);
/* resident variant for roofline measurement: upload once, run `reps` times */
int kml_hamming_knn2_bench(kml_handle* h, const uint8_t* q, int nq, const uint8_t* t,
                           int64_t nt, int reps, uint32_t* idx, uint16_t* dist, float* ms_avg);

/* Batched opengv::sac::Ransac<PointCloudSacProblem> (Arun): P problems with
 * N correspondences each, p1/p2 [P][N][3].  Evaluates the draws the
 * reference loop would consume; per problem: model[12], n_inliers,
 * iterations, best_draw, inlier mask bits [P][ceil(N/32)] (nullable).
 * full_hypotheses != 0 disables the adaptive stop (config C4: evaluate all
 * max_iterations+1 hypotheses, the winner is the first best). */
int kml_ransac_arun_batch(kml_handle* h, int P, int N, const double* p1, const double* p2,
                          int full_hypotheses, double* models, int32_t* n_inliers,
                          int32_t* iterations, int32_t* best_draw, uint32_t* inlier_mask,
                          float* ms_kernel);
/* Row f4, ransac_use_1point_3d3d (/root/reference/params/D455/LcdParams.yaml:58): the point-cloud
 * problem with the rotation given, R [P][9] row-major; a hypothesis is ONE correspondence,
 * model = [R | p1_i - R p2_i], residual and thresholds as kml_ransac_arun_batch. */
int kml_ransac_onepoint_batch(kml_handle* h, int P, int N, const double* p1, const double* p2, const double* R,
                              int full_hypotheses, double* models, int32_t* n_inliers, int32_t* iterations,
                              int32_t* best_draw, uint32_t* inlier_mask, float* ms_kernel);
/* Batched Ransac<CentralRelativePoseSacProblem>(NISTER): f1/f2 [P][N][3]. */
int kml_ransac_nister_batch(kml_handle* h, int P, int N, const double* f1, const double* f2,
                            int full_hypotheses, double* models, int32_t* n_inliers,
                            int32_t* iterations, int32_t* best_draw, uint32_t* inlier_mask,
                            float* ms_kernel);

/* ---- vocabulary transform (the step right before the path) ---
 * DBoW2::TemplatedVocabulary with k children per node and L levels (mit_voc.yml /
 * ORBvoc: k=10, L=6), TF-IDF weighting, L1 scoring.  node_desc[n_nodes][32] in
 * breadth-first order (the k level-1 nodes first, then k^2, ...; the children of
 * node i of a level are nodes i*k .. i*k+k-1 of the next); word id = index of the
 * leaf within the last level; word_weights[k^L] = the leaves' IDF weights. */
int kml_vocab_set(kml_handle* h, int k, int L, const uint8_t* node_desc, const double* word_weights);
/* TemplatedVocabulary::transform(features, BowVector) for B frames of F descriptors:
 * CSR output (out_off[B+1]; word ids ascending; L1-normalised TF-IDF values). */
int kml_transform_batch(kml_handle* h, int B, int F, const uint8_t* desc, int64_t* out_off,
                        uint32_t* out_ids, double* out_vals, int64_t cap, float* ms_kernel);

/* microbenchmarks for the roofline denominators (ops per second) */
int kml_peak_popc(kml_handle* h, double* popc32_per_s);
int kml_peak_fp64(kml_handle* h, double* flop_per_s);
/* device-side stopwatch across lanes (CUDA events): begin on h's stream; end waits for all
 * work enqueued on the listed lanes, returns elapsed device milliseconds */
int kml_timer_begin(kml_handle* h);
int kml_timer_end(kml_handle* h, kml_handle** lanes, int n_lanes, float* ms);
/* measurement hygiene: overwrite a 256 MiB device buffer (> 126 MB L2) on the
 * handle's stream so the next timed step starts with a cold L2 */
int kml_flush_l2(kml_handle* h);

/* ---- row f4 (part): shard persistence --------------------------------------
 * Everything addBowVector / addVLCFrame stored for this handle's robots (BoW vectors in
 * DBoW2 EntryId order, live frames with descriptors / bearings / points) to one file, and
 * back into an empty detector through the same add paths: queries after a reload return the
 * same records.  (The reference front end has no persistence; DBoW2's save()/load() cover
 * the database only.) */
int kml_save_shard(kml_handle* h, const char* path);
int kml_load_shard(kml_handle* h, const char* path);

/* the merge rule alone, on host blocks: nranks x { kml_result[B][cap_in]; int32 counts[B] } */
int kml_merge_shard_records(const void* blocks, int nranks, int B, int cap_in, int cap, kml_result* out,
                            int32_t* counts);

/* ---- row f3: post filters and wire layout around the path (host logic) ----
 * Kimera-VIO's detection flow "2 computeIslands() 3 checkTemporalConstraint()"
 * (/root/reference/images/kimera-multi.drawio:1565) with the
 * thresholds of /root/reference/params/D455/LcdParams.yaml:5,9-12. */
typedef struct kml_island {
  uint64_t start_id, end_id; /* first / last result id of the group */
  uint64_t best_id;          /* id with the largest score in the group */
  double island_score;       /* sum of the group's scores */
  double best_score;
} kml_island;
typedef struct kml_temporal_state { /* zero-initialise; one per (querying robot, database) */
  int32_t temporal_entries;
  int32_t pad;
  uint64_t latest_query_id;
  kml_island latest_island;
} kml_temporal_state;
/* LCDStatus values written to output_lcd_status.csv (/root/reference/evaluation/lc_result.py:143-162) */
#define KML_LCD_LOOP_DETECTED 0
#define KML_LCD_NO_MATCHES 1
#define KML_LCD_LOW_NSS_FACTOR 2
#define KML_LCD_LOW_SCORE 3
#define KML_LCD_NO_GROUPS 4
#define KML_LCD_FAILED_TEMPORAL_CONSTRAINT 5
#define KML_LCD_FAILED_GEOM_VERIFICATION 6
#define KML_LCD_FAILED_POSE_RECOVERY 7
/* LcdThirdPartyWrapper::computeIslands: ids/scores = surviving results of one database */
int kml_compute_islands(const uint64_t* ids, const double* scores, int n, int max_intraisland_gap,
                        int min_matches_per_island, kml_island* out, int cap, int* count);
/* LcdThirdPartyWrapper::checkTemporalConstraint: 1 = passes, 0 = fails, <0 error */
int kml_check_temporal_constraint(kml_temporal_state* st, uint64_t query_id, const kml_island* island,
                                  int max_nrFrames_between_queries, int max_nrFrames_between_islands,
                                  int min_temporal_matches);
/* detectLoopWithRobot + islands + temporal check: KML_OK and the best island's best entry,
 * or KML_NO_MATCH / KML_NSS_TOO_LOW with *lcd_status saying why */
int kml_detect_loop_islands(kml_handle* h, uint64_t robot, uint64_t q_robot, uint64_t q_pose,
                            const uint32_t* ids, const float* vals, int n, int max_intraisland_gap,
                            int min_matches_per_island, int max_nrFrames_between_islands,
                            int min_temporal_matches, kml_temporal_state* st, uint64_t* match_pose,
                            double* match_score, kml_island* best_island, int* lcd_status);
/* addVLCFrame from the VLCFrameMsg layout: float32 xyz clouds for versors and keypoints
 * (/root/reference/images/kimera-multi.drawio:385-414) */
int kml_add_frame_msg(kml_handle* h, uint64_t robot, uint64_t pose, const uint8_t* desc,
                      const float* versors_xyz, const float* keypoints_xyz, int F);

/* ---- multi-GPU (one process per GPU; databases sharded by robot) -------- */
#define KML_UNIQUE_ID_BYTES 128
int kml_comm_unique_id(void* id_out /* KML_UNIQUE_ID_BYTES */);
int kml_comm_init(kml_handle* h, int nranks, int rank, const void* unique_id);
/* every rank passes the same batch; each scores/verifies against its own
 * shard; per-rank top-k records are merged with one ncclAllGather and every
 * rank receives the merged, globally re-ranked list. */
int kml_query_batch_sharded(kml_handle* h, kml_result* out, int cap_per_query,
                            int32_t* counts);
/* Several query lanes of one rank exchanging concurrently (each lane has its own communicator):
 * `seq` is the batch's global sequence number, the same on every rank; a lane's all-gather is
 * enqueued only after the all-gathers of all lower sequence numbers of this detector family, so
 * every rank submits its collectives to the GPU in one order.  kml_comm_seq_reset sets the next
 * number expected (0 after kml_create). */
int kml_query_batch_sharded_seq(kml_handle* h, uint64_t seq, kml_result* out, int cap_per_query,
                                int32_t* counts);
int kml_comm_seq_reset(kml_handle* h, uint64_t next_seq);
/* kml_query_batch for the sharded detector: host buffers in (every rank passes the same batch),
 * merged records out; the copies are enqueued in front of the pipeline and covered by its single
 * wait.  use_seq = 0 ignores `seq` (one lane per rank). */
int kml_query_batch_sharded_host(kml_handle* h, int use_seq, uint64_t seq, int B, const uint64_t* q_robot,
                                 const uint64_t* q_pose, const int64_t* bow_off, const uint32_t* ids,
                                 const float* vals, const int64_t* prev_off, const uint32_t* prev_ids,
                                 const float* prev_vals, const uint8_t* desc, const double* bearings,
                                 const double* points, int F, kml_result* out, int cap_per_query, int32_t* counts);
/* the sharded query's device tail alone (merge_shards_kernel) on host blocks, laid out as for
 * kml_merge_shard_records */
int kml_merge_shard_records_device(kml_handle* h, const void* blocks, int nranks, int B, int cap_in, int cap,
                                   kml_result* out, int32_t* counts);

#ifdef __cplusplus
}
#endif
#endif
