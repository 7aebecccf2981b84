// kernels.h — host-callable launchers of the sm_100a kernels of libkml.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/kml.h"

namespace kml {

// ------------------------------------------------------------- hamming.cu
struct HamJob {
  const uint8_t* q;  // query descriptors [nq][32]
  const uint8_t* t;  // train descriptors [nt][32], nt <= 2^20
  uint32_t* keys;    // out [nq][2] packed (dist<<20 | trainIdx), 0xFFFFFFFF = none
  int nq, nt;
};
// norm: 0 = NORM_HAMMING (key = dist<<20 | idx, nt <= 2^20), 1 = NORM_L1 (key = dist<<18 | idx, nt <= 2^18)
inline int knn_key_shift(int norm) { return norm == 1 ? 18 : 20; }
void launch_hamming_jobs(const HamJob* d_jobs, int njobs, int norm, cudaStream_t s);
// NORM_HAMMING on the tensor cores (hamming_tc.cu): same keys as launch_hamming_jobs(.., 0, ..)
void launch_hamming_jobs_tc(const HamJob* d_jobs, int njobs, cudaStream_t s);
void launch_knn2_reduce(const uint32_t* partial, int nranges, int nq, int64_t range_len, int norm,
                        uint32_t* idx, uint16_t* dist, cudaStream_t s);
void launch_lowe_compact(const uint32_t* keys, const int* nq_arr, int key_stride, double lowe, int norm,
                         uint16_t* iq, uint16_t* im, int* M, int P, cudaStream_t s);
double measure_popc_peak(cudaStream_t s);
double measure_fp64_peak(cudaStream_t s);

// ----------------------------------------------------------------- bow.cu
// Device view of one robot's inverted file: row w (= word id) holds pool[rows[w].x, + rows[w].y).
struct BowDb {
  const uint2* rows;     // [W] {start (even: rows are 16-byte aligned), len}
  const uint2* postings; // the pool: {entry, float bits of weight}, every row ascending in entry
  uint32_t W;            // number of word rows
  uint32_t n_entries;
  // db_EntryId_to_PoseId_ and the dense frame-store index of every entry's keyframe (-1: the
  // keyframe's VLC frame has not been stored), read by the device-side candidate selection
  const uint64_t* entry_pose;   // [n_entries]
  const int32_t* entry_frame;   // [n_entries]
  uint64_t robot;
};
constexpr int kBowMaxWords = 1024;  // max words per BoW vector handled on device
constexpr int kBowMaxK = 128;       // max max_results handled on device
constexpr double kBowScale = 4611686018427387904.0;  // 2^62 fixed-point scale
// One CTA per (query, db, entry tile).  Query BoWs in CSR form (q_off).
// out_entry/out_score: [B][n_db][n_tiles][K]; out_count: [B][n_db][n_tiles].
// nss (nullable): [B] L1 score between query and prev BoW.
struct BowArgs {
  const BowDb* dbs;
  int n_db;
  int B;
  const int64_t* q_off;
  const uint32_t* q_ids;
  const float* q_vals;
  const int64_t* p_off;  // previous BoW (nullable with nss == nullptr)
  const uint32_t* p_ids;
  const float* p_vals;
  int K;
  const int32_t* max_id;  // [n_db] or nullptr (-1 = no limit)
  int tile_entries;       // entries per tile (shared-memory accumulators)
  int n_tiles;
  uint32_t* out_entry;
  double* out_score;
  int32_t* out_count;
  double* nss;
  unsigned long long* postings_touched;  // nullable: algorithmic postings counter
};
void launch_bow(const BowArgs& a, cudaStream_t s);
// incremental update of an inverted file (bow_merge.h BowUpdate): row copies, new postings, row table
void launch_bow_append(uint2* rows, uint2* pool, const uint4* copies, int n_copies, const uint4* posts, int n_posts,
                       const uint4* row_cmds, int n_rows, cudaStream_t s);

// -------------------------------------------------------------- ransac.cu
// Per-pair RANSAC state (opengv::sac::Ransac::computeModel locals)
struct SacState {
  int32_t iterations, skipped, draws, best, best_draw, done, exhausted;
  int32_t r_begin, r_end;  // draws [r_begin, r_end) are evaluated by the next chunk launch
  int32_t unit_bearings;   // mono: every bearing of the problem has | |f|^2 - 1 | <= 1e-12 (set by sac_init)
  double k;
};
struct SacArgs {
  int P;                 // problems
  const double* a;       // mono: query bearings; stereo: query points   [P][stride][3]
  const double* b;       // mono: match bearings; stereo: match points   [P][stride][3]
  const int32_t* N;      // correspondences per problem [P]
  int stride;            // row stride (max correspondences)
  const uint32_t* raw;   // pre-drawn mt19937()>>1 stream
  int raw_len;
  uint16_t* perm;        // [P][stride] persistent shuffle state
  int cap_draws;         // draws available per problem (raw_len / S)
  double* models;        // [P][kRoundCap][12] models of the draws of the current round
  int first, n_rounds;   // round schedule (sac_round_draws): draws of round 0, rounds enqueued blindly
  int alg;               // mono: 0 = NISTER, 1 = STEWENIUS (row f4)
  int fo_stride;         // mono: doubles of stage-1 output per draw (geom::kFrontOut / kFrontOutStew)
  double* fsol;          // mono: [P][kRoundCap][fo_stride] stage-1 output per draw of the round
  int32_t* nroot;        // mono: [P][kRoundCap] R0 | R1<<8 real-root counts of the two chains
  double* brk;           // mono: [P][kRoundCap][20][2] isolating brackets
  uint32_t* fb_list;     // mono: deferred root isolations of the round, slot*2 + chain
  unsigned int* fb_count;
  // mono: (draw, root) items of the round — one bracketed real root each.  A draw owns the
  // contiguous range item_base[slot] + [0, R0+R1) of the item arrays.
  unsigned int* item_count;
  uint32_t* item_base;   // [P][kRoundCap]
  uint32_t* item_list;   // [items] slot*32 + root
  double* item_q;        // [items] smallest candidate quality of the item (< 1e6), if status == 2
  double* item_model;    // [items][12] the candidate that attains it
  uint8_t* item_status;  // [items] 0 root not refined, 1 refined without usable candidate, 2 scored
  unsigned int item_cap; // capacity of the item arrays and of fb_list
  unsigned int* overflow;  // set when a round needed more than item_cap items (results of the batch are void)
  unsigned int* pending;   // nullable: sac_select counts the problems that are not done yet
  // Every problem starts from the identity permutation and the same pre-drawn stream, so its sample
  // sequence depends only on its number of correspondences N: sample_tab[N][draw][S] (N <=
  // tab_nmax) holds it for every draw the loop can consume, built once per (S, seed).  When it is
  // there the kernels read their samples from it and neither sac_init nor sac_replay draws
  // anything (the per-problem shuffle state `perm` and `samples` are unused).  Null for problems
  // larger than the table: the warp-serial per-problem sampler runs instead.
  const uint16_t* sample_tab;
  int tab_nmax;
  uint16_t* samples;     // [P][kRoundCap][S] samples of the current round's draws, slot = draw - r_begin
  int32_t* valid;        // [P][kRoundCap]
  int32_t* counts;       // [P][kRoundCap]
  // active lists of the rounds: active[(r & 1) * P + i], i < n_active[r & 1], are the problems round r
  // works on (sac_init writes list 0, the replay of round r writes the list of round r + 1)
  int32_t* active;       // [2][P]
  unsigned int* n_active;  // [2]
  SacState* st;          // [P]
  double* best_model;    // [P][12]
  const double* ktable;  // [(Nmax+1)*(Nmax+1)] k as function of (N, best count)
  int ktable_n;          // Nmax+1
  double threshold;      // mono: 1-cos sum threshold; stereo: metres
  double sq_crit;        // stereo: smallest s with sqrt(s) >= threshold
  int max_iterations;
  int full;              // evaluate every draw up to max_iterations+1 (no adaptive stop)
  int force_generic;     // test hooks: bit 0 (env KML_FORCE_GENERIC_ISOLATE) skip the register fast path of stage 2;
                         // bit 1 (env KML_NO_ROOT_GRID2) skip the 256-cell grid, every deferred chain is bisected
  // stereo, row f4 (ransac_use_1point_3d3d): the rotation is given per problem, prior[p] = row-major
  // 3x4 whose left 3x3 block is R (the mono model); a draw is ONE correspondence, model = [R | p1 - R p2]
  int onept;
  const double* prior;   // [P][12]
  uint32_t* inlier_mask; // [P][mask_words]
  int mask_words;
  int32_t* n_inliers;    // [P]
};
constexpr int kMonoChunk = 64;     // hypotheses per CTA (mono) and size of round 0
constexpr int kStereoChunk = 64;    // hypotheses per CTA (stereo) and size of round 0
constexpr int kStereoThreads = 128; // threads of a stereo CTA (4 counting warps)
constexpr int kRoundCap = 512;     // most new draws any round evaluates per problem
// Round schedule (SacArgs::first, SacArgs::n_rounds).  Round r evaluates at most `first` new draws
// per problem for r = 0, 1, then doubles (first << (r - 1)), capped at kRoundCap.  Throughput mode
// (batches of >= 16 queries): first = 32, 7 blind rounds = 32,32,64,128,256,512,512 — nearly half of
// the candidate pairs of a batch end within 16 trials, and a draw evaluated for nothing costs
// throughput.  Latency mode (small batches: the GPU is idle anyway and every round is a chain of
// six dependent kernel latencies): first = 128, 4 blind rounds = 128,128,256,512.  The outcome does
// not depend on the schedule (the replay consumes draws in order); more rounds are added by the host
// only while a problem is pending.
constexpr int kSacFirstThroughput = 32, kSacRoundsThroughput = 7;
constexpr int kSacFirstLatency = 128, kSacRoundsLatency = 4;
// A single query (<= 32 candidate pairs): 512, 512 — every round is a chain of six kernels whose latency does not
// depend on the number of draws at this size, and 16 pairs x 512 draws are a fraction of one wave.
constexpr int kSacFirstSingle = 512, kSacRoundsSingle = 2, kSacSingleMaxP = 32;
__host__ __device__ inline int sac_round_draws(int round, int first) {
  const int d = round == 0 ? first : (round > 9 ? first << 9 : first << (round - 1));
  return d < kRoundCap ? d : kRoundCap;
}
void launch_sac_init(const SacArgs& a, int sample_size, cudaStream_t s);
// tab[N][cap_draws][S] for N in [0, nmax]: the draws of SampleConsensusProblem::drawIndexSample
// from a fresh shuffle state (one warp per N)
void launch_sample_table(const uint32_t* raw, int cap_draws, int sample_size, int nmax, uint16_t* tab, cudaStream_t s);
constexpr int kSampleTabMaxN = 1024;  // larger problems keep the per-problem sampler
// one round = chunk kernel over the pending draw range + replay; returns #kernels launched
int launch_mono_round(const SacArgs& a, int round, cudaStream_t s);
int launch_stereo_round(const SacArgs& a, int round, cudaStream_t s);
// *a.pending += number of problems whose loop has not ended
void launch_sac_pending(const SacArgs& a, cudaStream_t s);
void launch_mono_select(const SacArgs& a, cudaStream_t s);
void launch_stereo_select(const SacArgs& a, cudaStream_t s);

// gather kernels between the stages of the verification pipeline
struct PairDesc {
  int32_t q_slot;   // index into the batch's query frames
  int32_t m_frame;  // index into the frame store
};
struct GatherArgs {
  int P;
  const PairDesc* pairs;
  // query frames of the batch
  const double* qb; const double* qp; int qF;       // bearings / points [B][qF][3]
  // frame store
  const double* sb; const double* sp;               // bearings / points [total_feat][3]
  const int64_t* s_off;                             // feature offset per stored frame
  // matches
  const uint16_t* iq; const uint16_t* im; const int32_t* M; int stride;
  // outputs
  double* a; double* b;                             // [P][stride][3]
  int32_t* N;                                       // [P]
};
void launch_gather_bearings(const GatherArgs& g, cudaStream_t s);
// stereo gather: from mono inlier mask keep pairs with both 3-D norms > 1e-3;
// also rewrites (iq, im) to the kept subset (kq, km).  Gated by mono success.
struct StereoGatherArgs {
  GatherArgs g;
  const uint32_t* mono_mask; int mask_words;
  const int32_t* mono_ok;                            // [P] 1 if mono verification passed
  uint16_t* kq; uint16_t* km;                        // [P][stride]
};
void launch_gather_points(const StereoGatherArgs& g, cudaStream_t s);
// Counters of one batch, accumulated by the kernels (device atomics) and read back with the
// records: the logLcdStat() counters and the algorithmic work the rooflines are computed from.
struct BatchStats {
  unsigned long long postings;       // inverted-file postings touched (bow_score_kernel)
  unsigned long long survivors;      // totalBoWMatches: candidates left after the alpha*nss cut
  unsigned long long pairs;          // candidate pairs that entered verification
  unsigned long long mono_ok;        // pairs that passed geometricVerificationNister
  unsigned long long hyp_m, hyp_s;   // draws consumed by the reference loop (mono / stereo)
  unsigned long long res_m, res_s;   // draws consumed x correspondences
  unsigned long long eval_m, eval_s; // draws evaluated by the round schedule
  unsigned int pending_m, pending_s; // problems whose loop has not ended after the enqueued rounds
  unsigned int item_overflow;        // a mono round produced more (draw, root) items than the list holds
  unsigned int pad;
};

// mono acceptance gate + final record assembly
struct FinalizeArgs {
  int P;
  const SacState* mono_st; const int32_t* mono_inl; const int32_t* M; const double* mono_model;
  const SacState* st3; const int32_t* inl3; const int32_t* N3; const double* model3;
  int min_inliers; double min_ratio_mono; double min_ratio_stereo;
  int32_t* mono_ok;  // out (phase 1)
  // phase 2 outputs: the kml_result records of the batch, pair slot p = b*K + i -> recs[b*cap + i]
  const PairDesc* pairs; int K; int cap;
  kml_result* recs;
  BatchStats* stats;
};
void launch_mono_gate(const FinalizeArgs& f, cudaStream_t s);
void launch_finalize(const FinalizeArgs& f, cudaStream_t s);

// ---------------------------------------------------------------- select.cu
// detectLoop = detectLoopWithRobot over every resident robot database, steps 3-7 (SURVEY.md A.3;
// /root/reference/images/kimera-multi.drawio:2571-2580) on the device: per query the per-database
// result lists of the scorer are cut at alpha*nss, filtered (inter_robot_only, dist_local),
// normalised by nss, and the top_k_verify best by (score desc, robot asc, pose asc) become the
// query's records and candidate pairs.  Pair slot p = b*K + i is fixed, so every later kernel of
// the batch has a static grid; a slot without a candidate (or without a stored frame) is inactive:
// m_frame = -1, nq = 0.
constexpr int kSelMaxK = 128;  // upper bound of top_k_verify and max_db_results on this path
struct SelectArgs {
  const BowDb* dbs; int n_db; int B; int n_tiles; int Kdb;
  const uint32_t* bow_entry; const double* bow_score; const int32_t* bow_count;  // [B][n_db][n_tiles][Kdb]
  const double* nss;                                                              // [B]
  const uint64_t* q_robot; const uint64_t* q_pose;                                // [B]
  double alpha, min_nss; int inter_robot_only; int dist_local;
  int K, cap;
  const uint8_t* s_desc; const int64_t* s_off; const int32_t* s_F;                // frame store
  const uint8_t* q_desc; int qF;                                                  // query frames [B][qF][32]
  kml_result* recs; int32_t* counts;                                              // [B][cap], [B]
  PairDesc* pairs; HamJob* jobs; int32_t* nq; uint32_t* keys; int key_stride;     // [B*K]
  BatchStats* stats;
};
void launch_select(const SelectArgs& a, cudaStream_t s);
// flag word of this rank's record block: 1 if RANSAC problems are pending or the item lists
// overflowed (the host continues / re-runs), else 0; also copies the overflow flag into the stats
void launch_block_flags(BatchStats* stats, const unsigned int* overflow, int32_t* flag, cudaStream_t s);

// Merge of the sharded query on the device: block r = { kml_result[B][cap_in]; int32 counts[B];
// int32 error; pad } at base + r*blk_stride, every list already ranked by (score desc, robot
// asc, pose asc); one warp per query keeps the best `cap` of the union.
struct MergeArgs {
  const uint8_t* base; size_t blk_stride; size_t counts_off; size_t err_off; int nranks; int B; int cap_in; int cap;
  kml_result* out; int32_t* counts;
  int32_t* err_out;  // max over the ranks' flag words
};
void launch_merge(const MergeArgs& a, cudaStream_t s);

}  // namespace kml
