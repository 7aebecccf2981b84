// handle.h — kml_handle: host-side state of the LoopClosureDetector replacement.
#pragma once
#include <time.h>

#include <condition_variable>
#include <map>
#include <mutex>
#include <memory>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/kml.h"
#include "bow_merge.h"
#include "common.cuh"
#include "kernels.h"

namespace kml {

typedef std::pair<uint64_t, uint64_t> RobotPoseId;
struct RobotPoseHash {
  size_t operator()(const RobotPoseId& id) const {
    uint64_t x = id.first * 0x9E3779B97F4A7C15ull ^ (id.second + 0x7F4A7C15ull + (id.first << 6));
    x ^= x >> 31; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 29;
    return (size_t)x;
  }
};

struct HostBow {
  std::vector<uint32_t> ids;
  std::vector<float> vals;
};

// One robot's DBoW2-style database: host append log + lazily rebuilt device CSR.
struct RobotDb {
  uint64_t robot = 0;
  // host log (entry order == insertion order == DBoW2 EntryId)
  std::vector<int64_t> off{0};
  std::vector<uint32_t> ids;
  std::vector<float> vals;
  std::vector<uint64_t> entry_to_pose;
  // dense frame index of the entry's keyframe, -1 = frame not stored (yet); kept current by
  // addBowVector / addVLCFrame
  std::vector<int32_t> entry_to_frame;
  std::map<uint64_t, uint32_t> pose_to_entry;
  // device inverted file: row table + posting pool (bow_merge.h BowInvFile is the host mirror).
  // dirty = the pool must be rebuilt from the log (first use, bulk load, or an append that did not
  // fit); otherwise the postings of the vectors added since the last query wait in `pend_*` and
  // are appended in place by the next query (cost independent of the database size).
  bool dirty = true;
  BowInvFile inv;
  DevBuf<uint2> rows;
  DevBuf<uint2> postings;
  std::vector<uint32_t> pend_word, pend_entry, pend_wbits;
  DevBuf<uint4> d_cmds;
  // device copies of entry_to_pose / entry_to_frame (read by select_candidates_kernel): entries
  // [0, synced) are on the device except the ones listed in frame_patches
  DevBuf<uint64_t> d_entry_pose;
  DevBuf<int32_t> d_entry_frame;
  size_t synced = 0;
  std::vector<uint32_t> frame_patches;
  uint32_t n_entries() const { return (uint32_t)entry_to_pose.size(); }
};

struct FrameRec {
  int64_t feat_off;  // offset into the feature arenas
  int32_t F;
  int32_t index;     // dense frame index
};

struct Comm;  // NCCL state (comm.cu)

}  // namespace kml

// Database, frame store, RANSAC constants and vocabulary: owned jointly by a handle and the
// query lanes cloned from it.  Lazily prepared device copies (CSR rebuild, k tables, frame
// offsets) are guarded by `mu`; mutating calls (add*, vocab_set) must not overlap queries.
struct kml_shared {
  std::mutex mu;
  std::vector<void*> retired;  // device tables replaced while another lane may still read them
  ~kml_shared() {
    for (void* p : retired) cudaFree(p);
  }
  // ---- BoW databases (one per robot, ordered: detectLoop visits ascending)
  std::map<uint64_t, std::unique_ptr<kml::RobotDb>> dbs;

  // ---- frame store (feature arenas in HBM)
  std::unordered_map<kml::RobotPoseId, kml::FrameRec, kml::RobotPoseHash> frames;
  std::vector<int64_t> frame_off_h;  // per dense frame index
  std::vector<int32_t> frame_F_h;
  int64_t n_feat = 0;
  kml::DevBuf<uint8_t> s_desc;     // [n_feat][32]
  kml::DevBuf<double> s_bear;      // [n_feat][3]
  kml::DevBuf<double> s_pts;       // [n_feat][3]
  kml::DevBuf<int64_t> s_off;      // [n_frames] feature offset of every stored frame
  kml::DevBuf<int32_t> s_F;        // [n_frames] features of every stored frame
  size_t frames_synced = 0;        // frames whose (offset, F) are on the device
  uint64_t version = 1;            // bumped by every add*: lanes re-upload their database views when it moved
  // kml_query_batch_sharded_seq: the lanes' all-gathers are enqueued in sequence-number order
  std::mutex seq_mu;
  std::condition_variable seq_cv;
  uint64_t seq_next = 0;

  // ---- pre-drawn sample stream + k tables
  std::vector<uint32_t> raw_h;
  kml::DevBuf<uint32_t> d_raw;
  // per-N sample tables (kernels.h SacArgs::sample_tab) for sample sizes 8, 3, 1
  kml::DevBuf<uint16_t> d_samptab[3];
  int samptab_nmax[3] = {-1, -1, -1};
  int ktable_n_mono = 0, ktable_n_stereo = 0, ktable_n_stereo1 = 0;
  kml::DevBuf<double> d_ktable_mono, d_ktable_stereo, d_ktable_stereo1;  // sample sizes 8, 3, 1

  // ---- vocabulary tree (row f1: TemplatedVocabulary::transform)
  int voc_k = 0, voc_L = 0;
  uint64_t voc_words = 0, voc_nodes = 0;
  kml::DevBuf<uint8_t> d_voc_nodes;  // [nodes][32], breadth-first, level 1 first
  kml::DevBuf<double> d_voc_w;       // [k^L] word weights (IDF)

};

struct kml_handle {
  kml_params prm;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[8] = {};
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_lane = nullptr;  // kml_timer_begin / kml_timer_end
  // Host waits of the batch path: several lanes per GPU and several GPUs per box put more waiting
  // threads on the host than it has cores, so a batch wait polls the stream and sleeps 20 us between
  // polls instead of spinning (a blocking-sync event was tried and doubled the 8-GPU step time);
  // single queries keep the low-latency spin of cudaStreamSynchronize.
  bool polite_wait = false;  // set for the duration of a batch of >= 16 queries
  void wait_stream() {
    if (!polite_wait) { KML_CUDA(cudaStreamSynchronize(stream)); return; }
    for (;;) {
      const cudaError_t e = cudaStreamQuery(stream);
      if (e == cudaSuccess) return;
      if (e != cudaErrorNotReady) KML_CUDA(e);
      struct timespec ts = {0, 20000};
      nanosleep(&ts, nullptr);
    }
  }
  // While a batch runs: B * top_k_verify, the most candidate pairs the batch can yield.  The
  // verification scratch is allocated for that many pairs, not for the pairs this batch happened
  // to produce, so equal-shaped batches never reallocate (cudaFree waits for the whole device,
  // i.e. for every other lane's kernels and pending collectives).  0 = size by the actual count.
  int pair_cap = 0;
  uint64_t views_version = 0;  // kml_shared::version the handle's d_dbs array was built from
  int views_n_db = 0, views_tile = 0, views_n_tiles = 0;
  std::string err;
  kml_stats stats = {};

  // ---- state shared by a handle and the query lanes cloned from it (kml_create_lane)
  std::shared_ptr<kml_shared> sh;
  kml::DevBuf<kml::BowDb> d_dbs;
  kml::DevBuf<int32_t> d_maxid;

  // ---- batch buffers (query side)
  int B = 0, qF = 0;
  std::vector<uint64_t> q_robot_h, q_pose_h;
  kml::DevBuf<int64_t> d_qoff, d_poff;
  kml::DevBuf<uint32_t> d_qids, d_pids;
  kml::DevBuf<float> d_qvals, d_pvals;
  kml::DevBuf<uint8_t> d_qdesc;
  kml::DevBuf<double> d_qbear, d_qpts;
  // BoW outputs
  kml::DevBuf<uint32_t> d_bow_entry;
  kml::DevBuf<double> d_bow_score, d_nss;
  kml::DevBuf<int32_t> d_bow_count;
  kml::DevBuf<unsigned long long> d_postings;
  kml::PinBuf<uint32_t> h_bow_entry;
  kml::PinBuf<double> h_bow_score, h_nss;
  kml::PinBuf<int32_t> h_bow_count;
  // verification buffers
  kml::DevBuf<kml::PairDesc> d_pairs;
  kml::DevBuf<kml::HamJob> d_jobs;
  kml::DevBuf<uint32_t> d_keys;
  kml::DevBuf<int32_t> d_nq, d_M, d_N3, d_mono_ok, d_status, d_out_mono, d_out_stereo;
  kml::DevBuf<uint16_t> d_iq, d_im, d_kq, d_km;
  kml::DevBuf<double> d_a, d_b, d_models, d_best_mono, d_best_stereo, d_outR, d_outT;
  kml::DevBuf<uint16_t> d_perm, d_samples;
  kml::DevBuf<int32_t> d_valid, d_counts, d_inl_mono, d_inl_stereo, d_nsol;
  kml::DevBuf<double> d_esol, d_brk;
  kml::DevBuf<uint32_t> d_fb_list;  // [0] = deferred counter, [1] = item counter, [2..] = deferred (slot, chain) list
  kml::DevBuf<uint32_t> d_item_base, d_item_list;  // (draw, root) items of a mono round (kernels.h SacArgs)
  kml::DevBuf<double> d_item_q, d_item_model;
  kml::DevBuf<uint8_t> d_item_status;
  kml::DevBuf<kml::SacState> d_st_mono, d_st_stereo;
  kml::DevBuf<uint32_t> d_mask_mono, d_mask_stereo;
  kml::DevBuf<int32_t> d_active;  // [0..1] counters, [2..] the two active lists of the RANSAC rounds
  kml::PinBuf<uint8_t> h_stage;  // generic pinned staging
  kml::DevBuf<uint8_t> d_scratch, d_scratch2;
  kml::DevBuf<double> d_prior;   // rotation prior of the single-pair recoverPose (row f4)
  // batch records: the all-gather buffer of the sharded query (rank r's block at r * blk; a single
  // GPU has one block), the merged records, the batch counters and their pinned host copies
  kml::DevBuf<uint8_t> d_blocks, d_merged;
  kml::DevBuf<kml::BatchStats> d_stats;
  kml::DevBuf<uint64_t> d_qrobot, d_qpose;
  kml::PinBuf<uint8_t> h_recs;
  kml::PinBuf<kml::BatchStats> h_stats;
  kml::PinBuf<uint8_t> h_in;     // pinned staging of a query batch that arrives in pageable memory
  kml::DevBuf<uint8_t> d_in;     // the packed query batch on the device
  size_t item_cap = 0;           // capacity of the mono rounds' item lists (grows after an overflow)
  bool capturing = false;        // a stream capture of the batch pipeline is in progress (lcd.cu)
  void* graph_cache = nullptr;   // lcd.cu GraphCache: the captured pipeline of small batches
  void* enqueue_graph = nullptr; // lcd.cu EnqueueGraph: the captured local pipeline of throughput batches
  bool graph_replayed = false;   // the batch in flight was enqueued by a graph replay (no stage events)

  kml::Comm* comm = nullptr;
};
