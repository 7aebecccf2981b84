// api_core.cu — handle lifecycle, Hamming kNN entry points, peak probes.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <random>

#include "handle.h"
#include "sac_host.h"

using namespace kml;

static thread_local std::string g_create_err;

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

extern "C" {

void kml_default_params(kml_params* p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  p->inter_robot_only = 0;
  p->alpha = 0.5;
  p->dist_local = 90;
  p->max_db_results = 50;
  p->min_nss_factor = 0.05;
  p->max_nrFrames_between_queries = 2;
  p->max_nrFrames_between_islands = 3;
  p->min_temporal_matches = 1;
  p->max_intraisland_gap = 3;
  p->min_matches_per_island = 1;
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
  p->ransac_randomize = 0;
  p->ransac_seed = 12345u;
  p->top_k_verify = 16;
  p->matcher_norm = 0;
  p->matcher_engine = 1;  // tensor cores: 2.1x the POPC kernel inside the batch, same keys
  p->mono_algorithm = 0;
  p->ransac_use_1point_3d3d = 0;
}

int kml_params_from_yaml(const char* path, int literal_matcher_enum, kml_params* p, int* n_mapped) {
  if (!path || !p) return KML_ERR_ARG;
  FILE* f = fopen(path, "r");
  if (!f) {
    g_create_err = std::string("kml_params_from_yaml: cannot open ") + path;
    return KML_ERR_ARG;
  }
  int n = 0, rc = KML_OK;
  char line[1024];
  while (fgets(line, sizeof line, f)) {
    char* hash = strchr(line, '#');  // comments
    if (hash) *hash = 0;
    if (line[0] == '%' || !strchr(line, ':')) continue;  // "%YAML:1.0" directive, blank lines
    char key[256];
    double v = 0.0;
    if (sscanf(line, " %255[^: \t] : %lf", key, &v) != 2) continue;
    const std::string k(key);
    const int iv = (int)v;
    ++n;
    if (k == "alpha") p->alpha = v;
    else if (k == "max_db_results") p->max_db_results = iv;
    else if (k == "min_nss_factor") p->min_nss_factor = v;
    else if (k == "min_temporal_matches") p->min_temporal_matches = iv;
    else if (k == "min_matches_per_island") p->min_matches_per_island = iv;
    else if (k == "max_intraisland_gap") p->max_intraisland_gap = iv;
    else if (k == "max_nrFrames_between_islands") p->max_nrFrames_between_islands = iv;
    else if (k == "max_nrFrames_between_queries") p->max_nrFrames_between_queries = iv;
    else if (k == "recent_frames_window") p->dist_local = iv;
    else if (k == "lowe_ratio") p->lowe_ratio = v;
    else if (k == "ransac_threshold_2d2d") p->ransac_threshold_mono = v;
    else if (k == "ransac_threshold_3d3d") p->ransac_threshold = v;
    else if (k == "ransac_max_iterations") { p->max_ransac_iterations_mono = iv; p->max_ransac_iterations = iv; }
    else if (k == "ransac_probability") { p->ransac_probability_mono = v; p->ransac_probability = v; }
    else if (k == "min_nr_3d3d_inliers") p->geometric_verification_min_inlier_count = iv;
    else if (k == "ransac_randomize") p->ransac_randomize = iv;
    else if (k == "ransac_use_1point_3d3d") p->ransac_use_1point_3d3d = iv ? 1 : 0;
    else if (k == "ransac_2d2d_algorithm") p->mono_algorithm = (iv == 0) ? 1 : 0;  // OpenGV: 0 STEWENIUS, 1 NISTER
    else if (k == "matcher_type") {
      const int hamming = literal_matcher_enum ? 4 : 3, l1 = literal_matcher_enum ? 3 : 2;
      if (iv == hamming) p->matcher_norm = 0;
      else if (iv == l1) p->matcher_norm = 1;
      else {
        g_create_err = "kml_params_from_yaml: matcher_type selects a matcher this library does not implement";
        rc = KML_ERR_ARG;
      }
    } else {
      --n;  // a key of the file that does not parameterise this path (ORB extractor, PGO, ...)
    }
  }
  fclose(f);
  if (p->matcher_norm != 0) p->matcher_engine = 0;
  if (n_mapped) *n_mapped = n;
  return rc;
}

int kml_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int kml_create(const kml_params* p, int device, kml_handle** out) {
  if (!out) return KML_ERR_ARG;
  *out = nullptr;
  kml_handle* h = nullptr;
  try {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
      cudaGetLastError();
      g_create_err = "no CUDA device: libkml has no CPU fallback";
      return KML_ERR_CUDA;
    }
    if (device < 0 || device >= n) {
      g_create_err = "bad device index";
      return KML_ERR_ARG;
    }
    cudaDeviceProp prop;
    KML_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
      g_create_err = std::string("device is sm_") + std::to_string(prop.major) +
                     std::to_string(prop.minor) + "; libkml kernels are built for sm_100a only";
      return KML_ERR_CUDA;
    }
    h = new kml_handle();
    h->sh = std::make_shared<kml_shared>();
    if (p) h->prm = *p; else kml_default_params(&h->prm);
    if (h->prm.ransac_randomize != 0) {
      g_create_err = "ransac_randomize must be 0 (pre-drawn sample stream)";
      delete h;
      return KML_ERR_ARG;
    }
    if (h->prm.matcher_norm != 0 && h->prm.matcher_norm != 1) {
      g_create_err = "matcher_norm must be 0 (HAMMING) or 1 (L1)";
      delete h;
      return KML_ERR_ARG;
    }
    if (h->prm.matcher_engine != 0 && h->prm.matcher_engine != 1) {
      g_create_err = "matcher_engine must be 0 (POPC pipe) or 1 (tensor cores)";
      delete h;
      return KML_ERR_ARG;
    }
    if (h->prm.matcher_norm != 0) h->prm.matcher_engine = 0;  // the tensor-core path computes NORM_HAMMING only
    if (h->prm.max_db_results > kBowMaxK || h->prm.max_db_results < 1) {
      g_create_err = "max_db_results must be in [1,128]";
      delete h;
      return KML_ERR_ARG;
    }
    if (h->prm.mono_algorithm != 0 && h->prm.mono_algorithm != 1) {
      g_create_err = "mono_algorithm must be 0 (NISTER) or 1 (STEWENIUS)";
      delete h;
      return KML_ERR_ARG;
    }
    if (h->prm.top_k_verify < 0 || h->prm.top_k_verify > kSelMaxK) {
      g_create_err = "top_k_verify must be in [0,128]";
      delete h;
      return KML_ERR_ARG;
    }
    if (h->prm.max_ransac_iterations_mono < 0 || h->prm.max_ransac_iterations < 0 ||
        h->prm.max_ransac_iterations_mono > 100000 || h->prm.max_ransac_iterations > 100000) {
      g_create_err = "max_ransac_iterations(_mono) must be in [0,100000]";
      delete h;
      return KML_ERR_ARG;
    }
    h->device = device;
    KML_CUDA(cudaSetDevice(device));
    KML_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (auto& e2 : h->ev) KML_CUDA(cudaEventCreate(&e2));
    // pre-drawn sample stream: mt19937(seed)() >> 1  
    const int max_it = std::max(h->prm.max_ransac_iterations_mono, h->prm.max_ransac_iterations);
    fill_raw_stream((uint32_t)h->prm.ransac_seed, max_it, &h->sh->raw_h);
    h->sh->d_raw.scratch(h->sh->raw_h.size());
    KML_CUDA(cudaMemcpy(h->sh->d_raw.p, h->sh->raw_h.data(), h->sh->raw_h.size() * 4, cudaMemcpyHostToDevice));
    *out = h;
    return KML_OK;
  } catch (const std::exception& e) {
    g_create_err = e.what();
    cudaGetLastError();
    delete h;
    return KML_ERR_CUDA;
  }
}

void kml_comm_destroy_internal(kml_handle* h);
void kml_graph_destroy_internal(kml_handle* h);

// A query lane: a second handle that SHARES the parent's databases, frame store, RANSAC
// constants and vocabulary (read-only while queries run) but owns its stream, events and
// batch buffers, so two host threads can keep two batches in flight on one GPU.
int kml_create_lane(kml_handle* parent, kml_handle** out) {
  if (!parent || !out) return KML_ERR_ARG;
  *out = nullptr;
  kml_handle* h = nullptr;
  try {
    KML_CUDA(cudaSetDevice(parent->device));
    h = new kml_handle();
    h->sh = parent->sh;
    h->prm = parent->prm;
    h->device = parent->device;
    KML_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (auto& e2 : h->ev) KML_CUDA(cudaEventCreate(&e2));
    *out = h;
    return KML_OK;
  } catch (const std::exception& e) {
    parent->err = e.what();
    cudaGetLastError();
    delete h;
    return KML_ERR_CUDA;
  }
}

int kml_destroy(kml_handle* h) {
  if (!h) return KML_ERR_ARG;
  cudaSetDevice(h->device);
  kml_comm_destroy_internal(h);
  kml_graph_destroy_internal(h);
  if (h->stream) {
    cudaStreamSynchronize(h->stream);
    cudaStreamDestroy(h->stream);
  }
  for (auto& e : h->ev)
    if (e) cudaEventDestroy(e);
  if (h->ev_t0) { cudaEventDestroy(h->ev_t0); cudaEventDestroy(h->ev_t1); cudaEventDestroy(h->ev_lane); }
  delete h;
  return KML_OK;
}

const char* kml_last_error(const kml_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int kml_get_stats(kml_handle* h, kml_stats* out) {
  if (!h || !out) return KML_ERR_ARG;
  *out = h->stats;
  return KML_OK;
}

// ------------------------------------------------------------- Hamming kNN
static void knn2_device(kml_handle* h, int norm, const uint8_t* d_q, int nq, const uint8_t* d_t, int64_t nt,
                        uint32_t* d_idx, uint16_t* d_dist, int reps, float* ms_avg) {
  // split the train set into ranges (one CTA each); >= 2 CTAs per SM when large
  int64_t range_len = (nt + 2 * kNumSMs - 1) / (2 * kNumSMs);
  range_len = ((range_len + 511) / 512) * 512;
  range_len = std::max<int64_t>(512, std::min<int64_t>(range_len, (int64_t)1 << knn_key_shift(norm)));
  const int nranges = nt > 0 ? (int)((nt + range_len - 1) / range_len) : 1;
  h->d_keys.scratch((size_t)nranges * nq * 2);
  std::vector<HamJob> jobs(nranges);
  for (int r = 0; r < nranges; ++r) {
    jobs[r].q = d_q;
    jobs[r].nq = nq;
    jobs[r].t = d_t + (size_t)r * range_len * 32;
    jobs[r].nt = (int)std::max<int64_t>(0, std::min<int64_t>(range_len, nt - (int64_t)r * range_len));
    jobs[r].keys = h->d_keys.p + (size_t)r * nq * 2;
  }
  h->d_jobs.scratch(nranges);
  KML_CUDA(cudaMemcpyAsync(h->d_jobs.p, jobs.data(), sizeof(HamJob) * nranges,
                           cudaMemcpyHostToDevice, h->stream));
  KML_CUDA(cudaStreamSynchronize(h->stream));
  float total = 0.f;
  for (int rep = 0; rep < reps; ++rep) {
    KML_CUDA(cudaEventRecord(h->ev[0], h->stream));
    if (h->prm.matcher_engine == 1 && norm == 0) launch_hamming_jobs_tc(h->d_jobs.p, nranges, h->stream);
    else launch_hamming_jobs(h->d_jobs.p, nranges, norm, h->stream);
    launch_knn2_reduce(h->d_keys.p, nranges, nq, range_len, norm, d_idx, d_dist, h->stream);
    KML_CUDA(cudaEventRecord(h->ev[1], h->stream));
    KML_CUDA(cudaGetLastError());
    KML_CUDA(cudaEventSynchronize(h->ev[1]));
    float ms = 0;
    KML_CUDA(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
    total += ms;
    h->stats.kernel_launches += 2;
  }
  if (ms_avg) *ms_avg = total / std::max(reps, 1);
}

static int knn2_host(kml_handle* h, int norm, const uint8_t* q, int nq, const uint8_t* t, int64_t nt,
                     int reps, uint32_t* idx, uint16_t* dist, float* ms) {
  if (nq < 0 || nt < 0 || (nq > 0 && !q) || (nt > 0 && !t) || !idx || !dist) {
    h->err = "kml_hamming_knn2: bad argument";
    return KML_ERR_ARG;
  }
  if (nq == 0) return KML_OK;
  h->d_scratch.scratch((size_t)nq * 32 + 16);
  h->d_scratch2.scratch((size_t)std::max<int64_t>(nt, 1) * 32 + 16);
  KML_CUDA(cudaMemcpyAsync(h->d_scratch.p, q, (size_t)nq * 32, cudaMemcpyHostToDevice, h->stream));
  if (nt > 0)
    KML_CUDA(cudaMemcpyAsync(h->d_scratch2.p, t, (size_t)nt * 32, cudaMemcpyHostToDevice, h->stream));
  DevBuf<uint32_t> d_idx;
  DevBuf<uint16_t> d_dist;
  d_idx.scratch((size_t)nq * 2);
  d_dist.scratch((size_t)nq * 2);
  knn2_device(h, norm, h->d_scratch.p, nq, h->d_scratch2.p, nt, d_idx.p, d_dist.p, reps, ms);
  KML_CUDA(cudaMemcpyAsync(idx, d_idx.p, sizeof(uint32_t) * nq * 2, cudaMemcpyDeviceToHost, h->stream));
  KML_CUDA(cudaMemcpyAsync(dist, d_dist.p, sizeof(uint16_t) * nq * 2, cudaMemcpyDeviceToHost, h->stream));
  KML_CUDA(cudaStreamSynchronize(h->stream));
  return KML_OK;
}

int kml_hamming_knn2(kml_handle* h, const uint8_t* q, int nq, const uint8_t* t, int64_t nt,
                     uint32_t* idx, uint16_t* dist, float* ms_kernel) {
  KML_API_BEGIN(h)
  return knn2_host(h, 0, q, nq, t, nt, 1, idx, dist, ms_kernel);
  KML_API_END(h)
}

int kml_hamming_knn2_bench(kml_handle* h, const uint8_t* q, int nq, const uint8_t* t, int64_t nt,
                           int reps, uint32_t* idx, uint16_t* dist, float* ms_avg) {
  KML_API_BEGIN(h)
  return knn2_host(h, 0, q, nq, t, nt, std::max(reps, 1), idx, dist, ms_avg);
  KML_API_END(h)
}

int kml_l1_knn2(kml_handle* h, const uint8_t* q, int nq, const uint8_t* t, int64_t nt, uint32_t* idx,
                uint16_t* dist, float* ms_kernel) {
  KML_API_BEGIN(h)
  return knn2_host(h, 1, q, nq, t, nt, 1, idx, dist, ms_kernel);
  KML_API_END(h)
}

// Device-side stopwatch over several lanes: begin records an event on h's stream; end makes
// h's stream wait for everything enqueued so far on the listed lanes, records the end event,
// synchronises and returns the elapsed device time.
int kml_timer_begin(kml_handle* h) {
  KML_API_BEGIN(h)
  if (!h->ev_t0) {
    KML_CUDA(cudaEventCreate(&h->ev_t0));
    KML_CUDA(cudaEventCreate(&h->ev_t1));
    KML_CUDA(cudaEventCreateWithFlags(&h->ev_lane, cudaEventDisableTiming));
  }
  KML_CUDA(cudaStreamSynchronize(h->stream));
  KML_CUDA(cudaEventRecord(h->ev_t0, h->stream));
  return KML_OK;
  KML_API_END(h)
}
int kml_timer_end(kml_handle* h, kml_handle** lanes, int n_lanes, float* ms) {
  KML_API_BEGIN(h)
  if (!h->ev_t0 || !ms || n_lanes < 0 || (n_lanes > 0 && !lanes)) return KML_ERR_ARG;
  for (int i = 0; i < n_lanes; ++i) {
    if (!lanes[i] || lanes[i] == h) continue;
    KML_CUDA(cudaEventRecord(h->ev_lane, lanes[i]->stream));
    KML_CUDA(cudaStreamWaitEvent(h->stream, h->ev_lane, 0));
  }
  KML_CUDA(cudaEventRecord(h->ev_t1, h->stream));
  KML_CUDA(cudaEventSynchronize(h->ev_t1));
  KML_CUDA(cudaEventElapsedTime(ms, h->ev_t0, h->ev_t1));
  return KML_OK;
  KML_API_END(h)
}

int kml_peak_popc(kml_handle* h, double* out) {
  KML_API_BEGIN(h)
  if (!out) return KML_ERR_ARG;
  *out = measure_popc_peak(h->stream);
  h->stats.kernel_launches += 5;
  return KML_OK;
  KML_API_END(h)
}
int kml_flush_l2(kml_handle* h) {
  KML_API_BEGIN(h)
  static thread_local void* buf = nullptr;
  static thread_local int buf_dev = -1;
  const size_t bytes = 256ull << 20;
  if (!buf || buf_dev != h->device) {
    KML_CUDA(cudaMalloc(&buf, bytes));
    buf_dev = h->device;
  }
  KML_CUDA(cudaMemsetAsync(buf, 0x5a, bytes, h->stream));
  KML_CUDA(cudaStreamSynchronize(h->stream));
  return KML_OK;
  KML_API_END(h)
}
int kml_peak_fp64(kml_handle* h, double* out) {
  KML_API_BEGIN(h)
  if (!out) return KML_ERR_ARG;
  *out = measure_fp64_peak(h->stream);
  h->stats.kernel_launches += 5;
  return KML_OK;
  KML_API_END(h)
}

}  // extern "C"
