// vocab.cu — DBoW2::TemplatedVocabulary::transform on sm_100a:
// ORB-256 descriptor -> word by descending the k-ary vocabulary tree (child with
// the strictly smallest Hamming distance, first child wins ties), TF-IDF
// accumulation per word and L1 normalisation.  This is the step immediately
// before the hot path (Kimera-VIO computes it and ships `bow_query`,
// /root/reference/images/kimera-multi.drawio:807-825; the vocabulary file is the
// one loaded at /root/reference/docker/copy/kimera_multi_lcd.patch:38 and named by
// /root/reference/launch/kimera_vio_jackal.launch:40-41).
//
// vocab_descend_kernel   thread = descriptor: L levels x k children, each child
//                        32 B read as two 128-bit loads, 8 x (LOP3 + POPC).
// bow_assemble_kernel    CTA = frame: bitonic sort of the word ids in shared
//                        memory, run-length accumulation of the IDF weights
//                        (sequential adds, as BowVector::addWeight), L1 norm
//                        summed in ascending word order by one thread
//                        (BowVector::normalize), IEEE division.
#include <algorithm>
#include <cstring>

#include "handle.h"

using namespace kml;

namespace kml {

__global__ void __launch_bounds__(256) vocab_descend_kernel(const uint8_t* __restrict__ desc, int64_t n,
                                                            const uint8_t* __restrict__ nodes, int k, int L,
                                                            uint32_t* __restrict__ words) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4* dp = reinterpret_cast<const uint4*>(desc) + 2 * i;
  const uint4 qa = __ldg(dp), qb = __ldg(dp + 1);
  uint64_t idx = 0, level_off = 0, level_n = 1;
  for (int l = 0; l < L; ++l) {
    level_n *= (uint64_t)k;
    const uint4* cp = reinterpret_cast<const uint4*>(nodes) + 2 * (level_off + idx * (uint64_t)k);
    int best = 0x7FFFFFFF, bc = 0;
    for (int c = 0; c < k; ++c) {
      const uint4 a = __ldg(cp + 2 * c), b = __ldg(cp + 2 * c + 1);
      const int d = __popc(qa.x ^ a.x) + __popc(qa.y ^ a.y) + __popc(qa.z ^ a.z) + __popc(qa.w ^ a.w) +
                    __popc(qb.x ^ b.x) + __popc(qb.y ^ b.y) + __popc(qb.z ^ b.z) + __popc(qb.w ^ b.w);
      if (d < best) { best = d; bc = c; }
    }
    idx = idx * (uint64_t)k + (uint64_t)bc;
    level_off += level_n;
  }
  words[i] = (uint32_t)idx;
}

constexpr int kAsmThreads = 512;
constexpr int kAsmMaxF = 1024;
__global__ void __launch_bounds__(kAsmThreads) bow_assemble_kernel(const uint32_t* __restrict__ words, int F,
                                                                   const double* __restrict__ weights,
                                                                   uint32_t* __restrict__ out_ids,
                                                                   double* __restrict__ out_vals,
                                                                   int32_t* __restrict__ out_cnt) {
  __shared__ uint32_t s_w[kAsmMaxF];
  __shared__ uint32_t s_ids[kAsmMaxF];
  __shared__ double s_v[kAsmMaxF];
  __shared__ int s_n;
  __shared__ double s_norm;
  const int f = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < kAsmMaxF; i += kAsmThreads) {
    uint32_t w = 0xFFFFFFFFu;
    if (i < F) {
      w = words[(size_t)f * F + i];
      if (!(weights[w] > 0.0)) w = 0xFFFFFFFFu;  // if (w > 0) v.addWeight(id, w)
    }
    s_w[i] = w;
  }
  __syncthreads();
  for (int size = 2; size <= kAsmMaxF; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < kAsmMaxF / 2; i += kAsmThreads) {
        const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
        const bool up = (lo & size) == 0;
        const uint32_t a = s_w[lo], b = s_w[hi];
        if ((a > b) == up) { s_w[lo] = b; s_w[hi] = a; }
      }
      __syncthreads();
    }
  if (tid == 0) s_n = 0;
  __syncthreads();
  // run heads: element i starts a run if it differs from its predecessor
  for (int i = tid; i < kAsmMaxF; i += kAsmThreads) {
    const uint32_t w = s_w[i];
    if (w != 0xFFFFFFFFu && (i == 0 || s_w[i - 1] != w)) {
      int c = 1;
      while (i + c < kAsmMaxF && s_w[i + c] == w) ++c;
      const double wt = weights[w];
      double v = wt;                                  // first addWeight inserts, the rest add
      for (int r = 1; r < c; ++r) v = v + wt;
      // rank of this run among the runs = number of distinct smaller ids: count heads before i
      int rank = 0;
      for (int j = 1; j <= i; ++j) rank += (s_w[j] != s_w[j - 1]);
      s_ids[rank] = w;
      s_v[rank] = v;
      atomicAdd(&s_n, 1);
    }
  }
  __syncthreads();
  const int n = s_n;
  if (tid == 0) {
    double norm = 0.0;
    for (int i = 0; i < n; ++i) norm = norm + fabs(s_v[i]);
    s_norm = norm;
  }
  __syncthreads();
  const double norm = s_norm;
  for (int i = tid; i < n; i += kAsmThreads) {
    out_ids[(size_t)f * F + i] = s_ids[i];
    out_vals[(size_t)f * F + i] = (norm > 0.0) ? s_v[i] / norm : s_v[i];
  }
  if (tid == 0) out_cnt[f] = n;
}

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

extern "C" {

int kml_vocab_set(kml_handle* h, int k, int L, const uint8_t* node_desc, const double* word_weights) {
  KML_API_BEGIN(h)
  if (k < 2 || L < 1 || !node_desc || !word_weights) { h->err = "vocab_set: bad argument"; return KML_ERR_ARG; }
  uint64_t nodes = 0, lvl = 1;
  for (int l = 0; l < L; ++l) {
    lvl *= (uint64_t)k;
    nodes += lvl;
    if (lvl > (1ull << 31)) { h->err = "vocab_set: more than 2^31 words"; return KML_ERR_CAPACITY; }
  }
  h->sh->voc_k = k; h->sh->voc_L = L; h->sh->voc_words = lvl; h->sh->voc_nodes = nodes;
  h->sh->d_voc_nodes.scratch(nodes * 32 + 32);
  h->sh->d_voc_w.scratch(lvl);
  KML_CUDA(cudaMemcpyAsync(h->sh->d_voc_nodes.p, node_desc, nodes * 32, cudaMemcpyHostToDevice, h->stream));
  KML_CUDA(cudaMemcpyAsync(h->sh->d_voc_w.p, word_weights, lvl * 8, cudaMemcpyHostToDevice, h->stream));
  KML_CUDA(cudaStreamSynchronize(h->stream));
  return KML_OK;
  KML_API_END(h)
}

int kml_transform_batch(kml_handle* h, int B, int F, const uint8_t* desc, int64_t* out_off, uint32_t* out_ids,
                        double* out_vals, int64_t cap, float* ms_kernel) {
  KML_API_BEGIN(h)
  if (h->sh->voc_nodes == 0) { h->err = "transform: no vocabulary (kml_vocab_set)"; return KML_ERR_ARG; }
  if (B < 0 || F < 0 || F > kAsmMaxF || !out_off || (B > 0 && F > 0 && (!desc || !out_ids || !out_vals))) {
    h->err = "transform: bad argument (F <= 1024)";
    return KML_ERR_ARG;
  }
  out_off[0] = 0;
  if (B == 0 || F == 0) {
    for (int b = 0; b < B; ++b) out_off[b + 1] = 0;
    return KML_OK;
  }
  const int64_t n = (int64_t)B * F;
  cudaStream_t s = h->stream;
  h->d_scratch.scratch((size_t)n * 32 + 32);
  DevBuf<uint32_t> d_words, d_ids;
  DevBuf<double> d_vals;
  DevBuf<int32_t> d_cnt;
  d_words.scratch(n); d_ids.scratch(n); d_vals.scratch(n); d_cnt.scratch(B);
  KML_CUDA(cudaMemcpyAsync(h->d_scratch.p, desc, (size_t)n * 32, cudaMemcpyHostToDevice, s));
  KML_CUDA(cudaEventRecord(h->ev[0], s));
  KML_LAUNCH((vocab_descend_kernel), (unsigned)((n + 255) / 256), 256, 0, s, h->d_scratch.p, n, h->sh->d_voc_nodes.p, h->sh->voc_k, h->sh->voc_L, d_words.p);
  KML_LAUNCH((bow_assemble_kernel), B, kAsmThreads, 0, s, d_words.p, F, h->sh->d_voc_w.p, d_ids.p, d_vals.p, d_cnt.p);
  KML_CUDA(cudaEventRecord(h->ev[1], s));
  KML_CUDA(cudaGetLastError());
  h->stats.kernel_launches += 2;
  std::vector<uint32_t> ids(n);
  std::vector<double> vals(n);
  std::vector<int32_t> cnt(B);
  KML_CUDA(cudaMemcpyAsync(ids.data(), d_ids.p, 4 * n, cudaMemcpyDeviceToHost, s));
  KML_CUDA(cudaMemcpyAsync(vals.data(), d_vals.p, 8 * n, cudaMemcpyDeviceToHost, s));
  KML_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt.p, 4 * B, cudaMemcpyDeviceToHost, s));
  KML_CUDA(cudaStreamSynchronize(s));
  if (ms_kernel) KML_CUDA(cudaEventElapsedTime(ms_kernel, h->ev[0], h->ev[1]));
  int64_t o = 0;
  for (int b = 0; b < B; ++b) {
    if (o + cnt[b] > cap) { h->err = "transform: output capacity"; return KML_ERR_CAPACITY; }
    memcpy(out_ids + o, ids.data() + (size_t)b * F, 4 * (size_t)cnt[b]);
    memcpy(out_vals + o, vals.data() + (size_t)b * F, 8 * (size_t)cnt[b]);
    o += cnt[b];
    out_off[b + 1] = o;
  }
  return KML_OK;
  KML_API_END(h)
}

}  // extern "C"
