// hamming.cu — brute-force ORB-256 Hamming kNN (k=2) + Lowe ratio test.
//
// Replaces cv::BFMatcher(NORM_HAMMING)::knnMatch as called from
// LoopClosureDetector::computeMatchedIndices (SURVEY.md A.4, B.1;
// /root/reference/images/kimera-multi.drawio:2583-2586, 2638; matcher created
// at /root/reference/docker/copy/kimera_multi_lcd.patch:34-35).
//
// Layout: one CTA per job (= one query frame x one train range).  Each thread
// keeps one 256-bit query descriptor in 8 registers; train descriptors are
// staged tile by tile (512 x 32 B = 16 KB) into shared memory by the TMA
// engine (cp.async.bulk + mbarrier, double buffered) and read back as
// warp-uniform LDS.128 broadcasts.  Distance = 8 x (LOP3 xor + POPC); best /
// second best are kept as packed keys (dist << 20 | trainIdx) so that the
// update is three integer min/max and ties resolve to the lowest trainIdx,
// which is BFMatcher's observed order (distance asc, trainIdx asc).
// Bound: POPC pipe (8 POPC32 per compare algorithmically, 5 executed after a carry-save
// reduction on the LOP3 pipe), see DESIGN.md §5.2.
#include "common.cuh"
#include "kernels.h"

namespace kml {

constexpr int kHamThreads = 512;
constexpr int kHamTile = 512;  // train descriptors per shared-memory stage

// L1 = false: NORM_HAMMING (8 XOR + 6 carry-save LOP3 + 5 POPC, key = dist << 20 | idx);
// L1 = true : NORM_L1 over the 32 bytes (8 x VABSDIFF4 sum-of-absolute-differences, dist <= 8160,
//             key = dist << 18 | idx) — what upstream's cv::DescriptorMatcher::create(3) selects
//             (/root/reference/docker/copy/kimera_multi_lcd.patch:34-35).
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
#ifdef KML_HOST_EMULATION
  return a ^ b ^ c;
#else
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
#endif
}
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) {
#ifdef KML_HOST_EMULATION
  return (a & b) | (a & c) | (b & c);
#else
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
#endif
}

template <bool L1>
__device__ __forceinline__ void ham_update(uint32_t q0, uint32_t q1, uint32_t q2, uint32_t q3,
                                           uint32_t q4, uint32_t q5, uint32_t q6, uint32_t q7,
                                           const uint4& a, const uint4& b, uint32_t idx,
                                           uint32_t& best, uint32_t& second) {
  uint32_t d;
  if (L1)
    d = __vsadu4(q0, a.x) + __vsadu4(q1, a.y) + __vsadu4(q2, a.z) + __vsadu4(q3, a.w) +
        __vsadu4(q4, b.x) + __vsadu4(q5, b.y) + __vsadu4(q6, b.z) + __vsadu4(q7, b.w);
  else {
    // The POPC pipe (16 lanes/clk/SM, a quarter of the LOP3 rate) is the bound, so three
    // carry-save adders fold seven of the eight XOR words into two weight-1 and three weight-2
    // words first: 5 POPC + 14 LOP3 per compare instead of 8 POPC + 8 LOP3, same integer.
    const uint32_t x0 = q0 ^ a.x, x1 = q1 ^ a.y, x2 = q2 ^ a.z, x3 = q3 ^ a.w;
    const uint32_t x4 = q4 ^ b.x, x5 = q5 ^ b.y, x6 = q6 ^ b.z, x7 = q7 ^ b.w;
    const uint32_t s1 = xor3(x0, x1, x2), c1 = maj3(x0, x1, x2);
    const uint32_t s2 = xor3(x3, x4, x5), c2 = maj3(x3, x4, x5);
    const uint32_t s3 = xor3(s1, s2, x6), c3 = maj3(s1, s2, x6);
    d = (__popc(s3) + __popc(x7)) + 2u * (__popc(c1) + __popc(c2) + __popc(c3));
  }
  uint32_t key = d * (L1 ? (1u << 18) : (1u << 20)) + idx;  // = (d << shift) | idx, idx < 2^shift: one IMAD off the LOP3 pipe
  uint32_t mx = max(best, key);
  best = min(best, key);
  second = min(second, mx);
}

template <bool L1>
__global__ void __launch_bounds__(kHamThreads, 2) hamming_knn2_kernel(const HamJob* __restrict__ jobs) {
  __shared__ __align__(128) uint4 tile[2][kHamTile * 2];
  __shared__ __align__(8) uint64_t full[2];
  const HamJob job = jobs[blockIdx.x];
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int ntiles = (job.nt + kHamTile - 1) / kHamTile;
  uint32_t it = 0;  // tiles issued so far by this CTA (stage = it&1, parity = (it>>1)&1)
  for (int qb = 0; qb < job.nq; qb += kHamThreads) {
    const int qi = qb + tid;
    const bool active = qi < job.nq;
    uint4 qa = make_uint4(0, 0, 0, 0), qc = make_uint4(0, 0, 0, 0);
    if (active) {
      const uint4* qp = reinterpret_cast<const uint4*>(job.q) + 2 * (size_t)qi;
      qa = __ldg(qp);
      qc = __ldg(qp + 1);
    }
    uint32_t best = 0xFFFFFFFFu, second = 0xFFFFFFFFu;
    if (tid == 0 && ntiles > 0) {
      const int n0 = min(kHamTile, job.nt);
      mbar_expect_tx(&full[it & 1], n0 * 32);
      bulk_g2s(tile[it & 1], job.t, n0 * 32, &full[it & 1]);
    }
    for (int k = 0; k < ntiles; ++k, ++it) {
      if (tid == 0 && k + 1 < ntiles) {
        const int n1 = min(kHamTile, job.nt - (k + 1) * kHamTile);
        const uint32_t nx = it + 1;
        mbar_expect_tx(&full[nx & 1], n1 * 32);
        bulk_g2s(tile[nx & 1], job.t + (size_t)(k + 1) * kHamTile * 32, n1 * 32, &full[nx & 1]);
      }
      mbar_wait(&full[it & 1], (it >> 1) & 1);
      const uint4* tp = tile[it & 1];
      const int n = min(kHamTile, job.nt - k * kHamTile);
      const uint32_t base = (uint32_t)(k * kHamTile);
      int j = 0;
#pragma unroll 1
      for (; j + 4 <= n; j += 4) {
        uint4 a0 = tp[2 * j + 0], b0 = tp[2 * j + 1];
        uint4 a1 = tp[2 * j + 2], b1 = tp[2 * j + 3];
        uint4 a2 = tp[2 * j + 4], b2 = tp[2 * j + 5];
        uint4 a3 = tp[2 * j + 6], b3 = tp[2 * j + 7];
        ham_update<L1>(qa.x, qa.y, qa.z, qa.w, qc.x, qc.y, qc.z, qc.w, a0, b0, base + j + 0, best, second);
        ham_update<L1>(qa.x, qa.y, qa.z, qa.w, qc.x, qc.y, qc.z, qc.w, a1, b1, base + j + 1, best, second);
        ham_update<L1>(qa.x, qa.y, qa.z, qa.w, qc.x, qc.y, qc.z, qc.w, a2, b2, base + j + 2, best, second);
        ham_update<L1>(qa.x, qa.y, qa.z, qa.w, qc.x, qc.y, qc.z, qc.w, a3, b3, base + j + 3, best, second);
      }
      for (; j < n; ++j) {
        uint4 a0 = tp[2 * j + 0], b0 = tp[2 * j + 1];
        ham_update<L1>(qa.x, qa.y, qa.z, qa.w, qc.x, qc.y, qc.z, qc.w, a0, b0, base + j, best, second);
      }
      __syncthreads();  // everyone is done with this stage before it is refilled
    }
    if (active) {
      job.keys[2 * (size_t)qi + 0] = best;
      job.keys[2 * (size_t)qi + 1] = second;
    }
  }
}

// Merge per-range top-2 keys of the sweep (config C3) into BFMatcher output.
// partial: [nranges][nq][2]; range r covers train indices [r*range_len, ...).
__global__ void knn2_reduce_kernel(const uint32_t* __restrict__ partial, int nranges, int nq,
                                   int64_t range_len, int shift, uint32_t* __restrict__ idx,
                                   uint16_t* __restrict__ dist) {
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  uint64_t best = ~0ull, second = ~0ull;
  for (int r = 0; r < nranges; ++r) {
    const uint32_t k0 = partial[((size_t)r * nq + qi) * 2 + 0];
    const uint32_t k1 = partial[((size_t)r * nq + qi) * 2 + 1];
    const uint64_t off = (uint64_t)r * (uint64_t)range_len;
    const uint32_t im = (1u << shift) - 1u;
    uint64_t g0 = (k0 == 0xFFFFFFFFu) ? ~0ull : (((uint64_t)(k0 >> shift) << 40) | (off + (k0 & im)));
    uint64_t g1 = (k1 == 0xFFFFFFFFu) ? ~0ull : (((uint64_t)(k1 >> shift) << 40) | (off + (k1 & im)));
    uint64_t mx = max(best, g0);
    best = min(best, g0);
    second = min(min(second, mx), g1);
  }
  idx[2 * qi + 0] = (best == ~0ull) ? 0xFFFFFFFFu : (uint32_t)(best & 0xFFFFFFFFFFull);
  idx[2 * qi + 1] = (second == ~0ull) ? 0xFFFFFFFFu : (uint32_t)(second & 0xFFFFFFFFFFull);
  dist[2 * qi + 0] = (best == ~0ull) ? 0xFFFF : (uint16_t)(best >> 40);
  dist[2 * qi + 1] = (second == ~0ull) ? 0xFFFF : (uint16_t)(second >> 40);
}

// Lowe ratio test + ordered compaction (computeMatchedIndices):
// keep query i iff it has two neighbours and (double)d0 < lowe * (double)d1.
// One CTA per pair; output ascending in query index.
__global__ void __launch_bounds__(256) lowe_compact_kernel(const uint32_t* __restrict__ keys,
                                                           const int* __restrict__ nq_arr,
                                                           int key_stride, double lowe, int shift,
                                                           uint16_t* __restrict__ iq,
                                                           uint16_t* __restrict__ im,
                                                           int* __restrict__ M) {
  __shared__ int warp_cnt[8];
  __shared__ int base_s;
  const int p = blockIdx.x;
  const int nq = nq_arr[p];
  const uint32_t* kp = keys + (size_t)p * key_stride * 2;
  uint16_t* oq = iq + (size_t)p * key_stride;
  uint16_t* om = im + (size_t)p * key_stride;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (int qb = 0; qb < nq; qb += 256) {
    const int qi = qb + threadIdx.x;
    bool keep = false;
    uint32_t k0 = 0;
    if (qi < nq) {
      k0 = kp[2 * qi];
      const uint32_t k1 = kp[2 * qi + 1];
      if (k1 != 0xFFFFFFFFu) {
        const double d0 = (double)(float)(k0 >> shift), d1 = (double)(float)(k1 >> shift);
        keep = d0 < lowe * d1;
      }
    }
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, keep);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int off = base_s;
    for (int w = 0; w < warp; ++w) off += warp_cnt[w];
    if (keep) {
      const int pos = off + __popc(bal & ((1u << lane) - 1u));
      oq[pos] = (uint16_t)qi;
      om[pos] = (uint16_t)(k0 & ((1u << shift) - 1u));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 8; ++w) tot += warp_cnt[w];
      base_s += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) M[p] = base_s;
}

void launch_hamming_jobs(const HamJob* d_jobs, int njobs, int norm, cudaStream_t s) {
  if (njobs <= 0) return;
  if (norm == 1) KML_LAUNCH((hamming_knn2_kernel<true>), njobs, kHamThreads, 0, s, d_jobs);
  else KML_LAUNCH((hamming_knn2_kernel<false>), njobs, kHamThreads, 0, s, d_jobs);
}
void launch_knn2_reduce(const uint32_t* partial, int nranges, int nq, int64_t range_len, int norm,
                        uint32_t* idx, uint16_t* dist, cudaStream_t s) {
  if (nq <= 0) return;
  KML_LAUNCH((knn2_reduce_kernel), (nq + 127) / 128, 128, 0, s, partial, nranges, nq, range_len, knn_key_shift(norm), idx, dist);
}
void launch_lowe_compact(const uint32_t* keys, const int* nq_arr, int key_stride, double lowe, int norm,
                         uint16_t* iq, uint16_t* im, int* M, int P, cudaStream_t s) {
  if (P <= 0) return;
  KML_LAUNCH((lowe_compact_kernel), P, 256, 0, s, keys, nq_arr, key_stride, lowe, knn_key_shift(norm), iq, im, M);
}

// ------------------------------------------------------------ peak probes
#ifdef KML_HOST_EMULATION  // tests/emu/: pipe peaks mean nothing on a host
double measure_popc_peak(cudaStream_t) { return 0.0; }
double measure_fp64_peak(cudaStream_t) { return 0.0; }
#else
__global__ void popc_peak_kernel(uint32_t* out, int iters) {
  uint32_t a0 = threadIdx.x, a1 = a0 * 3 + 1, a2 = a0 * 5 + 2, a3 = a0 * 7 + 3;
  uint32_t a4 = a0 * 11 + 4, a5 = a0 * 13 + 5, a6 = a0 * 17 + 6, a7 = a0 * 19 + 7;
  const uint32_t m = blockIdx.x * 2654435761u;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      a0 = __popc(a0 ^ m); a1 = __popc(a1 ^ m); a2 = __popc(a2 ^ m); a3 = __popc(a3 ^ m);
      a4 = __popc(a4 ^ m); a5 = __popc(a5 ^ m); a6 = __popc(a6 ^ m); a7 = __popc(a7 ^ m);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void fp64_peak_kernel(double* out, int iters, double x, double y) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
         a6 = a0 + 6, a7 = a0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
      a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

double measure_popc_peak(cudaStream_t s) {
  const int blocks = kNumSMs * 8, threads = 256, iters = 2048;
  uint32_t* d = nullptr;
  KML_CUDA(cudaMalloc(&d, sizeof(uint32_t) * blocks * threads));
  cudaEvent_t e0, e1;
  KML_CUDA(cudaEventCreate(&e0));
  KML_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    KML_CUDA(cudaEventRecord(e0, s));
    popc_peak_kernel<<<blocks, threads, 0, s>>>(d, iters);
    KML_CUDA(cudaEventRecord(e1, s));
    KML_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    KML_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    double ops = (double)blocks * threads * iters * 32.0;
    if (rep > 0) best = fmax(best, ops / (ms * 1e-3));
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return best;
}
double measure_fp64_peak(cudaStream_t s) {
  const int blocks = kNumSMs * 8, threads = 256, iters = 2048;
  double* d = nullptr;
  KML_CUDA(cudaMalloc(&d, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  KML_CUDA(cudaEventCreate(&e0));
  KML_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    KML_CUDA(cudaEventRecord(e0, s));
    fp64_peak_kernel<<<blocks, threads, 0, s>>>(d, iters, 0.999999, 1e-9);
    KML_CUDA(cudaEventRecord(e1, s));
    KML_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    KML_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    double flops = (double)blocks * threads * iters * 32.0 * 2.0;
    if (rep > 0) best = fmax(best, flops / (ms * 1e-3));
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return best;
}

#endif  // KML_HOST_EMULATION

}  // namespace kml
