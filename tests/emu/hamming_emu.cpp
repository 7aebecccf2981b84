// hamming_emu.cpp — hamming_knn2_kernel / knn2_reduce_kernel / lowe_compact_kernel
// (kimera-multi_b200/csrc/hamming.cu, the source nvcc compiles) run on the host under
// tests/emu/cuda_emu.h.  Test infrastructure (tests/test_emulated_kernels.py).
#include "cuda_emu.h"

#include "../../kimera-multi_b200/csrc/hamming.cu"

extern "C" {

// BFMatcher(norm).knnMatch(q, t, 2): the train set is cut into ranges of `range_len` descriptors
// (one CTA each, as the C3 sweep does) and the per-range top-2 keys are merged by
// knn2_reduce_kernel.  idx / dist: [nq][2], 0xFFFFFFFF / 0xFFFF = no such neighbour.
void hamemu_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int norm, int range_len, uint32_t* idx,
                 uint16_t* dist) {
  const int nranges = std::max(1, (nt + range_len - 1) / range_len);
  std::vector<uint32_t> partial((size_t)nranges * nq * 2 + 2, 0x12345678u);
  std::vector<kml::HamJob> jobs(nranges);
  for (int r = 0; r < nranges; ++r) {
    jobs[r].q = q; jobs[r].nq = nq;
    jobs[r].t = t + (size_t)r * range_len * 32;
    jobs[r].nt = std::max(0, std::min(range_len, nt - r * range_len));
    jobs[r].keys = partial.data() + (size_t)r * nq * 2;
  }
  kml_emu::Idx3 grid;
  grid.x = (unsigned)nranges; grid.y = 1;
  const kml::HamJob* jp = jobs.data();
  if (norm == 1) kml_emu::launch(grid, kml::kHamThreads, 0, [&] { kml::hamming_knn2_kernel<true>(jp); });
  else kml_emu::launch(grid, kml::kHamThreads, 0, [&] { kml::hamming_knn2_kernel<false>(jp); });
  if (nq <= 0) return;
  grid.x = (unsigned)((nq + 127) / 128);
  kml_emu::launch(grid, 128, 0, [&] {
    kml::knn2_reduce_kernel(partial.data(), nranges, nq, (int64_t)range_len, kml::knn_key_shift(norm), idx, dist);
  });
}

// computeMatchedIndices for P pairs that share one query side size: keys from one CTA per pair,
// then the Lowe test + ordered compaction.  q: [P][nq][32], t: [P][nt][32]; iq / im: [P][nq]; M: [P].
void hamemu_match_lowe(const uint8_t* q, int nq, const uint8_t* t, int nt, int P, int norm, double lowe,
                       uint16_t* iq, uint16_t* im, int* M) {
  const int stride = std::max(nq, 1);
  std::vector<uint32_t> keys((size_t)P * stride * 2, 0x12345678u);
  std::vector<kml::HamJob> jobs(P);
  std::vector<int> nqs(P, nq);
  for (int p = 0; p < P; ++p) {
    jobs[p].q = q + (size_t)p * nq * 32; jobs[p].nq = nq;
    jobs[p].t = t + (size_t)p * nt * 32; jobs[p].nt = nt;
    jobs[p].keys = keys.data() + (size_t)p * stride * 2;
  }
  kml_emu::Idx3 grid;
  grid.x = (unsigned)P; grid.y = 1;
  const kml::HamJob* jp = jobs.data();
  if (norm == 1) kml_emu::launch(grid, kml::kHamThreads, 0, [&] { kml::hamming_knn2_kernel<true>(jp); });
  else kml_emu::launch(grid, kml::kHamThreads, 0, [&] { kml::hamming_knn2_kernel<false>(jp); });
  kml_emu::launch(grid, 256, 0, [&] {
    kml::lowe_compact_kernel(keys.data(), nqs.data(), stride, lowe, kml::knn_key_shift(norm), iq, im, M);
  });
}

}  // extern "C"
