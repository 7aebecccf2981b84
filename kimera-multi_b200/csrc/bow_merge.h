// bow_merge.h — host side of the BoW scorer (bow.cu): the CSR inverted file built from a robot's
// insertion log, how many entry tiles a launch needs, and the merge of the tiles' ranked lists
// into one list per (query, database).
// Plain C++ (no CUDA types) so that the CPU suite can run it behind the emulated kernel
// (tests/emu/).
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <utility>
#include <vector>

namespace kml {

constexpr int kBowTileCap = 24576;  // entries per CTA: 192 KB of u64 accumulators in shared memory

// entries per tile (a multiple of 256, at most kBowTileCap) and tiles per database
inline void bow_tiling(uint32_t max_entries, int* tile_entries, int* n_tiles) {
  int tile = (int)std::min<uint32_t>(std::max<uint32_t>(max_entries, 1u), (uint32_t)kBowTileCap);
  tile = std::max(256, ((tile + 255) / 256) * 256);
  *tile_entries = tile;
  *n_tiles = (int)((std::max<uint32_t>(max_entries, 1u) + tile - 1) / tile);
}

// One posting of the inverted file: {entry id, float32 bits of the word weight}; same layout as
// the uint2 the kernel loads.
struct BowPosting {
  uint32_t entry, weight_bits;
};

// CSR inverted file of one robot database from its insertion log (entry e holds words
// ids[off[e] .. off[e+1]) with weights vals[...]): counting sort by word id; every row comes
// out ascending in entry id because the entries are visited in order.  Returns W = 1 + the
// largest word id (row_ptr has W + 1 elements).
inline uint32_t build_bow_csr(const std::vector<int64_t>& off, const std::vector<uint32_t>& ids,
                              const std::vector<float>& vals, uint32_t n_entries,
                              std::vector<uint32_t>* row_ptr_out, std::vector<BowPosting>* post_out) {
  uint32_t W = 0;
  for (uint32_t w : ids) W = std::max(W, w + 1);
  std::vector<uint32_t>& row_ptr = *row_ptr_out;
  row_ptr.assign((size_t)W + 1, 0);
  for (uint32_t w : ids) row_ptr[w + 1]++;
  for (uint32_t w = 0; w < W; ++w) row_ptr[w + 1] += row_ptr[w];
  post_out->resize(ids.size());
  std::vector<uint32_t> cur(row_ptr.begin(), row_ptr.end() - (W ? 1 : 0));
  if (W == 0) cur.clear();
  for (uint32_t e = 0; e < n_entries; ++e)
    for (int64_t k = off[e]; k < off[e + 1]; ++k) {
      uint32_t bits;
      memcpy(&bits, &vals[k], 4);
      (*post_out)[cur[ids[k]]++] = BowPosting{e, bits};
    }
  return W;
}

// in: the kernel's output, [B][n_db][n_tiles][K] entries / scores (each tile's list sorted
// best-first, ties in ascending entry id) and [B][n_db][n_tiles] counts.
// out: [B][n_db][K] / [B][n_db], score descending, equal scores in ascending entry id.
inline void merge_bow_tiles(const uint32_t* t_entry, const double* t_score, const int32_t* t_count, int B,
                            int n_db, int n_tiles, int K, uint32_t* out_entry, double* out_score,
                            int32_t* out_count) {
  std::vector<std::pair<double, uint32_t>> tmp;
  for (int b = 0; b < B; ++b)
    for (int d = 0; d < n_db; ++d) {
      const size_t l0 = ((size_t)b * n_db + d) * n_tiles;
      uint32_t* oe = out_entry + ((size_t)b * n_db + d) * K;
      double* os = out_score + ((size_t)b * n_db + d) * K;
      if (n_tiles == 1) {
        const int c = t_count[l0];
        memcpy(oe, t_entry + l0 * K, sizeof(uint32_t) * c);
        memcpy(os, t_score + l0 * K, sizeof(double) * c);
        out_count[(size_t)b * n_db + d] = c;
        continue;
      }
      tmp.clear();
      for (int t = 0; t < n_tiles; ++t)
        for (int i = 0; i < t_count[l0 + t]; ++i)
          tmp.emplace_back(-t_score[(l0 + t) * K + i], t_entry[(l0 + t) * K + i]);
      std::sort(tmp.begin(), tmp.end());  // score desc, entry asc
      const int c = (int)std::min<size_t>(tmp.size(), (size_t)K);
      for (int i = 0; i < c; ++i) { oe[i] = tmp[i].second; os[i] = -tmp[i].first; }
      out_count[(size_t)b * n_db + d] = c;
    }
}

}  // namespace kml
