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

constexpr int kBowTileCap = 20480;  // entries per CTA: 160 KB of u64 accumulators + 40 KB touched list in shared memory

// entries per tile (a multiple of 256, at most kBowTileCap) and tiles per database; a database
// wider than one tile is split into equal tiles
inline void bow_tiling(uint32_t max_entries, int* tile_entries, int* n_tiles) {
  const uint32_t n = std::max<uint32_t>(max_entries, 1u);
  const int nt = (int)((n + kBowTileCap - 1) / kBowTileCap);
  int tile = (int)((n + nt - 1) / nt);
  tile = std::max(256, ((tile + 255) / 256) * 256);
  *tile_entries = tile;
  *n_tiles = (int)((n + tile - 1) / tile);
}

// One posting of the inverted file: {entry id, float32 bits of the word weight}; same layout as
// the uint2 the kernel loads.
struct BowPosting {
  uint32_t entry, weight_bits;
};
// One row (= word) of the inverted file: its postings are pool[start, start + len), ascending in
// entry id; start is even, so a row is 16-byte aligned and read with 128-bit loads.
struct BowRow {
  uint32_t start, len;
};
// Commands of one incremental update (bow_append_kernel): copies of relocated rows, the new
// postings, and the rows whose (start, len) changed.
struct BowCopyCmd { uint32_t src, dst, n, pad; };
struct BowPostCmd { uint32_t dst, entry, weight_bits, pad; };
struct BowRowCmd { uint32_t row, start, len, pad; };
struct BowUpdate {
  std::vector<BowCopyCmd> copies;
  std::vector<BowPostCmd> posts;
  std::vector<BowRowCmd> rows;
  void clear() { copies.clear(); posts.clear(); rows.clear(); }
  bool empty() const { return posts.empty(); }
};

// Host mirror of one robot's inverted file: where every row lives in the device pool and how much
// room it has.  DBoW2::TemplatedDatabase::add appends {entry, weight} to the row of every word of
// the vector; here a bulk load builds the pool with a counting sort (rows packed,
// even-aligned), and later vectors are appended in place: a row that is full moves to the end of
// the pool with twice the room (one copy of its postings), the space it leaves is garbage until
// the next rebuild.  Cost of an append: O(words of the vector), independent of the database.
class BowInvFile {
 public:
  uint32_t W = 0;                       // rows of the device table
  std::vector<uint32_t> start, len, cap;
  uint64_t pool_top = 0, pool_cap = 0, garbage = 0, live = 0;

  static uint32_t round_rows(uint32_t w) { return (w + 65535u) / 65536u * 65536u; }

  // Full build from the insertion log (entry e holds words ids[off[e] .. off[e+1]) with weights
  // vals[...]); rows come out ascending in entry id because the entries are visited in order.
  // pool_out holds the pool_top postings in use; the device pool is allocated with pool_cap
  // (1.5 x that + 64 K) so that appends have room.
  void build(const std::vector<int64_t>& off, const std::vector<uint32_t>& ids, const std::vector<float>& vals,
             uint32_t n_entries, std::vector<BowRow>* rows_out, std::vector<BowPosting>* pool_out) {
    uint32_t wmax = 0;
    for (uint32_t w : ids) wmax = std::max(wmax, w + 1);
    W = std::max(round_rows(wmax), W);
    start.assign(W, 0); len.assign(W, 0); cap.assign(W, 0);
    for (uint32_t w : ids) len[w]++;
    uint64_t top = 0;
    for (uint32_t w = 0; w < W; ++w) {
      start[w] = (uint32_t)top;
      cap[w] = (len[w] + 1u) & ~1u;
      top += cap[w];
    }
    live = ids.size(); garbage = 0; pool_top = top;
    pool_cap = std::max<uint64_t>(top + top / 2 + 65536, 65536);
    pool_out->assign((size_t)pool_top, BowPosting{0xFFFFFFFFu, 0u});
    std::vector<uint32_t> cur(start);
    for (uint32_t e = 0; e < n_entries; ++e)
      for (int64_t k = off[e]; k < off[e + 1]; ++k) {
        uint32_t bits;
        memcpy(&bits, &vals[k], 4);
        (*pool_out)[cur[ids[k]]++] = BowPosting{e, bits};
      }
    rows_out->resize(W);
    for (uint32_t w = 0; w < W; ++w) (*rows_out)[w] = BowRow{start[w], len[w]};
  }

  // Plans the append of `n_post` postings given as (word, entry, weight bits) triples sorted by
  // (word, entry).  Returns false when the update cannot be done in place — a word beyond the
  // table, the pool full, or more garbage than live postings — and the caller rebuilds instead.
  bool plan_append(const uint32_t* words, const uint32_t* entries, const uint32_t* wbits, size_t n_post, BowUpdate* up) {
    if (W == 0) return false;
    for (size_t i = 0; i < n_post; ++i)
      if (words[i] >= W) return false;
    if (garbage > live + 65536) return false;
    // first pass: room needed at the end of the pool
    uint64_t need = 0;
    for (size_t i = 0; i < n_post;) {
      size_t j = i;
      while (j < n_post && words[j] == words[i]) ++j;
      const uint32_t w = words[i], k = (uint32_t)(j - i);
      if (len[w] + k > cap[w]) need += std::max<uint32_t>(4u, (2u * (len[w] + k) + 1u) & ~1u);
      i = j;
    }
    if (pool_top + need > pool_cap || pool_top + need >= 0xFFFFFFF0ull) return false;
    for (size_t i = 0; i < n_post;) {
      size_t j = i;
      while (j < n_post && words[j] == words[i]) ++j;
      const uint32_t w = words[i], k = (uint32_t)(j - i);
      if (len[w] + k > cap[w]) {  // the row moves to the end of the pool with twice the room
        const uint32_t ncap = std::max<uint32_t>(4u, (2u * (len[w] + k) + 1u) & ~1u);
        if (len[w]) up->copies.push_back(BowCopyCmd{start[w], (uint32_t)pool_top, len[w], 0u});
        garbage += cap[w];
        start[w] = (uint32_t)pool_top;
        cap[w] = ncap;
        pool_top += ncap;
      }
      for (size_t q = i; q < j; ++q) up->posts.push_back(BowPostCmd{start[w] + len[w] + (uint32_t)(q - i), entries[q], wbits[q], 0u});
      len[w] += k;
      live += k;
      up->rows.push_back(BowRowCmd{w, start[w], len[w], 0u});
      i = j;
    }
    return true;
  }
};

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
