// bow_emu.cpp — bow_score_kernel (kimera-multi_b200/csrc/bow.cu, the source nvcc compiles) run on
// the host under tests/emu/cuda_emu.h, behind the product's own host code for the CSR build, the
// tiling and the tile merge (csrc/bow_merge.h).  Test infrastructure (tests/test_emulated_kernels.py).
#include "cuda_emu.h"

#include "../../kimera-multi_b200/csrc/bow.cu"
#include "../../kimera-multi_b200/csrc/bow_merge.h"

namespace {
struct EmuDb {
  std::vector<int64_t> off{0};
  std::vector<uint32_t> ids;
  std::vector<float> vals;
  uint32_t n_entries = 0;
  bool dirty = true;
  kml::BowInvFile inv;
  std::vector<kml::BowRow> rows;
  std::vector<kml::BowPosting> pool;
  std::vector<uint32_t> pend_word, pend_entry, pend_wbits;
  int rebuilds = 0, appends = 0;
};
// rebuild_invfile / flush_appends of lcd.cu on host vectors ("device" memory = the vectors)
void sync_db(EmuDb* db) {
  if (db->n_entries == 0) return;
  if (!db->dirty && !db->pend_word.empty()) {
    const size_t n = db->pend_word.size();
    std::vector<uint32_t> order(n);
    for (size_t i = 0; i < n; ++i) order[i] = (uint32_t)i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return db->pend_word[x] < db->pend_word[y]; });
    std::vector<uint32_t> w(n), e(n), b(n);
    for (size_t i = 0; i < n; ++i) { w[i] = db->pend_word[order[i]]; e[i] = db->pend_entry[order[i]]; b[i] = db->pend_wbits[order[i]]; }
    kml::BowUpdate up;
    if (db->inv.plan_append(w.data(), e.data(), b.data(), n, &up)) {
      static_assert(sizeof(kml::BowCopyCmd) == 16 && sizeof(kml::BowPostCmd) == 16 && sizeof(kml::BowRowCmd) == 16, "command layout");
      kml::launch_bow_append(reinterpret_cast<uint2*>(db->rows.data()), reinterpret_cast<uint2*>(db->pool.data()),
                             reinterpret_cast<const uint4*>(up.copies.data()), (int)up.copies.size(),
                             reinterpret_cast<const uint4*>(up.posts.data()), (int)up.posts.size(),
                             reinterpret_cast<const uint4*>(up.rows.data()), (int)up.rows.size(), nullptr);
      db->appends++;
    } else {
      db->dirty = true;
    }
    db->pend_word.clear(); db->pend_entry.clear(); db->pend_wbits.clear();
  }
  if (db->dirty) {
    db->inv.build(db->off, db->ids, db->vals, db->n_entries, &db->rows, &db->pool);
    db->pool.resize((size_t)db->inv.pool_cap, kml::BowPosting{0xDEADBEEFu, 0xDEADBEEFu});  // device pool: unused slots are garbage
    db->pend_word.clear(); db->pend_entry.clear(); db->pend_wbits.clear();
    db->dirty = false;
    db->rebuilds++;
  }
}
}  // namespace

extern "C" {

void* bowemu_db_create() { return new EmuDb(); }
void bowemu_db_destroy(void* p) { delete (EmuDb*)p; }
int bowemu_db_rebuilds(void* p) { return ((EmuDb*)p)->rebuilds; }
int bowemu_db_appends(void* p) { return ((EmuDb*)p)->appends; }
// `count` vectors in CSR form, as kml_add_bow_bulk takes them (add_bow_host of lcd.cu: once the
// inverted file exists the postings are queued for the in-place append)
void bowemu_db_add(void* p, int count, const int64_t* off, const uint32_t* ids, const float* vals) {
  EmuDb* db = (EmuDb*)p;
  for (int i = 0; i < count; ++i) {
    const uint32_t entry = db->n_entries;
    db->ids.insert(db->ids.end(), ids + off[i], ids + off[i + 1]);
    db->vals.insert(db->vals.end(), vals + off[i], vals + off[i + 1]);
    db->off.push_back((int64_t)db->ids.size());
    db->n_entries++;
    if (!db->dirty)
      for (int64_t k = off[i]; k < off[i + 1]; ++k) {
        uint32_t bits;
        memcpy(&bits, &vals[k], 4);
        db->pend_word.push_back(ids[k]);
        db->pend_entry.push_back(entry);
        db->pend_wbits.push_back(bits);
      }
  }
}

// What run_bow (lcd.cu) does around the launch: views, tiling, one CTA per (query, db, tile),
// tile merge.  q_off/q_ids/q_vals: B query vectors; p_*: previous vectors for the NSS factor
// (nullable); max_id: [n_db] or null.  out_entry/out_score [B][n_db][K], out_count [B][n_db],
// nss [B] (nullable).  tile_cap_override > 0 shrinks the tile (to exercise many tiles on small
// databases).  Returns the number of tiles used.
int bowemu_query(void** dbs_, int n_db, int B, const int64_t* q_off, const uint32_t* q_ids, const float* q_vals,
                 const int64_t* p_off, const uint32_t* p_ids, const float* p_vals, int K, const int32_t* max_id,
                 int tile_cap_override, uint32_t* out_entry, double* out_score, int32_t* out_count, double* nss,
                 unsigned long long* postings_touched) {
  std::vector<kml::BowDb> views(n_db);
  uint32_t max_entries = 1;
  for (int i = 0; i < n_db; ++i) {
    EmuDb* db = (EmuDb*)dbs_[i];
    sync_db(db);
    views[i].rows = reinterpret_cast<const uint2*>(db->rows.data());
    views[i].postings = reinterpret_cast<const uint2*>(db->pool.data());
    views[i].W = db->n_entries ? db->inv.W : 0u;
    views[i].n_entries = db->n_entries;
    views[i].entry_pose = nullptr; views[i].entry_frame = nullptr; views[i].robot = (uint64_t)i;
    max_entries = std::max(max_entries, db->n_entries);
  }
  int tile = 0, n_tiles = 0;
  kml::bow_tiling(max_entries, &tile, &n_tiles);
  if (tile_cap_override > 0) {
    tile = std::max(256, (tile_cap_override + 255) / 256 * 256);
    n_tiles = (int)((max_entries + tile - 1) / tile);
  }
  const size_t nlist = (size_t)B * n_db * n_tiles;
  // device output buffers are not cleared by run_bow either: poison them
  std::vector<uint32_t> t_entry(nlist * K, 0xDEADBEEFu);
  std::vector<double> t_score(nlist * K, -1.0);
  std::vector<int32_t> t_count(nlist, -12345);
  unsigned long long touched = 0;
  kml::BowArgs a;
  a.dbs = views.data(); a.n_db = n_db; a.B = B;
  a.q_off = q_off; a.q_ids = q_ids; a.q_vals = q_vals;
  a.p_off = p_off; a.p_ids = p_ids; a.p_vals = p_vals;
  a.K = K; a.max_id = max_id; a.tile_entries = tile; a.n_tiles = n_tiles;
  a.out_entry = t_entry.data(); a.out_score = t_score.data(); a.out_count = t_count.data();
  a.nss = (p_off && nss) ? nss : nullptr;
  a.postings_touched = &touched;
  kml_emu::Idx3 grid;
  grid.x = (unsigned)(B * n_db * n_tiles); grid.y = 1;
  kml::launch_bow(a, nullptr);
  kml::merge_bow_tiles(t_entry.data(), t_score.data(), t_count.data(), B, n_db, n_tiles, K, out_entry, out_score,
                       out_count);
  if (postings_touched) *postings_touched = touched;
  return n_tiles;
}

}  // extern "C"
