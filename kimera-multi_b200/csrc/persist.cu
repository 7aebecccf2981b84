// persist.cu — SURVEY.md §8 row f4 (part): database shard save / load.
// One file holds what addBowVector / addVLCFrame built for every robot of this handle: the BoW
// vectors in entry order (= DBoW2 EntryId order, so query results and tie-breaks are unchanged
// after a reload) and the live frames with their descriptors, bearings and 3-D points read back
// from the HBM arenas.  Loading replays them through the same add paths into an empty detector.
// Layout (little endian): "KMLSHARD" u32 version=1 u32 n_robots, per robot { u64 robot, u64
// n_entries, i64 off[n+1], u64 pose[n], u32 ids[nnz], f32 vals[nnz] }, u64 n_frames, per frame
// { u64 robot, u64 pose, i32 F, u8 desc[F][32], f64 bearings[F][3], f64 points[F][3] }.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include "handle.h"

using namespace kml;

namespace {
struct File {
  FILE* f;
  explicit File(FILE* p) : f(p) {}
  ~File() { if (f) fclose(f); }
  bool w(const void* p, size_t n) { return n == 0 || fwrite(p, 1, n, f) == n; }
  bool r(void* p, size_t n) { return n == 0 || fread(p, 1, n, f) == n; }
};
const char kMagic[8] = {'K', 'M', 'L', 'S', 'H', 'A', 'R', 'D'};
}  // namespace

extern "C" {

int kml_save_shard(kml_handle* h, const char* path) {
  if (!h || !path) return KML_ERR_ARG;
  try {
    KML_CUDA(cudaSetDevice(h->device));
    File out(fopen(path, "wb"));
    if (!out.f) { h->err = std::string("kml_save_shard: cannot open ") + path; return KML_ERR_IO; }
    const uint32_t version = 1, n_robots = (uint32_t)h->sh->dbs.size();
    bool ok = out.w(kMagic, 8) && out.w(&version, 4) && out.w(&n_robots, 4);
    for (auto& kv : h->sh->dbs) {
      const RobotDb& db = *kv.second;
      const uint64_t robot = db.robot, n = db.n_entries();
      ok = ok && out.w(&robot, 8) && out.w(&n, 8) && out.w(db.off.data(), 8 * (n + 1)) &&
           out.w(db.entry_to_pose.data(), 8 * n) && out.w(db.ids.data(), 4 * db.ids.size()) &&
           out.w(db.vals.data(), 4 * db.vals.size());
    }
    const uint64_t n_frames = h->sh->frames.size();
    ok = ok && out.w(&n_frames, 8);
    KML_CUDA(cudaStreamSynchronize(h->stream));
    std::vector<uint8_t> desc;
    std::vector<double> bear, pts;
    // dense frame order = insertion order, so a reload rebuilds the arenas in the same order
    std::vector<std::pair<int32_t, RobotPoseId>> order;
    for (auto& kv : h->sh->frames) order.emplace_back(kv.second.index, kv.first);
    std::sort(order.begin(), order.end(), [](const std::pair<int32_t, RobotPoseId>& a,
                                             const std::pair<int32_t, RobotPoseId>& b) { return a.first < b.first; });
    for (auto& it : order) {
      const FrameRec& fr = h->sh->frames.at(it.second);
      const uint64_t robot = it.second.first, pose = it.second.second;
      const int32_t F = fr.F;
      desc.resize((size_t)F * 32); bear.resize((size_t)F * 3); pts.resize((size_t)F * 3);
      if (F) {
        KML_CUDA(cudaMemcpy(desc.data(), h->sh->s_desc.p + (size_t)fr.feat_off * 32, desc.size(), cudaMemcpyDeviceToHost));
        KML_CUDA(cudaMemcpy(bear.data(), h->sh->s_bear.p + (size_t)fr.feat_off * 3, 8 * bear.size(), cudaMemcpyDeviceToHost));
        KML_CUDA(cudaMemcpy(pts.data(), h->sh->s_pts.p + (size_t)fr.feat_off * 3, 8 * pts.size(), cudaMemcpyDeviceToHost));
      }
      ok = ok && out.w(&robot, 8) && out.w(&pose, 8) && out.w(&F, 4) && out.w(desc.data(), desc.size()) &&
           out.w(bear.data(), 8 * bear.size()) && out.w(pts.data(), 8 * pts.size());
    }
    if (!ok) { h->err = "kml_save_shard: write failed"; return KML_ERR_IO; }
    return KML_OK;
  } catch (const std::exception& e) {
    h->err = e.what();
    return KML_ERR_CUDA;
  }
}

// The file is not trusted: every count is bounded by what is left of the file before anything is
// allocated, offsets must be non-decreasing with at most 1 024 words per vector (what addBowVector
// accepts), and nothing is thrown across the C boundary.
int kml_load_shard(kml_handle* h, const char* path) {
  if (!h || !path) return KML_ERR_ARG;
  try {
    File in(fopen(path, "rb"));
    if (!in.f) { h->err = std::string("kml_load_shard: cannot open ") + path; return KML_ERR_IO; }
    if (fseek(in.f, 0, SEEK_END) != 0) { h->err = "kml_load_shard: cannot seek"; return KML_ERR_IO; }
    const long fsize = ftell(in.f);
    rewind(in.f);
    auto left = [&]() -> uint64_t { const long at = ftell(in.f); return (fsize >= 0 && at >= 0 && at <= fsize) ? (uint64_t)(fsize - at) : 0; };
    auto corrupt = [&](const char* what) { h->err = std::string("kml_load_shard: ") + what; return KML_ERR_IO; };
    char magic[8];
    uint32_t version = 0, n_robots = 0;
    if (!in.r(magic, 8) || memcmp(magic, kMagic, 8) != 0 || !in.r(&version, 4) || version != 1 || !in.r(&n_robots, 4))
      return corrupt("not a shard file of version 1");
    for (uint32_t r = 0; r < n_robots; ++r) {
      uint64_t robot = 0, n = 0;
      if (!in.r(&robot, 8) || !in.r(&n, 8)) return corrupt("truncated");
      if (n > 0x7FFFFFFFull || (n + 1) * 8 + n * 8 > left()) return corrupt("entry count beyond the file");
      std::vector<int64_t> off(n + 1);
      std::vector<uint64_t> poses(n);
      if (!in.r(off.data(), 8 * (n + 1)) || !in.r(poses.data(), 8 * n) || off[0] != 0) return corrupt("truncated");
      for (uint64_t i = 0; i < n; ++i)
        if (off[i + 1] < off[i] || off[i + 1] - off[i] > 1024) return corrupt("corrupt vector offsets");
      if ((uint64_t)off[n] * 8 > left()) return corrupt("word count beyond the file");
      std::vector<uint32_t> ids((size_t)off[n]);
      std::vector<float> vals((size_t)off[n]);
      if (!in.r(ids.data(), 4 * ids.size()) || !in.r(vals.data(), 4 * vals.size())) return corrupt("truncated");
      const int rc = kml_add_bow_bulk(h, robot, poses.data(), (int)n, off.data(), ids.data(), vals.data());
      if (rc != KML_OK) return rc;
    }
    uint64_t n_frames = 0;
    if (!in.r(&n_frames, 8)) return corrupt("truncated");
    std::vector<uint8_t> desc;
    std::vector<double> bear, pts;
    for (uint64_t i = 0; i < n_frames; ++i) {
      uint64_t robot = 0, pose = 0;
      int32_t F = 0;
      if (!in.r(&robot, 8) || !in.r(&pose, 8) || !in.r(&F, 4) || F < 0 || F > 65535) return corrupt("truncated");
      if ((uint64_t)F * 80 > left()) return corrupt("frame beyond the file");
      desc.resize((size_t)F * 32); bear.resize((size_t)F * 3); pts.resize((size_t)F * 3);
      if (!in.r(desc.data(), desc.size()) || !in.r(bear.data(), 8 * bear.size()) || !in.r(pts.data(), 8 * pts.size()))
        return corrupt("truncated");
      const int rc = kml_add_frame(h, robot, pose, desc.data(), bear.data(), pts.data(), F);
      if (rc != KML_OK) return rc;
    }
    return KML_OK;
  } catch (const kml::CudaError& e) {
    h->err = e.what();
    cudaGetLastError();
    return KML_ERR_CUDA;
  } catch (const std::exception& e) {
    h->err = e.what();
    return KML_ERR_IO;
  }
}

}  // extern "C"
