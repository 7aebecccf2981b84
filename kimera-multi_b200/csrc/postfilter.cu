// postfilter.cu — SURVEY.md §8 row f3: the steps around the hot path inside the front end.
// Host-only logic on <= max_db_results results per query (no kernels): island grouping and the
// temporal-consistency check of the Kimera-VIO detection flow (variant ii;
// "2 computeIslands() 3 checkTemporalConstraint()", /root/reference/images/kimera-multi.drawio:1565;
// thresholds /root/reference/params/D455/LcdParams.yaml:5,9-11), and the VLCFrameMsg wire layout
// (float32 point clouds for versors / keypoints, drawio:385-414) as an add-frame entry point.
#include <algorithm>
#include <cstring>
#include <vector>

#include "handle.h"

using namespace kml;

extern "C" {

// Groups results whose ids are closer than max_intraisland_gap.  Results are visited in
// ascending id; an island's score is the sum of its results' scores, its best entry the first
// one with the strictly largest score; islands spanning fewer than min_matches_per_island ids
// (end - start + 1) are dropped.  Output order = ascending start id.
int kml_compute_islands(const uint64_t* ids, const double* scores, int n, int max_intraisland_gap,
                        int min_matches_per_island, kml_island* out, int cap, int* count) {
  if (!count || n < 0 || (n > 0 && (!ids || !scores)) || cap < 0 || (cap > 0 && !out)) return KML_ERR_ARG;
  *count = 0;
  if (n == 0) return KML_OK;
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return ids[a] < ids[b]; });
  std::vector<kml_island> isl;
  auto open_island = [&](int i) {
    kml_island k;
    k.start_id = k.end_id = k.best_id = ids[i];
    k.island_score = k.best_score = scores[i];
    return k;
  };
  if (n == 1) {  // a single result is an island on its own, whatever the length threshold
    isl.push_back(open_island(order[0]));
  } else {
    kml_island cur = open_island(order[0]);
    auto close_island = [&]() {
      if ((int64_t)(cur.end_id - cur.start_id) + 1 >= (int64_t)min_matches_per_island) isl.push_back(cur);
    };
    for (int j = 1; j < n; ++j) {
      const int i = order[j];
      if ((int64_t)ids[i] - (int64_t)cur.end_id < (int64_t)max_intraisland_gap) {
        cur.end_id = ids[i];
        cur.island_score = cur.island_score + scores[i];
        if (scores[i] > cur.best_score) { cur.best_score = scores[i]; cur.best_id = ids[i]; }
      } else {
        close_island();
        cur = open_island(i);
      }
    }
    close_island();
  }
  if ((int)isl.size() > cap) return KML_ERR_CAPACITY;
  for (size_t i = 0; i < isl.size(); ++i) out[i] = isl[i];
  *count = (int)isl.size();
  return KML_OK;
}

// Counts consecutive queries whose best islands are consistent (overlapping, or closer than
// max_nrFrames_between_islands) and that arrive within max_nrFrames_between_queries of each
// other; returns 1 when the count exceeds min_temporal_matches.  Always records the island.
int kml_check_temporal_constraint(kml_temporal_state* st, uint64_t query_id, const kml_island* island,
                                  int max_nrFrames_between_queries, int max_nrFrames_between_islands,
                                  int min_temporal_matches) {
  if (!st || !island) return KML_ERR_ARG;
  if (st->temporal_entries == 0 ||
      (int64_t)query_id - (int64_t)st->latest_query_id > (int64_t)max_nrFrames_between_queries) {
    st->temporal_entries = 1;
  } else {
    const int64_t a1 = (int64_t)st->latest_island.start_id, a2 = (int64_t)st->latest_island.end_id;
    const int64_t b1 = (int64_t)island->start_id, b2 = (int64_t)island->end_id;
    const bool overlap = (b1 <= a1 && a1 <= b2) || (a1 <= b1 && b1 <= a2);
    bool near = false;
    if (!overlap) {
      const int64_t d = (a1 > b2) ? a1 - b2 : b1 - a2;
      near = d <= (int64_t)max_nrFrames_between_islands;
    }
    if (overlap || near) st->temporal_entries += 1; else st->temporal_entries = 1;
  }
  st->latest_island = *island;
  st->latest_query_id = query_id;
  return st->temporal_entries > min_temporal_matches ? 1 : 0;
}

// detectLoopWithRobot followed by the two post filters: the candidate is the best entry of the
// best island (largest island score, first on ties) if the temporal check passes.
int kml_detect_loop_islands(kml_handle* h, uint64_t robot, uint64_t q_robot, uint64_t q_pose,
                            const uint32_t* ids, const float* vals, int n, int max_intraisland_gap,
                            int min_matches_per_island, int max_nrFrames_between_islands,
                            int min_temporal_matches, kml_temporal_state* st, uint64_t* match_pose,
                            double* match_score, kml_island* best_island, int* lcd_status) {
  if (!h || !st || !match_pose || !match_score || !lcd_status) return KML_ERR_ARG;
  *lcd_status = KML_LCD_NO_MATCHES;
  const int cap = std::max(1, h->prm.max_db_results);
  std::vector<uint64_t> r(cap), p(cap);
  std::vector<double> s(cap);
  int cnt = 0;
  const int rc = kml_detect_loop_with_robot(h, robot, q_robot, q_pose, ids, vals, n, r.data(), p.data(), s.data(), cap, &cnt);
  if (rc < 0) return rc;
  if (rc == KML_NSS_TOO_LOW) { *lcd_status = KML_LCD_LOW_NSS_FACTOR; return rc; }
  if (rc == KML_NO_MATCH) { *lcd_status = KML_LCD_LOW_SCORE; return rc; }
  if (rc != KML_OK) return rc;  // no database / no previous vector / inter-robot only
  std::vector<kml_island> isl(cnt);
  int ni = 0;
  int rc2 = kml_compute_islands(p.data(), s.data(), cnt, max_intraisland_gap, min_matches_per_island, isl.data(), cnt, &ni);
  if (rc2 != KML_OK) return rc2;
  if (ni == 0) { *lcd_status = KML_LCD_NO_GROUPS; return KML_NO_MATCH; }
  int bi = 0;
  for (int i = 1; i < ni; ++i)
    if (isl[bi].island_score < isl[i].island_score) bi = i;
  if (best_island) *best_island = isl[bi];
  *match_pose = isl[bi].best_id;
  *match_score = isl[bi].best_score;
  if (!kml_check_temporal_constraint(st, q_pose, &isl[bi], h->prm.max_nrFrames_between_queries,
                                     max_nrFrames_between_islands, min_temporal_matches)) {
    *lcd_status = KML_LCD_FAILED_TEMPORAL_CONSTRAINT;
    return KML_NO_MATCH;
  }
  *lcd_status = KML_LCD_LOOP_DETECTED;
  return KML_OK;
}

// VLCFrameMsg carries versors and keypoints as float32 point clouds: widen and store.
int kml_add_frame_msg(kml_handle* h, uint64_t robot, uint64_t pose, const uint8_t* desc,
                      const float* versors_xyz, const float* keypoints_xyz, int F) {
  if (!h || F < 0 || (F > 0 && (!desc || !versors_xyz || !keypoints_xyz))) return KML_ERR_ARG;
  std::vector<double> b((size_t)F * 3), k((size_t)F * 3);
  for (size_t i = 0; i < (size_t)F * 3; ++i) { b[i] = (double)versors_xyz[i]; k[i] = (double)keypoints_xyz[i]; }
  return kml_add_frame(h, robot, pose, desc, b.data(), k.data(), F);
}

}  // extern "C"
