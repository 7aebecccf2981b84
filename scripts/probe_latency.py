"""Single-query latency (BASELINE.json configs[0] shape: one query against a 2 000-keyframe database)
through kml_query_batch with B = 1: eager enqueue (KML_NO_GRAPH=1) vs the captured graph, results equal."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kimera-multi_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import kml, kml_oracle as ko
from kml import synth
world = synth.World(500, F=500)
det, ref = kml.LoopClosureDetector(), ko.LoopClosureDetector()
for ch in synth.build_database(world, [0], 2000, chunk=1000):
    det.addBowVectors(ch["robot"], ch["poses"], ch["bow_off"], ch["bow_ids"], ch["bow_vals"])
    det.addVLCFrames(ch["robot"], ch["poses"], ch["desc"], ch["bearings"], ch["points"])
    for i, p in enumerate(ch["poses"]):
        o0, o1 = ch["bow_off"][i], ch["bow_off"][i + 1]
        ref.addBowVector(0, int(p), ch["bow_ids"][o0:o1], ch["bow_vals"][o0:o1]); ref.addVLCFrame(0, int(p), ch["desc"][i], ch["bearings"][i], ch["points"][i])
q = synth.make_queries(world, 24, 2000, 1); fq, fp = q["frames"], q["prev"]
ts, same = [], True
for b in range(24):
    o0, o1 = fq["bow_off"][b], fq["bow_off"][b + 1]; p0, p1 = fp["bow_off"][b], fp["bow_off"][b + 1]
    a = (q["q_robot"][b:b+1] + 7, q["q_pose"][b:b+1], np.array([0, o1 - o0]), fq["bow_ids"][o0:o1], fq["bow_vals"][o0:o1], np.array([0, p1 - p0]),
         fp["bow_ids"][p0:p1], fp["bow_vals"][p0:p1], fq["desc"][b:b+1], fq["bearings"][b:b+1], fq["points"][b:b+1])
    t0 = time.perf_counter(); r1, c1 = det.query_batch(*a); ts.append(time.perf_counter() - t0)
    r0, c0 = ref.query_batch(*a, threads=1)
    same = same and np.array_equal(c0, c1) and all(np.array_equal(r0[k], r1[k]) for k in ("m_pose", "mono_inliers", "stereo_inliers", "status", "n_matches"))
print("graph" if not os.environ.get("KML_NO_GRAPH") else "eager", "per-query ms:", " ".join("%.3f" % (t * 1e3) for t in ts[:6]), "... median of the last 16: %.3f ms" % (np.median(ts[8:]) * 1e3),
      "records match oracle:", same, "device ms_total %.3f" % det.stats().ms_total)
