"""One C2 step for profilers: build the 6 x N DB, run `warm` warm-up batches and `steps` batches."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "kimera-multi_b200"))
import bench, kml
nkf = int(os.environ.get("KML_NKF", "5000")); warm = int(os.environ.get("KML_WARM", "1")); steps = int(os.environ.get("KML_STEPS", "1"))
bench.N_KEYFRAMES = nkf
bench.N_ROBOTS = int(os.environ.get("KML_NROBOTS", str(bench.N_ROBOTS)))
prm = kml.default_params(); prm.matcher_engine = int(os.environ.get("KML_MATCHER_ENGINE", "1"))
det = kml.LoopClosureDetector(prm)
world, robots = bench.build_world(0, lambda m: None)
bench.fill_detector(det, world, robots, lambda m: None)
batches = bench.make_batches(world, warm + steps, bench.N_ROBOTS)
for i, b in enumerate(batches):
    det.query_batch_upload(*b); det.flush_l2()
    out, cnt = det.query_batch_run()
    st = det.stats()
    print("step", i, "ms_total %.2f bow %.2f match %.2f mono %.2f stereo %.2f pairs %d hyp %d" % (st.ms_total, st.ms_bow, st.ms_match, st.ms_mono, st.ms_stereo, st.pairs_last, st.mono_hypotheses_last))
