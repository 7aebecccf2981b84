import sys, time
sys.path.insert(0, "kimera-multi_b200"); sys.path.insert(0, "oracle")
import numpy as np, kml, kml_oracle as ko
prm = kml.default_params(); prm.matcher_engine = 1
det = kml.LoopClosureDetector(prm)
p0 = kml.default_params(); p0.matcher_engine = 0
d0 = kml.LoopClosureDetector(p0)
rng = np.random.default_rng(1)
for nq, nt in [(128, 256), (500, 500), (513, 255), (64, 20000)]:
    q = rng.integers(0, 256, (nq, 32), np.uint8); t = rng.integers(0, 256, (nt, 32), np.uint8)
    i1, d1, _ = det.hamming_knn2(q, t)
    i0, dd0 = ko.hamming_knn2(q, t)
    ok = np.array_equal(i0, i1) and np.array_equal(dd0, d1)
    print(nq, nt, "ok" if ok else "MISMATCH", flush=True)
    if not ok:
        bad = np.nonzero((i0 != i1).any(1) | (dd0 != d1).any(1))[0]
        print(" bad rows", bad[:10], "of", len(bad)); print(i0[bad[:4]], dd0[bad[:4]]); print(i1[bad[:4]], d1[bad[:4]])
for nt in (1000, 10000, 100000, 1000000):
    q = rng.integers(0, 256, (500, 32), np.uint8); t = rng.integers(0, 256, (nt, 32), np.uint8)
    _, _, ms1 = det.hamming_knn2(q, t, reps=10)
    _, _, ms0 = d0.hamming_knn2(q, t, reps=10)
    print("sweep nt=%d: tensor %.4f ms  popc %.4f ms  (%.2f T compares/s vs %.2f)" % (nt, ms1, ms0, 500*nt/ms1/1e9, 500*nt/ms0/1e9), flush=True)
