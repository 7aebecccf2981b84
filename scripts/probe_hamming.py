"""GPU probe: Hamming kNN parity vs the oracle + POPC / FP64 pipe peaks."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kimera-multi_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import kml, kml_oracle as ko
det = kml.LoopClosureDetector()
out = {}
out["popc32_per_s"] = det.peak_popc(); out["fp64_flop_per_s"] = det.peak_fp64()
print("popc peak %.3e /s -> %.3f T compares/s ; fp64 %.2f TFLOP/s" % (out["popc32_per_s"], out["popc32_per_s"]/8e12, out["fp64_flop_per_s"]/1e12))
rng = np.random.default_rng(1)
for nq, nt in [(500, 500), (500, 1), (500, 0), (7, 1300), (500, 10000), (1024, 5000)]:
    q = rng.integers(0, 256, (nq, 32), np.uint8); t = rng.integers(0, 256, (nt, 32), np.uint8)
    if nt > 10: t[5] = t[3]; q[0] = t[3]
    i1, d1, ms = det.hamming_knn2(q, t)
    i0, d0 = ko.hamming_knn2(q, t)
    ok = np.array_equal(i0, i1) and np.array_equal(d0, d1)
    print(nq, nt, "parity", ok, "ms", ms)
    assert ok
res = []
for nt in [1000, 10000, 100000, 1000000]:
    q = rng.integers(0, 256, (500, 32), np.uint8); t = rng.integers(0, 256, (nt, 32), np.uint8)
    i1, d1, ms = det.hamming_knn2(q, t, reps=10)
    cps = 500 * nt / (ms * 1e-3)
    res.append(dict(nt=nt, ms=ms, compares_per_s=cps, frac_popc_algorithmic=cps * 8 / out["popc32_per_s"],
                    frac_popc_executed=cps * 5 / out["popc32_per_s"]))
    print(res[-1])
out["sweep"] = res
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_hamming.json"), "w"), indent=1)
