"""Measurements of the BASELINE.json configs that are not the bench line: C1 (single-robot query latency),
C3 (Hamming sweep vs POPC peak, with cv2 and the oracle timed beside it) and C4 (batched stereo Arun RANSAC,
4 096 problems x 1 001 hypotheses x 500 correspondences).  Writes gpurun_out/configs_r01.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kimera-multi_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import cv2
import kml, kml_oracle as ko
from kml import synth
out = {}
det = kml.LoopClosureDetector()
popc = det.peak_popc(); fp64 = det.peak_fp64()
out["peaks"] = {"popc32_per_s": popc, "fp64_flop_per_s": fp64}
rng = np.random.default_rng(5)
# ---------------------------------------------------------------- C3
sweep = []
for nt in (1000, 10000, 100000, 1000000):
    q = rng.integers(0, 256, (500, 32), np.uint8); t = rng.integers(0, 256, (nt, 32), np.uint8)
    _, _, ms = det.hamming_knn2(q, t, reps=10)
    row = {"nt": nt, "gpu_ms": ms, "gpu_compares_per_s": 500 * nt / (ms * 1e-3), "frac_of_popc_peak": 500 * nt * 8 / (ms * 1e-3) / popc}
    if nt < (1 << 18):   # cv2's BFMatcher asserts trainDescCollection rows < IMGIDX_ONE (2^18)
        t0 = time.perf_counter(); cv2.setNumThreads(1); cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, 2); row["cv2_1thread_ms"] = (time.perf_counter() - t0) * 1e3
        cv2.setNumThreads(0); t0 = time.perf_counter(); cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, 2); row["cv2_allthreads_ms"] = (time.perf_counter() - t0) * 1e3
    if nt <= 100000:
        t0 = time.perf_counter(); ko.hamming_knn2(q, t); row["oracle_ms"] = (time.perf_counter() - t0) * 1e3
    sweep.append(row); print(row, flush=True)
out["C3_hamming_sweep"] = sweep
# ---------------------------------------------------------------- C4
P, N = 4096, 500
X = np.stack([rng.uniform(-5, 5, (P, N)), rng.uniform(-5, 5, (P, N)), rng.uniform(2, 12, (P, N))], axis=2)
from scipy.spatial.transform import Rotation as Rot
R = Rot.from_rotvec(rng.normal(size=(P, 3)) * 0.2).as_matrix(); tt = rng.uniform(-1, 1, (P, 3))
X2 = np.einsum("pnj,pjk->pnk", X - tt[:, None, :], R) + rng.normal(size=X.shape) * 0.03
outl = rng.random((P, N)) < 0.35
X2[outl] = rng.uniform(-8, 8, (int(outl.sum()), 3))
det.ransac_arun_batch(X[:64], X2[:64], full_hypotheses=True)
g = det.ransac_arun_batch(X, X2, full_hypotheses=True)
hyp = float(P) * 1001; res = hyp * N
c4 = {"problems": P, "hypotheses_each": 1001, "correspondences": N, "gpu_ms": g["ms"], "hypotheses_per_s": hyp / (g["ms"] * 1e-3),
      "residuals_per_s": res / (g["ms"] * 1e-3), "algorithmic_tflops": (hyp * 1.5e3 + res * 27) / (g["ms"] * 1e-3) / 1e12,
      "frac_of_fp64_peak": (hyp * 1.5e3 + res * 27) / (g["ms"] * 1e-3) / fp64}
t0 = time.perf_counter()
for p in range(4):
    ko.ransac_arun(X[p], X2[p], 0.5, 1.0 - 1e-300, 1000, 12345)   # p ~ 1 keeps k large: all 1 001 trials
c4["oracle_ms_per_problem_1thread"] = (time.perf_counter() - t0) / 4 * 1e3
ga = det.ransac_arun_batch(X, X2, full_hypotheses=False)
c4["adaptive_gpu_ms"] = ga["ms"]; c4["adaptive_mean_iterations"] = float(ga["iterations"].mean())
out["C4_stereo_ransac"] = c4; print(c4, flush=True)
# ---------------------------------------------------------------- C1
world = synth.World(500, F=500)
ref = ko.LoopClosureDetector(); d1 = kml.LoopClosureDetector()
for ch in synth.build_database(world, [0], 2000, chunk=1000):
    d1.addBowVectors(ch["robot"], ch["poses"], ch["bow_off"], ch["bow_ids"], ch["bow_vals"])
    d1.addVLCFrames(ch["robot"], ch["poses"], ch["desc"], ch["bearings"], ch["points"])
    for i, p in enumerate(ch["poses"]):
        o0, o1 = ch["bow_off"][i], ch["bow_off"][i + 1]
        ref.addBowVector(0, int(p), ch["bow_ids"][o0:o1], ch["bow_vals"][o0:o1]); ref.addVLCFrame(0, int(p), ch["desc"][i], ch["bearings"][i], ch["points"][i])
q = synth.make_queries(world, 8, 2000, 1); fq, fp = q["frames"], q["prev"]
gpu_t, cpu_t = [], []
for b in range(8):
    o0, o1 = fq["bow_off"][b], fq["bow_off"][b + 1]; p0, p1 = fp["bow_off"][b], fp["bow_off"][b + 1]
    a = (q["q_robot"][b:b+1] + 7, q["q_pose"][b:b+1], np.array([0, o1 - o0]), fq["bow_ids"][o0:o1], fq["bow_vals"][o0:o1], np.array([0, p1 - p0]),
         fp["bow_ids"][p0:p1], fp["bow_vals"][p0:p1], fq["desc"][b:b+1], fq["bearings"][b:b+1], fq["points"][b:b+1])
    t0 = time.perf_counter(); r1, c1 = d1.query_batch(*a); gpu_t.append(time.perf_counter() - t0)
    t0 = time.perf_counter(); r0, c0 = ref.query_batch(*a, threads=1); cpu_t.append(time.perf_counter() - t0)
    assert np.array_equal(c0, c1) and np.array_equal(r0["mono_inliers"], r1["mono_inliers"])
out["C1_single_query"] = {"gpu_ms_per_query_median": float(np.median(gpu_t[1:]) * 1e3), "cpu_oracle_1thread_ms_per_query_median": float(np.median(cpu_t[1:]) * 1e3),
                          "candidates_verified": int(c1[0])}
print(out["C1_single_query"])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs_r01.json"), "w"), indent=1)
