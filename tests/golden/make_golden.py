"""Generates tests/golden/golden_v1.npz — known answers produced by INDEPENDENT implementations
(OpenCV's BFMatcher and 5-point solver through cv2, numpy's MT19937 and LAPACK SVD), not by
anything in this repository.  The reference repo holds no golden vectors for this path
(its arithmetic lives in un-vendored DBoW2 / OpenCV / OpenGV), so these stand in
for them: BFMatcher is the very matcher upstream calls, MT19937 the very generator.

    python tests/golden/make_golden.py        # cv2 4.x, numpy; deterministic
"""
import os

import cv2
import numpy as np
from scipy.spatial.transform import Rotation as Rot

HERE = os.path.dirname(os.path.abspath(__file__))


def knn2(norm, q, t):
    m = cv2.BFMatcher(norm).knnMatch(q, t, 2)
    idx = np.full((len(q), 2), 0xFFFFFFFF, np.uint32)
    dist = np.full((len(q), 2), 0xFFFF, np.uint16)
    for i, row in enumerate(m):
        for k, x in enumerate(row):
            idx[i, k], dist[i, k] = x.trainIdx, int(x.distance)
    return idx, dist


def main():
    rng = np.random.default_rng(20260101)
    out = {}
    # ---- BFMatcher(NORM_HAMMING / NORM_L1).knnMatch(k=2), with duplicates and exact hits
    q = rng.integers(0, 256, (64, 32), np.uint8)
    t = rng.integers(0, 256, (300, 32), np.uint8)
    t[7] = t[3]; t[299] = t[0]; q[0] = t[3]; q[5] = t[150]
    out["knn_q"], out["knn_t"] = q, t
    out["ham_idx"], out["ham_dist"] = knn2(cv2.NORM_HAMMING, q, t)
    out["l1_idx"], out["l1_dist"] = knn2(cv2.NORM_L1, q, t)
    # Lowe ratio 0.8 in double on the float distances, strict <
    keep = [(i, int(out["ham_idx"][i, 0])) for i in range(len(q))
            if float(np.float32(out["ham_dist"][i, 0])) < 0.8 * float(np.float32(out["ham_dist"][i, 1]))]
    out["lowe08_pairs"] = np.array(keep, np.uint32).reshape(-1, 2)
    # ---- std::mt19937 == numpy MT19937 with legacy seeding: raw stream and the >>1 draws
    for seed in (12345, 1, 5489):
        mt = np.random.MT19937()
        mt._legacy_seeding(seed)
        out["mt_raw_%d" % seed] = mt.random_raw(64).astype(np.uint32)
    # persistent partial Fisher-Yates driven by that stream (N = 23, sample sizes 8 and 3)
    raw = out["mt_raw_12345"] >> 1
    for S in (8, 3):
        perm = list(range(23)); rows = []
        for d in range(64 // S):
            for i in range(S):
                j = i + int(raw[d * S + i]) % (23 - i)
                perm[i], perm[j] = perm[j], perm[i]
            rows.append(perm[:S])
        out["fy23_s%d" % S] = np.array(rows, np.uint16)
    # ---- LAPACK SVD / Kabsch for 3-point alignment
    P1 = rng.normal(size=(20, 3, 3)); Rs = []; ts = []; P2 = np.zeros_like(P1)
    for k in range(20):
        R = Rot.from_rotvec(rng.normal(size=3)).as_matrix(); tt = rng.normal(size=3)
        P2[k] = (P1[k] - tt) @ R            # p1 = R p2 + t
        Rs.append(R); ts.append(tt)
    out["arun_p1"], out["arun_p2"] = P1, P2
    out["arun_R"], out["arun_t"] = np.array(Rs), np.array(ts)
    A = rng.normal(size=(20, 3, 3))
    out["svd_A"] = A
    out["svd_S"] = np.array([np.linalg.svd(a, compute_uv=False) for a in A])
    # ---- cv2's 5-point solver: all essential matrices of 5 correspondences (normalised, sign-fixed)
    f1s, f2s, Es, nE = [], [], [], []
    for k in range(40):
        X = np.c_[rng.uniform(-5, 5, 5), rng.uniform(-5, 5, 5), rng.uniform(2, 12, 5)]
        R = Rot.from_rotvec(rng.normal(size=3) * 0.15).as_matrix(); tt = rng.uniform(-1, 1, 3)
        X2 = (X - tt) @ R
        f1 = X / np.linalg.norm(X, axis=1, keepdims=True); f2 = X2 / np.linalg.norm(X2, axis=1, keepdims=True)
        E, _ = cv2.findEssentialMat(f2[:, :2] / f2[:, 2:], f1[:, :2] / f1[:, 2:], np.eye(3), cv2.RANSAC, 0.999, 1e-9)
        E = np.zeros((0, 3, 3)) if E is None else E.reshape(-1, 3, 3)
        pad = np.zeros((10, 3, 3))
        for i, e in enumerate(E):
            e = e / np.linalg.norm(e)
            pad[i] = e if e.flat[np.argmax(np.abs(e))] > 0 else -e
        f1s.append(f1); f2s.append(f2); Es.append(pad); nE.append(len(E))
    out["five_f1"], out["five_f2"] = np.array(f1s), np.array(f2s)
    out["five_E_cv2"], out["five_nE_cv2"] = np.array(Es), np.array(nE, np.int32)
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    print("wrote golden_v1.npz with", sorted(out))


if __name__ == "__main__":
    main()
