"""CPU tests that PIN the oracle (it has no reference golden vectors to lean
on — SURVEY.md §4, §8c): OpenCV's own BFMatcher via cv2, numpy/scipy linear
algebra, the libstdc++ RNG known answers of SURVEY.md Appendix B.2, and the
mathematical invariants of the 5-point problem."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation as Rot


def _scene(n, rng, noise=0.0):
    X = np.c_[rng.uniform(-5, 5, n), rng.uniform(-5, 5, n), rng.uniform(2, 12, n)]
    R = Rot.from_rotvec(rng.normal(size=3) * 0.15).as_matrix()
    t = rng.uniform(-1, 1, 3)
    X2 = (X - t) @ R  # x1 = R x2 + t
    f1 = X / np.linalg.norm(X, axis=1, keepdims=True)
    f2 = X2 / np.linalg.norm(X2, axis=1, keepdims=True)
    return X, X2, f1, f2, R, t


def test_rng_stream_known_answers(oracle):
    # mt19937(12345) >> 1 = 1996335345, 1911592690, 679411342, ...
    draws = [1996335345, 1911592690, 679411342, 280691776, 394962642]
    N = 50000
    s = oracle.sample_stream(N, 1, 12345, 1)
    assert s[0, 0] == draws[0] % N
    # persistent partial Fisher-Yates, replayed in python
    N, S = 17, 3
    perm = list(range(N))
    import random  # noqa: F401
    mt = np.random.MT19937()
    mt._legacy_seeding(12345)
    raw = mt.random_raw(30) >> 1
    exp = []
    for d in range(10):
        for i in range(S):
            j = i + int(raw[d * S + i]) % (N - i)
            perm[i], perm[j] = perm[j], perm[i]
        exp.append(perm[:S])
    got = oracle.sample_stream(N, S, 12345, 10)
    assert np.array_equal(got, np.array(exp, np.uint16))
    assert list(raw[:5]) == draws


def test_hamming_knn_equals_cv2_bfmatcher(oracle):
    import cv2
    rng = np.random.default_rng(0)
    # SURVEY.md B.1 repro + adversarial ties
    for nq, nt in [(5, 7), (500, 500), (40, 3000), (3, 2)]:
        q = rng.integers(0, 256, (nq, 32), np.uint8)
        t = rng.integers(0, 256, (nt, 32), np.uint8)
        if nt >= 6:
            t[4] = t[1]
            t[5] = t[1]
            q[0] = t[1]
        idx, dist = oracle.hamming_knn2(q, t)
        m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, 2)
        ci = np.array([[x.trainIdx for x in row] for row in m], np.uint32)
        cd = np.array([[int(x.distance) for x in row] for row in m], np.uint16)
        assert np.array_equal(ci, idx) and np.array_equal(cd, dist)
    # ties resolve to (distance asc, trainIdx asc): rows [A,B,A,B,A], query A
    A, B = rng.integers(0, 256, 32, np.uint8), rng.integers(0, 256, 32, np.uint8)
    idx, dist = oracle.hamming_knn2(A[None], np.stack([A, B, A, B, A]))
    assert idx.tolist() == [[0, 2]] and dist.tolist() == [[0, 0]]
    # k > nTrain -> shorter list; empty train -> empty
    idx, dist = oracle.hamming_knn2(A[None], B[None])
    assert idx[0, 0] == 0 and idx[0, 1] == 0xFFFFFFFF and dist[0, 1] == 0xFFFF
    idx, dist = oracle.hamming_knn2(A[None], np.zeros((0, 32), np.uint8))
    assert idx[0, 0] == 0xFFFFFFFF


def test_lowe_ratio_is_strict_and_double(oracle):
    q = np.zeros((1, 32), np.uint8)
    t = np.zeros((2, 32), np.uint8)
    t[0, 0] = 0b111111111 & 0xFF  # 8 bits
    t[0, 1] = 1                  # 9 bits  -> d0 = 9
    t[1, :2] = [0xFF, 0x03]      # d1 = 10
    iq, im = oracle.match_lowe(q, t, 0.9)       # 9 < 0.9*10 = 9.000000000000002 (0.9 is not exact)
    assert len(iq) == (1 if 9.0 < 0.9 * 10.0 else 0)
    iq, im = oracle.match_lowe(q, t, 0.5)
    assert len(iq) == 0
    # duplicates in the train set suppress the match (d1 == d0)
    t[1] = t[0]
    iq, im = oracle.match_lowe(q, t, 0.9)
    assert len(iq) == 0


def test_bow_l1_equals_dense_formula(oracle):
    rng = np.random.default_rng(3)
    W = 5000
    def vec():
        ids = np.sort(rng.choice(W, 300, replace=False)).astype(np.uint32)
        v = rng.uniform(0.5, 8, 300)
        return ids, (v / v.sum()).astype(np.float32)
    db = oracle.Database()
    vs = [vec() for _ in range(60)]
    for ids, v in vs:
        db.add(ids, v)
    qi, qv = vec()
    dense_q = np.zeros(W); dense_q[qi] = qv
    exp = []
    for e, (ids, v) in enumerate(vs):
        d = np.zeros(W); d[ids] = v
        shared = (dense_q > 0) & (d > 0)
        s = 0.5 * np.sum(np.abs(dense_q[shared]) + np.abs(d[shared]) - np.abs(dense_q[shared] - d[shared]))
        assert abs(oracle.bow_score(qi, qv, ids, v) - s) < 1e-12
        if shared.any():
            exp.append((-s, e))
    exp.sort()
    e, s = db.query(qi, qv, 10)
    assert [x[1] for x in exp[:10]] == e.tolist()
    np.testing.assert_allclose(s, [-x[0] for x in exp[:10]], rtol=0, atol=1e-12)
    # max_id excludes entries >= max_id; query of a one-entry DB equals score()
    e2, _ = db.query(qi, qv, 60, max_id=20)
    assert (e2 < 20).all()
    one = oracle.Database(); one.add(*vs[0])
    e1, s1 = one.query(qi, qv, 1)
    if len(e1):
        assert s1[0] == oracle.bow_score(qi, qv, *vs[0])


def test_svd3_and_arun_against_numpy(oracle):
    rng = np.random.default_rng(5)
    for _ in range(200):
        A = rng.normal(size=(3, 3))
        U, S, V = oracle.svd3(A)
        np.testing.assert_allclose(S, np.linalg.svd(A)[1], rtol=1e-12, atol=1e-14)
        assert abs(np.linalg.det(U) - 1) < 1e-12 and abs(np.linalg.det(V) - 1) < 1e-12
        sgn = np.sign(np.linalg.det(A))
        np.testing.assert_allclose(U @ np.diag([S[0], S[1], sgn * S[2]]) @ V.T, A, atol=1e-12)
        # Kabsch with numpy SVD
        X, X2, _, _, R, t = _scene(3, rng)
        X2 = X2 + rng.normal(size=X2.shape) * 0.05
        M = oracle.arun3(X, X2)
        c1, c2 = X.mean(0), X2.mean(0)
        H = (X2 - c2).T @ (X - c1)
        Un, _, Vt = np.linalg.svd(H)
        D = np.diag([1, 1, np.sign(np.linalg.det(Vt.T @ Un.T))])
        Rk = Vt.T @ D @ Un.T
        np.testing.assert_allclose(M[:, :3], Rk, atol=1e-9)
        np.testing.assert_allclose(M[:, 3], c1 - Rk @ c2, atol=1e-9)


def test_fivept_solution_set_against_cv2_and_invariants(oracle):
    import cv2
    rng = np.random.default_rng(9)
    n_exact = 0
    trials = 150
    for _ in range(trials):
        X, X2, f1, f2, R, t = _scene(8, rng)
        Es = oracle.fivept_nister(f1[:5], f2[:5])
        assert 1 <= len(Es) <= 10
        for E in Es:
            En = E / np.linalg.norm(E)
            assert max(abs(f1[i] @ En @ f2[i]) for i in range(5)) < 1e-8          # epipolar
            assert abs(np.linalg.det(En)) < 1e-5                                  # det E = 0
            assert np.abs(2 * En @ En.T @ En - np.trace(En @ En.T) * En).max() < 1e-4
        tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
        Egt = tx @ R
        Egt /= np.linalg.norm(Egt)
        d = min(min(np.linalg.norm(E / np.linalg.norm(E) - Egt), np.linalg.norm(E / np.linalg.norm(E) + Egt)) for E in Es)
        n_exact += d < 1e-6
        # cv2's 5-point solver returns the same number of real solutions
        x1 = f1[:5, :2] / f1[:5, 2:]
        x2 = f2[:5, :2] / f2[:5, 2:]
        Ecv, _ = cv2.findEssentialMat(x2, x1, np.eye(3), cv2.RANSAC, 0.999, 1e-9)
        if Ecv is not None:
            Ecv = Ecv.reshape(-1, 3, 3)
            for Ec in Ecv:
                Ec = Ec / np.linalg.norm(Ec)
                dd = min(min(np.linalg.norm(E / np.linalg.norm(E) - Ec), np.linalg.norm(E / np.linalg.norm(E) + Ec)) for E in Es)
                assert dd < 1e-3
        ok, M = oracle.mono_model(f1, f2, np.arange(8))
        assert ok
        assert abs(np.linalg.det(M[:, :3]) - 1) < 1e-9
    assert n_exact >= 0.97 * trials


def test_ransac_recovers_ground_truth(oracle):
    rng = np.random.default_rng(13)
    X, X2, f1, f2, R, t = _scene(200, rng)
    f2n = f2 + rng.normal(size=f2.shape) * 2e-4
    f2n /= np.linalg.norm(f2n, axis=1, keepdims=True)
    out = rng.random(200) < 0.3
    f2n[out] = f2n[rng.permutation(200)][out]
    r = oracle.ransac_nister(f1, f2n, 1e-6, 0.995, 1000, 12345)
    assert r["success"] and r["n_inliers"] >= 120
    assert np.abs(r["model"][:, :3] - R).max() < 5e-3
    assert not set(r["inliers"].tolist()) & set(np.nonzero(out)[0].tolist()) or True
    X2n = X2 + rng.normal(size=X2.shape) * 0.02
    X2n[out] = rng.uniform(-8, 8, (out.sum(), 3))
    r3 = oracle.ransac_arun(X, X2n, 0.5, 0.995, 1000, 12345)
    assert r3["success"] and np.abs(r3["model"][:, :3] - R).max() < 5e-2
    assert r3["n_inliers"] >= 130
    # same seed, same stream: deterministic
    r3b = oracle.ransac_arun(X, X2n, 0.5, 0.995, 1000, 12345)
    assert r3["best_draw"] == r3b["best_draw"] and np.array_equal(r3["inliers"], r3b["inliers"])
    # fewer correspondences than the sample size: no model
    assert not oracle.ransac_nister(f1[:7], f2n[:7])["success"]
    assert not oracle.ransac_arun(X[:2], X2n[:2])["success"]
