"""Committed known answers (tests/golden/golden_v1.npz, written by tests/golden/make_golden.py from
OpenCV's BFMatcher / 5-point solver, numpy's MT19937 and LAPACK) against the CPU oracle and, on a
GPU box, against the CUDA path through the C ABI.  Nothing here needs cv2 at run time."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz"))


def test_oracle_matches_bfmatcher_vectors(oracle):
    idx, dist = oracle.hamming_knn2(G["knn_q"], G["knn_t"])
    assert np.array_equal(idx, G["ham_idx"]) and np.array_equal(dist, G["ham_dist"])
    idx, dist = oracle.l1_knn2(G["knn_q"], G["knn_t"])
    assert np.array_equal(idx, G["l1_idx"]) and np.array_equal(dist, G["l1_dist"])
    iq, im = oracle.match_lowe(G["knn_q"], G["knn_t"], 0.8)
    assert np.array_equal(np.c_[iq, im], G["lowe08_pairs"])


def test_oracle_sampler_matches_mt19937_vectors(oracle):
    for S in (8, 3):
        exp = G["fy23_s%d" % S]
        got = oracle.sample_stream(23, S, 12345, len(exp))
        assert np.array_equal(got, exp)
    # the raw generator: first draw of a 1-element sample over a huge range is raw >> 1
    for seed in (12345, 1, 5489):
        raw = G["mt_raw_%d" % seed]
        n = 65535  # uint16 sample indices
        got = oracle.sample_stream(n, 1, seed, 1)
        assert int(got[0, 0]) == int(raw[0] >> 1) % n


def test_oracle_linear_algebra_matches_lapack_vectors(oracle):
    for k in range(len(G["svd_A"])):
        U, S, V = oracle.svd3(G["svd_A"][k])
        assert np.abs(np.sort(S)[::-1] - G["svd_S"][k]).max() < 1e-12
        sgn = np.sign(np.linalg.det(G["svd_A"][k]))  # U, V are proper rotations: the sign sits on s2
        assert np.abs(U @ np.diag([S[0], S[1], sgn * S[2]]) @ V.T - G["svd_A"][k]).max() < 1e-12
        M = oracle.arun3(G["arun_p1"][k], G["arun_p2"][k])
        assert np.abs(M[:, :3] - G["arun_R"][k]).max() < 1e-9
        assert np.abs(M[:, 3] - G["arun_t"][k]).max() < 1e-9


def test_oracle_fivept_matches_cv2_solution_sets(oracle):
    """Same number of real solutions as OpenCV's 5-point solver on every case, and every OpenCV
    solution that itself satisfies the essential-matrix constraints (its eigen-solver loses digits on
    ill-conditioned roots; ours refines each root) is in the oracle's set."""
    n_checked = 0
    for k in range(len(G["five_f1"])):
        Es = oracle.fivept_nister(G["five_f1"][k], G["five_f2"][k])
        assert len(Es) == int(G["five_nE_cv2"][k]), k
        for i in range(int(G["five_nE_cv2"][k])):
            Ec = G["five_E_cv2"][k][i]
            if abs(np.linalg.det(Ec)) > 1e-7 or np.abs(2 * Ec @ Ec.T @ Ec - np.trace(Ec @ Ec.T) * Ec).max() > 1e-6:
                continue
            d = min(min(np.linalg.norm(E / np.linalg.norm(E) - Ec), np.linalg.norm(E / np.linalg.norm(E) + Ec))
                    for E in Es)
            assert d < 1e-4, (k, i, d)
            n_checked += 1
    assert n_checked >= 150


@pytest.mark.gpu
def test_cuda_matcher_matches_bfmatcher_vectors():
    import kml
    det = kml.LoopClosureDetector()
    idx, dist, _ = det.hamming_knn2(G["knn_q"], G["knn_t"])
    assert np.array_equal(idx, G["ham_idx"]) and np.array_equal(dist, G["ham_dist"])
    idx, dist, _ = det.l1_knn2(G["knn_q"], G["knn_t"])
    assert np.array_equal(idx, G["l1_idx"]) and np.array_equal(dist, G["l1_dist"])
    # computeMatchedIndices on two stored frames: BFMatcher + Lowe 0.8 (double, strict)
    p = kml.default_params()
    p.lowe_ratio = 0.8
    det2 = kml.LoopClosureDetector(params=p)
    nq, nt = len(G["knn_q"]), len(G["knn_t"])
    det2.addVLCFrame(0, 1, G["knn_q"], np.zeros((nq, 3)), np.zeros((nq, 3)))
    det2.addVLCFrame(1, 2, G["knn_t"], np.zeros((nt, 3)), np.zeros((nt, 3)))
    iq, im = det2.computeMatchedIndices(0, 1, 1, 2)
    assert np.array_equal(np.c_[iq, im], G["lowe08_pairs"])
    det.close(); det2.close()
